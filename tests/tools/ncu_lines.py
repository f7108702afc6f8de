"""Per-source-line summary of an `ncu --set full --import-source on` capture.

    python tests/tools/ncu_lines.py report.ncu-rep [--top 40] [--ranges 300-370:solve,810-1030:explicit]

Reads `ncu -i <rep> --page source --csv --print-source cuda,sass` (first kernel of the report) and prints, per CUDA source line, the
warp-stall samples, executed warp instructions, FP64 instructions (DFMA/DADD/DMUL) and shared-memory wavefronts, plus totals per
named line range.  Used for profiles/*.md; not part of the product.
"""
import csv, subprocess, sys, collections, argparse

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    files = []          # (filename, header, rows)
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            cur = [r[1], None, []]; files.append(cur)
        elif len(r) > 4 and r[0] == "Line No":
            if cur is None:
                cur = ["?", None, []]; files.append(cur)
            cur[1] = r
        elif cur is not None and cur[1] is not None and len(r) == len(cur[1]):
            cur[2].append(r)
    return files

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("--top", type=int, default=30); ap.add_argument("--ranges", default="")
    ap.add_argument("--file", default="qc_kernel_impl.cuh"); ap.add_argument("--sort", default="samples")
    a = ap.parse_args()
    files = load(a.rep)
    tot = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    stall_names = None
    for fn, hdr, rows in files:
        idx = {}
        for i, h in enumerate(hdr):
            idx.setdefault(h, i)
        src_i = [i for i, h in enumerate(hdr) if h == "Source"][1]
        if stall_names is None:
            stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        cur_key = (fn.split("/")[-1], -1)
        for r in rows:
            def f(k):
                try: return float(r[idx[k]])
                except Exception: return 0.0
            addr_i = idx["Address"]
            if r[addr_i] == "-":                      # CUDA-line row: aggregated samples / instructions / wavefronts of that line
                cur_key = (fn.split("/")[-1], int(r[0]) if r[0].isdigit() else -1)
                c = per[cur_key]
                c["samples"] += f("# Samples"); c["inst"] += f("Instructions Executed"); c["wf"] += f("L1 Wavefronts Shared")
                for s in stall_names: c[s] += f(s)
                continue
            key = cur_key                             # SASS row below its CUDA line: classify the opcode
            ops = r[src_i].split()
            op = ""
            if ops: op = (ops[1] if ops[0].startswith("@") and len(ops) > 1 else ops[0]).split(".")[0]
            c = per[key]
            if op in ("DFMA", "DADD", "DMUL"): c["fp64"] += f("Instructions Executed")
            if op in ("LDL", "STL"): c["spill"] += f("Instructions Executed")
    for k, c in per.items():
        for kk, v in c.items(): tot[kk] += v
    print("total samples %d inst %.4g fp64 %.4g smem_wf %.4g spill_inst %.4g" % (tot["samples"], tot["inst"], tot["fp64"], tot["wf"], tot["spill"]))
    print("stalls: " + ", ".join("%s %.1f%%" % (s[6:], 100 * tot[s] / max(tot["samples"], 1)) for s in sorted(stall_names, key=lambda s: -tot[s])[:8]))
    print("\n top lines by samples")
    for k, c in sorted(per.items(), key=lambda kv: -kv[1][a.sort])[: a.top]:
        top = sorted(stall_names, key=lambda s: -c[s])[:3]
        print("%-22s:%5d  samples %5.1f%%  inst %5.1f%%  fp64 %5.1f%%  wf %5.1f%%  spill %5.1f%%   %s" % (k[0][:22], k[1], 100 * c["samples"] / tot["samples"], 100 * c["inst"] / tot["inst"],
              100 * c["fp64"] / max(tot["fp64"], 1), 100 * c["wf"] / max(tot["wf"], 1), 100 * c["spill"] / max(tot["spill"], 1), " ".join("%s %.0f%%" % (s[6:], 100 * c[s] / max(c["samples"], 1)) for s in top)))
    if a.ranges:
        print("\n ranges (%s)" % a.file)
        for spec in a.ranges.split(","):
            rng, name = spec.split(":"); lo, hi = map(int, rng.split("-"))
            c = collections.Counter()
            for k, cc in per.items():
                if k[0] == a.file and lo <= k[1] <= hi:
                    for kk, v in cc.items(): c[kk] += v
            top = sorted(stall_names, key=lambda s: -c[s])[:4]
            print("%-14s %5d-%5d samples %5.1f%% inst %5.1f%% fp64 %5.1f%% wf %5.1f%% spill %5.1f%%  %s" % (name, lo, hi, 100 * c["samples"] / tot["samples"], 100 * c["inst"] / tot["inst"],
                  100 * c["fp64"] / max(tot["fp64"], 1), 100 * c["wf"] / max(tot["wf"], 1), 100 * c["spill"] / max(tot["spill"], 1),
                  " ".join("%s %.0f%%" % (s[6:], 100 * c[s] / max(c["samples"], 1)) for s in top)))

if __name__ == "__main__":
    main()
