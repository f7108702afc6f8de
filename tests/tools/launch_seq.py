"""Development aid: per-kernel times of an `ncu --metrics gpu__time_duration.sum --csv` launch list.   python tests/tools/launch_seq.py file.csv [n_tail]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    seq.append((r[ki].split('(')[0].split('::')[-1][:28], v))
d = collections.defaultdict(list)
for n, v in seq: d[n].append(v)
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])): print("%-30s n=%4d mean %9.1f us  min %9.1f  max %9.1f" % (k, len(v), sum(v) / len(v), min(v), max(v)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("tail:", [(a, round(b)) for a, b in seq[-n:]])
