"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time, share."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    if "gpu__time_duration" not in r[ix["Metric Name"]]:
        continue
    name = r[ix["Kernel Name"]].split("(")[0][:90]
    tot[name] += float(r[ix["Metric Value"]].replace(",", "")) / 1e6
    cnt[name] += 1
total = sum(tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "mean_us", "share_of_all_kernel_time"])
for k in sorted(tot, key=lambda k: -tot[k]):
    w.writerow([k, cnt[k], "%.3f" % tot[k], "%.1f" % (1e3 * tot[k] / cnt[k]), "%.4f" % (tot[k] / total)])
