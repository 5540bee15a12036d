"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (development tool).
usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= iv:
        continue
    k = r[ik].split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} launches={c:4d} total={t / 1e6:10.3f} ms  share={100 * t / tot:5.1f}%")
