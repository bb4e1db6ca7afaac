"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel, launches, total ms, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ci["Kernel Name"]].split("(")[0][-60:]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(unit, 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in agg.items():
    print(f"{k:62s} {c:4d} {t:9.3f} ms {100 * t / tot:5.1f}%  ({t / c:.3f} ms each)")
