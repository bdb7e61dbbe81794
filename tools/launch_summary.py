"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{k[:64]:64s} n={len(v):4d} mean={sum(v) / len(v) / 1000:9.1f}us tot={sum(v) / 1e6:8.2f}ms")
