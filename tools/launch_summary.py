"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share, count, mean duration."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:90]].append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v) / tot * 100:5.1f}% n={len(v):4d} avg={sum(v) / len(v) / 1000:8.1f} us  {k}")
print(f"total {tot / 1000:.1f} us over {sum(len(v) for v in d.values())} launches")
