"""Print the SASS instructions with the most stall samples from an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index("# Samples"); so = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((int(r[si]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, k, r in sorted(data, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}%  #{k:4d} exec={r[ie]:>7s}  {r[so].strip()[:70]:70s} {st}")
