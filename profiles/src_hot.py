"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` (reads the csv on stdin)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{100.0 * n / tot:5.1f}% {r[ix['Address']][-5:]} {r[ix['Source']][:70]:70s} " + " ".join(f"{c}:{v}" for v, c in st if v))
