"""Top stall sites of an `ncu --page source --csv --print-source sass` export (per SASS instruction)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
si, src, ex = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source'), hdr.index('Instructions Executed')
L, tot = [], 0
for i, r in enumerate(rows[2:]):
    try:
        v = int(r[si])
    except Exception:
        continue
    tot += v
    L.append((v, i, r[src].strip()[:100], r[ex]))
print("kernel:", rows[0][1], "total samples:", tot, "instructions:", len(L))
for v, i, s, e in sorted(L, reverse=True)[:top]:
    print(f"{v:6d} {100*v/tot:5.1f}%  #{i:5d} x{e:>8}  {s}")
