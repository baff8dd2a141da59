"""Top source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[2]
li, si, wi = 0, 1, h.index("Warp Stall Sampling (All Samples)")
ii = h.index("Instructions Executed")
items = []
for r in rows[3:]:
    if len(r) <= wi or r[li] == "": continue
    try: items.append((float(r[wi]), float(r[ii]), r[li], r[si].strip()[:120]))
    except ValueError: pass
tot = sum(i[0] for i in items)
print("total samples", tot)
for v, n, ln, src in sorted(items, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{100*v/tot:5.1f}%  inst {n:10.0f}  L{ln}: {src}")
