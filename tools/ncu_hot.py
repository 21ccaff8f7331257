"""Top stall locations from `ncu --page source --csv` (SASS view): usage ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
si, src = hdr.index("# Samples"), hdr.index("Source")
data = []
for i, r in enumerate(rows[2:]):
    if len(r) > si and r[si].isdigit():
        data.append((int(r[si]), i, r[src].strip()))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, i, t in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {s/tot:6.2%} idx {i:5d} {t[:110]}")
# cumulative by coarse region: print running share at barriers
acc = 0
print("--- samples between barriers (BAR.SYNC) ---")
seg = 0
for s, i, t in data:
    acc += s
    if "BAR.SYNC" in t or "WARPSYNC" in t or "EXIT" in t:
        print(f"  up to idx {i:5d} ({t[:40]:40s}): {acc - seg:7d} {100.0*(acc-seg)/tot:5.1f}%")
        seg = acc
