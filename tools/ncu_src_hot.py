"""Aggregates warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = []
for r in rows[3:]:
    if not r or not r[0].strip().isdigit():
        continue
    for i in range(1, len(r) - 3):
        if r[i] == '-' and r[i + 1] == '-' and r[i + 2].isdigit():
            agg.append((int(r[i + 2]), int(r[0]), ','.join(r[1:i]).strip()[:100], r[i + 5] if i + 5 < len(r) else ''))
            break
tot = sum(a[0] for a in agg) or 1
print("total samples", tot)
for s, l, src, inst in sorted(agg, reverse=True)[:top]:
    print(f"{s:7d} {100*s/tot:5.1f}%  L{l:>4} inst={inst:>10} {src}")
