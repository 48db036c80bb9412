"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_list_summary.py
<launches.csv> <out.csv> "<title>" "<command>"."""
import collections
import csv
import re
import sys

src, dst, title, cmd = sys.argv[1:5]
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.reader(open(src, errors="ignore")):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n# command: {cmd}\n# per-launch times are cold-cache and serialised: compare SHARES with bench.py's per_class_ms, not absolutes\n")
    f.write(f"# total device time of the {sum(v[0] for v in agg.values())} launches: {tot / 1e3:.1f} ms\nkernel,launches,total_us,share,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f},{v[1] / v[0]:.2f}\n")
print(open(dst).read()[:3000])
