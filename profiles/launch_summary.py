"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0) if r[iu] in ("ns", "us", "ms") else 1e-3
    agg[r[ik]][0] += 1
    agg[r[ik]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms of device time (per-launch times are cold-cache/serialised: compare SHARES)")
print("share  launches  avg_us  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:5.1f}%  {n:6d}  {t / n:9.1f}  {k[:110]}")
