"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share.
    python scripts/launch_summary.py gpurun_out/launches_r02.csv "<command>" > profiles/r02_launches.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    us = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
    t = tot[r[ik]]
    t[0] += 1
    t[1] += us
total = sum(t[1] for t in tot.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none : {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"{sum(t[0] for t in tot.values())} launches, {total:.1f} us total; cold-cache, serialised per-launch times: compare SHARES with bench.py's live numbers, not absolutes\n")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us {100 * us / total:5.1f}% x{n:3d} avg {us / n:8.2f} us  {name[:120]}")
