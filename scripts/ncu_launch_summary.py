"""Per-kernel-family totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list.

    python scripts/ncu_launch_summary.py gpurun_out/launches.csv
"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    rows.append((r["Kernel Name"], us))
fam = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    key = re.sub(r"_[0-9a-f]{12}$", "", name)
    fam[key][0] += 1
    fam[key][1] += us
total = sum(v[1] for v in fam.values())
print(f"Total device time {total / 1e3:.1f} ms over {len(rows)} launches.\n")
print("| kernel family | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {us / 1e3:.2f} | {100 * us / total:.1f}% | {us / n:.1f} |")
