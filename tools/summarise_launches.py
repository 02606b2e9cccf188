"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python tools/summarise_launches.py launches.csv [skip_first_n] > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = name.replace("vb200::", "")
    rows.append((name + " grid=" + r.get("Grid Size", "?"), us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(lambda: [0.0, 0])
for n, us in rows:
    agg[n][0] += us
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"# {len(rows)} launches, {tot:.1f} us in total (each launch serialised and cold under ncu: shares, not absolutes)")
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:12.1f} us {100 * us / tot:5.1f}%  n={c:4d} avg={us / c:9.2f} us  {n}")
