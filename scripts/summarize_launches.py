"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv` launch list per kernel name (time share and, when present, DRAM bytes per launch)."""
import csv
import re
import sys
from collections import defaultdict

with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
UNIT_T = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        agg[name]["n"] += 1
        agg[name]["us"] += val * UNIT_T.get(unit, 1e-3)
    elif m == "dram__bytes_read.sum":
        agg[name]["rd"] += val * UNIT_B.get(unit, 1.0)
    elif m == "dram__bytes_write.sum":
        agg[name]["wr"] += val * UNIT_B.get(unit, 1.0)
total = sum(v["us"] for v in agg.values()) or 1.0
print(f"{'kernel':90s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s} {'dram_rd_MB/launch':>18s} {'dram_wr_MB/launch':>18s}")
for name, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = max(v["n"], 1)
    print(f"{name:90s} {v['n']:8d} {v['us']:12.1f} {v['us']/n:10.1f} {100*v['us']/total:6.1f}% "
          f"{v['rd']/n/1e6:18.1f} {v['wr']/n/1e6:18.1f}")
