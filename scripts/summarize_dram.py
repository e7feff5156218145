"""Per-kernel totals of an `ncu --csv` log with duration / DRAM bytes / L2 hit rate metrics."""
import csv, sys, collections

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0][:60]
    key = (r["ID"], name)
    agg.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * \
        {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "usecond": 1e-6, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1, "%": 1}.get(r["Metric Unit"], 1)
tot = collections.OrderedDict()
for (_id, name), m in agg.items():
    t = tot.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    t[0] += 1
    t[1] += m.get("gpu__time_duration.sum", 0.0)
    t[2] += m.get("dram__bytes_read.sum", 0.0)
    t[3] += m.get("dram__bytes_write.sum", 0.0)
    t[4] += m.get("lts__t_sector_hit_rate.pct", 0.0)
print(f"{'kernel':60s} {'n':>4s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'L2 hit %':>8s}")
for name, (n, s, rd, wr, hit) in tot.items():
    print(f"{name:60s} {n:4d} {s * 1e6:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {hit / max(n, 1):8.1f}")
print(f"{'total':60s} {sum(t[0] for t in tot.values()):4d} {sum(t[1] for t in tot.values()) * 1e6:9.1f} "
      f"{sum(t[2] for t in tot.values()) / 1e6:9.1f} {sum(t[3] for t in tot.values()) / 1e6:9.1f}")
