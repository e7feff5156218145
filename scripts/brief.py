import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print({k: (round(d[k], 3) if isinstance(d[k], float) else d[k]) for k in ("n_gpus", "value", "ms_per_step", "mass_drift") if k in d},
          d.get("config", {}).get("workload"), "| e2e", round(d.get("e2e", {}).get("value", 0)), "| step frac", round(d.get("roofline", {}).get("step", d.get("roofline", {})).get("frac", 0), 3))
