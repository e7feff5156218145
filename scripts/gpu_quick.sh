#!/bin/bash
# quick kernel timing session: z-pass residency variants, per-kernel times from bench.py
for v in 0 5 6; do
  echo "EVX_Z_OCC=$v"; EVX_Z_OCC=$v timeout 300 python bench.py --steps 50 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k.split(' ')[0]:round(v['ms']*1000,1) for k,v in d['roofline']['kernels'].items()})"
done
