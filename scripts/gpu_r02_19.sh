#!/bin/bash
OUT=gpurun_out/r02_19; mkdir -p $OUT
echo "== rhs tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "ch_golden or ch_rhs or ch_step_512 or chained" -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu --no-extras --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['clocks'].get('sm_mhz_timed_region'), {k.split(' ')[0]: round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"
done
echo "== done"
