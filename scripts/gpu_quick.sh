#!/bin/bash
# quick session on the GPU box: stencil/FFT parity tests + per-kernel times
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py 512 | tail -1 | cut -c1-1500
timeout 300 python bench.py --steps 50 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k.split(' ')[0]:round(v['ms']*1000,1) for k,v in d['roofline']['kernels'].items()})"
