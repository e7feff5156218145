#!/bin/bash
# quick kernel timing session: AC parity tests, then bench_kernels with the two AC variants
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ac or allen or AC or rk4 or euler" 2>&1 | tail -4
for v in 1 4; do echo "AC_V=$v"; EVX_AC_V=$v timeout 600 python scripts/bench_kernels.py 512 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d.items() if k.startswith('ac_')})"; done
timeout 600 python scripts/bench_kernels.py 512 | tail -1 | tee $OUT/kernels.json | cut -c1-900
