#!/bin/bash
# quick kernel timing session: general-BC CH kernel at 2 vs 3 CTAs per SM + stencil parity tests
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d.items() if k in ('ch_rhs_periodic_ms','ch_rhs_neumann_ms','ac_euler_ms','ch_step_native_ms')})"; }
for v in "EVX_CH_GOCC=2" "EVX_CH_GOCC=3"; do
  echo "$v"; env $v timeout 300 python scripts/bench_kernels.py 512 | tail -1 | pick
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
