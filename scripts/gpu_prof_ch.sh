#!/bin/bash
OUT=gpurun_out/${1:-pch}; mkdir -p $OUT
for occ in 2 3; do echo "CH_OCC=$occ"; EVX_CH_OCC=$occ python scripts/bench_kernels.py 512 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d.items() if k.startswith('ch_rhs')})"; done
bash scripts/gpu_profile_cmd.sh ${1:-pch} "ch_rhs_kernel|fft_pass_kernel|fft_pipe_kernel" 12 7 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 | tail -16
