#!/bin/bash
# Round-2 check of the L2 sub-chunking inside the multi-GPU plan (evx_dist_plan_set_l2_planes).
# Usage (gpurun --gpus 2, repo root):  bash scripts/gpu_round2_n2.sh [tag]
TAG=${1:-r02_dist_l2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
for L2 in 0 16; do      # forced settings (the search is skipped)
  echo "== dist tests, EVX_DIST_L2_PLANES=$L2"
  EVX_DIST_L2_PLANES=$L2 timeout 900 python -m pytest tests/test_gpu_distributed.py -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_dist_l2_$L2.log | tail -4
done
echo "== bench --gpus $N, collective search (default)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29540 bench.py --gpus $N --steps 50 --warmup 5 2>$OUT/bench_n${N}_tuned.err | tee $OUT/bench_n${N}_tuned.json | cut -c1-400
for L2 in 0 8 16 32; do
  echo "== bench --gpus $N, EVX_DIST_L2_PLANES=$L2"
  EVX_DIST_L2_PLANES=$L2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 2>$OUT/bench_n${N}_l2_$L2.err | tee $OUT/bench_n${N}_l2_$L2.json | cut -c1-260
done
