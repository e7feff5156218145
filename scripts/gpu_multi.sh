#!/bin/bash
# Multi-GPU session: distributed parity tests + weak-scaling bench lines.
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt
echo "== pytest (all gpu tests)"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== kernels"; timeout 300 python scripts/bench_kernels.py 512 2>&1 | tail -1 | tee $OUT/kernels.json | cut -c1-900
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench --gpus $n"
    if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>$OUT/bench1.err | tee $OUT/bench_n1.json | cut -c1-400
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 2>$OUT/bench$n.err | tee $OUT/bench_n$n.json | cut -c1-600; tail -3 $OUT/bench$n.err; fi
  fi
done
echo "== done"
