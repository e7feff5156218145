#!/bin/bash
# Scaling session on one multi-GPU box: distributed parity at every N, then bench N=1,2,4,8.
TAG=${1:-scale}; NMAX=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== distributed parity"; timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_dist.log | tail -6
for n in 1 2 4 8; do
  if [ $n -le $NMAX ]; then
    for tr in ce p2p nccl; do
      if [ $n -eq 1 ] && [ $tr != ce ]; then continue; fi
      echo "== bench --gpus $n --transport $tr"
      if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu 2>$OUT/bench1.err | tee $OUT/bench_n1.json | python scripts/brief.py
      else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 50 --warmup 5 --transport $tr 2>$OUT/bench${n}_$tr.err | tee $OUT/bench_n${n}_$tr.json | python scripts/brief.py; tail -2 $OUT/bench${n}_$tr.err | cut -c1-300; fi
    done
  fi
done
echo "== done"
