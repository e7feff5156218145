#!/bin/bash
# 2-GPU session: distributed parity tests + bench lines per transport
TAG=${1:-multi2}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== dist tests"; timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6
for tr in p2p ce; do for ch in 4 8; do
  echo "== bench --gpus $N --transport $tr --overlap-chunks $ch"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 5 --transport $tr --overlap-chunks $ch --no-cpu 2>$OUT/bench_$tr$ch.err | tee $OUT/bench_$tr$ch.json | python scripts/brief.py; tail -2 $OUT/bench_$tr$ch.err | cut -c1-300
done; done
