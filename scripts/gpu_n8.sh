#!/bin/bash
# 8-GPU session: bench lines of the 'ce' transport for several pipeline depths
OUT=gpurun_out/${1:-n8}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for cfg in "4 1" "2 1" "4 2"; do set -- $cfg
  echo "== bench ce --overlap-chunks $1 --mid-chunks $2"
  timeout 200 $TR --master-port 29513 bench.py --gpus 8 --steps 30 --warmup 5 --transport ce --overlap-chunks $1 --mid-chunks $2 --no-cpu 2>$OUT/bench_$1_$2.err | tee $OUT/bench_n8_ce_$1_$2.json | python scripts/brief.py
done
