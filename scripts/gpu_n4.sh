#!/bin/bash
# 4-GPU session: distributed parity worker + bench line with the default transport
OUT=gpurun_out/${1:-n4}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
echo "== parity worker"; timeout 200 $TR --master-port 29511 tests/dist_gpu_worker.py 2>&1 | grep -E "DIST-GPU|Error|assert|Traceback" | head -5
echo "== bench (default transport)"
timeout 200 $TR --master-port 29513 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu 2>$OUT/bench.err | tee $OUT/bench_n4_ce.json | python scripts/brief.py
tail -2 $OUT/bench.err | cut -c1-200
