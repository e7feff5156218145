#!/bin/bash
OUT=gpurun_out/r02_n2b; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_worker.py > $OUT/worker.log 2>&1
grep -v "^W1017\|OMP_NUM_THREADS\|^\*\*\*" $OUT/worker.log | grep -B30 -m1 "Error\|error" | head -60
tail -3 $OUT/worker.log
echo "== done"
