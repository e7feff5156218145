#!/bin/bash
# N GPUs: A/B of the transports (one session) + event timeline of the default step
N=${1:-8}
OUT=gpurun_out/r02_22_n$N; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/exp/dist_ab.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee $OUT/ab_n$N.txt | tail -20
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 scripts/dist_timeline.py 512 4 > $OUT/timeline_n$N.txt 2>$OUT/timeline.err
cat $OUT/timeline_n$N.txt
