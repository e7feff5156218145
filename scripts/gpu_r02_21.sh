#!/bin/bash
# N-GPU bench line (driver form) + event timeline of one step, after the 1024-point TMA-tiled passes
N=${1:-2}
OUT=gpurun_out/r02_21_n$N; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/bench_n$N.err | tail -1 > $OUT/bench_n$N.json
tail -3 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("ms_per_step","value","parity","other_configs"):
    print(k, json.dumps(d.get(k))[:600])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 scripts/dist_timeline.py 512 4 > $OUT/timeline_n$N.txt 2>$OUT/timeline.err
cat $OUT/timeline_n$N.txt
