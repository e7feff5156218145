#!/bin/bash
# Round-2 final multi-GPU bench line, launched exactly like the driver does (N ranks under torchrun).
N=${1:-8}
OUT=gpurun_out/r02_n${N}_final; mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/bench_n$N.err | tail -1 > $OUT/bench_n$N.json
tail -3 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("ms_per_step","value","clocks","parity","other_configs","mass_drift"):
    print(k, json.dumps(d.get(k))[:900])
PY
