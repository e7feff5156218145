#!/bin/bash
# Round 2: eight GPUs - the bench line as the driver runs it (parity record, config 3) plus config 4
OUT=gpurun_out/r02_n8; mkdir -p $OUT
nvidia-smi -L | wc -l
echo "== bench --gpus 8 --config4"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --config4 2>$OUT/bench_n8.err > $OUT/bench_n8.json
tail -5 $OUT/bench_n8.err
python - $OUT/bench_n8.json <<'PY'
import json,sys
txt=open(sys.argv[1]).read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith("{")][-1])
for k in ("ms_per_step","value","clocks","parity","other_configs","e2e"):
    print(k, json.dumps(d.get(k))[:900])
PY
echo "== pytest distributed (2,4,8)"; timeout 900 python -m pytest tests/test_gpu_distributed.py -q -x --tb=short -p no:cacheprovider -k "matches_single_gpu" 2>&1 | tail -4
echo "== done"
