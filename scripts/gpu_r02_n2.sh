#!/bin/bash
# Round 2: two GPUs - distributed parity tests and the bench line with its parity / config-3 records
OUT=gpurun_out/r02_n2; mkdir -p $OUT
nvidia-smi -L | head -3
echo "== pytest distributed"; timeout 600 python -m pytest tests/test_gpu_distributed.py -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>$OUT/bench_n2.err > $OUT/bench_n2.json
tail -5 $OUT/bench_n2.err
python - $OUT/bench_n2.json <<'PY'
import json,sys
txt=open(sys.argv[1]).read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith("{")][-1])
for k in ("ms_per_step","value","clocks","parity","other_configs","e2e","config"):
    print(k, json.dumps(d.get(k))[:600])
PY
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-500
echo "== done"
