#!/bin/bash
# One GPU-box session: smoke, gpu tests, bench, ncu launch list (+ optional full capture).
# Usage (from repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [ncu-kernel-regex]
TAG=${1:-run}
KREGEX=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -40
echo "== kernels"; for t in 14 16 30; do EVX_CH_TILE=$t timeout 300 python scripts/bench_kernels.py 512 2>&1 | tail -1 | tee -a $OUT/kernels.jsonl | cut -c1-700; done
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-1500
tail -5 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt | head -40
if [ -n "$KREGEX" ]; then
  echo "== ncu full capture of $KREGEX"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 4 -c 2 -f -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_full.log 2>&1
  tail -3 $OUT/ncu_full.log
fi
echo "== done"
