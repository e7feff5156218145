#!/bin/bash
# Round 2: first hardware run of the TMA-tiled strided passes (tests, then A/B timings)
OUT=gpurun_out/r02b; mkdir -p $OUT
echo "== pytest tma"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tma_tiled or ch_golden or ch_step_512 or imex_step_vs_live" -p no:cacheprovider 2>&1 | tail -15
for cfg in "0 8" "1 8" "1 16"; do
  set -- $cfg
  echo "== bench EVX_FFT_TMA=$1 KZ=$2"
  EVX_FFT_TMA=$1 EVX_FFT_TMA_KZ=$2 timeout 300 python bench.py --no-cpu --steps 50 2>$OUT/bench_$1_$2.err > $OUT/bench_$1_$2.json
  tail -3 $OUT/bench_$1_$2.err
  python - $OUT/bench_$1_$2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms_per_step", round(d["ms_per_step"],4), {k.split(" ")[0]: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("no bench line:", e)
PY
done
echo "== done"
