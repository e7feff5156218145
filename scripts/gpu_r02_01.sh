#!/bin/bash
# Round 2, call 1: full GPU suite on the committed TMA-tiled passes, A/B bench, ncu capture.
OUT=gpurun_out/r02_01; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -8
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
echo "== ncu full capture (TMA on, default KZ)"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ch_rhs_kernel|fft_pass_kernel|fft_pipe_kernel|fft_line_kernel|ac_tile_kernel' -f -o $OUT/prof python scripts/profile_kernels.py 512 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
echo "== done"
