#!/bin/bash
# ncu full captures of the library's kernels during a short bench run.
# Usage: bash scripts/gpu_profile.sh <tag> <kernel-regex> [count]
TAG=${1:-prof}; KREGEX=${2:-fft_pass_kernel}; CNT=${3:-6}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $CNT -c $CNT -f -o $OUT/prof \
   python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
