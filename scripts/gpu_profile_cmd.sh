#!/bin/bash
# ncu full capture of kernels matching a regex while running an arbitrary command.
# Usage: bash scripts/gpu_profile_cmd.sh <tag> <kernel-regex> <skip> <count> <cmd...>
TAG=$1; KREGEX=$2; SKIP=$3; CNT=$4; shift 4
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $CNT -f -o $OUT/prof "$@" > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
