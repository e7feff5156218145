#!/bin/bash
# quick kernel timing session: launch list of one bench_kernels run + env-var variants
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
for occ in 1 2; do echo "AC_OCC=$occ"; EVX_AC_OCC=$occ python scripts/bench_kernels.py 512 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d.items() if k.startswith('ac_')})"; done
python scripts/bench_kernels.py 512 | tail -1 | tee $OUT/kernels.json | cut -c1-900
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt | head -12
