#!/bin/bash
# First GPU call of the next round: does the L2-blocked schedule pay?
#   1. gpu tests of the schedule (bit-identity, graph capture, tuner)
#   2. every tuner candidate's time at 512^3 (EVX_TUNE_VERBOSE) + the bench line, tuned and untuned
#   3. ncu DRAM bytes + duration per launch of ONE step under a forced chunked schedule
#      (--cache-control none: ncu must not flush L2 between launches, that is the effect measured)
#      (the y pass of a chunk should read ~0 bytes from DRAM if the L2 blocking works)
# Usage (under gpurun, repo root):  bash scripts/gpu_round2.sh [tag] ["X,streams,flags"]
TAG=${1:-r02_sched}; FORCED=${2:-16,1,3}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest schedule"; timeout 900 python -m pytest tests/test_gpu_schedule.py -q -s --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_schedule.log | tail -8
echo "== bench tuned"; EVX_TUNE_VERBOSE=1 timeout 600 python bench.py --no-cpu 2>$OUT/bench_tuned.err | tee $OUT/bench_tuned.log | grep -v '^{' | sort -t' ' -k5 -n | head -40
grep '^{' $OUT/bench_tuned.log | cut -c1-900
echo "== bench untuned"; EVX_TUNE=0 timeout 600 python bench.py --no-cpu 2>$OUT/bench_untuned.err | tee $OUT/bench_untuned.json | cut -c1-300
echo "== ncu dram bytes per launch, schedule $FORCED"
EVX_SCHEDULE=$FORCED timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
  --clock-control none --cache-control none --profile-from-start off --csv --log-file $OUT/sched_dram.csv python scripts/profile_step.py 512 > $OUT/ncu_sched.log 2>&1
python scripts/summarize_dram.py $OUT/sched_dram.csv | tee $OUT/sched_dram_summary.txt | head -40
echo "== same, one launch per pass"
EVX_SCHEDULE=0,1,0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
  --clock-control none --cache-control none --profile-from-start off --csv --log-file $OUT/serial_dram.csv python scripts/profile_step.py 512 > $OUT/ncu_serial.log 2>&1
python scripts/summarize_dram.py $OUT/serial_dram.csv | tee $OUT/serial_dram_summary.txt | head -20
echo "== done"
