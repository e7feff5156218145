#!/bin/bash
# One GPU-box session: smoke, gpu tests, bench, ncu launch list, ncu full capture of every
# kernel on the hot path.  Usage (from repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -6
echo "== kernels"; timeout 300 python scripts/bench_kernels.py 512 2>&1 | tail -1 | tee $OUT/kernels.json | cut -c1-1500
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-600
tail -3 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt | head -14
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ch_rhs_kernel|fft_pass_kernel|fft_pipe_kernel|ac_tile_kernel|rd_rhs_kernel' -f -o $OUT/prof python scripts/profile_kernels.py 512 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
echo "== done"
