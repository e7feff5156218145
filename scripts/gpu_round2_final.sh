#!/bin/bash
# Round-2 evidence run on one B200: smoke, gpu tests, full bench line (all legs), reference arm,
# ncu launch list with DRAM bytes, ncu full capture of every hot-path kernel.
# Usage (under gpurun, from the repo root):  bash scripts/gpu_round2_final.sh [tag]
TAG=${1:-final}
OUT=gpurun_out/r02_$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -4
echo "== bench (all legs)"; timeout 1500 python bench.py 2>$OUT/bench.err > $OUT/bench.json; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("ms_per_step","value","clocks","e2e","cpu_baseline","reference_gpu","other_configs"):
    print(k, json.dumps(d.get(k))[:700])
print({k.split(" ")[0]: (round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items()})
print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","frac","traffic")}, d["roofline"]["step"])
PY
echo "== bench --steps 20 --warmup 5 (driver form)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['clocks'])"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
echo "== kernels"; timeout 300 python scripts/bench_kernels.py 512 2>&1 | tail -1 | tee $OUT/kernels.json | cut -c1-1200
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt | head -8
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ch_rhs_kernel|ch_rhs_tma_kernel|fft_pass_kernel|fft_pipe_kernel|fft_line_kernel|fft_line_ws_kernel|fft_chain_kernel|ac_tile_kernel|rd_rhs_kernel' -f -o $OUT/prof python scripts/profile_kernels.py 512 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
python scripts/ncu_traffic.py $OUT/prof_raw.csv "profiles/r02_ncu_full_summary_final.txt (ncu --set full --clock-control none, one launch per kernel, scripts/profile_kernels.py 512; scripts/gpu_round2_final.sh)" > /dev/null
echo "== done"
