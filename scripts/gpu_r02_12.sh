#!/bin/bash
# Round 2, call 13: chain on/off in the full step, ncu launch list + full capture of the chained step
OUT=gpurun_out/r02_12; mkdir -p $OUT
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --no-extras --steps 50 2>$OUT/bench_$tag.err > $OUT/bench_$tag.json
  python - $OUT/bench_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms_per_step", round(d["ms_per_step"],4), d["clocks"].get("sm_mhz_timed_region"), {k.split(" ")[0]: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(sys.argv[2], "no bench line:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
}
run sep EVX_FFT_CHAIN=0
run chain EVX_FFT_CHAIN=1
run chain2 EVX_FFT_CHAIN=1 EVX_FFT_CHAIN_NBUF=2
run sep200 EVX_FFT_CHAIN=0
echo "== long runs (power cap): 2000 steps"
EVX_FFT_CHAIN=0 timeout 300 python bench.py --no-cpu --no-extras --steps 2000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sep 2000 steps', round(d['ms_per_step'],4), d['clocks'].get('sm_mhz_timed_region'), d['clocks'].get('power_w_max'))"
EVX_FFT_CHAIN=1 timeout 300 python bench.py --no-cpu --no-extras --steps 2000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chain 2000 steps', round(d['ms_per_step'],4), d['clocks'].get('sm_mhz_timed_region'), d['clocks'].get('power_w_max'))"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt | head -16
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ch_rhs_kernel|fft_pass_kernel|fft_pipe_kernel|fft_line_kernel|fft_chain_kernel' -f -o $OUT/prof python scripts/profile_kernels.py 512 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/prof_raw.csv | tee $OUT/prof_summary.txt
echo "== done"
