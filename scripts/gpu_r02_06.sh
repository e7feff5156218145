#!/bin/bash
# Round 2, call 6: chain kernels with the two-lines-per-group z program
OUT=gpurun_out/r02_06; mkdir -p $OUT
echo "== dbg nx=512"; LAGS=12,24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | tail -30
echo "== seq nx=8"; timeout 300 python -X faulthandler scripts/dbg_chain_seq.py 8 2>&1 | tail -4
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -v -x -s -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -40
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
run chain24 EVX_FFT_CHAIN=1 EVX_FFT_CHAIN_LAG=24
echo "== done"
