#!/bin/bash
OUT=gpurun_out/r02_16; mkdir -p $OUT
echo "== tma tests (ws default)"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tma_tiled or chained" -p no:cacheprovider 2>&1 | tail -3
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
run ws0 EVX_FFT_LINE_WS=0
run ws3 EVX_FFT_LINE_WS=3
run ws4 EVX_FFT_LINE_WS=4
run ws5 EVX_FFT_LINE_WS=5
run sep_ws0 EVX_FFT_CHAIN=0 EVX_FFT_LINE_WS=0
run sep_ws3 EVX_FFT_CHAIN=0 EVX_FFT_LINE_WS=3
run sep_ws4 EVX_FFT_CHAIN=0 EVX_FFT_LINE_WS=4
echo "== done"
