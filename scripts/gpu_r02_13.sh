#!/bin/bash
OUT=gpurun_out/r02_13; mkdir -p $OUT
echo "== dbg nx=512"; LAGS=24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | grep -E "^lag|fwd lag|inv lag|chain =="
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -4
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
run chain EVX_FFT_CHAIN=1
run sep EVX_FFT_CHAIN=0
echo "== done"
