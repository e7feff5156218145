#!/bin/bash
# Round 2, call 2: TMA-vs-cp.async diagnosis, L2 prefetch variants, first run of the chained z/y kernels
OUT=gpurun_out/r02_02; mkdir -p $OUT
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -12
echo "== diag"; EVX_FFT_CHAIN=0 timeout 600 python scripts/diag_tma.py 2>&1 | tee $OUT/diag.log | tail -40
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 50 2>$OUT/bench_$tag.err > $OUT/bench_$tag.json
  python - $OUT/bench_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms_per_step", round(d["ms_per_step"],4), {k.split(" ")[0]: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(sys.argv[2], "no bench line:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
}
run base EVX_FFT_CHAIN=0
run pf1 EVX_FFT_CHAIN=0 EVX_FFT_TMA_PF=1
run pf2 EVX_FFT_CHAIN=0 EVX_FFT_TMA_PF=2
run pf4 EVX_FFT_CHAIN=0 EVX_FFT_TMA_PF=4
run pf2_z740 EVX_FFT_CHAIN=0 EVX_FFT_TMA_PF=2 EVX_FFT_Z_PF=740
run pf2_z2960 EVX_FFT_CHAIN=0 EVX_FFT_TMA_PF=2 EVX_FFT_Z_PF=2960
run chain EVX_FFT_CHAIN=1 EVX_FFT_TMA_PF=2
run chain_lag4 EVX_FFT_CHAIN=1 EVX_FFT_TMA_PF=2 EVX_FFT_CHAIN_LAG=4
run chain_lag24 EVX_FFT_CHAIN=1 EVX_FFT_TMA_PF=2 EVX_FFT_CHAIN_LAG=24
run chain_a0 EVX_FFT_CHAIN=1 EVX_FFT_TMA_PF=2 EVX_FFT_CHAIN_AHEAD=0
run chain_a1 EVX_FFT_CHAIN=1 EVX_FFT_TMA_PF=2 EVX_FFT_CHAIN_AHEAD=1
echo "== done"
