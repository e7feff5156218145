#!/bin/bash
# Round 2, call 3: chain kernels - memcheck of the small case, bit comparison, cycle accounting
OUT=gpurun_out/r02_03; mkdir -p $OUT
echo "== dbg nx=8 under compute-sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dbg_chain.py 8 2>&1 | tee $OUT/memcheck8.log | tail -25
echo "== dbg nx=32"; timeout 300 python scripts/dbg_chain.py 32 2>&1 | tee $OUT/dbg32.log | tail -8
echo "== dbg nx=512"; LAGS=12,24,48 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | tail -30
echo "== tma bit test"; EVX_FFT_CHAIN=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tma_tiled" -p no:cacheprovider 2>&1 | tail -4
echo "== done"
