#!/bin/bash
# Round 2, call 4: chain kernels with staged z inputs - parity test, cycle accounting, bench
OUT=gpurun_out/r02_04; mkdir -p $OUT
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -12
echo "== dbg nx=512"; LAGS=8,12,24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | tail -30
echo "== bench"; EVX_FFT_TMA_PF=1 timeout 600 python bench.py --no-cpu --no-extras --steps 100 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-1800
tail -5 $OUT/bench.err
echo "== done"
