#!/bin/bash
OUT=gpurun_out/r02_07; mkdir -p $OUT
echo "== dbg nx=512"; LAGS=12,24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | tail -30
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -5
echo "== done"
