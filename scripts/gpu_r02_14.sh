#!/bin/bash
OUT=gpurun_out/r02_14; mkdir -p $OUT
for tw in 0 1 2; do
echo "== TW=$tw"; EVX_FFT_CHAIN_TW=$tw LAGS=24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512_tw$tw.log | grep -E "^lag|fwd lag|inv lag"
done
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tail -3
echo "== done"
