#!/bin/bash
OUT=gpurun_out/r02_18; mkdir -p $OUT
for tw in 3 1 2 0; do
echo "== TW=$tw"; EVX_FFT_CHAIN_TW=$tw LAGS=24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512_tw$tw.log | grep -E "^lag|inverse chain"
done
echo "== done"
