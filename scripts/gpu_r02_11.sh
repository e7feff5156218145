#!/bin/bash
OUT=gpurun_out/r02_11; mkdir -p $OUT
for nb in 4 5 3; do
echo "== dbg nx=512 NBUF=$nb"; EVX_FFT_CHAIN_NBUF=$nb LAGS=16,24,40 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512_$nb.log | grep -E "^lag|fwd lag|inv lag|chain =="
done
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -5
echo "== done"
