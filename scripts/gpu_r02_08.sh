#!/bin/bash
OUT=gpurun_out/r02_08; mkdir -p $OUT
echo "== seq nx=8 (memcheck)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -X faulthandler scripts/dbg_chain_seq.py 8 2>&1 | tail -8
echo "== dbg nx=512"; LAGS=6,12,24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512.log | tail -30
echo "== chain test"; EVX_FFT_CHAIN=1 timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -q -x -k "chained" -p no:cacheprovider 2>&1 | tee $OUT/pytest_chain.log | tail -5
echo "== done"
