#!/bin/bash
OUT=gpurun_out/r02_15; mkdir -p $OUT
for cfg in "3 0" "0 1" "3 1"; do
set -- $cfg
echo "== TW=$1 DISCARD=$2"; EVX_FFT_CHAIN_TW=$1 EVX_FFT_CHAIN_DISCARD=$2 LAGS=24 timeout 300 python scripts/dbg_chain.py 512 2>&1 | tee $OUT/dbg512_tw$1_d$2.log | grep -E "^lag|inverse chain"
done
echo "== ncu dram bytes with discard"
EVX_FFT_CHAIN_DISCARD=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'fft_chain_kernel' -c 8 --csv --log-file $OUT/launches_discard.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches_discard.csv | head -5
echo "== chain test with discard"; EVX_FFT_CHAIN_DISCARD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chained or ch_step_512 or readme" -p no:cacheprovider 2>&1 | tail -3
echo "== done"
