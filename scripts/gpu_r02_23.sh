#!/bin/bash
# ncu full capture of the four-stage TMA-tiled kernel (1 GPU)
OUT=gpurun_out/r02_23; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_line4 -s 6 -c 3 -f -o $OUT/line4 \
   python scripts/exp/line4_ncu_target.py > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $OUT/line4.ncu-rep --page raw --csv > $OUT/line4_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/line4_raw.csv | tee $OUT/line4_summary.txt
rm -f $OUT/line4.ncu-rep.tmp
