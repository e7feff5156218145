#!/bin/bash
# ncu full capture of the shipped 1024-point passes (y: eight-point form, x: sixteen-point form)
OUT=gpurun_out/r02_24b; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_line4 -s 6 -c 3 -f -o $OUT/line16 \
   python scripts/exp/line4_ncu_target.py > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $OUT/line16.ncu-rep --page raw --csv > $OUT/line16_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/line16_raw.csv | tee $OUT/line16_summary.txt
python - $OUT/line16_raw.csv <<'PY' | tee $OUT/line16_smem.txt
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[0]; kn=h.index('Kernel Name')
cols=['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','memory_l1_wavefronts_shared_ideal','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','gpu__time_duration.sum','smsp__inst_executed.sum','launch__registers_per_thread']
for r in rows[2:]:
    print(r[kn][:70], ' '.join('%s=%s' % (c.split('.')[0].split('__')[-1], r[h.index(c)]) for c in cols if c in h))
PY
