#!/bin/bash
# quick kernel timing session
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 600 python scripts/bench_kernels.py 512 | tail -1 | tee $OUT/kernels.json | cut -c1-1400
