#!/bin/bash
# quick kernel timing session: stencil parity tests, then bench_kernels
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python scripts/bench_kernels.py 512 | tail -1 | tee $OUT/kernels.json | cut -c1-900
