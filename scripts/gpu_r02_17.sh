#!/bin/bash
OUT=gpurun_out/r02_17; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== done"
