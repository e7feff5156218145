#!/bin/bash
OUT=gpurun_out/r02_20; mkdir -p $OUT
echo "== memcheck: chained kernels (32x512x512) and all line-kernel forms (512x512x64)"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "chained_zy_passes_match and shape0 or line_pass_forms" -p no:cacheprovider 2>&1 | tee $OUT/memcheck.log | tail -8
echo "== racecheck (shared memory hazards): warp-specialised x pass, 512x512x64"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "line_pass_forms and 3" -p no:cacheprovider 2>&1 | tee $OUT/racecheck.log | tail -12
echo "== done"
