#!/bin/bash
# quick kernel timing session: FFT parity tests + one bench run with per-kernel roofline
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft or imex or step or crd or etd1" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu 2>/dev/null | tail -1 | tee $OUT/bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step',d['ms_per_step']); print({k:round(v['us'],1) for k,v in d['roofline'].get('kernels',{}).items()} if 'kernels' in d['roofline'] else d['roofline'])"
