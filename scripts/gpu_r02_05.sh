#!/bin/bash
# Round 2, call 5: full GPU suite + full bench line (all legs) + the nx=8 chain sequence under memcheck
OUT=gpurun_out/r02_05; mkdir -p $OUT
echo "== seq nx=8 (memcheck)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -X faulthandler scripts/dbg_chain_seq.py 8 2>&1 | tee $OUT/seq8.log | tail -30
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_chained_zy_passes_match_one_kernel_per_pass_bit_for_bit 2>&1 | tee $OUT/pytest_gpu.log | tail -6
echo "== bench (all legs)"; EVX_FFT_TMA_PF=1 timeout 1500 python bench.py 2>$OUT/bench.err > $OUT/bench.json; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("ms_per_step","clocks","cpu_baseline","reference_gpu","other_configs"):
    print(k, json.dumps(d.get(k))[:900])
print({k.split(" ")[0]: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()})
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-900
echo "== done"
