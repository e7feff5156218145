#!/bin/bash
# Late round 2: how the records under profiles/ named below were produced (each line is one gpurun call).
#   1 GPU:  scripts/exp/line4_check.py                          -> profiles/r02_line4_check.json
#           scripts/gpu_r02_23.sh, scripts/gpu_r02_24.sh        -> profiles/r02_ncu_full_summary_line4.txt, _line16.txt, r02_ncu_smem_pipe.txt
#           scripts/exp/ac_time.py                              -> Allen-Cahn stage timing
#           python bench.py                                     -> profiles/r02_bench_n1_final3.json
#   N GPUs: gpurun --gpus N -- scripts/gpu_r02_n8_final.sh N    -> profiles/r02_bench_n2_final2.json, r02_bench_n4_final.json
#           gpurun --gpus N -- torchrun ... scripts/exp/dist_ab.py   -> profiles/r02_transport_ab_n{2,4,8}*.json
#           gpurun --gpus 8 -- scripts/gpu_r02_22.sh 8          -> profiles/r02_timeline_n8_ce_line4.txt
# Usage: scripts/gpu_r02_late.sh <N>   (runs the A/B of the transports on N GPUs of this box)
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
  scripts/exp/dist_ab.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -20
