"""Launch ONE of every hot-path kernel at 512^3 between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full` (scripts/gpu_round.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
u = 0.5 + 0.1 * torch.rand((n, n, n), device=dev)
out = torch.empty_like(u)
u2 = torch.stack([torch.rand((n, n, n), device=dev), 0.5 * torch.rand((n, n, n), device=dev)])
per = (("periodic", None),) * 3
neu = (("neumann", None),) * 3
plan = _native.ImexPlan((n, n, n), torch.float32, "cuda", _native.FFT_NATIVE)


def once():
    plan.ch_step(u, out, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)          # rhs + five FFT passes
    _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, neu)               # general-BC instantiation
    _native.ac_stage(u, (1, 1, 1), 2.0, 1.0, 1.0, 0.0, 0.01, neu, base=u, y_out=out, alpha=0.05)
    _native.rd2_rhs(u2, (1, 1, 1), 1.0, 0.5, 0.055, 0.117)


for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
