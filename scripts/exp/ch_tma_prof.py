"""One launch of the CH rhs kernel per configuration, between profiler start/stop (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from evoxels_b200 import _native
n = 512
u = 0.5 + 0.1 * torch.rand((n, n, n), device="cuda")
out = torch.empty_like(u)
per = (("periodic", None),) * 3
cfgs = [int(a) for a in sys.argv[1:]] or [0, 1]
def once():
    for c in cfgs:
        os.environ["EVX_CH_TMA"] = "0" if c < 0 else "1"
        os.environ["EVX_CH_TMA_CFG"] = str(max(c, 0))
        _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, per)
for _ in range(2): once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
