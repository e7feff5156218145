"""One CH step at a small grid between profiler start/stop (ncu launch list), plus event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from evoxels_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
u = 0.5 + 0.1 * torch.rand((n, n, n), device="cuda")
out = torch.empty_like(u)
plan = _native.ImexPlan((n, n, n), torch.float32, "cuda", _native.FFT_AUTO)
print("backend", plan.backend_name)
def once():
    plan.ch_step(u, out, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)
for _ in range(5): once()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): once()
b.record(); torch.cuda.synchronize()
print("us/step", a.elapsed_time(b) / 200 * 1e3)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): once()
g.replay(); torch.cuda.synchronize()
a.record()
for _ in range(20): g.replay()
b.record(); torch.cuda.synchronize()
print("us/step (graph of 10)", a.elapsed_time(b) / 200 * 1e3)
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
