"""BASELINE config 4 on ONE GPU: Cahn-Hilliard IMEX at n^3 (default 2048^3, 137 GB of HBM).
   python scripts/bench_big.py [n=2048] [steps=5]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
plan = _native.ImexPlan((n, n, n), torch.float32, dev, _native.FFT_NATIVE)
u = torch.empty((n, n, n), device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
for i in range(0, n, 64):          # fill in slices: torch.rand of 34 GB would need a temporary
    u[i:i + 64] = 0.5 + 0.1 * torch.rand((min(64, n - i), n, n), device=dev, generator=gen)
out = torch.empty_like(u)
m0 = float(u[::8].double().mean())
for _ in range(2):
    plan.ch_step(u, out, (1, 1, 1), 0.1, 3.0, 1.0, 0.25); u, out = out, u
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    plan.ch_step(u, out, (1, 1, 1), 0.1, 3.0, 1.0, 0.25); u, out = out, u
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
full_mean0 = None
print(json.dumps({"config": f"CH IMEX {n}^3 fp32 periodic, 1 GPU", "ms_per_step": ms,
                  "voxel_updates_per_s": n**3 / (ms * 1e-3), "GBs_at_60B": 60 * n**3 / (ms * 1e-3) / 1e9,
                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                  "finite": bool(torch.isfinite(u[::16]).all()), "mean_sample": float(u[::8].double().mean()), "mean_sample0": m0}))
