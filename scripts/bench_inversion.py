"""BASELINE config 5: forward + backward through N Cahn-Hilliard steps (hand-written adjoint).
   python scripts/bench_inversion.py [size=256] [steps=100]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evoxels_b200 as evo
from evoxels_b200.problem_definition import CahnHilliard
from evoxels_b200.timesteppers import PseudoSpectralIMEX
from evoxels_b200.voxelgrid import VoxelGridTorch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
vf = evo.VoxelFields((n, n, n), (float(n),) * 3)
vg = VoxelGridTorch(vf.grid_info(), device="cuda")
u0 = 0.5 + 0.1 * torch.rand((1, n, n, n), device="cuda")
obs_at = {steps // 3, 2 * steps // 3, steps}
with torch.no_grad():
    ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=3.0, D=1.0), 0.1)
    v, obs = u0, {}
    for i in range(1, steps + 1):
        v = ts.step(0.0, v)
        if i in obs_at:
            obs[i] = v.clone()

def fwd_bwd():
    D = torch.tensor(2.0, dtype=torch.float64, device="cuda", requires_grad=True)
    eps = torch.tensor(2.0, dtype=torch.float64, device="cuda", requires_grad=True)
    ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=eps, D=D), 0.1)
    v, loss = u0.clone().requires_grad_(True), 0.0
    for i in range(1, steps + 1):
        v = ts.step(0.0, v)
        if i in obs_at:
            loss = loss + ((v - obs[i]) ** 2).sum()
    gD, ge = torch.autograd.grad(loss, (D, eps))
    return float(loss), float(gD), float(ge)

fwd_bwd()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.reset_peak_memory_stats()
a.record(); res = fwd_bwd(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({"config": f"inversion fwd+bwd {steps} CH steps at {n}^3", "ms_total": ms,
                  "ms_per_step_fwd_bwd": ms / steps, "voxel_updates_per_s": n**3 * steps / (ms * 1e-3),
                  "GBs_at_160B_per_voxel": 160 * n**3 * steps / (ms * 1e-3) / 1e9,
                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                  "loss": res[0], "dL_dD": res[1], "dL_deps": res[2]}))
