"""Allen-Cahn Euler stage at 512^3 (periodic and Neumann): event timing."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evoxels_b200 import _native
from evoxels_b200.problem_definition import normalize_bc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
phi = torch.rand((n, n, n), device="cuda")
out = torch.empty_like(phi)
res = {}
for name, bc in (("periodic", ("periodic",) * 3), ("neumann", ("neumann",) * 3)):
    bcn = normalize_bc(bc)
    def run():
        _native.ac_stage(phi, (1.0, 1.0, 1.0), 2.0, 1.0, 1.0, 0.0, 0.01, bcn, base=phi, y_out=out, alpha=0.01)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 20
print(json.dumps(res))
