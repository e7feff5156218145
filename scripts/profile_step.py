"""ONE Cahn-Hilliard IMEX step at n^3 between cudaProfilerStart/Stop, through the public
stepper (so EVX_SCHEDULE / EVX_TUNE apply), for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evoxels_b200 as evo
from evoxels_b200.problem_definition import CahnHilliard
from evoxels_b200.timesteppers import PseudoSpectralIMEX
from evoxels_b200.voxelgrid import VoxelGridTorch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vf = evo.VoxelFields((n, n, n), (float(n),) * 3)
vg = VoxelGridTorch(vf.grid_info(), device="cuda")
ts = PseudoSpectralIMEX(CahnHilliard(vg), 0.1)
u = 0.5 + 0.1 * torch.rand((1, n, n, n), device="cuda")
for _ in range(3):
    u = ts.step(0.0, u)
torch.cuda.synchronize()
torch.cuda.profiler.start()
u = ts.step(0.0, u)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
