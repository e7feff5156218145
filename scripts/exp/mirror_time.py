"""CH IMEX step with a zero-flux x axis: native mirrored x pass against the explicit 2 Nx extension."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import evoxels_b200 as evo
from evoxels_b200.problem_definition import CahnHilliard
from evoxels_b200.timesteppers import PseudoSpectralIMEX
from evoxels_b200.voxelgrid import VoxelGridTorch
for n in (128, 256, 512, 100):
    shape = (n, n, n)
    vf = evo.VoxelFields(shape, tuple(float(v) for v in shape))
    vg = VoxelGridTorch(vf.grid_info(), device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob = CahnHilliard(vg, bc=("neumann", "periodic", "periodic"))
    u = 0.5 + 0.1 * torch.rand((1,) + shape, device="cuda")
    res = {}
    for name, off in (("mirror", False), ("extension", True)):
        ts = PseudoSpectralIMEX(prob, 0.1)
        ts._no_native_mirror = off
        v = u
        for _ in range(3): v = ts.step(0.0, v)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): v = ts.step(0.0, v)
        b.record(); torch.cuda.synchronize()
        res[name] = (round(a.elapsed_time(b) / 10, 3), v)
    d = float((res["mirror"][1] - res["extension"][1]).abs().max())
    print(n, {k: v[0] for k, v in res.items()}, "max diff after 13 steps", d)
