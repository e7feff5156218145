import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evoxels_b200 import _native
SP = (1.0, 0.5, 2.0)
for shape in [(1024, 8, 16), (1024, 64, 32)]:
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    for power, name in ((2, "imex"), (1 | _native.FILTER_ETD1, "etd1"), (1, "imex_p1")):
        outs = []
        for flag in ("0", "1", "1"):
            os.environ["EVX_FFT_LINE4"] = flag
            out = torch.full_like(u, float("nan"))
            plan.apply(u, r, out, SP, 0.1, 1.5, power)
            torch.cuda.synchronize()
            outs.append(out)
        d = (outs[0] - outs[1]).abs()
        print(shape, name, "equal", torch.equal(outs[0], outs[1]), "repeatable", torch.equal(outs[1], outs[2]),
              "max", float(d.max()), "n_diff", int((d > 0).sum()), "max|out|", float(outs[0].abs().max()),
              "first idx", (d > 0).nonzero()[:3].tolist() if (d > 0).any() else None, flush=True)
