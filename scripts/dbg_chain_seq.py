"""Replays tests/test_gpu_parity.py::test_chained_zy_passes... for one shape, step by step with
a synchronize after every call (GPU box).  python scripts/dbg_chain_seq.py [nx=8]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shape = (nx, 512, 512)
gen = torch.Generator(device="cuda").manual_seed(5)
u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
r = torch.randn(shape, device="cuda", generator=gen)
plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
outs = {}
for chain in ("0", "1", "1"):
    os.environ["EVX_FFT_CHAIN"] = chain
    res = []
    for what in ("apply_u", "apply_none", "ch_step"):
        out = torch.full_like(u, float("nan"))
        if what == "apply_u":
            plan.apply(u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        elif what == "apply_none":
            plan.apply(None, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        else:
            plan.ch_step(u, out, (1.0, 1.0, 1.0), 0.1, 3.0, 1.0, 0.25)
        torch.cuda.synchronize()
        print("chain", chain, what, "ok finite", bool(torch.isfinite(out).all()), flush=True)
        res.append(out)
    outs.setdefault(chain, []).append(res)
for run in outs["1"]:
    print("equal:", [bool(torch.equal(a, b)) for a, b in zip(outs["0"][0], run)], flush=True)
