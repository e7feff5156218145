"""GPU probe: warp-specialised CH rhs (ch_rhs_tma.cu) against the cp.async form, then timings."""
import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from evoxels_b200 import _native

dev = torch.device("cuda")
per = (("periodic", None),) * 3
neu = (("neumann", None),) * 3
mix = (("dirichlet", (0.2, 0.6)), ("neumann", None), ("periodic", None))
mix2 = (("periodic", None), ("dirichlet", (0.1, 0.9)), ("dirichlet", (0.3, 0.5)))

def run(u, bc, tma, cfg=0, halos=(None, None)):
    os.environ["EVX_CH_TMA"] = str(tma)
    os.environ["EVX_CH_TMA_CFG"] = str(cfg)
    out = torch.full_like(u, float("nan"))
    _native.ch_rhs(u, out, (1.0, 0.5, 2.0), 3.0, 1.0, bc, halo_lo=halos[0], halo_hi=halos[1])
    torch.cuda.synchronize()
    return out

ok = True
g = torch.Generator(device="cuda").manual_seed(1)
for shape in [(64, 64, 64), (8, 32, 128), (5, 16, 256), (100, 100, 100), (48, 40, 132), (20, 18, 260), (130, 34, 512), (2, 4, 64)]:
    u = -0.1 + 1.2 * torch.rand(shape, device=dev, generator=g)
    for bcn, bc in (("per", per), ("neu", neu), ("mix", mix), ("mix2", mix2)):
        ref = run(u, bc, 0)
        for cfg in (0, 1, 2, 4):
            got = run(u, bc, 1, cfg)
            err = float((got - ref).norm() / ref.norm())
            bad = not (err < 2e-6) or bool(torch.isnan(got).any())
            if bad or cfg == 0:
                print(shape, bcn, cfg, "rel", err, "BAD" if bad else "")
            ok &= not bad
# x halos (slab of a larger periodic / neumann field)
for bcn, bc in (("per", per), ("neu", neu)):
    full = -0.1 + 1.2 * torch.rand((24, 32, 128), device=dev, generator=g)
    ref = run(full, bc, 0)
    for a, b in ((0, 8), (8, 16), (16, 24)):
        sl = full[a:b].contiguous()
        if bc is per:
            lo = torch.stack([full[(a - 2) % 24], full[(a - 1) % 24]]).contiguous()
            hi = torch.stack([full[b % 24], full[(b + 1) % 24]]).contiguous()
        else:
            lo = full[a - 2:a].contiguous() if a >= 2 else None
            hi = full[b:b + 2].contiguous() if b + 2 <= 24 else None
        got = run(sl, bc, 1, 0, (lo, hi))
        err = float((got - ref[a:b]).norm() / ref[a:b].norm())
        print("halo", bcn, a, b, err)
        ok &= err < 2e-6
print("PARITY", "OK" if ok else "FAIL")

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n = 512
u = 0.5 + 0.1 * torch.rand((n, n, n), device=dev)
out = torch.empty_like(u)
res = {}
os.environ["EVX_CH_TMA"] = "0"
res["old_per"] = timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, per))
res["old_neu"] = timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, neu))
os.environ["EVX_CH_TMA"] = "1"
for cfg in range(7):
    for chunk in (32, 64, 128):
        os.environ["EVX_CH_TMA_CFG"] = str(cfg)
        os.environ["EVX_CH_TMA_CHUNK"] = str(chunk)
        res[f"tma_cfg{cfg}_chunk{chunk}_per"] = round(timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, per)), 4)
    os.environ["EVX_CH_TMA_CHUNK"] = "64"
    res[f"tma_cfg{cfg}_neu"] = round(timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, neu)), 4)
print(json.dumps(res, indent=1))
