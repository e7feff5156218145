"""GPU probe: warp-specialised Allen-Cahn stage (ac_tma.cu) against the cp.async tile kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from evoxels_b200 import _native

dev = torch.device("cuda")
per = (("periodic", None),) * 3
neu = (("neumann", None),) * 3
mix = (("dirichlet", (0.0, 1.0)), ("neumann", None), ("periodic", None))
mix2 = (("periodic", None), ("dirichlet", (0.1, 0.9)), ("dirichlet", (0.3, 0.5)))
SP = (1.0, 0.5, 2.0)

def run(u, bc, tma, cfg=0, halos=(None, None), mode="euler"):
    os.environ["EVX_AC_TMA"] = str(tma)
    os.environ["EVX_AC_TMA_CFG"] = str(cfg)
    out = torch.full_like(u, float("nan"))
    kw = dict(halo_lo=halos[0], halo_hi=halos[1])
    if mode == "euler":
        _native.ac_stage(u, SP, 2.0, 1.0, 1.0, 0.3, 0.01, bc, base=u, y_out=out, alpha=0.05, **kw)
        res = (out,)
    else:   # rk-like stage: k, y = base + a k, acc = acc_in + b k with separate base / acc fields
        g = torch.Generator(device="cuda").manual_seed(9)
        base = torch.rand(u.shape, device=dev, generator=g)
        acc = torch.rand(u.shape, device=dev, generator=g)
        k = torch.full_like(u, float("nan")); acc_out = torch.full_like(u, float("nan"))
        _native.ac_stage(u, SP, 2.0, 1.0, 1.0, 0.3, 0.01, bc, k_out=k, base=base, y_out=out, alpha=0.025,
                         acc_in=acc, acc_out=acc_out, beta=0.0125, **kw)
        res = (out, k, acc_out)
    torch.cuda.synchronize()
    return res

ok = True
g = torch.Generator(device="cuda").manual_seed(1)
for shape in [(64, 64, 64), (8, 32, 128), (5, 16, 256), (100, 100, 100), (48, 40, 132), (20, 18, 260), (130, 34, 512), (2, 4, 64), (1, 8, 128)]:
    u = -0.1 + 1.2 * torch.rand(shape, device=dev, generator=g)
    for bcn, bc in (("per", per), ("neu", neu), ("mix", mix), ("mix2", mix2)):
        for mode in ("euler", "rk"):
            ref = run(u, bc, 0, mode=mode)
            for cfg in (0, 1, 2, 3):
                got = run(u, bc, 1, cfg, mode=mode)
                for a, b in zip(got, ref):
                    err = float((a - b).abs().max() / b.abs().max())
                    bad = not (err < 1e-6) or bool(torch.isnan(a).any())
                    if bad or (cfg == 0 and mode == "euler" and bcn == "neu"):
                        print(shape, bcn, mode, cfg, "maxrel", err, "BAD" if bad else "", bool(torch.equal(a, b)))
                    ok &= not bad
for bcn, bc in (("per", per), ("neu", neu), ("mix", mix)):
    full = -0.1 + 1.2 * torch.rand((24, 32, 128), device=dev, generator=g)
    ref = run(full, bc, 0)[0]
    for a, b in ((0, 8), (8, 16), (16, 24)):
        sl = full[a:b].contiguous()
        if bc is per:
            lo = full[(a - 1) % 24][None].contiguous(); hi = full[b % 24][None].contiguous()
        else:
            lo = full[a - 1:a].contiguous() if a >= 1 else None
            hi = full[b:b + 1].contiguous() if b + 1 <= 24 else None
        got = run(sl, bc, 1, 0, (lo, hi))[0]
        err = float((got - ref[a:b]).abs().max())
        print("halo", bcn, a, b, err, bool(torch.equal(got, ref[a:b])))
        ok &= err < 1e-6
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
for bcn, bc in (("neu", neu), ("per", per)):
    os.environ["EVX_AC_TMA"] = "0"
    res[f"old_{bcn}"] = round(timed(lambda: _native.ac_stage(u, (1, 1, 1), 2.0, 1.0, 1.0, 0.0, 0.01, bc, base=u, y_out=out, alpha=0.05)), 4)
    os.environ["EVX_AC_TMA"] = "1"
    for cfg in range(4):
        for chunk in (32, 64, 128):
            os.environ["EVX_AC_TMA_CFG"] = str(cfg)
            os.environ["EVX_AC_TMA_CHUNK"] = str(chunk)
            res[f"tma_cfg{cfg}_chunk{chunk}_{bcn}"] = round(timed(lambda: _native.ac_stage(u, (1, 1, 1), 2.0, 1.0, 1.0, 0.0, 0.01, bc, base=u, y_out=out, alpha=0.05)), 4)
print(json.dumps(res, indent=1))
