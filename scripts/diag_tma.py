"""Pass-by-pass comparison of the TMA-tiled strided passes with the cp.async passes (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native

shapes = [(512, 512, 32), (16, 512, 64), (512, 32, 16), (512, 512, 512)]
for shape in shapes:
    for kz in ("8", "16"):
        os.environ["EVX_FFT_TMA_KZ"] = kz
        gen = torch.Generator(device="cuda").manual_seed(2)
        u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
        r = torch.randn(shape, device="cuda", generator=gen)
        out = torch.empty_like(u)
        plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
        plan.workspace.zero_()
        args = (u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        os.environ["EVX_FFT_TMA"] = "0"
        plan.native_pass(0, *args)
        torch.cuda.synchronize()
        state = plan.workspace.clone()
        nx, ny, nz = shape
        P = ((nz // 2 + 1 + 7) // 8) * 8
        real_bytes = plan.workspace.numel() - nx * ny * P * 8
        for which in (1, 2, 3):
            res = {}
            for tag, tma in (("A", "0"), ("B", "1"), ("B2", "1")):
                os.environ["EVX_FFT_TMA"] = tma
                plan.workspace.copy_(state)
                plan.native_pass(which, *args)
                torch.cuda.synchronize()
                res[tag] = plan.workspace.clone()
            a = res["A"][real_bytes:real_bytes + nx * ny * P * 8].view(torch.float32).view(nx, ny, P, 2)
            b = res["B"][real_bytes:real_bytes + nx * ny * P * 8].view(torch.float32).view(nx, ny, P, 2)
            b2 = res["B2"][real_bytes:real_bytes + nx * ny * P * 8].view(torch.float32).view(nx, ny, P, 2)
            ne = (a.view(torch.int32) != b.view(torch.int32))
            nz_ = int(ne.sum())
            msg = f"shape={shape} kz={kz} pass={which}: differing floats {nz_}/{a.numel()}"
            if nz_:
                d = (a - b).abs()
                msg += f" max|d|={float(d.max()):.3e} max|a|={float(a.abs().max()):.3e}"
                idx = ne.nonzero()
                msg += f" x[{int(idx[:,0].min())},{int(idx[:,0].max())}] y[{int(idx[:,1].min())},{int(idx[:,1].max())}] kz[{int(idx[:,2].min())},{int(idx[:,2].max())}]"
                msg += f" per-kz counts {ne.sum(dim=(0,1,3)).tolist()[:20]}"
            msg += f" | TMA run-to-run equal: {bool(torch.equal(b.view(torch.int32), b2.view(torch.int32)))}"
            print(msg, flush=True)
            state = res["A"]
        del plan
