"""Time individual library kernels with CUDA events (GPU box). Usage: python scripts/bench_kernels.py [n]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
u = 0.5 + 0.1 * torch.rand((n, n, n), device=dev)
out = torch.empty_like(u)
per = (("periodic", None),) * 3
neu = (("neumann", None),) * 3

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {"n": n}
ms = timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, per))
res["ch_rhs_periodic_ms"] = ms; res["ch_rhs_periodic_GBs"] = 8 * n**3 / ms / 1e6
ms = timed(lambda: _native.ch_rhs(u, out, (1, 1, 1), 3.0, 1.0, neu))
res["ch_rhs_neumann_ms"] = ms
ms = timed(lambda: _native.ac_stage(u, (1, 1, 1), 2.0, 1.0, 1.0, 0.0, 0.01, neu, base=u, y_out=out, alpha=0.05))
res["ac_euler_ms"] = ms; res["ac_euler_GBs"] = 8 * n**3 / ms / 1e6
for name, code in (("cufft", _native.FFT_CUFFT), ("native", _native.FFT_NATIVE)):
    try:
        plan = _native.ImexPlan((n, n, n), torch.float32, "cuda", code)
        r = torch.randn_like(u)
        ms = timed(lambda: plan.apply(u, r, out, (1, 1, 1), 0.1, 1.5, 2), reps=10)
        res[f"imex_apply_{name}_ms"] = ms
        ms = timed(lambda: plan.ch_step(u, out, (1, 1, 1), 0.1, 3.0, 1.0, 0.25), reps=10)
        res[f"ch_step_{name}_ms"] = ms
        del plan
    except Exception as exc:
        res[f"imex_apply_{name}_ms"] = str(exc)
# two-species reaction-diffusion (16 B/voxel rhs) and the ETD1 step
u2 = torch.stack([torch.rand((n, n, n), device=dev), 0.5 * torch.rand((n, n, n), device=dev)])
ms = timed(lambda: _native.rd2_rhs(u2, (1, 1, 1), 1.0, 0.5, 0.055, 0.117), reps=10)
res["rd2_rhs_ms"] = ms; res["rd2_rhs_GBs"] = 16 * n**3 / ms / 1e6
try:
    plan = _native.ImexPlan((n, n, n), torch.float32, "cuda", _native.FFT_NATIVE)
    r = torch.randn_like(u)
    ms = timed(lambda: plan.apply(u, r, out, (1, 1, 1), 0.5, 1.0, 1 | _native.FILTER_ETD1), reps=10)
    res["etd1_apply_native_ms"] = ms
    ms = timed(lambda: plan.apply(u, r, out, (1, 1, 1), 0.5, 1.0, 1), reps=10)
    res["imex_apply_native_p1_ms"] = ms
    del plan, r
except Exception as exc:
    res["etd1_apply_native_ms"] = str(exc)
del u2
ms = timed(lambda: out.copy_(u))
res["copy_ms"] = ms; res["copy_GBs"] = 8 * n**3 / ms / 1e6
print(json.dumps(res))
