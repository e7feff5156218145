"""Time the experimental pass variants of scripts/exp/exp_passes.cu (GPU box).

    python scripts/exp/run_passes.py [out.json]

Every variant is checked bit for bit against the shipped form of the same pass, then timed
with CUDA events around single launches (full 512^3 grid, data larger than L2).  Also measures
what the z / y passes cost when their data is L2-resident (small plane counts) and the chunked
z->y pair, which bounds what any L2 fusion of the two passes can gain.
"""
import ctypes
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libevx_exp.so"))
vp, ci, cd, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong
lib.exp_z.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, cll, ci, vp]
lib.exp_strided.argtypes = [ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, cd, cd, vp]

dev = torch.device("cuda")
N, P, M = 512, 264, 256


def roots(n, count):
    m = np.arange(count, dtype=np.float64)
    a = -2.0 * math.pi * m / n
    return torch.from_numpy(np.stack([np.cos(a), np.sin(a)], -1).astype(np.float32)).to(dev).contiguous()


tw512, tw256, twr = roots(512, 512), roots(256, 256), roots(512, 257)
st = lambda: vp(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: vp(t.data_ptr()) if t is not None else None


def z(inv, ws, minb, real_in, real_out, spec, rows):
    rc = lib.exp_z(inv, ws, minb, ptr(real_in), ptr(real_out), ptr(spec), ptr(tw256), ptr(twr), rows, P, st())
    assert rc == 0, ("exp_z", inv, ws, minb, rc)


def strided(mode, kz, twreg, l2, spec, nx=N):
    rc = lib.exp_strided(mode, kz, twreg, l2, ptr(spec), ptr(tw512), nx, N, P, M + 1, 0.1, 1.5, st())
    assert rc == 0, ("exp_strided", mode, kz, twreg, l2, rc)


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3      # median, us


res = {}
gen = torch.Generator(device=dev).manual_seed(0)
r = (torch.rand((N, N, N), device=dev, generator=gen) - 0.5) * 1e-12
u = torch.rand((N, N, N), device=dev, generator=gen)
spec0 = torch.zeros((N, N, P, 2), dtype=torch.float32, device=dev)
spec1 = torch.zeros_like(spec0)
out0, out1 = torch.empty_like(u), torch.empty_like(u)

# ---- z passes ---------------------------------------------------------------------------
z(0, 0, 5, r, None, spec0, N * N)
for ws, minb in ((1, 5), (1, 4), (1, 6)):
    spec1.zero_()
    z(0, ws, minb, r, None, spec1, N * N)
    assert torch.equal(spec0, spec1), ("zfwd mismatch", ws, minb)
res["zfwd_sync_mb5"] = timed(lambda: z(0, 0, 5, r, None, spec1, N * N))
for ws, minb in ((1, 5), (1, 4), (1, 6)):
    res[f"zfwd_warpsync_mb{minb}"] = timed(lambda: z(0, ws, minb, r, None, spec1, N * N))
z(1, 0, 4, u, out0, spec0, N * N)
for ws, minb in ((1, 4), (1, 5)):
    out1.zero_()
    z(1, ws, minb, u, out1, spec0, N * N)
    assert torch.equal(out0, out1), ("zinv mismatch", ws, minb)
res["zinv_sync_mb4"] = timed(lambda: z(1, 0, 4, u, out1, spec0, N * N))
for ws, minb in ((1, 4), (1, 5)):
    res[f"zinv_warpsync_mb{minb}"] = timed(lambda: z(1, ws, minb, u, out1, spec0, N * N))
print(json.dumps(res), flush=True)

# ---- strided passes -----------------------------------------------------------------------
def check_and_time(tag, mode, variants, base):
    ref = spec0.clone()
    strided(mode, *base, ref)
    for v in variants:
        got = spec0.clone()
        strided(mode, *v, got)
        same = bool(torch.equal(ref, got))
        del got
        work = spec0.clone()
        t = timed(lambda: strided(mode, *v, work), reps=8)
        del work
        res[f"{tag}_kz{v[0]}_tw{v[1]}_l2{v[2]}"] = t
        if not same:
            res[f"{tag}_kz{v[0]}_tw{v[1]}_l2{v[2]}_MISMATCH"] = True
    del ref

yv = [(8, 0, 0), (8, 1, 0), (8, 0, 128), (8, 0, 256), (8, 1, 256)]
check_and_time("yfwd", 0, yv, (8, 0, 0))
check_and_time("yinv", 1, [(8, 0, 0), (8, 1, 0)], (8, 0, 0))
xv = [(16, 0, 0), (16, 1, 0), (16, 0, 256), (8, 0, 0), (8, 0, 128), (8, 0, 256), (8, 1, 128)]
check_and_time("xmid", 2, xv, (16, 0, 0))
print(json.dumps(res), flush=True)

# ---- L2-resident passes: per-plane cost on small plane counts ------------------------------
for planes in (8, 16, 32, 64):
    rows = planes * N
    work = spec0[:planes].clone()
    res[f"l2_zfwd_{planes}pl_us_per_plane"] = timed(lambda: z(0, 0, 5, r, None, work, rows), reps=20, warm=5) / planes
    res[f"l2_yfwd_{planes}pl_us_per_plane"] = timed(lambda: strided(0, 8, 0, 0, work, nx=planes), reps=20, warm=5) / planes
    res[f"l2_yinv_{planes}pl_us_per_plane"] = timed(lambda: strided(1, 8, 0, 0, work, nx=planes), reps=20, warm=5) / planes
    res[f"l2_zinv_{planes}pl_us_per_plane"] = timed(lambda: z(1, 0, 4, u, out1, work, rows), reps=20, warm=5) / planes
    del work
res["full_zfwd_us_per_plane"] = res["zfwd_sync_mb5"] / N
res["full_yfwd_us_per_plane"] = res["yfwd_kz8_tw0_l20"] / N
print(json.dumps(res), flush=True)

# ---- chunked z->y forward pair over the whole grid (what the L2-blocked schedule launches) ---
def pair(chunk):
    for x0 in range(0, N, chunk):
        z(0, 0, 5, r[x0:x0 + chunk], None, spec1[x0:x0 + chunk], chunk * N)
        strided(0, 8, 0, 0, spec1[x0:x0 + chunk], nx=chunk)

def pair_inv(chunk):
    for x0 in range(0, N, chunk):
        strided(1, 8, 0, 0, spec1[x0:x0 + chunk], nx=chunk)
        z(1, 0, 4, u[x0:x0 + chunk], out1[x0:x0 + chunk], spec1[x0:x0 + chunk], chunk * N)

def timed_region(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for chunk in (512, 64, 32, 16, 8):
    res[f"pair_fwd_chunk{chunk}_us"] = timed_region(lambda: pair(chunk))
    res[f"pair_inv_chunk{chunk}_us"] = timed_region(lambda: pair_inv(chunk))
# same launches captured in a CUDA graph (no host launch gaps)
for chunk in (32, 16):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        pair(chunk)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            pair(chunk)
    torch.cuda.synchronize()
    res[f"pair_fwd_chunk{chunk}_graph_us"] = timed_region(g.replay)

res["copy_us"] = timed(lambda: out1.copy_(u))
res["copy_GBs"] = 8 * N ** 3 / res["copy_us"] / 1e3
print(json.dumps(res, indent=1), flush=True)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(res, f, indent=1)
