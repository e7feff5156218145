"""Chained z/y kernels: bit comparison with the one-kernel-per-pass pipeline and thread-0 cycle
accounting per block (GPU box).   python scripts/dbg_chain.py [nx=512] [reps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoxels_b200 import _native

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape = (nx, 512, 512)
gen = torch.Generator(device="cuda").manual_seed(5)
u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
r = torch.randn(shape, device="cuda", generator=gen)
plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
args = ((1.0, 0.5, 2.0), 0.1, 1.5, 2)

def run(which):
    out = torch.full_like(u, float("nan"))
    plan.native_pass(which, u, r, out, *args)
    torch.cuda.synchronize()
    return out

def timed(which, n=10):
    out = torch.empty_like(u)
    for _ in range(2):
        plan.native_pass(which, u, r, out, *args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        plan.native_pass(which, u, r, out, *args)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def stats(tag):
    st = plan.workspace[-65536:].view(torch.int64).view(-1, 16)[:296].double().cpu()
    st = st[st[:, 4] > 0]            # blocks that ran (one or two per SM)
    names = ["loader:wait_buffer", "loader:wait_plane", "loader:issue", "items", "total", "", "", "",
             "retirer:wait_done", "retirer:retire", "", "", "", "", "", ""]
    print(tag, " | ".join(f"{n} {st[:, i].mean():.0f}" for i, n in enumerate(names) if n), flush=True)

os.environ["EVX_FFT_CHAIN"] = "1"
os.environ["EVX_FFT_CHAIN_STATS"] = "1"
def zero_stats():
    plan.workspace[-65536:].zero_()
# forward: separate z + y
spec_off = plan.workspace.numel()
run(0); run(1)
ref_spec = plan.workspace.clone()
plan.workspace.zero_()
run(5)
stats("fwd stats (cycles):")
n = ref_spec.numel() - 65536 - 4096
print("forward chain == z + y passes:", bool(torch.equal(ref_spec[:n], plan.workspace[:n])), flush=True)
# inverse from the same spectrum
plan.workspace.copy_(ref_spec)
run(3); o_ref = run(4)
plan.workspace.copy_(ref_spec)
o_chain = run(6)
stats("inv stats (cycles):")
print("inverse chain == y + z passes:", bool(torch.equal(o_ref, o_chain)), "finite", bool(torch.isfinite(o_chain).all()), flush=True)
if nx >= 64:
    for lag in os.environ.get("LAGS", "12").split(","):
        os.environ["EVX_FFT_CHAIN_LAG"] = lag
        print(f"lag {lag}: zy fwd {timed(5):.4f} ms, yz inv {timed(6):.4f} ms | separate: z {timed(0):.4f} y {timed(1):.4f} yinv {timed(3):.4f} zinv {timed(4):.4f}", flush=True)
        zero_stats(); run(5); stats(f"  fwd lag {lag}:")
        zero_stats(); run(6); stats(f"  inv lag {lag}:")
