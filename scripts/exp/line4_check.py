"""GPU check + timing of the four-stage TMA-tiled pass for 1024-point lines (StridedLine4):
bit-identity against the cp.async passes on single-GPU grids and on the x-slab plan with virtual
ranks on one GPU, then event timings of the passes at the sizes of the multi-GPU runs."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evoxels_b200 import _native  # noqa: E402

SP = (1.0, 0.5, 2.0)
res = {}


def single(shape):
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    outs = {}
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag
        o = []
        for power in (2, 1 | _native.FILTER_ETD1):
            out = torch.full_like(u, float("nan"))
            plan.apply(u, r, out, SP, 0.1, 1.5, power)
            o.append(out)
        torch.cuda.synchronize()
        outs[flag] = o
    ok = all(torch.equal(a, b) and bool(torch.isfinite(b).all()) for a, b in zip(outs["0"], outs["1"]))
    print("single", shape, "bit-identical" if ok else "MISMATCH", flush=True)
    return ok


def virtual(shape, world, chunks):
    gen = torch.Generator(device="cuda").manual_seed(11)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    os.environ["EVX_FFT_LINE4"] = "0"
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, SP, 0.1, 1.5, 2)
    nxl, nyl = shape[0] // world, shape[1] // world
    ok = True
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag
        plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
        send = [p.new_buffer().zero_() for p in plans]
        A = [p.new_buffer().fill_(float("nan")) for p in plans]
        B = [p.new_buffer().fill_(float("nan")) for p in plans]
        spec = [p.new_buffer() for p in plans]
        xb = [round(i * nxl / chunks) for i in range(chunks + 1)]
        for k, p in enumerate(plans):
            rl = r[k * nxl:(k + 1) * nxl].contiguous()
            if chunks == 1:
                p.forward(rl, spec[k], send[k])
                B[k][k].copy_(send[k][k])
            else:
                for i in range(chunks):
                    p.forward_chunk(rl, spec[k], send[k], xb[i], xb[i + 1] - xb[i], self_block=B[k])
        for k in range(world):
            for j in range(world):
                if j != k:
                    B[k][j].copy_(send[j][k])
        bounds = [round(i * nyl / chunks) for i in range(chunks + 1)]
        for k, p in enumerate(plans):
            for i in range(chunks):
                p.middle_chunk(B[k], bounds[i], bounds[i + 1] - bounds[i], SP, 0.1, 1.5, 2, self_block=A[k])
        for k in range(world):
            for j in range(world):
                if j != k:
                    A[j][k].copy_(B[k][j])
        out = torch.empty_like(u)
        for k, p in enumerate(plans):
            o = torch.empty((nxl,) + tuple(shape[1:]), device="cuda")
            p.backward(A[k], spec[k], u[k * nxl:(k + 1) * nxl].contiguous(), o)
            out[k * nxl:(k + 1) * nxl] = o
        torch.cuda.synchronize()
        eq = torch.equal(out, ref)
        print("virtual", shape, "W", world, "chunks", chunks, "LINE4", flag, "bit-identical" if eq else "MISMATCH",
              float((out - ref).abs().max()), flush=True)
        ok = ok and eq
    return ok


def virtual_p2p(shape, world, chunks):
    """peer-store transport with W plans on one GPU: the passes write into each other's buffers"""
    gen = torch.Generator(device="cuda").manual_seed(12)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    os.environ["EVX_FFT_LINE4"] = "0"
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, SP, 0.1, 1.5, 2)
    nxl = shape[0] // world
    ok = True
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag
        plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
        A = [p.new_buffer().fill_(float("nan")) for p in plans]
        B = [p.new_buffer().fill_(float("nan")) for p in plans]
        spec = [p.new_buffer() for p in plans]
        xb = [round(i * nxl / chunks) for i in range(chunks + 1)]
        for k, p in enumerate(plans):
            p.set_p2p_ctas(0 if chunks == 1 else 148)
            rl = r[k * nxl:(k + 1) * nxl].contiguous()
            if chunks == 1:
                p.forward_p2p(rl, spec[k], [b.data_ptr() for b in B])
            else:
                for i in range(chunks):
                    p.forward_chunk_p2p(rl, spec[k], [b.data_ptr() for b in B], xb[i], xb[i + 1] - xb[i], parts=1)
                    p.forward_chunk_p2p(rl, spec[k], [b.data_ptr() for b in B], xb[i], xb[i + 1] - xb[i], parts=2)
        for k, p in enumerate(plans):
            p.middle_p2p(B[k], [a.data_ptr() for a in A], SP, 0.1, 1.5, 2)
        out = torch.empty_like(u)
        for k, p in enumerate(plans):
            o = torch.empty((nxl,) + tuple(shape[1:]), device="cuda")
            p.backward(A[k], spec[k], u[k * nxl:(k + 1) * nxl].contiguous(), o)
            out[k * nxl:(k + 1) * nxl] = o
        torch.cuda.synchronize()
        eq = torch.equal(out, ref)
        print("virtual p2p", shape, "W", world, "chunks", chunks, "LINE4", flag,
              "bit-identical" if eq else "MISMATCH", float((out - ref).abs().max()), flush=True)
        ok = ok and eq
    return ok


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def time_passes(shape, passes):
    u = torch.rand(shape, device="cuda")
    r = torch.randn(shape, device="cuda")
    out = torch.empty_like(u)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    os.environ["EVX_FFT_CHAIN"] = "0"
    row = {}
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag
        plan.apply(u, r, out, SP, 0.1, 1.5, 2)
        for which in passes:
            row[f"pass{which}_line4={flag}_ms"] = timed(lambda: plan.native_pass(which, u, r, out, SP, 0.1, 1.5, 2))
        row[f"apply_line4={flag}_ms"] = timed(lambda: plan.apply(u, r, out, SP, 0.1, 1.5, 2))
    os.environ.pop("EVX_FFT_CHAIN")
    print("timing", shape, json.dumps(row), flush=True)
    res["timing %s" % (shape,)] = row
    del plan, u, r, out
    torch.cuda.empty_cache()


def time_dist_middle(shape, world):
    """x pass of one rank's y-pencils ([nx][ny/W][P], in place) in 4 chunks, as the ce transport runs it"""
    plan = _native.DistPlan(shape, world, 0, "cuda")
    B = plan.new_buffer().zero_()
    A = plan.new_buffer().zero_()
    nyl = shape[1] // world
    bounds = [round(i * nyl / 4) for i in range(5)]
    row = {}
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag

        def run():
            for i in range(4):
                plan.middle_chunk(B, bounds[i], bounds[i + 1] - bounds[i], SP, 0.1, 1.5, 2, self_block=A)
        row[f"middle_4chunks_line4={flag}_ms"] = timed(run)
    print("dist middle", shape, "W", world, json.dumps(row), flush=True)
    res["dist middle %s W%d" % (shape, world)] = row
    del plan, A, B
    torch.cuda.empty_cache()


def time_dist_fwd_bwd(shape, world):
    """z + y forward in 4 x chunks and y + z inverse of one rank's slab, as the ce transport runs them"""
    plan = _native.DistPlan(shape, world, 0, "cuda")
    send, B, spec = plan.new_buffer().zero_(), plan.new_buffer().zero_(), plan.new_buffer()
    nxl = shape[0] // world
    rl = torch.randn((nxl,) + tuple(shape[1:]), device="cuda")
    out = torch.empty_like(rl)
    xb = [round(i * nxl / 4) for i in range(5)]
    row = {}
    for flag in ("0", "1"):
        os.environ["EVX_FFT_LINE4"] = flag

        def fwd():
            for i in range(4):
                plan.forward_chunk(rl, spec, send, xb[i], xb[i + 1] - xb[i], self_block=B)
        row[f"forward_4chunks_line4={flag}_ms"] = timed(fwd)
        row[f"backward_line4={flag}_ms"] = timed(lambda: plan.backward(B, spec, rl, out))
    print("dist fwd/bwd", shape, "W", world, json.dumps(row), flush=True)
    res["dist fwd/bwd %s W%d" % (shape, world)] = row
    del plan, send, B, spec, rl, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ok = True
    for shape in [(1024, 64, 32), (32, 1024, 64), (1024, 1024, 16), (1024, 8, 16)]:
        ok &= single(shape)
    for world, chunks in [(2, 1), (4, 2), (8, 3)]:
        ok &= virtual((1024, 64, 32), world, chunks)
    for shape, world, chunks in [((64, 1024, 32), 4, 2), ((64, 1024, 32), 8, 1), ((64, 1024, 32), 2, 2),
                                 ((1024, 1024, 16), 8, 4), ((1024, 1024, 16), 4, 1)]:
        ok &= virtual(shape, world, chunks)
    for shape, world, chunks in [((1024, 64, 32), 2, 1), ((1024, 1024, 16), 8, 2), ((1024, 1024, 16), 4, 1),
                                 ((64, 1024, 32), 8, 4)]:
        ok &= virtual_p2p(shape, world, chunks)
    res["bit_identical"] = bool(ok)
    time_dist_fwd_bwd((1024, 1024, 1024), 8)
    time_dist_fwd_bwd((1024, 1024, 512), 4)
    time_dist_middle((1024, 512, 512), 2)
    time_dist_middle((1024, 1024, 1024), 8)
    time_passes((1024, 512, 512), [2])
    time_passes((512, 1024, 512), [1, 3])
    os.makedirs("gpurun_out/r02_line4", exist_ok=True)
    json.dump(res, open("gpurun_out/r02_line4/line4_check.json", "w"), indent=1)
    sys.exit(0 if ok else 1)
