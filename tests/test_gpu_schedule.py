"""L2-blocked launch schedules of the native pipeline (evx_imex_plan_set_schedule) on the GPU:
every schedule runs the same kernels on sub-ranges, so its result must equal the
one-launch-per-pass result BIT FOR BIT - for the fused CH step, for the plain application,
under CUDA-graph capture (the two-stream form forks and joins inside the capture) and on a
non-default stream.  The tuner must end on a schedule with that property."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import evx_oracle as O

pytestmark = pytest.mark.gpu

import evoxels_b200 as evo  # noqa: E402
from evoxels_b200 import _native  # noqa: E402
from evoxels_b200.problem_definition import CahnHilliard  # noqa: E402
from evoxels_b200.timesteppers import PseudoSpectralIMEX  # noqa: E402
from evoxels_b200.voxelgrid import VoxelGridTorch  # noqa: E402

R, C = _native.SCHED_RING_INV, _native.SCHED_CHUNK_RHS
SCHEDULES = [(8, 1, 0), (8, 1, R), (8, 2, 0), (8, 2, R), (8, 2, R | C), (8, 1, R | C), (16, 2, R | C),
             (5, 2, R | C), (6, 1, C), (2, 2, R | C), (31, 2, R | C), (13, 2, R),
             (8, 3, R | C), (16, 3, C), (5, 3, R | C), (2, 3, R | C), (8, 3, R)]
SP = (1.0, 0.5, 2.0)


def ch_args():
    return SP, 0.1, 3.0, 1.0, 0.25


@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 32, 64), (32, 64, 256)])
def test_every_schedule_is_bit_identical(cuda_device, shape):
    g = torch.Generator(device="cuda").manual_seed(3)
    u = 0.5 + 0.6 * (torch.rand(shape, device="cuda", generator=g) - 0.5)    # leaves [0,1] in places
    r = torch.randn(shape, device="cuda", generator=g)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    base_step = plan.ch_step(u, torch.empty_like(u), *ch_args()).clone()
    base_apply = plan.apply(u, r, torch.empty_like(u), SP, 0.1, 1.5, 2).clone()
    base_upd = plan.apply(None, r, torch.empty_like(u), SP, 0.1, 0.7, 1 | _native.FILTER_ETD1).clone()
    want = O.CHOracle(shape, SP, 0.1, 3.0, 1.0, 0.25).step(u.cpu()[None])[0]
    assert rel_l2(base_step.cpu().numpy(), want.numpy()) <= 1e-5
    u_before, r_before = u.clone(), r.clone()
    for sched in SCHEDULES:
        if sched[0] >= shape[0]:
            continue
        plan.set_schedule(*sched)
        assert plan.schedule() == sched
        for _ in range(2):                       # twice: ring slots / events are reused across calls
            got = plan.ch_step(u, torch.full_like(u, float("nan")), *ch_args())
            assert torch.equal(got, base_step), sched
        got = plan.apply(u, r, torch.full_like(u, float("nan")), SP, 0.1, 1.5, 2)
        assert torch.equal(got, base_apply), sched
        got = plan.apply(None, r, torch.full_like(u, float("nan")), SP, 0.1, 0.7, 1 | _native.FILTER_ETD1)
        assert torch.equal(got, base_upd), sched
    assert torch.equal(u, u_before) and torch.equal(r, r_before)
    plan.set_schedule(0, 1, 0)
    assert torch.equal(plan.ch_step(u, torch.empty_like(u), *ch_args()), base_step)


def test_schedule_argument_errors(cuda_device):
    plan = _native.ImexPlan((32, 32, 32), torch.float32, "cuda", _native.FFT_NATIVE)
    for bad in [(-1, 1, 0), (8, 0, 0), (8, 4, 0), (8, 1, 4)]:
        with pytest.raises(_native.NativeLibraryError):
            plan.set_schedule(*bad)
    other = _native.ImexPlan((20, 20, 20), torch.float32, "cuda")       # mixed-radix back end
    with pytest.raises(_native.NativeLibraryError):
        other.set_schedule(4, 1, 0)


def test_two_stream_schedule_on_side_stream_and_in_a_graph(cuda_device):
    shape = (64, 32, 64)
    g = torch.Generator(device="cuda").manual_seed(4)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=g)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)

    def three_steps(out_a, out_b):
        plan.ch_step(u, out_a, *ch_args())
        plan.ch_step(out_a, out_b, *ch_args())
        plan.ch_step(out_b, out_a, *ch_args())

    a0, b0 = torch.empty_like(u), torch.empty_like(u)
    three_steps(a0, b0)
    want = a0.clone()
    plan.set_schedule(8, 3, R | C)
    # caller's stream is not the default stream
    s = torch.cuda.Stream()
    a1, b1 = torch.empty_like(u), torch.empty_like(u)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        three_steps(a1, b1)
    torch.cuda.current_stream().wait_stream(s)
    assert torch.equal(a1, want)
    # captured into a CUDA graph and replayed
    a2, b2 = torch.empty_like(u), torch.empty_like(u)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        three_steps(a2, b2)
    a2.fill_(float("nan"))
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(a2, want)


def test_tuner_keeps_a_bit_identical_schedule(cuda_device):
    shape = (256, 256, 256)                      # = PseudoSpectralIMEX.TUNE_MIN_VOXELS
    vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
    vg = VoxelGridTorch(vf.grid_info(), device="cuda")
    ts = PseudoSpectralIMEX(CahnHilliard(vg), 0.1)
    g = torch.Generator(device="cuda").manual_seed(0)
    u = 0.5 + 0.1 * torch.rand((1,) + shape, device="cuda", generator=g)
    got = ts.step(0.0, u)
    plan = next(iter(ts._plans.values()))
    assert plan.tuned and plan.tune_report is not None
    rep = plan.tune_report
    assert rep["candidates"], rep
    assert all(c["bit_identical"] for c in rep["candidates"]), rep
    assert plan.schedule() == tuple(rep["chosen"])
    assert rep["chosen_ms"] <= rep["baseline_ms"]
    chosen = plan.schedule()
    plan.set_schedule(0, 1, 0)
    want = ts.step(0.0, u)
    assert torch.equal(got, want)
    plan.set_schedule(*chosen)
    ref = O.CHOracle(shape, vf.spacing, 0.1).step(u.cpu())
    assert rel_l2(got.cpu().numpy(), ref.numpy()) <= 1e-5
    print("tuned schedule", chosen, "%.3f -> %.3f ms" % (rep["baseline_ms"], rep["chosen_ms"]))


@pytest.mark.parametrize("world,l2", [(2, 0), (4, 4), (8, 2), (2, 8)])
def test_block_copy_transport_with_virtual_ranks_and_l2_blocking(cuda_device, world, l2):
    """The copy-engine entry points (forward_chunk / middle_chunk with the self block written in
    place) for W plans on ONE GPU, blocks moved with plain tensor copies as the DMA engines do
    between real ranks, pipeline chunks and the L2 sub-chunking of the transform pairs: the
    assembled result equals the single-GPU spectral stage bit for bit."""
    shape, sp = (64, 32, 64), (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(12)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, sp, 0.1, 1.5, 2)
    plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
    for p in plans:
        p.set_l2_planes(l2)
    nan = float("nan")
    A = [p.new_buffer().fill_(nan) for p in plans]
    B = [p.new_buffer().fill_(nan) for p in plans]
    spec = [p.new_buffer().fill_(nan) for p in plans]
    nxl, nyl = shape[0] // world, shape[1] // world
    fwd_chunks = 2 if nxl >= 2 else 1
    for k, p in enumerate(plans):                       # forward: blocks j != k into A[k], block k into B[k]
        rk = r[k * nxl:(k + 1) * nxl].contiguous()
        for i in range(fwd_chunks):
            x0, x1 = i * nxl // fwd_chunks, (i + 1) * nxl // fwd_chunks
            p.forward_chunk(rk, spec[k], A[k], x0, x1 - x0, self_block=B[k])
    for k in range(world):                              # "DMA": block j of A[k] -> block k of B[j]
        for j in range(world):
            if j != k:
                B[j][k].copy_(A[k][j])
    for a in A:
        a.fill_(nan)
    mid_chunks = 2 if nyl >= 2 else 1
    for k, p in enumerate(plans):                       # x pass: in place in B[k], block k into A[k]
        for i in range(mid_chunks):
            y0, y1 = i * nyl // mid_chunks, (i + 1) * nyl // mid_chunks
            p.middle_chunk(B[k], y0, y1 - y0, sp, 0.1, 1.5, 2, self_block=A[k])
    for k in range(world):
        for j in range(world):
            if j != k:
                A[j][k].copy_(B[k][j])
    out = torch.empty_like(u)
    for k, p in enumerate(plans):
        o = torch.empty((nxl,) + shape[1:], device="cuda")
        p.backward(A[k], spec[k].fill_(nan), u[k * nxl:(k + 1) * nxl].contiguous(), o)
        out[k * nxl:(k + 1) * nxl] = o
    assert torch.equal(out, ref)
