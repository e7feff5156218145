"""Multi-GPU (x-slab) parity: launched through torchrun on 2 GPUs when the box has them."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_matches_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "DIST-GPU OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_single_rank_distributed_plan_equals_single_gpu_path(cuda_device):
    """world == 1: the distributed stage calls (block layout with one block) must reproduce
    the single-GPU native step bit for bit."""
    import evoxels_b200 as evo
    from evoxels_b200.distributed import DistributedCahnHilliardIMEX
    from evoxels_b200.problem_definition import CahnHilliard
    from evoxels_b200.timesteppers import PseudoSpectralIMEX
    from evoxels_b200.voxelgrid import VoxelGridTorch
    shape, spacing = (64, 32, 128), (1.0, 1.0, 1.0)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda")
    vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
    vg = VoxelGridTorch(vf.grid_info(), device="cuda")
    ref = PseudoSpectralIMEX(CahnHilliard(vg), 0.1, fft_backend="native").step(0.0, u[None])[0]
    got = DistributedCahnHilliardIMEX(shape, spacing, 0.1, device="cuda").step(u)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_p2p_block_addressing_with_virtual_ranks(cuda_device, world):
    """Peer-store mode of the distributed passes on ONE GPU: W plans ("virtual ranks") write
    into each other's buffers exactly as real ranks do over NVLink; the assembled result must
    equal the single-GPU spectral stage bit for bit."""
    from evoxels_b200 import _native
    shape, sp = (64, 32, 64), (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(11)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, sp, 0.1, 1.5, 2)
    plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
    A = [p.new_buffer().zero_() for p in plans]
    B = [p.new_buffer().zero_() for p in plans]
    spec = [p.new_buffer() for p in plans]
    nxl = shape[0] // world
    for k, p in enumerate(plans):
        p.forward_p2p(r[k * nxl:(k + 1) * nxl].contiguous(), spec[k], [b.data_ptr() for b in B])
    for k, p in enumerate(plans):
        p.middle_p2p(B[k], [a.data_ptr() for a in A], sp, 0.1, 1.5, 2)
    out = torch.empty_like(u)
    for k, p in enumerate(plans):
        o = torch.empty((nxl,) + shape[1:], device="cuda")
        p.backward(A[k], spec[k], u[k * nxl:(k + 1) * nxl].contiguous(), o)
        out[k * nxl:(k + 1) * nxl] = o
    assert torch.equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,world,chunks", [((1024, 64, 32), 2, 1), ((1024, 64, 32), 8, 3),
                                                ((64, 1024, 32), 4, 2), ((64, 1024, 32), 2, 2),
                                                ((1024, 1024, 16), 8, 4)])
@pytest.mark.parametrize("transport", ["local", "p2p"])
def test_four_stage_tma_passes_in_the_distributed_plan_with_virtual_ranks(cuda_device, monkeypatch, shape,
                                                                            world, chunks, transport):
    """1024-point y / x lines of the x-slab plan through the four-stage TMA-tiled kernel, W plans
    ("virtual ranks") on ONE GPU: `local` = block buffers + explicit block exchange (what the
    copy-engine / NCCL transports do, own block through the second buffer), `p2p` = the TMA stores
    of the passes land in the other plans' buffers (per-destination tensor maps).  Both must equal
    the single-GPU spectral stage bit for bit, with and without the four-stage kernel."""
    from evoxels_b200 import _native
    sp = (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(11)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    monkeypatch.setenv("EVX_FFT_LINE4", "0")
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, sp, 0.1, 1.5, 2)
    nxl, nyl = shape[0] // world, shape[1] // world
    xb = [round(i * nxl / chunks) for i in range(chunks + 1)]
    yb = [round(i * nyl / chunks) for i in range(chunks + 1)]
    for flag in ("0", "1"):
        monkeypatch.setenv("EVX_FFT_LINE4", flag)
        plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
        send = [p.new_buffer().zero_() for p in plans]
        A = [p.new_buffer().fill_(float("nan")) for p in plans]
        B = [p.new_buffer().fill_(float("nan")) for p in plans]
        spec = [p.new_buffer() for p in plans]
        for k, p in enumerate(plans):
            rl = r[k * nxl:(k + 1) * nxl].contiguous()
            for i in range(chunks):
                if transport == "p2p":
                    p.set_p2p_ctas(148 if chunks > 1 else 0)
                    p.forward_chunk_p2p(rl, spec[k], [b.data_ptr() for b in B], xb[i], xb[i + 1] - xb[i], parts=1)
                    p.forward_chunk_p2p(rl, spec[k], [b.data_ptr() for b in B], xb[i], xb[i + 1] - xb[i], parts=2)
                else:
                    p.forward_chunk(rl, spec[k], send[k], xb[i], xb[i + 1] - xb[i], self_block=B[k])
        if transport == "local":
            for k in range(world):
                for j in range(world):
                    if j != k:
                        B[k][j].copy_(send[j][k])
        for k, p in enumerate(plans):
            if transport == "p2p":
                p.middle_p2p(B[k], [a.data_ptr() for a in A], sp, 0.1, 1.5, 2)
            else:
                for i in range(chunks):
                    p.middle_chunk(B[k], yb[i], yb[i + 1] - yb[i], sp, 0.1, 1.5, 2, self_block=A[k])
        if transport == "local":
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
        assert torch.equal(out, ref), (flag, transport)


@pytest.mark.parametrize("transport,chunks", [("local", 1), ("local", 2), ("p2p", 1)])
def test_chained_zy_kernels_in_the_distributed_plan_with_virtual_ranks(cuda_device, monkeypatch, transport, chunks):
    """2 ranks, 512-point y and z lines: the chained z/y kernels write / read the all-to-all block
    layout through one tensor map per 256-row box (fft_chain.cu, ChainMaps::blk) - local block
    buffers with an explicit exchange, or straight into the other plan's buffer.  Bit-identical to
    the single-GPU spectral stage and to the un-chained distributed passes."""
    from evoxels_b200 import _native
    shape, world, sp = (64, 512, 512), 2, (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(21)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    ref = torch.empty_like(u)
    _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE).apply(u, r, ref, sp, 0.1, 1.5, 2)
    nxl = shape[0] // world
    xb = [round(i * nxl / chunks) for i in range(chunks + 1)]
    for chain in ("0", "1"):
        monkeypatch.setenv("EVX_FFT_CHAIN", chain)
        plans = [_native.DistPlan(shape, world, k, "cuda") for k in range(world)]
        send = [p.new_buffer().zero_() for p in plans]
        A = [p.new_buffer().fill_(float("nan")) for p in plans]
        B = [p.new_buffer().fill_(float("nan")) for p in plans]
        spec = [p.new_buffer() for p in plans]
        for k, p in enumerate(plans):
            rl = r[k * nxl:(k + 1) * nxl].contiguous()
            if transport == "p2p":
                p.forward_p2p(rl, spec[k], [b.data_ptr() for b in B])
            else:
                for i in range(chunks):
                    p.forward_chunk(rl, spec[k], send[k], xb[i], xb[i + 1] - xb[i], self_block=B[k])
        if transport == "local":
            for k in range(world):
                B[k][1 - k].copy_(send[1 - k][k])
        for k, p in enumerate(plans):
            p.middle_p2p(B[k], [a.data_ptr() for a in A], sp, 0.1, 1.5, 2)
        out = torch.empty_like(u)
        for k, p in enumerate(plans):
            o = torch.empty((nxl,) + tuple(shape[1:]), device="cuda")
            p.backward(A[k], spec[k], u[k * nxl:(k + 1) * nxl].contiguous(), o)
            out[k * nxl:(k + 1) * nxl] = o
        torch.cuda.synchronize()
        assert torch.equal(out, ref), (chain, transport, chunks)
