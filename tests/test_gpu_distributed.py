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
