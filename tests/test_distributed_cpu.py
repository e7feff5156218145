"""Multi-rank orchestration (x-slab halos, slab<->pencil all-to-all, global offsets) on CPU:
world_size 2 and 4 with the gloo backend and an oracle stand-in for the kernels."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,shape", [(2, "16x8x12"), (4, "16x16x10"), (1, "8x8x8")])
def test_distributed_steps_match_single_domain_oracle(world, shape):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), str(world),
           str(_free_port()), shape]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_slab_geometry_and_validation():
    from evoxels_b200.distributed import Slab
    s = Slab((16, 8, 4), 4, 2)
    assert (s.nxl, s.nyl, s.x0, s.local_shape) == (4, 2, 8, (4, 8, 4))
    with pytest.raises(ValueError):
        Slab((10, 8, 4), 4, 0)
