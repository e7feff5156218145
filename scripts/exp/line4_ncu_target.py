"""ncu target: the three strided passes of a 1024 x 1024 x 256 grid through the four-stage TMA-tiled
kernel (y forward, x forward*weight*inverse, y inverse), a few launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evoxels_b200 import _native  # noqa: E402

shape = (1024, 1024, 256)
u = torch.rand(shape, device="cuda")
r = torch.randn(shape, device="cuda")
out = torch.empty_like(u)
plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
os.environ["EVX_FFT_CHAIN"] = "0"
plan.apply(u, r, out, (1.0, 1.0, 1.0), 0.1, 1.5, 2)
for _ in range(3):
    for which in (1, 2, 3):
        plan.native_pass(which, u, r, out, (1.0, 1.0, 1.0), 0.1, 1.5, 2)
torch.cuda.synchronize()
print("done")
