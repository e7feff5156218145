"""BASELINE configs 3 and 5 at their full sizes, plus Cahn-Hilliard at 1024^3 on one GPU.
The CPU oracle cannot run these sizes as a whole, so the checks are size-independent properties
and parity on a sub-problem the oracle can do:
  * Allen-Cahn 1024^3, Neumann, ForwardEuler (config 3): determinism, a uniform field stays
    uniform and moves by the closed-form rate, and - the stencil has radius 1 - the corner block
    of the result equals the oracle's step of the corner block of the input;
  * inversion (config 5): gradients of a three-observation misfit through 100 steps at 256^3
    against central finite differences of the same loss;
  * Cahn-Hilliard IMEX 1024^3: mass conservation, determinism, translation equivariance.
(The file sorts last on purpose: these are the slowest GPU tests.)"""
import warnings

import pytest
import torch

from conftest import rel_l2
from oracle import evx_oracle as O

pytestmark = pytest.mark.gpu

import evoxels_b200 as evo  # noqa: E402
from evoxels_b200.problem_definition import CahnHilliard, TwoPhaseAllenCahn  # noqa: E402
from evoxels_b200.timesteppers import ForwardEuler, PseudoSpectralIMEX  # noqa: E402
from evoxels_b200.voxelgrid import VoxelGridTorch  # noqa: E402


def _grid(n):
    vf = evo.VoxelFields((n, n, n), (float(n),) * 3)          # spacing 1
    return VoxelGridTorch(vf.grid_info(), device="cuda")


def _need_memory(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * (1 << 30):
        pytest.skip(f"needs {gib} GiB of free device memory")


def test_config3_allen_cahn_1024_neumann_euler(cuda_device):
    _need_memory(40)
    n, dt = 1024, 0.05
    vg = _grid(n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob = TwoPhaseAllenCahn(vg)                         # defaults: eps=2, gab=1, M=1, Neumann x3
    ts = ForwardEuler(prob, dt)
    gen = torch.Generator(device="cuda").manual_seed(1)
    phi = torch.rand((1, n, n, n), device="cuda", generator=gen)
    before = phi[:, :8, :8, :8].clone()
    v = ts.step(0.0, phi)
    assert torch.equal(phi[:, :8, :8, :8], before), "step must not mutate its input"
    assert torch.equal(v, ts.step(0.0, phi))
    assert bool(torch.isfinite(v).all())
    # parity on the corner block: 19-point stencil, radius 1 -> block of b cells is exact on b-1
    b = 66
    ref = O.ACOracle((b, b, b), (1.0, 1.0, 1.0), dt).step(phi[:, :b, :b, :b].cpu().contiguous())
    got = v[:, :b - 1, :b - 1, :b - 1].cpu()
    assert rel_l2(got.numpy(), ref[:, :b - 1, :b - 1, :b - 1].numpy()) <= 1e-6
    # ... and on the opposite corner (high-side ghosts)
    ref = O.ACOracle((b, b, b), (1.0, 1.0, 1.0), dt).step(phi[:, -b:, -b:, -b:].cpu().contiguous())
    got = v[:, -(b - 1):, -(b - 1):, -(b - 1):].cpu()
    assert rel_l2(got.numpy(), ref[:, 1:, 1:, 1:].numpy()) <= 1e-6
    del v, ref, got
    # uniform field: every voxel does the same arithmetic; rate = -g(c)/(2 eps), g = 18/eps c(1-c)(1-2c)
    c = 0.3
    w = ts.step(0.0, torch.full((1, n, n, n), c, device="cuda"))
    assert float(w.max()) - float(w.min()) <= 1e-7
    assert abs(float(w[0, 0, 0, 0]) - (c + dt * (-(9.0 * c * (1 - c) * (1 - 2 * c)) / 4.0))) <= 1e-6


def test_config5_inversion_256_100_steps(cuda_device):
    _need_memory(40)
    n, dt, nsteps, obs_at = 256, 0.1, 100, (33, 66, 100)
    vg = _grid(n)
    gen = torch.Generator(device="cuda").manual_seed(0)
    u0 = 0.5 + 0.1 * torch.rand((1, n, n, n), device="cuda", generator=gen)

    def run(D, eps, keep):
        ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=eps, D=D), dt)
        v, out = u0, []
        for i in range(1, nsteps + 1):
            v = ts.step(0.0, v)
            if i in keep:
                out.append(v)
        return out

    with torch.no_grad():
        obs = [o.clone() for o in run(1.0, 3.0, obs_at)]       # "measurements": true D = 1, eps = 3

    def loss_of(states):
        return sum(((s.double() - o.double()) ** 2).sum() for s, o in zip(states, obs))

    D = torch.tensor(2.0, dtype=torch.float64, device="cuda", requires_grad=True)
    eps = torch.tensor(2.0, dtype=torch.float64, device="cuda", requires_grad=True)
    loss = loss_of(run(D, eps, obs_at))
    gD, ge = torch.autograd.grad(loss, (D, eps))
    assert bool(torch.isfinite(gD)) and bool(torch.isfinite(ge)) and float(loss) > 0
    h = 1e-2
    with torch.no_grad():
        fdD = (float(loss_of(run(2.0 + h, 2.0, obs_at))) - float(loss_of(run(2.0 - h, 2.0, obs_at)))) / (2 * h)
        fde = (float(loss_of(run(2.0, 2.0 + h, obs_at))) - float(loss_of(run(2.0, 2.0 - h, obs_at)))) / (2 * h)
    # fp32 states through 100 steps of a linearly unstable (spinodal) regime: 5 % is the bar
    assert abs(float(gD) - fdD) <= 5e-2 * abs(fdD) + 1e-6, (float(gD), fdD)
    assert abs(float(ge) - fde) <= 5e-2 * abs(fde) + 1e-6, (float(ge), fde)


def test_cahn_hilliard_1024_properties(cuda_device):
    _need_memory(60)
    n = 1024
    vg = _grid(n)
    ts = PseudoSpectralIMEX(CahnHilliard(vg), 0.1)
    gen = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand((1, n, n, n), device="cuda", generator=gen).mul_(0.1).add_(0.5)
    v = ts.step(0.0, u)
    m0, m1 = float(u.sum(dtype=torch.float64)), float(v.sum(dtype=torch.float64))
    assert abs(m1 - m0) <= 4e-7 * abs(m0)
    assert torch.equal(v, ts.step(0.0, u))
    shift = (3, 129, 64)
    vs = ts.step(0.0, torch.roll(u, shift, (1, 2, 3)))
    vs -= torch.roll(v, shift, (1, 2, 3))
    assert float(vs.norm() / v.norm()) <= 1e-6
