"""Hand-written adjoint (evoxels_b200.autograd) against (a) gradients recorded from the
reference by torch autograd (golden ch_grad_f64), (b) live autograd through the CPU oracle,
(c) central finite differences.  BASELINE config 5 (inversion) in miniature."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import evx_oracle as O

pytestmark = pytest.mark.gpu

import evoxels_b200 as evo  # noqa: E402
from evoxels_b200.problem_definition import CahnHilliard  # noqa: E402
from evoxels_b200.timesteppers import PseudoSpectralIMEX  # noqa: E402
from evoxels_b200.voxelgrid import VoxelGridTorch  # noqa: E402


def _grid(shape, spacing, precision):
    dom = tuple(float(n * h) for n, h in zip(shape, spacing))
    vf = evo.VoxelFields(tuple(int(n) for n in shape), dom)
    vf.precision = precision
    return VoxelGridTorch(vf.grid_info(), precision=precision, device="cuda")


def _loss_and_grads(vg, u0, target, D0, eps0, dt, nsteps, dtype):
    D = torch.tensor(D0, dtype=torch.float64, device="cuda", requires_grad=True)
    eps = torch.tensor(eps0, dtype=torch.float64, device="cuda", requires_grad=True)
    u = torch.as_tensor(u0, dtype=dtype, device="cuda")[None].requires_grad_(True)
    ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=eps, D=D), dt)
    v = u
    for _ in range(nsteps):
        v = ts.step(0.0, v)
    loss = ((v - torch.as_tensor(target, dtype=dtype, device="cuda")[None]) ** 2).sum()
    gu, gD, ge = torch.autograd.grad(loss, (u, D, eps))
    return float(loss), gu[0].cpu().numpy(), float(gD), float(ge), v.detach()


def test_gradients_match_reference_autograd_record(cuda_device):
    g = load_golden("ch_grad_f64")
    vg = _grid(g["u0"].shape, g["spacing"], "float64")
    loss, gu, gD, ge, v = _loss_and_grads(vg, g["u0"], g["target"], g["D"], g["eps"], g["dt"],
                                          g["nsteps"], torch.float64)
    assert rel_l2(v[0].cpu().numpy(), g["final"]) <= 1e-7
    assert abs(loss - g["loss"]) <= 1e-6 * abs(g["loss"])
    assert rel_l2(gu, g["grad_u0"]) <= 1e-6
    assert abs(gD - g["grad_D"]) <= 1e-5 * abs(g["grad_D"])
    assert abs(ge - g["grad_eps"]) <= 1e-5 * abs(g["grad_eps"])


@pytest.mark.parametrize("shape,nsteps", [((32, 32, 32), 10), ((16, 32, 64), 5)])
def test_fp32_gradients_match_oracle_autograd(cuda_device, shape, nsteps):
    spacing, dt, D0, eps0 = (1.0, 1.0, 1.0), 0.1, 2.0, 2.0
    u0 = O.noise_field(shape, seed=5, lo=0.1, amp=0.8)[0].numpy()
    tgt = O.noise_field(shape, seed=6, lo=0.45, amp=0.1)[0].numpy()
    # oracle: float64 autograd on CPU (the product sets torch's default device to cuda)
    with torch.device("cpu"):
        D = torch.tensor(D0, dtype=torch.float64, requires_grad=True)
        eps = torch.tensor(eps0, dtype=torch.float64, requires_grad=True)
        u = torch.from_numpy(u0).double()[None].requires_grad_(True)
        v = u
        for _ in range(nsteps):
            v = O.ch_imex_step(v, spacing, dt, eps, D, 0.25)
        loss_ref = ((v - torch.from_numpy(tgt).double()[None]) ** 2).sum()
        ru, rD, re = torch.autograd.grad(loss_ref, (u, D, eps))
    vg = _grid(shape, spacing, "float32")
    loss, gu, gD, ge, _ = _loss_and_grads(vg, u0, tgt, D0, eps0, dt, nsteps, torch.float32)
    assert abs(loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert rel_l2(gu, ru[0].numpy()) <= 1e-4
    assert abs(gD - float(rD)) <= 2e-3 * abs(float(rD)) + 1e-6
    assert abs(ge - float(re)) <= 2e-3 * abs(float(re)) + 1e-6


def test_parameter_gradients_match_finite_differences(cuda_device):
    shape, spacing, dt, nsteps = (16, 16, 16), (1.0, 1.0, 1.0), 0.1, 4
    vg = _grid(shape, spacing, "float64")
    u0 = 0.2 + 0.6 * np.random.default_rng(8).random(shape)
    tgt = np.full(shape, 0.5)

    def loss_at(D0, eps0):
        ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=eps0, D=D0), dt)
        v = torch.as_tensor(u0, dtype=torch.float64, device="cuda")[None]
        for _ in range(nsteps):
            v = ts.step(0.0, v)
        return float(((v - 0.5) ** 2).sum())

    _, _, gD, ge, _ = _loss_and_grads(vg, u0, tgt, 1.5, 2.5, dt, nsteps, torch.float64)
    h = 1e-2     # the float32 wavenumber/prefactor arithmetic makes the loss noisy at ~1e-7
    fdD = (loss_at(1.5 + h, 2.5) - loss_at(1.5 - h, 2.5)) / (2 * h)
    fde = (loss_at(1.5, 2.5 + h) - loss_at(1.5, 2.5 - h)) / (2 * h)
    assert abs(gD - fdD) <= 5e-3 * abs(fdD) + 1e-9, (gD, fdD)
    assert abs(ge - fde) <= 5e-3 * abs(fde) + 1e-9, (ge, fde)


def test_rhs_alone_is_differentiable(cuda_device):
    shape, spacing = (12, 10, 8), (1.0, 0.5, 2.0)
    vg = _grid(shape, spacing, "float64")
    u0 = (-0.2 + 1.4 * np.random.default_rng(2).random(shape))
    w0 = np.random.default_rng(3).standard_normal(shape)
    D = torch.tensor(1.3, dtype=torch.float64, device="cuda", requires_grad=True)
    eps = torch.tensor(2.5, dtype=torch.float64, device="cuda", requires_grad=True)
    u = torch.as_tensor(u0, device="cuda")[None].requires_grad_(True)
    R = CahnHilliard(vg, eps=eps, D=D).rhs(0.0, u)
    gu, gD, ge = torch.autograd.grad((R * torch.as_tensor(w0, device="cuda")[None]).sum(), (u, D, eps))
    with torch.device("cpu"):
        ut = torch.from_numpy(u0)[None].requires_grad_(True)
        Dc = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
        ec = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
        Rc = O.ch_rhs(ut, spacing, ec, Dc)
        ru, rD, re = torch.autograd.grad((Rc * torch.from_numpy(w0)[None]).sum(), (ut, Dc, ec))
    assert rel_l2(gu[0].cpu().numpy(), ru[0].numpy()) <= 1e-11
    assert abs(float(gD) - float(rD)) <= 1e-9 * abs(float(rD))
    assert abs(float(ge) - float(re)) <= 1e-9 * abs(float(re))


def test_slab_form_of_the_adjoint_kernels(cuda_device):
    """The x-slab form of the backward pass on one GPU: (a) DistributedCahnHilliardIMEX with one
    rank reproduces the single-GPU autograd step; (b) the stencil adjoint on a slab extended by
    ADJ_HALO periodic image planes with the dL/deps sum restricted to the slab's own planes
    (evx_ch_adjoint_combine_range) - the slabs' fields tile the full-domain dL/du and their
    partial sums add up to the full-domain dL/deps."""
    from evoxels_b200 import _native
    from evoxels_b200.distributed import DistributedCahnHilliardIMEX
    shape, spacing = (32, 16, 64), (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(4)
    u = -0.05 + 1.1 * torch.rand(shape, device="cuda", generator=gen)
    w = torch.randn(shape, device="cuda", generator=gen)
    lam_ref, deps_ref = _native.ch_rhs_vjp(u, w, spacing, 2.5, 1.3)
    H = DistributedCahnHilliardIMEX.ADJ_HALO
    nxl, total = 8, 0.0
    for a in range(0, shape[0], nxl):
        idx = torch.arange(a - H, a + nxl + H, device="cuda") % shape[0]
        lam, deps = _native.ch_rhs_vjp(u[idx].contiguous(), w[idx].contiguous(), spacing, 2.5, 1.3,
                                       deps_planes=(H, H + nxl))
        assert torch.allclose(lam[H:H + nxl], lam_ref[a:a + nxl], rtol=0, atol=1e-6 * float(lam_ref.abs().max()))
        total += float(deps)
    assert abs(total - float(deps_ref)) <= 1e-9 * abs(float(deps_ref)) + 1e-12
    # (a) one rank
    tgt = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    D1 = torch.tensor(1.3, dtype=torch.float64, device="cuda", requires_grad=True)
    e1 = torch.tensor(2.5, dtype=torch.float64, device="cuda", requires_grad=True)
    st = DistributedCahnHilliardIMEX(shape, spacing, 0.1, eps=2.5, D=1.3, device="cuda")
    x = u.clone().requires_grad_(True)
    y = st.step_autograd(st.step_autograd(x, D1, e1), D1, e1)
    ((y - tgt) ** 2).sum().backward()
    import evoxels_b200 as evo
    from evoxels_b200.problem_definition import CahnHilliard
    from evoxels_b200.timesteppers import PseudoSpectralIMEX
    from evoxels_b200.voxelgrid import VoxelGridTorch
    vf = evo.VoxelFields(shape, tuple(float(n * h) for n, h in zip(shape, spacing)))
    vg = VoxelGridTorch(vf.grid_info(), device="cuda")
    D2 = torch.tensor(1.3, dtype=torch.float64, device="cuda", requires_grad=True)
    e2 = torch.tensor(2.5, dtype=torch.float64, device="cuda", requires_grad=True)
    ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=e2, D=D2), 0.1, fft_backend="native")
    x2 = u[None].clone().requires_grad_(True)
    y2 = ts.step(0.0, ts.step(0.0, x2))
    ((y2[0] - tgt) ** 2).sum().backward()
    assert float((x.grad - x2.grad[0]).norm() / x2.grad.norm()) < 1e-6
    assert abs(float(D1.grad) - float(D2.grad)) <= 1e-6 * abs(float(D2.grad))
    assert abs(float(e1.grad) - float(e2.grad)) <= 1e-6 * abs(float(e2.grad))
