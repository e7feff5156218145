"""Differentiable Cahn-Hilliard step: `torch.autograd.Function`s whose backward passes are
the hand-written adjoint kernels (csrc/adjoint_core.h) plus the same spectral filter as the
forward pass.

Capability parity: the reference estimates (D, eps) by differentiating N Cahn-Hilliard
steps with JAX/diffrax (evoxels/inversion.py:51-124, `InversionModel`).  On torch its step
is differentiable only through ~60 autograd nodes per step; here one node per step stores
just the step's input and output (u_n, u_{n+1}) and evaluates

    w      = G lam+                        G = F^-1 diag(P) F  (self-adjoint, same filter)
    lam    = lam+ + dR/du^T w              three stencil kernels
    dL/dD   = <w, u+ - u> / (dt D)
    dL/deps = < z, d mu / d eps > + < w/dt - lam+, u+ - u > / eps

(derivation in DESIGN.md section 8; uses dP/ds = -P^2 and s P = 1 - P/dt, so no second
filtered transform and no stored spectra are needed).  Fully periodic grids and the default
double-well potential only; anything else raises.
"""
from __future__ import annotations

import torch

from . import _native


def _as_param(value, like):
    if isinstance(value, torch.Tensor):
        return value
    return torch.tensor(float(value), dtype=torch.float64, device=like.device)


def _check_supported(problem):
    if problem.bc_type != ("periodic",) * 3:
        raise NotImplementedError("the hand-written adjoint supports fully periodic grids only")
    if not getattr(problem, "_default_mu", True):
        raise NotImplementedError("the hand-written adjoint supports the default mu_hom only")


class _CHRhsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, D, eps, spacing):
        out = torch.empty_like(u)
        per = (("periodic", None),) * 3
        for ch in range(u.shape[0]):
            _native.ch_rhs(u[ch], out[ch], spacing, float(eps), float(D), per)
        ctx.save_for_backward(u, out, D, eps)
        ctx.spacing = spacing
        return out

    @staticmethod
    def backward(ctx, w):
        u, R, D, eps = ctx.saved_tensors
        w = w.contiguous()
        lam = torch.empty_like(u)
        g_eps = torch.zeros((), dtype=torch.float64, device=u.device)
        for ch in range(u.shape[0]):
            l, de = _native.ch_rhs_vjp(u[ch], w[ch], ctx.spacing, float(eps), float(D))
            lam[ch] = l
            g_eps = g_eps + de
        g_D = (w.double() * R.double()).sum() / float(D)
        return lam, g_D.to(D.dtype), g_eps.to(eps.dtype), None


class _CHImexStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, D, eps, stepper):
        prob = stepper.problem
        out = torch.empty_like(u)
        plan = stepper._plan(u.shape[1:], u.dtype, u.device)
        for ch in range(u.shape[0]):
            plan.ch_step(u[ch], out[ch], prob.vg.spacing, stepper.dt, float(eps), float(D), prob.A)
        ctx.save_for_backward(u, out, D, eps)
        ctx.stepper = stepper
        return out

    @staticmethod
    def backward(ctx, lam_plus):
        u, u_plus, D, eps = ctx.saved_tensors
        stepper = ctx.stepper
        prob = stepper.problem
        spacing, dt = prob.vg.spacing, stepper.dt
        lam_plus = lam_plus.contiguous()
        plan = stepper._plan(u.shape[1:], u.dtype, u.device)
        coef = 2.0 * float(eps) * float(D) * float(prob.A)
        lam = torch.empty_like(u)
        w = torch.empty_like(u)
        g_eps = torch.zeros((), dtype=torch.float64, device=u.device)
        for ch in range(u.shape[0]):
            plan.apply(None, lam_plus[ch], w[ch], spacing, dt, coef, 2)        # w = G lam+
            l, de = _native.ch_rhs_vjp(u[ch], w[ch], spacing, float(eps), float(D),
                                       lam_in=lam_plus[ch])
            lam[ch] = l
            g_eps = g_eps + de
        du = (u_plus - u).double()
        wd = w.double()
        g_D = (wd * du).sum() / (dt * float(D))
        g_eps = g_eps + ((wd / dt - lam_plus.double()) * du).sum() / float(eps)
        return lam, g_D.to(D.dtype), g_eps.to(eps.dtype), None


def ch_rhs_autograd(problem, c):
    _check_supported(problem)
    _native.require_cuda(c)
    return _CHRhsFn.apply(c.contiguous(), _as_param(problem.D, c), _as_param(problem.eps, c),
                          problem.vg.spacing)


def ch_imex_step_autograd(stepper, u):
    from .problem_definition import CahnHilliard
    prob = stepper.problem
    if not isinstance(prob, CahnHilliard):
        raise NotImplementedError("differentiable IMEX steps are implemented for CahnHilliard")
    _check_supported(prob)
    _native.require_cuda(u)
    return _CHImexStepFn.apply(u.contiguous(), _as_param(prob.D, u), _as_param(prob.eps, u), stepper)
