"""Single-step time integrators with the reference's `TimeStepper.step(t, u) -> u` seam.

Interface parity with evoxels/timesteppers.py: `TimeStepper` (:9-29), `ForwardEuler`
(:32-43), `RungeKutta4` (:46-61), `PseudoSpectralIMEX` (:64-89) - dataclasses constructed
as `(problem, dt)`, pure functions of `(t, u)` that never mutate `u`.

How a step runs here:
* `PseudoSpectralIMEX` + `CahnHilliard` on a fully periodic grid: ONE C call
  (`evx_ch_imex_step_*`): fused rhs kernel -> forward FFT -> on-the-fly filter -> inverse
  FFT -> `u + update`.  No prefactor array is stored.
* `PseudoSpectralIMEX` + any `SemiLinearODE` with a closed-form symbol: `problem.rhs`
  followed by `evx_imex_apply_*` (mirror extension in x for Neumann/Dirichlet).
* explicit steppers + a problem that offers `fused_stage` (TwoPhaseAllenCahn): one kernel
  per stage computing rhs and the stage's axpy together.
* anything else: the textbook composition on CUDA tensors.
* `ExponentialEuler` (:136-202): same pipeline as the IMEX stepper with
  the weight dt*phi1(dt*symbol) evaluated inside the x pass (`EVX_FILTER_ETD1`).
The diffrax clone and the RKC steppers of the reference are outside the accelerated path
and not provided.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any

import torch

from . import _native
from .problem_definition import ODE, CahnHilliard, SemiLinearODE, is_stock

State = Any


class TimeStepper(ABC):
    @property
    @abstractmethod
    def order(self) -> int:
        """Temporal order of accuracy."""

    @abstractmethod
    def step(self, t: float, u: State) -> State:
        """Advance `u` from t to t + dt and return the new state (input untouched)."""


@dataclass
class ForwardEuler(TimeStepper):
    problem: ODE
    dt: float

    @property
    def order(self) -> int:
        return 1

    def step(self, t, u):
        if getattr(self.problem.vg, "slab", None) is not None:
            return _distributed_ac_euler(self, u)
        stage = getattr(self.problem, "fused_stage", None)
        if stage is None or u.requires_grad:
            return u + self.dt * self.problem.rhs(t, u)
        out = torch.empty_like(u, memory_format=torch.contiguous_format)
        u = u.contiguous()
        stage(u, base=u, alpha=self.dt, y_out=out)
        return out


@dataclass
class RungeKutta4(TimeStepper):
    problem: ODE
    dt: float

    @property
    def order(self) -> int:
        return 4

    def step(self, t, u):
        dt, f = self.dt, self.problem.rhs
        stage = getattr(self.problem, "fused_stage", None)
        if stage is None or u.requires_grad:
            k1 = f(t, u)
            k2 = f(t + 0.5 * dt, u + 0.5 * dt * k1)
            k3 = f(t + 0.5 * dt, u + 0.5 * dt * k2)
            k4 = f(t + dt, u + dt * k3)
            return u + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
        # fused: every stage kernel writes the next stage input AND accumulates the result
        u = u.contiguous()
        ya, yb, acc = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
        stage(u, base=u, alpha=0.5 * dt, y_out=ya, acc_in=u, beta=dt / 6, acc_out=acc)
        stage(ya, base=u, alpha=0.5 * dt, y_out=yb, acc_in=acc, beta=dt / 3, acc_out=acc)
        stage(yb, base=u, alpha=dt, y_out=ya, acc_in=acc, beta=dt / 3, acc_out=acc)
        stage(ya, acc_in=acc, beta=dt / 6, acc_out=acc)
        return acc


# ---- x-slab decomposed grids (VoxelGridTorch.slab): the stock problem / stepper pairs run the
# distributed kernels of distributed.py behind the same step(t, u) call -----------------------
def _require_no_grad_distributed(u):
    if u.requires_grad:
        raise NotImplementedError("backprop through the multi-GPU step is not available; "
                                  "run the inversion on one GPU")


def _distributed_ac_euler(ts, u):
    from .distributed import DistributedAllenCahnEuler
    from .problem_definition import TwoPhaseAllenCahn, is_stock
    prob = ts.problem
    _require_no_grad_distributed(u)
    if not is_stock(prob, TwoPhaseAllenCahn, ("rhs",)) or not getattr(prob, "_default_potential", False):
        raise NotImplementedError("on an x-slab decomposed grid ForwardEuler steps the stock "
                                  "TwoPhaseAllenCahn (default potential) only")
    if getattr(ts, "_dist", None) is None:
        vg = prob.vg
        ts._dist = DistributedAllenCahnEuler(vg.shape, vg.spacing, ts.dt, eps=float(prob.eps), gab=float(prob.gab),
                                             M=float(prob.M), force=float(prob.force),
                                             curvature=float(prob.curvature), bc=prob.bc, device=vg.device)
    return torch.stack([ts._dist.step(ch) for ch in u.contiguous()], 0)


def _distributed_ch_imex(ts, u):
    from .distributed import DistributedCahnHilliardIMEX
    prob = ts.problem
    traced = u.requires_grad or any(isinstance(v, torch.Tensor) and v.requires_grad
                                    for v in (getattr(prob, "eps", None), getattr(prob, "D", None)))
    periodic = prob.bc_type == ("periodic",) * 3
    if not (periodic and is_stock(prob, CahnHilliard, ("rhs", "fourier_symbol", "spectral_form", "hom_field"))):
        raise NotImplementedError("on an x-slab decomposed grid PseudoSpectralIMEX steps the stock, "
                                  "fully periodic CahnHilliard only")
    if getattr(ts, "_dist", None) is None:
        vg = prob.vg
        try:
            ts._dist = DistributedCahnHilliardIMEX(vg.shape, vg.spacing, ts.dt, eps=float(prob.eps),
                                                   D=float(prob.D), A=float(prob.A), device=vg.device,
                                                   hom_fn=None if prob._default_mu else prob.hom_field)
        except _native.NativeLibraryError as exc:
            raise NotImplementedError(
                f"the multi-GPU spectral step needs power-of-two extents (8..2048) divisible by the number "
                f"of ranks along x and y; grid {vg.shape}: {exc}.  Build the grid with distributed=False "
                "to let every rank solve the whole problem.") from exc
    if traced:
        # backward = the distributed adjoint (slab gradient; dL/dD, dL/deps all-reduced)
        if not prob._default_mu:
            raise NotImplementedError("the hand-written adjoint supports the default mu_hom only")
        D = prob.D if isinstance(prob.D, torch.Tensor) else None
        eps = prob.eps if isinstance(prob.eps, torch.Tensor) else None
        return torch.stack([ts._dist.step_autograd(ch, D, eps) for ch in u.contiguous()], 0)
    return torch.stack([ts._dist.step(ch) for ch in u.contiguous()], 0)


def _defining_class(obj, name):
    for klass in type(obj).__mro__:
        if name in vars(klass):
            return klass
    return None


def _symbol_matches_form(problem):
    return _defining_class(problem, "spectral_form") is _defining_class(problem, "fourier_symbol")


@dataclass
class PseudoSpectralIMEX(TimeStepper):
    """First-order semi-implicit Fourier spectral scheme (Zhu & Chen 1999):
    u+ = u + F^-1[ dt / (1 - dt * symbol) * F[ rhs(u) ] ]."""
    problem: SemiLinearODE
    dt: float
    fft_backend: str = "auto"      # 'auto' | 'cufft' | 'native' | 'native-mixed'

    def __post_init__(self):
        self.problem.verify_fft_bc_config()
        self.pad = self.problem.pad_fft_bc
        self._no_native_mirror = False
        self._plans = {}
        self._prefac = None

    @property
    def order(self) -> int:
        return 1

    # the reference bakes this array in __post_init__ (timesteppers.py:77); here it is only
    # built if somebody asks for it or the problem has no closed-form symbol
    @property
    def _fft_prefac(self):
        if self._prefac is None:
            self._prefac = self.dt / (1 - self.dt * self.problem.fourier_symbol)
        return self._prefac

    def _plan(self, shape, dtype, device):
        key = (tuple(shape), dtype, str(device))
        if key not in self._plans:
            code = {"auto": _native.FFT_AUTO, "cufft": _native.FFT_CUFFT,
                    "native": _native.FFT_NATIVE,
                    "native-mixed": _native.FFT_NATIVE_MIXED}[self.fft_backend]
            self._plans[key] = _native.ImexPlan(shape, dtype, device, code)
        return self._plans[key]

    def step(self, t, u):
        _native.require_cuda(u)
        prob = self.problem
        if getattr(prob.vg, "slab", None) is not None:
            return _distributed_ch_imex(self, u)
        traced = u.requires_grad or any(
            isinstance(v, torch.Tensor) and v.requires_grad
            for v in (getattr(prob, "eps", None), getattr(prob, "D", None)))
        if traced:
            from .autograd import ch_imex_step_autograd
            return ch_imex_step_autograd(self, u)
        u = u.contiguous()
        spacing = prob.vg.spacing
        periodic = prob.bc_type == ("periodic",) * 3
        out = torch.empty_like(u)

        if periodic and is_stock(prob, CahnHilliard, ("rhs", "fourier_symbol", "spectral_form", "hom_field")):
            plan = self._plan(u.shape[1:], u.dtype, u.device)
            hom = prob.hom_field(u)
            for ch in range(u.shape[0]):
                plan.ch_step(u[ch], out[ch], spacing, self.dt, prob.eps, prob.D, prob.A,
                             hom=None if hom is None else hom[ch])
            return out

        return self._spectral_update(t, u, out, periodic, 0, lambda: self._fft_prefac)

    def _spectral_update(self, t, u, out, periodic, kind_flag, stored_weight):
        """u + irfftn(W * rfftn(pad(rhs))) with W evaluated on the fly inside the FFT
        (kind_flag: 0 = IMEX prefactor, FILTER_ETD1 = exponential-Euler weight)."""
        prob = self.problem
        spacing = prob.vg.spacing
        # the closed form stands in for problem.fourier_symbol: only trust it when the class
        # that defines spectral_form also defines the symbol (a subclass overriding one of the
        # two falls back to the stored-array path, which reads problem.fourier_symbol)
        form = prob.spectral_form() if _symbol_matches_form(prob) else None
        rhs = prob.rhs(t, u)
        # non-periodic x with the stock mirror padding (boundary_conditions.py:65-71): the native
        # x pass synthesises the mirror image of every line in shared memory, so the transforms
        # run on the un-extended field - no concatenated 2 Nx array, half the y / z work
        mirror = {"neumann": _native.FILTER_MIRROR_EVEN, "dirichlet": _native.FILTER_MIRROR_ODD}.get(
            prob.bc_type[0], 0)
        if (form is not None and mirror and not getattr(self, "_no_native_mirror", False)
                and _defining_class(prob, "pad_fft_bc") is SemiLinearODE):
            coef, power = form
            plan = self._plan(u.shape[1:], u.dtype, u.device)
            r = rhs.contiguous()
            try:
                for ch in range(u.shape[0]):
                    plan.apply(u[ch], r[ch], out[ch], spacing, self.dt, coef, power | kind_flag | mirror)
                return out
            except _native.NativeLibraryError as exc:
                if getattr(exc, "code", None) != _native.ERR_UNSUPPORTED:
                    raise
                self._no_native_mirror = True      # cuFFT back end / extents beyond the native x pass
        r = self.pad(rhs).contiguous()
        if form is None:
            # user-defined symbol: stored weight array, cuFFT through torch
            upd = torch.fft.irfftn(stored_weight() * torch.fft.rfftn(r, s=r.shape), s=r.shape)
            return u + upd[:, :u.shape[1]]
        coef, power = form
        plan = self._plan(r.shape[1:], r.dtype, r.device)
        if periodic:
            for ch in range(u.shape[0]):
                plan.apply(u[ch], r[ch], out[ch], spacing, self.dt, coef, power | kind_flag)
            return out
        upd = torch.empty_like(r)
        for ch in range(u.shape[0]):
            plan.apply(None, r[ch], upd[ch], spacing, self.dt, coef, power | kind_flag)
        return u + upd[:, :u.shape[1]]


@dataclass
class ExponentialEuler(PseudoSpectralIMEX):
    """First-order exponential Euler (ETD1; Hochbruck, Lubich, Selhofer 1998):
    u+ = u + F^-1[ dt * phi1(dt * symbol) * F[ rhs(u) ] ],  phi1(z) = (exp(z) - 1) / z.
    Mirrors the reference's ExponentialEuler (timesteppers.py:136-202); the weight is
    evaluated inside the native x pass (EVX_FILTER_ETD1) instead of being stored."""

    def phi1(self, z):
        """(exp(z) - 1) / z with the reference's (6,6) Pade branch for |z| < 0.5."""
        N = [1, 1 / 26, 5 / 156, 1 / 858, 1 / 5720, 1 / 205920, 1 / 8648640]
        D = [1, -6 / 13, 5 / 52, -5 / 429, 1 / 1144, -1 / 25740, 1 / 1235520]
        phi = (torch.exp(z) - 1) / z
        small = torch.abs(z) < 0.5
        return torch.where(small, self.phiPade(torch.where(small, z, torch.zeros_like(z)), 6, N, D), phi)

    def phiPade(self, z, Q, Ncoeff, Dcoeff):
        numerator = Ncoeff[Q]
        denominator = Dcoeff[Q]
        for k in range(Q - 1, -1, -1):
            numerator = numerator * z + Ncoeff[k]
            denominator = denominator * z + Dcoeff[k]
        return numerator / denominator

    # the reference bakes this array in __post_init__ (timesteppers.py:153); built on demand
    @property
    def phi_1_k_squared(self):
        # own cache: `_prefac` belongs to the inherited IMEX prefactor
        if getattr(self, "_phi1_cache", None) is None:
            self._phi1_cache = self.phi1(self.dt * self.problem.fourier_symbol)
        return self._phi1_cache

    def step(self, t, u):
        _native.require_cuda(u)
        u = u.contiguous()
        periodic = self.problem.bc_type == ("periodic",) * 3
        out = torch.empty_like(u)
        return self._spectral_update(t, u, out, periodic, _native.FILTER_ETD1,
                                     lambda: self.dt * self.phi_1_k_squared)
