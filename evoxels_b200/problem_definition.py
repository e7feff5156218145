"""PDE right-hand sides with the reference's `problem_definition` interface.

Interface parity with evoxels/problem_definition.py: `ODE` (:13-120), `SemiLinearODE`
(:123-180), `CahnHilliard` (:290-371), `TwoPhaseAllenCahn` (:374-447) and
`ReactionDiffusion` (:191-229) keep their constructor signatures, attributes (`bc`,
`bc_type`, `order`, `fourier_symbol`, `pad_bc`, `pad_fft_bc`, ...) and error behaviour, so
`solver._init_stepper` (solvers.py:52-72) can build them unchanged.

What differs is *how* `rhs` is evaluated: one fused CUDA kernel per call
(`evx_ch_rhs_*`, `evx_ac_stage_*`) instead of ~60 / ~125 tensor operations on ghost-padded
copies.  There is no CPU path - `rhs` raises for non-CUDA tensors.

`CoupledReactionDiffusion` (:576-633) is fused the same way
(`evx_rd2_rhs_*`).  The remaining problem classes of the reference (ReactionDiffusionSBM,
MultiPhaseAllenCahn) are different physics and out of scope; users can still write them
against `VoxelGridTorch`'s operator API.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any, Callable

import torch

from . import _native
from .voxelgrid import VoxelGrid

_KINDS = ("periodic", "dirichlet", "neumann")
_PERIODIC3 = (("periodic", None),) * 3


def normalize_bc(bc, convention="cell_center"):
    """Bring a user `bc` into the canonical ((kind, values), (kind, values), (kind, values)).

    Accepts what the reference accepts (problem_definition.py:60-102): None or
    'fully_periodic'; per axis either 'periodic' / 'neumann' or a pair (kind, values) where
    Dirichlet needs two values.  Raises ValueError with the reference's wording otherwise
    and warns about the order loss of Dirichlet conditions on a cell-centred grid.
    """
    if bc is None or bc == "fully_periodic":
        return _PERIODIC3
    if len(bc) != 3:
        raise ValueError("bc must contain exactly three axis entries ordered as (x, y, z).")
    out = []
    for entry in bc:
        if isinstance(entry, str):
            if entry == "dirichlet":
                raise ValueError("Dirichlet BCs require explicit values ('dirichlet', (left, right)).")
            if entry not in ("periodic", "neumann"):
                raise ValueError(f"Unsupported BC type: {entry}")
            out.append((entry, None))
            continue
        if len(entry) != 2:
            raise ValueError("Each axis boundary specification must be either a string or "
                             "a tuple like ('dirichlet', (left, right)).")
        kind, values = entry
        if kind not in _KINDS:
            raise ValueError(f"Unsupported BC type: {kind}")
        if kind == "dirichlet":
            if values is None or len(values) != 2:
                raise ValueError("Dirichlet BCs require two boundary values.")
            if convention == "cell_center":
                warnings.warn("Applying Dirichlet BCs on a cell_center grid "
                              "reduces the spatial order of convergence to 0.5!")
            out.append((kind, tuple(values)))
        else:
            if values is not None:
                raise ValueError(f"{kind} BCs do not accept boundary values.")
            out.append((kind, None))
    return tuple(out)


def _call_closure(fn, *args):
    """User closures may or may not take the trailing `lib` argument (reference
    problem_definition.py:315-320)."""
    try:
        return fn(*args)
    except TypeError:
        return fn(*args[:-1])


class ODE(ABC):
    # True when rhs(t, u) does not depend on t.  Only then may the solver capture a few
    # steps in a CUDA graph (jit=True) - a graph freezes the time argument.  User problem
    # classes are non-autonomous unless they say otherwise.
    autonomous = False

    @property
    @abstractmethod
    def order(self) -> int:
        """Spatial order of convergence of the discrete right-hand side."""

    @abstractmethod
    def rhs_analytic(self, t, u):
        """Sympy expression of the continuous right-hand side."""

    @abstractmethod
    def rhs(self, t, u):
        """Discrete right-hand side, same shape/type as `u`."""

    def initialize_boundary_conditions(self):
        self.bc = normalize_bc(getattr(self, "bc", None), self.vg.convention)

    @property
    def bc_type(self):
        return tuple(kind for kind, _ in self.bc)

    def pad_bc(self, u):
        """Ghost-pad `u` with this problem's boundary conditions (one kernel launch)."""
        return self.vg.bc.pad_bc(u, self.bc)


class SemiLinearODE(ODE):
    @property
    @abstractmethod
    def fourier_symbol(self):
        """Symbol of the stiff linear operator treated implicitly by FFT steppers."""

    def spectral_form(self):
        """(coef, power) such that fourier_symbol == -coef * |k|^(2*power), or None.
        When available, FFT steppers recompute the symbol on the fly inside the filter
        kernel instead of reading a stored array."""
        return None

    def verify_fft_bc_config(self):
        kinds = self.bc_type
        off = [ax for ax, kind in zip("xyz", kinds) if kind != "periodic"]
        if len(off) > 1:
            raise ValueError("FFT-based timesteppers currently support at most one "
                             f"non-periodic axis, got {kinds}.")
        if off and off[0] != "x":
            raise NotImplementedError(
                "FFT-based timesteppers currently only implement the single non-periodic axis "
                f"case for x; got {kinds}. Axis permutation is not implemented yet.")
        pads = {"periodic": self.vg.bc.pad_fft_periodic,
                "dirichlet": self.vg.bc.pad_fft_dirichlet_periodic,
                "neumann": self.vg.bc.pad_fft_zero_flux_periodic}
        if kinds[0] not in pads:
            raise ValueError("FFT-based timesteppers only support periodic, dirichlet, or "
                             f"neumann boundary conditions in x, got {kinds[0]}.")
        self._pad_fft_bc = pads[kinds[0]]

    def k_squared(self):
        if self.bc_type[0] in ("dirichlet", "neumann"):
            return self.vg.rfft_k_squared_nonperiodic()
        return self.vg.rfft_k_squared()

    def pad_fft_bc(self, u):
        return self._pad_fft_bc(u)


def is_stock(problem, cls, names=("rhs", "fourier_symbol", "spectral_form")):
    """True if `problem` is a `cls` whose listed members are the ones defined by `cls` itself.
    The fused kernels implement exactly those; a subclass that overrides one of them (say an
    rhs with a source term) must go through its own methods, as in the reference, where every
    stepper calls problem.rhs / problem.fourier_symbol."""
    if not isinstance(problem, cls):
        return False
    return all(getattr(type(problem), n, None) is getattr(cls, n, None) for n in names)


def _is_traced(*values):
    return any(isinstance(v, torch.Tensor) and v.requires_grad for v in values)


@dataclass
class CahnHilliard(SemiLinearODE):
    """dc/dt = div( D c(1-c) grad mu ),  mu = mu_hom(c) - 2 eps lap(c)."""
    autonomous = True
    vg: VoxelGrid
    eps: float = 3.0
    D: float = 1.0
    mu_hom: Callable | None = None
    A: float = 0.25
    bc: tuple = ("periodic", "periodic", "periodic")
    _fourier_symbol: Any = field(init=False, repr=False, default=None)

    def __post_init__(self):
        self.initialize_boundary_conditions()
        self._default_mu = self.mu_hom is None
        if self._default_mu:
            self.mu_hom = lambda c, lib=None: 18 / self.eps * c * (1 - c) * (1 - 2 * c)

    @property
    def order(self):
        return 2

    @property
    def fourier_symbol(self):
        # materialised only on request; the stepper uses spectral_form() instead
        if self._fourier_symbol is None:
            self._fourier_symbol = -2 * self.eps * self.D * self.A * self.k_squared() ** 2
        return self._fourier_symbol

    def spectral_form(self):
        return 2.0 * float(self.eps) * float(self.D) * float(self.A), 2

    def _eval_mu(self, c, lib):
        return _call_closure(self.mu_hom, c, lib)

    def rhs_analytic(self, t, c):
        import sympy as sp
        import sympy.vector as spv
        mu = self._eval_mu(c, sp) - 2 * self.eps * spv.laplacian(c)
        return spv.divergence(self.D * c * (1 - c) * spv.gradient(mu))

    def hom_field(self, c):
        """mu_hom evaluated on clip(c) for user-supplied potentials (None for the default,
        which the kernel evaluates itself)."""
        if self._default_mu:
            return None
        return self._eval_mu(torch.clip(c, 0, 1), torch).to(c.dtype).contiguous()

    def rhs(self, t, c):
        _native.require_cuda(c)
        if _is_traced(c, self.eps, self.D):
            from .autograd import ch_rhs_autograd
            return ch_rhs_autograd(self, c)
        c = c.contiguous()
        hom = self.hom_field(c)
        out = torch.empty_like(c)
        for ch in range(c.shape[0]):
            _native.ch_rhs(c[ch], out[ch], self.vg.spacing, float(self.eps), float(self.D),
                           self.bc, hom=None if hom is None else hom[ch])
        return out


@dataclass
class TwoPhaseAllenCahn(SemiLinearODE):
    """dphi/dt = M [ gab ( curv lap(phi) + (1-curv) d2phi/dn2 - g(phi)/(2 eps) )
                     + 3/eps phi (1-phi) force ]."""
    autonomous = True
    vg: VoxelGrid
    eps: float = 2.0
    gab: float = 1.0
    M: float = 1.0
    force: float = 0.0
    curvature: float = 0.01
    potential: Callable | None = None
    bc: tuple = ("neumann", "neumann", "neumann")
    _fourier_symbol: Any = field(init=False, repr=False, default=None)

    def __post_init__(self):
        self.initialize_boundary_conditions()
        self._default_potential = self.potential is None
        if self._default_potential:
            self.potential = lambda u, lib=None: 18 / self.eps * u * (1 - u) * (1 - 2 * u)

    @property
    def order(self):
        return 2

    @property
    def fourier_symbol(self):
        if self._fourier_symbol is None:
            self._fourier_symbol = -self.M * self.gab * self.k_squared()
        return self._fourier_symbol

    def spectral_form(self):
        return float(self.M) * float(self.gab), 1

    def _eval_potential(self, phi, lib):
        return _call_closure(self.potential, phi, lib)

    def rhs_analytic(self, t, phi):
        import sympy as sp
        import sympy.vector as spv
        grad = spv.gradient(phi)
        norm = sp.sqrt(grad.dot(grad))
        curv = norm * spv.divergence(grad / norm)
        n_laplace = spv.laplacian(phi) - (1 - self.curvature) * curv
        df = self.gab * (n_laplace - self._eval_potential(phi, sp) / (2 * self.eps)) \
            + 3 / self.eps * phi * (1 - phi) * self.force
        return self.M * df

    def fused_stage(self, y, *, base=None, alpha=0.0, y_out=None, acc_in=None, beta=0.0,
                    acc_out=None, k_out=None):
        """k = rhs(y); optionally y_out = base + alpha k, acc_out = acc_in + beta k, k_out = k.
        One kernel launch per channel (see evx_ac_stage_* in include/evoxels_b200.h)."""
        _native.require_cuda(y)
        y = y.contiguous()
        pot = None
        if not self._default_potential:
            pot = self._eval_potential(torch.clip(y, 0, 1), torch).to(y.dtype).contiguous()
        for ch in range(y.shape[0]):
            pick = lambda t: None if t is None else t[ch]  # noqa: E731
            _native.ac_stage(y[ch], self.vg.spacing, float(self.eps), float(self.gab),
                             float(self.M), float(self.force), float(self.curvature), self.bc,
                             pot=pick(pot), k_out=pick(k_out), base=pick(base),
                             y_out=pick(y_out), alpha=alpha, acc_in=pick(acc_in),
                             acc_out=pick(acc_out), beta=beta)

    def rhs(self, t, phi):
        out = torch.empty_like(phi, memory_format=torch.contiguous_format)
        self.fused_stage(phi, k_out=out)
        return out


@dataclass
class ReactionDiffusion(SemiLinearODE):
    """du/dt = D lap(u) + f(t, u) - the reference's test vehicle for the Laplacian and the
    ghost rules (problem_definition.py:191-229, tests/test_laplace.py)."""
    vg: VoxelGrid
    D: float
    f: Callable | None = None
    A: float = 0.25
    bc: tuple = ("periodic", "periodic", "periodic")
    _fourier_symbol: Any = field(init=False, repr=False, default=None)

    def __post_init__(self):
        self.autonomous = self.f is None      # a user source term f(t, u, lib) may depend on t
        if self.f is None:
            self.f = lambda c=None, t=None, lib=None: 0
        self.initialize_boundary_conditions()

    @property
    def order(self):
        return 2

    @property
    def fourier_symbol(self):
        if self._fourier_symbol is None:
            self._fourier_symbol = -self.D * self.A * self.k_squared()
        return self._fourier_symbol

    def spectral_form(self):
        return float(self.D) * float(self.A), 1

    def _eval_f(self, t, c, lib):
        return _call_closure(self.f, t, c, lib)

    def rhs_analytic(self, t, u):
        import sympy as sp
        import sympy.vector as spv
        return self.D * spv.laplacian(u) + self._eval_f(t, u, sp)

    def rhs(self, t, u):
        return self.D * self.vg.laplace(self.pad_bc(u)) + self._eval_f(t, u, self.vg.lib)


@dataclass
class CoupledReactionDiffusion(SemiLinearODE):
    """Two-species reaction-diffusion system of Gray-Scott type; the species are the two
    batch channels of `u` (reference problem_definition.py:576-633).  Always fully periodic,
    like the reference (the class has no `bc` field).  With the default interaction
    u0*u1**2 the whole right-hand side is ONE fused kernel (`evx_rd2_rhs_*`); a user
    `interaction` closure is evaluated with torch and handed to the same kernel as a field."""
    autonomous = True
    vg: VoxelGrid
    D_A: float = 1.0
    D_B: float = 0.5
    feed: float = 0.055
    kill: float = 0.117
    interaction: Callable | None = None
    _fourier_symbol: Any = field(init=False, repr=False, default=None)

    def __post_init__(self):
        self.initialize_boundary_conditions()
        self._default_interaction = self.interaction is None
        if self.interaction is None:
            self.interaction = lambda u, lib=None: u[0] * u[1] ** 2

    @property
    def order(self):
        return 2

    @property
    def fourier_symbol(self):
        if self._fourier_symbol is None:
            self._fourier_symbol = -max(self.D_A, self.D_B) * self.k_squared()
        return self._fourier_symbol

    def spectral_form(self):
        return float(max(self.D_A, self.D_B)), 1

    def _eval_interaction(self, u, lib):
        return _call_closure(self.interaction, u, lib)

    def rhs_analytic(self, t, u):
        import sympy as sp
        import sympy.vector as spv
        interaction = self._eval_interaction(u, sp)
        dc_A = self.D_A * spv.laplacian(u[0]) - interaction + self.feed * (1 - u[0])
        dc_B = self.D_B * spv.laplacian(u[1]) + interaction - self.kill * u[1]
        return (dc_A, dc_B)

    def rhs(self, t, u):
        _native.require_cuda(u)
        if u.dim() != 4 or u.shape[0] != 2:
            raise ValueError("CoupledReactionDiffusion expects a state of shape [2, Nx, Ny, Nz]")
        u = u.contiguous()
        inter = None
        if not self._default_interaction:
            inter = self._eval_interaction(u, self.vg.lib).to(u.dtype).contiguous()
        return _native.rd2_rhs(u, self.vg.spacing, self.D_A, self.D_B, self.feed, self.kill, inter)
