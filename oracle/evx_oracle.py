"""CPU oracle for the evoxels hot path (Cahn-Hilliard IMEX step, Allen-Cahn rhs/Euler/RK4).

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the
package `evoxels_b200` never does (tests/test_host_logic.py::test_product_never_imports_the_oracle
checks it).

This is a restatement, in plain torch-on-CPU tensor arithmetic, of what the reference
computes (cites are into /root/reference/evoxels/):

  ghost_pad            boundary_conditions.py:9-59  (+ voxelgrid.py:194-195 circular pad)
  laplace7             fd_stencils.py:62-75
  normal_laplace19     fd_stencils.py:44-60, 77-103
  ch_rhs               problem_definition.py:328-371 (default mu_hom :305)
  ac_rhs               problem_definition.py:421-447 (default potential :391)
  k_squared            voxelgrid.py:84-90, 110-124
  ch_symbol/ac_symbol  problem_definition.py:303, 389
  imex_prefactor       timesteppers.py:77
  imex_step            timesteppers.py:85-89 (+ boundary_conditions.py:61-71 mirror pads)
  euler_step, rk4_step timesteppers.py:42-43, 56-61

Parity is PINNED: tests/test_oracle_golden.py checks every function against fixtures in
tests/golden/*.npz that were produced by running the unmodified reference in the build
container (tests/golden/make_golden.py, through oracle/ref_shim.py), and
tests/test_oracle_vs_reference.py re-runs the reference live whenever /root/reference
is present.  The third-party arithmetic underneath (torch.fft.rfftn/irfftn, "backward"
norm; torch>=2.1 un-pinned in the reference's pyproject.toml:41-43, 2.11.0 installed)
is the same library call here as in the reference, so the FFT boundary is pinned by
those fixtures too.

All tensors are created on the CPU explicitly: the product mirrors the reference's global
`torch.set_default_device(cuda)` side effect (voxelgrid.py:178), which must not leak in here.

The operation structure (ghost-padded copies + strided slices, ~60 elementwise passes
per Cahn-Hilliard step) deliberately follows the reference so that timing this module
on the host cores is a fair stand-in for the reference's torch CPU path.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import torch

PERIODIC, NEUMANN, DIRICHLET = "periodic", "neumann", "dirichlet"
_FULLY_PERIODIC = ((PERIODIC, None),) * 3


# ----------------------------------------------------------------------------------
# boundary handling
# ----------------------------------------------------------------------------------
def normalize_bc(bc) -> tuple:
    """`bc` in any spelling the reference accepts -> ((kind, values),)*3.
    Follows problem_definition.py:60-102 (no warnings here; validation is tested on the
    product's own implementation)."""
    if bc is None or bc == "fully_periodic":
        return _FULLY_PERIODIC
    out = []
    for entry in bc:
        if isinstance(entry, str):
            out.append((entry, None))
        else:
            kind, vals = entry
            out.append((kind, None if vals is None else tuple(vals)))
    return tuple(out)


def _edge(axis: int, idx: int):
    """Index tuple selecting plane `idx` of spatial axis `axis` of a [C,X,Y,Z] tensor."""
    sel = [slice(None)] * 4
    sel[axis + 1] = idx
    return tuple(sel)


def ghost_pad(f: torch.Tensor, bc=_FULLY_PERIODIC) -> torch.Tensor:
    """One ghost layer around a [C,Nx,Ny,Nz] field.

    Start from the circular pad, then overwrite ghost planes axis by axis in the order
    x, y, z over the *whole* padded plane (so edge/corner ghosts inherit the rule of the
    later axis applied to the earlier axis' ghosts) - boundary_conditions.py:33-59.
    Neumann: ghost = first inner plane.  Dirichlet: ghost = 2*value - inner."""
    bc = normalize_bc(bc)
    g = torch.nn.functional.pad(f, (1, 1, 1, 1, 1, 1), mode="circular")
    for axis, (kind, vals) in enumerate(bc):
        if kind == PERIODIC:
            continue
        lo_g, lo_i, hi_g, hi_i = _edge(axis, 0), _edge(axis, 1), _edge(axis, -1), _edge(axis, -2)
        if kind == NEUMANN:
            g[lo_g] = g[lo_i]
            g[hi_g] = g[hi_i]
        elif kind == DIRICHLET:
            g[lo_g] = 2.0 * vals[0] - g[lo_i]
            g[hi_g] = 2.0 * vals[1] - g[hi_i]
        else:
            raise ValueError(f"Unsupported BC type: {kind}")
    return g


def _win(g: torch.Tensor, dx: int, dy: int, dz: int) -> torch.Tensor:
    """Interior-sized window of a ghost-padded field shifted by (dx,dy,dz) in {-1,0,1}."""
    nx, ny, nz = g.shape[1] - 2, g.shape[2] - 2, g.shape[3] - 2
    return g[:, 1 + dx:1 + dx + nx, 1 + dy:1 + dy + ny, 1 + dz:1 + dz + nz]


# ----------------------------------------------------------------------------------
# stencils
# ----------------------------------------------------------------------------------
def _inv_h(spacing, like: torch.Tensor):
    h = torch.tensor([float(s) for s in spacing], dtype=like.dtype, device=like.device)
    return 1.0 / h, 1.0 / h ** 2


def laplace7(g: torch.Tensor, spacing) -> torch.Tensor:
    """7-point Laplacian of a ghost-padded field, interior result (fd_stencils.py:62-75)."""
    _, ih2 = _inv_h(spacing, g)
    return ((_win(g, 1, 0, 0) + _win(g, -1, 0, 0)) * ih2[0]
            + (_win(g, 0, 1, 0) + _win(g, 0, -1, 0)) * ih2[1]
            + (_win(g, 0, 0, 1) + _win(g, 0, 0, -1)) * ih2[2]
            - 2 * _win(g, 0, 0, 0) * torch.sum(ih2))


def _grad_center(g, spacing, axis):
    ih, _ = _inv_h(spacing, g)
    d = [0, 0, 0]
    d[axis] = 1
    return 0.5 * (_win(g, *d) - _win(g, *[-k for k in d])) * ih[axis]


def normal_laplace19(g: torch.Tensor, spacing) -> torch.Tensor:
    """d^2 f / dn^2 with n = grad f/|grad f| on a ghost-padded field
    (fd_stencils.py:77-103): second differences weighted by squared centred gradients,
    plus mixed differences on the 12 edge neighbours, divided by |grad f|^2 with the
    `<= 1e-7 -> 1` guard."""
    ih, ih2 = _inv_h(spacing, g)
    gx, gy, gz = (_grad_center(g, spacing, a) for a in range(3))
    c = _win(g, 0, 0, 0)

    def second(axis):
        d = [0, 0, 0]
        d[axis] = 1
        return (_win(g, *d) - 2 * c + _win(g, *[-k for k in d])) * ih2[axis]

    def mixed(a, b):
        pp, mm, mp, pm = [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]
        pp[a], pp[b] = 1, 1
        mm[a], mm[b] = -1, -1
        mp[a], mp[b] = -1, 1
        pm[a], pm[b] = 1, -1
        return (_win(g, *pp) + _win(g, *mm) - _win(g, *mp) - _win(g, *pm)) * ih[a] * ih[b]

    num = (gx ** 2 * second(0) + gy ** 2 * second(1) + gz ** 2 * second(2)
           + 0.5 * gx * gy * mixed(0, 1)
           + 0.5 * gx * gz * mixed(0, 2)
           + 0.5 * gy * gz * mixed(1, 2))
    norm2 = gx ** 2 + gy ** 2 + gz ** 2
    norm2 = torch.where(norm2 <= 1e-7, torch.ones_like(norm2), norm2)
    return num / norm2


def double_well_prime(c, eps):
    """Default homogeneous chemical potential / phase-field potential derivative:
    18/eps * c (1-c) (1-2c)  (problem_definition.py:305, 391)."""
    return 18 / eps * c * (1 - c) * (1 - 2 * c)


def ch_rhs(u: torch.Tensor, spacing, eps=3.0, D=1.0, bc=_FULLY_PERIODIC,
           mu_hom: Optional[Callable] = None) -> torch.Tensor:
    """Cahn-Hilliard right-hand side  D div( c(1-c) grad mu ),  mu = mu_hom(c) - 2 eps lap c
    (problem_definition.py:350-371).  `u` is [C,Nx,Ny,Nz]; note that BOTH c and mu are
    ghost-padded with the same `bc` (so Dirichlet values apply to mu as well)."""
    ih, _ = _inv_h(spacing, u)
    c = torch.clip(u, 0, 1)
    cg = ghost_pad(c, bc)
    hom = double_well_prime(c, eps) if mu_hom is None else mu_hom(c)
    mu = hom - 2 * eps * laplace7(cg, spacing)
    mg = ghost_pad(mu, bc)
    nx, ny, nz = u.shape[1:]

    def face_flux_divergence(axis):
        # faces between padded cells k and k+1 along `axis`, interior cross-section
        sl = [slice(None), slice(1, 1 + nx), slice(1, 1 + ny), slice(1, 1 + nz)]
        lo, hi = list(sl), list(sl)
        n = (nx, ny, nz)[axis]
        lo[axis + 1] = slice(0, n + 1)
        hi[axis + 1] = slice(1, n + 2)
        lo, hi = tuple(lo), tuple(hi)
        cf = 0.5 * (cg[hi] + cg[lo])
        flux = cf * (1 - cf) * ((mg[hi] - mg[lo]) * ih[axis])
        a, b = [slice(None)] * 4, [slice(None)] * 4
        a[axis + 1] = slice(1, None)
        b[axis + 1] = slice(None, -1)
        return (flux[tuple(a)] - flux[tuple(b)]) * ih[axis]

    div = face_flux_divergence(0)
    div = div + face_flux_divergence(1)
    div = div + face_flux_divergence(2)
    return D * div


_NEUMANN3 = ((NEUMANN, None),) * 3


def ac_rhs(u: torch.Tensor, spacing, eps=2.0, gab=1.0, M=1.0, force=0.0, curvature=0.01,
           bc=_NEUMANN3, potential: Optional[Callable] = None) -> torch.Tensor:
    """Two-phase Allen-Cahn right-hand side (problem_definition.py:440-447)."""
    phi = torch.clip(u, 0, 1)
    pot = double_well_prime(phi, eps) if potential is None else potential(phi)
    g = ghost_pad(phi, bc)
    lap = curvature * laplace7(g, spacing)
    nlap = (1 - curvature) * normal_laplace19(g, spacing)
    df = gab * (lap + nlap - pot / 2 / eps) + 3 / eps * phi * (1 - phi) * force
    return M * df


# ----------------------------------------------------------------------------------
# spectral side
# ----------------------------------------------------------------------------------
def k_squared(shape: Sequence[int], spacing, mirrored_x: bool = False) -> torch.Tensor:
    """|k|^2 on the rfftn half-spectrum grid [Nx(or 2Nx), Ny, Nz//2+1].

    ALWAYS float32, like the reference (voxelgrid.py:84-90,110-124 never pass a dtype to
    fftfreq).  `mirrored_x` = the non-periodic-x variant that uses
    fftfreq(2*Nx) (cell_center convention, voxelgrid.py:116-118)."""
    nx, ny, nz = shape
    hx, hy, hz = (float(s) for s in spacing)
    kx = 2 * math.pi * torch.fft.fftfreq(2 * nx if mirrored_x else nx, hx, device="cpu")
    ky = 2 * math.pi * torch.fft.fftfreq(ny, hy, device="cpu")
    kz = 2 * math.pi * torch.fft.rfftfreq(nz, hz, device="cpu")
    KX, KY, KZ = torch.meshgrid(kx, ky, kz, indexing="ij")
    return KX ** 2 + KY ** 2 + KZ ** 2


def ch_symbol(shape, spacing, eps=3.0, D=1.0, A=0.25, mirrored_x=False):
    """-2 eps D A |k|^4  (problem_definition.py:303)."""
    return -2 * eps * D * A * k_squared(shape, spacing, mirrored_x) ** 2


def ac_symbol(shape, spacing, gab=1.0, M=1.0, mirrored_x=False):
    """-M gab |k|^2  (problem_definition.py:389)."""
    return -M * gab * k_squared(shape, spacing, mirrored_x)


def imex_prefactor(symbol: torch.Tensor, dt: float) -> torch.Tensor:
    """dt / (1 - dt * symbol)  (timesteppers.py:77)."""
    return dt / (1 - dt * symbol)


def fft_mirror_pad(r: torch.Tensor, x_kind: str) -> torch.Tensor:
    """Extension along x before the FFT (boundary_conditions.py:61-71, cell_center)."""
    if x_kind == PERIODIC:
        return r
    if x_kind == NEUMANN:
        return torch.cat((r, torch.flip(r, [1])), 1)
    if x_kind == DIRICHLET:
        return torch.cat((r, -torch.flip(r, [1])), 1)
    raise ValueError(x_kind)


def imex_step(u: torch.Tensor, rhs: torch.Tensor, prefac: torch.Tensor,
              x_kind: str = PERIODIC) -> torch.Tensor:
    """u + irfftn(prefac * rfftn(pad(rhs)))[:, :Nx]  (timesteppers.py:85-89).
    The transforms run over all four dims with s = padded shape, as in the reference
    (the size-C leading dim is transformed too; a no-op for C == 1)."""
    r = fft_mirror_pad(rhs, x_kind)
    spec = prefac * torch.fft.rfftn(r, s=r.shape)
    upd = torch.fft.irfftn(spec, s=r.shape)[:, :u.shape[1]]
    return u + upd


def ch_imex_step(u, spacing, dt, eps=3.0, D=1.0, A=0.25, bc=_FULLY_PERIODIC,
                 mu_hom=None, prefac=None):
    """One full reference step for CahnHilliard + PseudoSpectralIMEX."""
    bc = normalize_bc(bc)
    x_kind = bc[0][0]
    if prefac is None:
        prefac = imex_prefactor(
            ch_symbol(u.shape[1:], spacing, eps, D, A, mirrored_x=x_kind != PERIODIC), dt)
    return imex_step(u, ch_rhs(u, spacing, eps, D, bc, mu_hom), prefac, x_kind)


# ----------------------------------------------------------------------------------
# two-species reaction-diffusion (Gray-Scott) + exponential Euler
# ----------------------------------------------------------------------------------
def crd_rhs(u: torch.Tensor, spacing, D_A=1.0, D_B=0.5, feed=0.055, kill=0.117,
            interaction: Optional[Callable] = None) -> torch.Tensor:
    """CoupledReactionDiffusion.rhs (problem_definition.py:614-633): channels are the two
    species, always fully periodic (the class has no `bc` field)."""
    inter = u[0] * u[1] ** 2 if interaction is None else interaction(u)
    lap = laplace7(ghost_pad(u), spacing)
    dA = D_A * lap[0] - inter + feed * (1 - u[0])
    dB = D_B * lap[1] + inter - kill * u[1]
    return torch.stack((dA, dB), 0)


def crd_symbol(shape, spacing, D_A=1.0, D_B=0.5):
    """-max(D_A, D_B) |k|^2  (problem_definition.py:589)."""
    return -max(D_A, D_B) * k_squared(shape, spacing)


def rd_symbol(shape, spacing, D, A=0.25, mirrored_x=False):
    """-D A |k|^2  (problem_definition.py:206)."""
    return -D * A * k_squared(shape, spacing, mirrored_x)


_PADE_N = [1, 1 / 26, 5 / 156, 1 / 858, 1 / 5720, 1 / 205920, 1 / 8648640]
_PADE_D = [1, -6 / 13, 5 / 52, -5 / 429, 1 / 1144, -1 / 25740, 1 / 1235520]


def phi1(z: torch.Tensor) -> torch.Tensor:
    """varphi_1(z) = (exp(z) - 1) / z with the (6,6) Pade branch for |z| < 0.5
    (timesteppers.py:155-192)."""
    phi = (torch.exp(z) - 1) / z
    small = torch.abs(z) < 0.5
    zs = z[small]
    num = torch.full_like(zs, _PADE_N[6])
    den = torch.full_like(zs, _PADE_D[6])
    for k in range(5, -1, -1):
        num = num * zs + _PADE_N[k]
        den = den * zs + _PADE_D[k]
    phi = phi.clone()
    phi[small] = num / den
    return phi


def etd1_step(u: torch.Tensor, rhs: torch.Tensor, symbol: torch.Tensor, dt: float,
              x_kind: str = PERIODIC) -> torch.Tensor:
    """ExponentialEuler.step (timesteppers.py:198-202):
    u + irfftn(dt * phi1(dt * symbol) * rfftn(pad(rhs)))[:, :Nx]."""
    r = fft_mirror_pad(rhs, x_kind)
    spec = dt * phi1(dt * symbol) * torch.fft.rfftn(r, s=r.shape)
    return u + torch.fft.irfftn(spec, s=r.shape)[:, :u.shape[1]]


def euler_step(u, rhs_fn: Callable, dt: float):
    """u + dt * rhs(u)  (timesteppers.py:42-43)."""
    return u + dt * rhs_fn(u)


def rk4_step(u, rhs_fn: Callable, dt: float):
    """Classical RK4 (timesteppers.py:56-61)."""
    k1 = rhs_fn(u)
    k2 = rhs_fn(u + 0.5 * dt * k1)
    k3 = rhs_fn(u + 0.5 * dt * k2)
    k4 = rhs_fn(u + dt * k3)
    return u + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


# ----------------------------------------------------------------------------------
# convenience drivers used by tests / bench
# ----------------------------------------------------------------------------------
class CHOracle:
    """Holds the baked prefactor like PseudoSpectralIMEX.__post_init__ does."""

    def __init__(self, shape, spacing, dt, eps=3.0, D=1.0, A=0.25, bc=_FULLY_PERIODIC,
                 mu_hom=None):
        self.shape, self.spacing, self.dt = tuple(shape), tuple(spacing), dt
        self.eps, self.D, self.A, self.mu_hom = eps, D, A, mu_hom
        self.bc = normalize_bc(bc)
        self.x_kind = self.bc[0][0]
        self.prefac = imex_prefactor(
            ch_symbol(shape, spacing, eps, D, A, mirrored_x=self.x_kind != PERIODIC), dt)

    def rhs(self, u):
        return ch_rhs(u, self.spacing, self.eps, self.D, self.bc, self.mu_hom)

    def step(self, u):
        return imex_step(u, self.rhs(u), self.prefac, self.x_kind)


class ACOracle:
    def __init__(self, shape, spacing, dt, eps=2.0, gab=1.0, M=1.0, force=0.0,
                 curvature=0.01, bc=_NEUMANN3, potential=None, scheme="euler"):
        self.shape, self.spacing, self.dt = tuple(shape), tuple(spacing), dt
        self.kw = dict(eps=eps, gab=gab, M=M, force=force, curvature=curvature,
                       bc=normalize_bc(bc), potential=potential)
        self.scheme = scheme

    def rhs(self, u):
        return ac_rhs(u, self.spacing, **self.kw)

    def step(self, u):
        if self.scheme == "euler":
            return euler_step(u, self.rhs, self.dt)
        return rk4_step(u, self.rhs, self.dt)


def noise_field(shape, seed=0, lo=0.5, amp=0.1, dtype=torch.float32) -> torch.Tensor:
    """Synthetic initial condition of BASELINE.md section 3: lo + amp * U[0,1), numpy
    default_rng(seed), returned as [1,Nx,Ny,Nz]."""
    import numpy as np
    a = np.random.default_rng(seed).random(tuple(shape)).astype(np.float32)
    return (lo + amp * torch.from_numpy(a).to(dtype)).unsqueeze(0)
