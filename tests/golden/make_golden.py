"""Generate tests/golden/*.npz by running the UNMODIFIED reference (daubners/evoxels,
/root/reference) on CPU through oracle/ref_shim.py.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.

Each case stores the inputs (u0, spacing, parameters) and the reference outputs.
`bc` is stored as its repr() string and re-evaluated by the tests.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import load_reference  # noqa: E402

warnings.simplefilter("ignore")
ref = load_reference()
PD, TS, VG, VF = ref.problem_definition, ref.timesteppers, ref.voxelgrid, ref.voxelfields

P3 = ("periodic",) * 3
N3 = ("neumann",) * 3


def grid(shape, domain, precision="float32"):
    vf = VF.VoxelFields(shape, domain)
    vf.precision = precision
    return vf, VG.VoxelGridTorch(vf.grid_info(), precision, "cpu")


def field(shape, seed, lo, amp):
    return (lo + amp * np.random.default_rng(seed).random(shape)).astype(np.float32)


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    sz = os.path.getsize(os.path.join(HERE, name + ".npz"))
    print(f"{name:40s} {sz/1024:8.1f} KiB")


def npy(t):
    return t.detach().cpu().numpy()


# --------------------------------------------------------------------------------
# Cahn-Hilliard: rhs + IMEX steps
# --------------------------------------------------------------------------------
CH_CASES = [
    # name, shape, domain, bc, eps, D, A, dt, (lo, amp), nsteps
    ("ch_readme16", (16, 16, 16), (16, 16, 16), P3, 3.0, 1.0, 0.25, 0.1, (0.5, 0.1), 10),
    ("ch_odd_aniso", (12, 9, 7), (6.0, 9.0, 3.5), P3, 2.5, 1.3, 0.25, 0.05, (-0.2, 1.4), 3),
    ("ch_line16", (16, 1, 1), (16, 1, 1), P3, 3.0, 1.0, 0.25, 0.5, (-0.1, 1.2), 10),
    ("ch_pow2_small", (16, 32, 64), (16, 32, 64), P3, 3.0, 1.0, 0.25, 0.1, (0.5, 0.1), 2),
    ("ch_neumann_x", (10, 8, 6), (10, 8, 6), ("neumann", "periodic", "periodic"),
     3.0, 1.0, 0.25, 0.1, (0.3, 0.4), 3),
    ("ch_dirichlet_x", (10, 8, 6), (10, 8, 6),
     (("dirichlet", (0.3, 0.7)), "periodic", "periodic"), 3.0, 1.0, 0.25, 0.1, (0.3, 0.4), 3),
    ("ch_mixed_rhs_only", (9, 8, 7), (9, 4, 7),
     (("dirichlet", (0.2, 0.6)), "neumann", "periodic"), 3.0, 0.7, 0.25, 0.1, (-0.2, 1.4), 0),
    ("ch_neumann3_rhs_only", (8, 9, 10), (8, 9, 10), N3, 3.0, 1.0, 0.25, 0.1, (-0.2, 1.4), 0),
    ("ch_zdirichlet_rhs_only", (8, 6, 9), (8, 6, 9),
     ("periodic", "neumann", ("dirichlet", (0.1, 0.9))), 3.0, 1.0, 0.25, 0.1, (0.1, 0.8), 0),
]

for name, shape, dom, bc, eps, D, A, dt, (lo, amp), nsteps in CH_CASES:
    vf, vg = grid(shape, dom)
    u0 = field(shape, 7, lo, amp)
    u = vg.init_scalar_field(u0)
    prob = PD.CahnHilliard(vg, eps=eps, D=D, A=A, bc=bc)
    out = dict(u0=u0, spacing=np.array(vf.spacing), eps=eps, D=D, A=A, dt=dt, bc=repr(bc),
               rhs=npy(prob.rhs(0.0, u))[0], padded=npy(prob.pad_bc(u))[0])
    if nsteps:
        ts = TS.PseudoSpectralIMEX(prob, dt)
        out["prefac"] = npy(ts._fft_prefac)
        v = u
        for i in range(1, nsteps + 1):
            v = ts.step(0.0, v)
            if i in (1, nsteps):
                out[f"step{i}"] = npy(v)[0]
        out["nsteps"] = nsteps
    save(name, **out)

# custom mu_hom (log-type potential, as in docs/notebooks/01-using-solvers.ipynb cell 22)
vf, vg = grid((12, 10, 8), (12, 10, 8))
u0 = field((12, 10, 8), 3, 0.2, 0.6)
u = vg.init_scalar_field(u0)


def mu_log(c, lib=None):
    cc = torch.clip(c, 1e-4, 1 - 1e-4)
    return torch.log(cc / (1 - cc)) + 2.5 * (1 - 2 * c)


prob = PD.CahnHilliard(vg, eps=3.0, D=1.0, mu_hom=mu_log)
ts = TS.PseudoSpectralIMEX(prob, 0.05)
save("ch_custom_mu", u0=u0, spacing=np.array(vf.spacing), eps=3.0, D=1.0, A=0.25, dt=0.05,
     bc=repr(P3), rhs=npy(prob.rhs(0, u))[0], step1=npy(ts.step(0, u))[0])

# float64 case (k arrays stay float32 in the reference - SURVEY 8a row a13)
vf, vg = grid((12, 9, 7), (6.0, 9.0, 3.5), "float64")
u0 = (-0.2 + 1.4 * np.random.default_rng(7).random((12, 9, 7)))
u = vg.init_scalar_field(u0)
prob = PD.CahnHilliard(vg, eps=2.5, D=1.3)
ts = TS.PseudoSpectralIMEX(prob, 0.05)
save("ch_odd_aniso_f64", u0=u0, spacing=np.array(vf.spacing), eps=2.5, D=1.3, A=0.25, dt=0.05,
     bc=repr(P3), rhs=npy(prob.rhs(0, u))[0], step1=npy(ts.step(0, u))[0],
     prefac=npy(ts._fft_prefac))

# --------------------------------------------------------------------------------
# Allen-Cahn: rhs, Euler and RK4 steps
# --------------------------------------------------------------------------------
AC_CASES = [
    # name, shape, domain, bc, kwargs, dt, (lo, amp)
    ("ac_default_neumann", (20, 12, 10), (20, 12, 10), N3, {}, 0.05, (0.0, 1.0)),
    ("ac_curv_force", (12, 9, 7), (6.0, 9.0, 3.5), N3,
     dict(eps=3.0, curvature=0.5, force=1.0, gab=0.8, M=1.5), 0.02, (-0.2, 1.4)),
    ("ac_periodic", (16, 16, 16), (16, 16, 16), P3, dict(curvature=0.3), 0.05, (0.0, 1.0)),
    ("ac_mixed", (9, 8, 7), (9, 8, 7), (("dirichlet", (0.0, 1.0)), "neumann", "periodic"),
     dict(curvature=0.2, force=0.5), 0.05, (0.0, 1.0)),
    ("ac_line16", (16, 1, 1), (16, 1, 1), N3, dict(eps=3.0), 0.5, (0.0, 1.0)),
    ("ac_flat_bulk", (8, 8, 8), (8, 8, 8), N3, dict(curvature=0.5), 0.05, (0.5, 0.0)),
]
for name, shape, dom, bc, kw, dt, (lo, amp) in AC_CASES:
    vf, vg = grid(shape, dom)
    u0 = field(shape, 11, lo, amp)
    if name == "ac_flat_bulk":           # constant blocks exercise the |grad|^2 <= 1e-7 guard
        u0[:4] = 1.0
        u0[4:] = 0.25
    u = vg.init_scalar_field(u0)
    prob = PD.TwoPhaseAllenCahn(vg, bc=bc, **kw)
    full = dict(eps=2.0, gab=1.0, M=1.0, force=0.0, curvature=0.01)
    full.update(kw)
    save(name, u0=u0, spacing=np.array(vf.spacing), dt=dt, bc=repr(bc), **full,
         rhs=npy(prob.rhs(0, u))[0],
         euler1=npy(TS.ForwardEuler(prob, dt).step(0, u))[0],
         rk4_1=npy(TS.RungeKutta4(prob, dt).step(0, u))[0])

# --------------------------------------------------------------------------------
# ghost-layer rules and generic stencils on padded fields
# --------------------------------------------------------------------------------
vf, vg = grid((5, 4, 3), (2.5, 4.0, 1.5))
f0 = field((5, 4, 3), 5, -1.0, 2.0)
f = vg.init_scalar_field(f0)
pads = {}
for i, bc in enumerate([P3, N3, ("neumann", "periodic", "periodic"),
                        (("dirichlet", (1.0, -1.0)), "periodic", "periodic"),
                        (("dirichlet", (10.0, 20.0)), "neumann", "periodic"),
                        ("periodic", ("dirichlet", (0.5, 0.25)), "neumann"),
                        (("dirichlet", (1, 2)), ("dirichlet", (3, 4)), ("dirichlet", (5, 6)))]):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob = PD.ReactionDiffusion(vg, D=1.0, bc=bc)
    pads[f"bc{i}"] = repr(bc)
    pads[f"pad{i}"] = npy(prob.pad_bc(f))[0]
g = vg.bc.pad_bc(f, (("neumann", None),) * 3)
save("ghost_and_stencils", f0=f0, spacing=np.array(vf.spacing), n=7, **pads,
     laplace=npy(vg.laplace(g))[0], normal_laplace=npy(vg.normal_laplace(g))[0],
     gradnorm2=npy(vg.gradient_norm_squared(g))[0],
     x_face=npy(vg.to_x_face(g))[0], y_face=npy(vg.to_y_face(g))[0], z_face=npy(vg.to_z_face(g))[0],
     gx_face=npy(vg.grad_x_face(g))[0], gy_face=npy(vg.grad_y_face(g))[0],
     gz_face=npy(vg.grad_z_face(g))[0],
     k2=npy(vg.rfft_k_squared()), k2_mirror=npy(vg.rfft_k_squared_nonperiodic()))

# --------------------------------------------------------------------------------
# gradients through several CH steps (torch autograd on the reference, float64)
# --------------------------------------------------------------------------------
shape, dom = (12, 10, 8), (12.0, 10.0, 8.0)
vf, vg = grid(shape, dom, "float64")
u0 = 0.1 + 0.8 * np.random.default_rng(21).random(shape)
tgt = 0.5 + 0.1 * np.random.default_rng(22).random(shape)
D = torch.tensor(1.2, dtype=torch.float64, requires_grad=True)
eps = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
u = vg.init_scalar_field(u0).requires_grad_(True)
prob = PD.CahnHilliard(vg, eps=eps, D=D)
ts = TS.PseudoSpectralIMEX(prob, 0.1)
v = u
nst = 5
for _ in range(nst):
    v = ts.step(0, v)
loss = ((v - vg.init_scalar_field(tgt)) ** 2).sum()
gu, gD, geps = torch.autograd.grad(loss, (u, D, eps))
save("ch_grad_f64", u0=u0, target=tgt, spacing=np.array(vf.spacing), D=1.2, eps=2.5, A=0.25,
     dt=0.1, nsteps=nst, loss=float(loss), grad_u0=npy(gu)[0], grad_D=float(gD),
     grad_eps=float(geps), final=npy(v)[0])

# --------------------------------------------------------------------------------
# README configuration (README.md:98-115): 100^3, dt=0.1, 1000 steps, fp32.
# Stored as summary statistics + an every-4th-voxel subsample (full fields would be
# 4 MB each).
# --------------------------------------------------------------------------------
shape = (100, 100, 100)
vf, vg = grid(shape, shape)
u0 = (0.5 + 0.1 * np.random.default_rng(0).random(shape).astype(np.float32))
u = vg.init_scalar_field(u0)
prob = PD.CahnHilliard(vg, eps=3.0, D=1.0)
ts = TS.PseudoSpectralIMEX(prob, 0.1)
out = dict(seed=0, eps=3.0, D=1.0, A=0.25, dt=0.1, spacing=np.array(vf.spacing))
v = u
for i in range(1, 1001):
    v = ts.step(0, v)
    if i in (1, 10, 100, 1000):
        a = npy(v)[0]
        out[f"sub{i}"] = a[::4, ::4, ::4].copy()
        out[f"mean{i}"] = float(a.astype(np.float64).mean())
        out[f"l2_{i}"] = float(np.linalg.norm(a.astype(np.float64)))
        out[f"min{i}"], out[f"max{i}"] = float(a.min()), float(a.max())
out["mean0"] = float(u0.astype(np.float64).mean())
save("ch_readme100_1000steps", **out)
print("done")
