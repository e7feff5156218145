"""Parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures.
Needs a GPU: run with `pytest -m gpu` on the B200 box.

Tolerances (north_star): fp32 state rel-L2 <= 1e-5 per step, <= 1e-4 after 1000 steps,
mass conserved to fp32 rounding.  Stricter self-checks where the noise floor allows:
rhs <= 5e-6, update <= 2e-5."""
import ast
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import evx_oracle as O

pytestmark = pytest.mark.gpu

import evoxels_b200 as evo  # noqa: E402
from evoxels_b200 import _native  # noqa: E402
from evoxels_b200.problem_definition import (CahnHilliard, CoupledReactionDiffusion,  # noqa: E402
                                             ReactionDiffusion, TwoPhaseAllenCahn)
from evoxels_b200.solvers import TimeDependentSolver  # noqa: E402
from evoxels_b200.timesteppers import (ExponentialEuler, ForwardEuler, PseudoSpectralIMEX,  # noqa: E402
                                       RungeKutta4)
from evoxels_b200.voxelgrid import VoxelGridTorch  # noqa: E402

RHS_TOL = {torch.float32: 5e-6, torch.float64: 1e-12}
STEP_TOL = {torch.float32: 1e-5, torch.float64: 5e-8}   # fp64: wavenumbers/prefactor are fp32 upstream


def make_grid(shape, spacing, precision="float32"):
    dom = tuple(float(n * h) for n, h in zip(shape, spacing))
    vf = evo.VoxelFields(tuple(int(n) for n in shape), dom)
    vf.precision = precision
    return vf, VoxelGridTorch(vf.grid_info(), precision=precision, device="cuda")


def quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


CH_CASES = ["ch_readme16", "ch_odd_aniso", "ch_line16", "ch_pow2_small", "ch_neumann_x",
            "ch_dirichlet_x", "ch_mixed_rhs_only", "ch_neumann3_rhs_only",
            "ch_zdirichlet_rhs_only", "ch_odd_aniso_f64"]


@pytest.mark.parametrize("name", CH_CASES)
def test_ch_golden(cuda_device, name):
    g = load_golden(name)
    prec = "float64" if g["u0"].dtype == np.float64 else "float32"
    vf, vg = make_grid(g["u0"].shape, g["spacing"], prec)
    u = vg.init_scalar_field(g["u0"])
    prob = quiet(CahnHilliard, vg, eps=g["eps"], D=g["D"], A=g["A"], bc=g["bc"])
    rhs = prob.rhs(0.0, u)
    assert rel_l2(rhs[0].cpu().numpy(), g["rhs"]) <= RHS_TOL[u.dtype]
    if "padded" in g:
        assert np.array_equal(prob.pad_bc(u)[0].cpu().numpy(), g["padded"])
    if "step1" not in g:
        return
    ts = PseudoSpectralIMEX(prob, g["dt"])
    v = u
    u_before = u.clone()
    for i in range(1, g.get("nsteps", 1) + 1):
        v = ts.step(0.0, v)
        if f"step{i}" in g:
            assert rel_l2(v[0].cpu().numpy(), g[f"step{i}"]) <= STEP_TOL[u.dtype] * max(1, i ** 0.5)
    assert torch.equal(u, u_before), "step must not mutate its input"
    assert rel_l2(ts._fft_prefac.cpu().numpy(), g["prefac"]) <= 1e-6


def test_ch_custom_mu_hom(cuda_device):
    g = load_golden("ch_custom_mu")
    vf, vg = make_grid(g["u0"].shape, g["spacing"])

    def mu_log(c, lib=None):
        cc = torch.clip(c, 1e-4, 1 - 1e-4)
        return torch.log(cc / (1 - cc)) + 2.5 * (1 - 2 * c)

    prob = CahnHilliard(vg, eps=g["eps"], D=g["D"], mu_hom=mu_log)
    u = vg.init_scalar_field(g["u0"])
    assert rel_l2(prob.rhs(0, u)[0].cpu().numpy(), g["rhs"]) <= 2e-5
    out = PseudoSpectralIMEX(prob, g["dt"]).step(0, u)
    assert rel_l2(out[0].cpu().numpy(), g["step1"]) <= 1e-5


AC_CASES = ["ac_default_neumann", "ac_curv_force", "ac_periodic", "ac_mixed", "ac_line16",
            "ac_flat_bulk"]


@pytest.mark.parametrize("name", AC_CASES)
def test_ac_golden(cuda_device, name):
    g = load_golden(name)
    vf, vg = make_grid(g["u0"].shape, g["spacing"])
    prob = quiet(TwoPhaseAllenCahn, vg, eps=g["eps"], gab=g["gab"], M=g["M"], force=g["force"],
                 curvature=g["curvature"], bc=g["bc"])
    u = vg.init_scalar_field(g["u0"])
    tol = 2e-5   # the normal-Laplacian quotient amplifies fp32 rounding (oracle floor 1.7e-7..2.6e-6)
    assert rel_l2(prob.rhs(0, u)[0].cpu().numpy(), g["rhs"]) <= tol
    assert rel_l2(ForwardEuler(prob, g["dt"]).step(0, u)[0].cpu().numpy(), g["euler1"]) <= 1e-5
    assert rel_l2(RungeKutta4(prob, g["dt"]).step(0, u)[0].cpu().numpy(), g["rk4_1"]) <= 1e-5


def test_ghost_rules_and_padded_stencils(cuda_device):
    g = load_golden("ghost_and_stencils")
    vf, vg = make_grid(g["f0"].shape, g["spacing"])
    f = vg.init_scalar_field(g["f0"])
    for i in range(g["n"]):
        bc = ast.literal_eval(str(g[f"bc{i}"]))
        prob = quiet(ReactionDiffusion, vg, D=1.0, bc=bc)
        assert np.array_equal(prob.pad_bc(f)[0].cpu().numpy(), g[f"pad{i}"]), bc
    pad = vg.bc.pad_bc(f, (("neumann", None),) * 3)
    assert rel_l2(vg.laplace(pad)[0].cpu().numpy(), g["laplace"]) <= 5e-6
    assert rel_l2(vg.normal_laplace(pad)[0].cpu().numpy(), g["normal_laplace"]) <= 5e-5
    assert rel_l2(vg.gradient_norm_squared(pad)[0].cpu().numpy(), g["gradnorm2"]) <= 5e-6
    for nm, fn in [("x_face", vg.to_x_face), ("y_face", vg.to_y_face), ("z_face", vg.to_z_face),
                   ("gx_face", vg.grad_x_face), ("gy_face", vg.grad_y_face), ("gz_face", vg.grad_z_face)]:
        assert rel_l2(fn(pad)[0].cpu().numpy(), g[nm]) <= 1e-6
    assert rel_l2(vg.rfft_k_squared().cpu().numpy(), g["k2"]) <= 1e-7
    # the reference's own known-answer array (tests/test_solvers.py:105-126)
    vf2, vg2 = make_grid((2, 2, 2), (0.5, 0.5, 0.5))
    a = np.arange(1, 9, dtype=np.float32).reshape(2, 2, 2)
    prob = quiet(ReactionDiffusion, vg2, D=1.0, bc=(("dirichlet", (10.0, 20.0)), "neumann", "periodic"))
    expected = np.pad(a, 1, mode="wrap")
    expected[0] = 2.0 * 10.0 - expected[1]
    expected[-1] = 2.0 * 20.0 - expected[-2]
    expected[:, 0] = expected[:, 1]
    expected[:, -1] = expected[:, -2]
    assert np.allclose(prob.pad_bc(vg2.init_scalar_field(a))[0].cpu().numpy(), expected)


@pytest.mark.parametrize("shape", [(64, 64, 64), (33, 20, 18), (100, 100, 100), (48, 40, 132)])
@pytest.mark.parametrize("bc", [("periodic",) * 3, ("neumann",) * 3,
                                (("dirichlet", (0.2, 0.6)), "neumann", "periodic")])
def test_ch_rhs_vs_live_oracle(cuda_device, shape, bc):
    u = O.noise_field(shape, seed=3, lo=-0.1, amp=1.2)
    ref = O.ch_rhs(u, (1.0, 0.5, 2.0), 3.0, 1.0, bc)
    vf, vg = make_grid(shape, (1.0, 0.5, 2.0))
    prob = quiet(CahnHilliard, vg, bc=bc)
    got = prob.rhs(0, u.cuda())
    assert rel_l2(got.cpu().numpy(), ref.numpy()) <= 5e-6


@pytest.mark.parametrize("shape", [(64, 64, 64), (33, 20, 18), (40, 36, 132)])
@pytest.mark.parametrize("bc", [("neumann",) * 3, ("periodic",) * 3,
                                (("dirichlet", (0.0, 1.0)), "neumann", "periodic")])
def test_ac_vs_live_oracle(cuda_device, shape, bc):
    u = O.noise_field(shape, seed=1, lo=0.0, amp=1.0)
    orc = O.ACOracle(shape, (1.0, 1.0, 1.0), 0.05, bc=bc)
    vf, vg = make_grid(shape, (1.0, 1.0, 1.0))
    prob = quiet(TwoPhaseAllenCahn, vg, bc=bc)
    ts = ForwardEuler(prob, 0.05)
    v, w = u, u.cuda()
    for _ in range(3):
        v, w = orc.step(v), ts.step(0, w)
    assert rel_l2(w.cpu().numpy(), v.numpy()) <= 1e-5


@pytest.mark.parametrize("shape,backend", [((64, 64, 64), "cufft"), ((100, 100, 100), "cufft"),
                                           ((33, 20, 18), "cufft"), ((128, 64, 256), "cufft"),
                                           ((64, 64, 64), "native"), ((128, 64, 256), "native"),
                                           ((256, 256, 256), "native"), ((100, 100, 100), "native-mixed"),
                                           ((35, 20, 18), "native-mixed"), ((64, 64, 64), "native-mixed")])
def test_ch_imex_step_vs_live_oracle(cuda_device, shape, backend):
    u = O.noise_field(shape, seed=0)
    orc = O.CHOracle(shape, (1.0, 1.0, 1.0), 0.1)
    vf, vg = make_grid(shape, (1.0, 1.0, 1.0))
    ts = PseudoSpectralIMEX(CahnHilliard(vg), 0.1, fft_backend=backend)
    v, w = u, u.cuda()
    for i in range(3):
        v_new, w_new = orc.step(v), ts.step(0, w)
        # update-level check is ~100x stricter than the state-level one
        assert rel_l2((w_new - w).cpu().numpy(), (v_new - v).numpy()) <= 2e-5
        assert rel_l2(w_new.cpu().numpy(), v_new.numpy()) <= 1e-5
        v, w = v_new, w_new
    m0, m1 = float(u.double().mean()), float(w.double().mean())
    assert abs(m1 - m0) <= 2e-7 * abs(m0), "mass must be conserved to fp32 rounding"


@pytest.mark.parametrize("shape", [(8, 8, 16), (16, 32, 64), (64, 8, 32), (8, 128, 16), (32, 16, 128),
                                   (256, 64, 32), (64, 512, 64), (1024, 16, 32), (16, 16, 2048),
                                   (32, 2048, 16), (2048, 8, 16), (128, 128, 128)])
def test_native_fft_backend_matches_cufft_backend(cuda_device, shape):
    """The hand-written five-pass FFT path against the cuFFT path on identical inputs
    (both are checked against the oracle elsewhere; this sweeps every line length)."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    r = torch.randn(shape, device="cuda", generator=gen)
    u = torch.rand(shape, device="cuda", generator=gen)
    sp = (1.0, 0.5, 2.0)
    outs = {}
    for name, code in (("cufft", _native.FFT_CUFFT), ("native", _native.FFT_NATIVE)):
        plan = _native.ImexPlan(shape, torch.float32, "cuda", code)
        assert plan.backend_name == name
        out = torch.empty_like(u)
        plan.apply(u, r, out, sp, 0.1, 1.5, 2)
        upd = torch.empty_like(u)
        plan.apply(None, r, upd, sp, 0.1, 1.5, 2)
        assert torch.allclose(out, u + upd, rtol=0, atol=1e-6)
        outs[name] = (out - u).double()
    assert float((outs["native"] - outs["cufft"]).norm() / outs["cufft"].norm()) <= 2e-6


def test_on_the_fly_prefactor_is_the_reference_array(cuda_device):
    """The filter kernel recomputes dt/(1-dt*symbol) per element; against the reference's
    stored float32 array (golden `prefac`) it must agree to float32 rounding."""
    for name in ("ch_readme16", "ch_odd_aniso", "ch_pow2_small"):
        g = load_golden(name)
        shape = g["u0"].shape
        spec = torch.ones((shape[0], shape[1], shape[2] // 2 + 1), dtype=torch.complex64, device="cuda")
        _native.spectral_filter(spec, shape, g["spacing"], g["dt"], 2 * g["eps"] * g["D"] * g["A"], 2)
        got = spec.real.cpu().numpy()
        assert np.abs(got / g["prefac"] - 1).max() <= 2.5e-7, name
        assert (got == g["prefac"]).mean() >= 0.9, name      # bit-identical almost everywhere


def test_fft_backend_selection(cuda_device):
    P = _native.ImexPlan
    assert P((64, 64, 64), torch.float32, "cuda").backend_name == "native"          # radix-8 passes
    assert P((100, 100, 100), torch.float32, "cuda").backend_name == "native-mixed"  # 2^2 5^2
    assert P((64, 64, 64), torch.float64, "cuda").backend_name == "native-mixed"     # fp64
    assert P((12, 9, 7), torch.float32, "cuda").backend_name == "native-mixed"
    assert P((16, 1, 1), torch.float32, "cuda").backend_name == "native-mixed"       # degenerate axes
    assert P((22, 8, 8), torch.float32, "cuda").backend_name == "cufft"              # prime factor 11
    assert P((100, 100, 100), torch.float32, "cuda", _native.FFT_CUFFT).backend_name == "cufft"
    with pytest.raises(_native.NativeLibraryError):
        P((100, 64, 64), torch.float32, "cuda", _native.FFT_NATIVE)
    with pytest.raises(_native.NativeLibraryError):
        P((22, 8, 8), torch.float32, "cuda", _native.FFT_NATIVE_MIXED)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(100, 100, 100), (12, 9, 7), (16, 1, 1), (1, 6, 1), (48, 40, 126),
                                   (35, 27, 50), (64, 64, 64), (3, 2048, 5)])
def test_mixed_radix_fft_backend_matches_cufft_backend_and_oracle(cuda_device, shape, dtype):
    """The hand-written mixed-radix passes against the cuFFT back end (same filter arithmetic)
    and against the oracle's IMEX / ETD1 updates."""
    sp = (1.0, 0.5, 2.0)
    gen = torch.Generator(device="cuda").manual_seed(5)
    u = torch.rand(shape, device="cuda", generator=gen).to(dtype)
    r = torch.randn(shape, device="cuda", generator=gen).to(dtype)
    mixed = _native.ImexPlan(shape, dtype, "cuda", _native.FFT_NATIVE_MIXED)
    lib = _native.ImexPlan(shape, dtype, "cuda", _native.FFT_CUFFT)
    tol = 3e-6 if dtype == torch.float32 else 1e-10
    for dt, coef, power in ((0.1, 1.5, 2), (0.3, 0.7, 1), (0.3, 0.7, 1 | _native.FILTER_ETD1)):
        a, b = torch.empty_like(u), torch.empty_like(u)
        mixed.apply(u, r, a, sp, dt, coef, power)
        lib.apply(u, r, b, sp, dt, coef, power)
        assert rel_l2((a - u).cpu().numpy(), (b - u).cpu().numpy()) <= tol, (dt, coef, power)
    if max(shape) <= 132:
        with torch.device("cpu"):
            pref = O.imex_prefactor(O.ch_symbol(shape, sp, 3.0, 1.0, 0.25), 0.1)
            want = O.imex_step(u.cpu()[None], r.cpu()[None], pref)[0]
        a = torch.empty_like(u)
        mixed.apply(u, r, a, sp, 0.1, 1.5, 2)
        assert rel_l2((a - u).cpu().numpy(), (want - u.cpu()).numpy()) <= (3e-6 if dtype == torch.float32 else 2e-7)
    # update-only form (u = NULL) used by the mirrored-x path
    a, b = torch.empty_like(u), torch.empty_like(u)
    mixed.apply(None, r, a, sp, 0.1, 1.5, 2)
    lib.apply(None, r, b, sp, 0.1, 1.5, 2)
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= tol


def test_ch_nonperiodic_x_imex(cuda_device):
    for name in ("ch_neumann_x", "ch_dirichlet_x"):
        g = load_golden(name)
        vf, vg = make_grid(g["u0"].shape, g["spacing"])
        prob = quiet(CahnHilliard, vg, eps=g["eps"], D=g["D"], A=g["A"], bc=g["bc"])
        out = PseudoSpectralIMEX(prob, g["dt"]).step(0, vg.init_scalar_field(g["u0"]))
        assert rel_l2(out[0].cpu().numpy(), g["step1"]) <= 1e-5


def test_readme_config_1000_steps(cuda_device):
    """README.md:98-115: 100^3, dt=0.1, 1000 steps, against the reference record."""
    g = load_golden("ch_readme100_1000steps")
    vf = evo.VoxelFields((100, 100, 100), (100, 100, 100))
    u0 = 0.5 + 0.1 * np.random.default_rng(0).random((100, 100, 100)).astype(np.float32)
    vf.add_field("c", u0)
    vg = VoxelGridTorch(vf.grid_info(), device="cuda")
    ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=3.0, D=1.0), 0.1)
    v = vg.init_scalar_field(u0)
    for i in range(1, 1001):
        v = ts.step(0, v)
        if i in (1, 10, 100, 1000):
            a = v[0].cpu().numpy()
            tol = 1e-5 if i < 1000 else 1e-4
            assert rel_l2(a[::4, ::4, ::4], g[f"sub{i}"]) <= tol, i
            assert abs(a.astype(np.float64).mean() - g["mean0"]) <= 2e-7, "mass drift"


def test_solver_drivers(cuda_device):
    # seam test of the reference (tests/test_solvers.py:14-27)
    vf = evo.VoxelFields((4, 4, 4))
    vf.add_field("a", np.ones(vf.shape))
    vf.add_field("b", np.zeros(vf.shape))
    TimeDependentSolver(vf, ["a", "b"], backend="torch", step_fn=lambda t, u: u + 1,
                        device="cuda").solve(frames=1, max_iters=1, verbose=False, jit=False)
    assert np.allclose(vf.fields["a"], 2) and np.allclose(vf.fields["b"], 1)
    # 1-D tanh equilibrium (tests/test_solvers.py:29-82), both drivers on CUDA
    Nx = 16
    vf = evo.VoxelFields((Nx, 1, 1), domain_size=(Nx, 1, 1))
    phi = np.zeros((Nx, 1, 1), dtype=np.float32)
    phi[: Nx // 2] = 1.0
    vf.add_field("phi1", phi.copy())
    vf.add_field("phi2", phi.copy())
    eps = 3.0
    evo.run_allen_cahn_solver(vf, "phi1", backend="torch", device="cuda", frames=1, max_iters=10,
                              time_increment=0.5, eps=eps, jit=False, verbose=False)
    evo.run_cahn_hilliard_solver(vf, "phi2", backend="torch", device="cuda", frames=1,
                                 max_iters=10, time_increment=0.5, eps=eps, jit=True, verbose=False)
    x = np.arange(Nx) + 0.5
    ana = 0.5 - 0.5 * np.tanh(3 * (x - 0.5 * Nx) / 2 / eps)
    assert np.linalg.norm(vf.fields["phi1"].squeeze() - ana) < 0.05
    sel = (x > 5) & (x < 11)
    assert np.linalg.norm(vf.fields["phi2"].squeeze()[sel] - ana[sel]) < 0.05


def test_cuda_graph_solver_loop_matches_eager(cuda_device):
    """solve(jit=True) replays CUDA graphs of several steps; the result must equal the eager
    loop bit for bit (same kernels, same order)."""
    res = {}
    for jit in (False, True):
        vf = evo.VoxelFields((100, 100, 100), (100, 100, 100))
        vf.add_field("c", 0.5 + 0.1 * np.random.default_rng(0).random((100, 100, 100)).astype(np.float32))
        s = evo.run_cahn_hilliard_solver(vf, "c", backend="torch", device="cuda", jit=jit, frames=2,
                                         max_iters=100, time_increment=0.1, verbose=False)
        res[jit] = (vf.fields["c"].copy(), s.computation_time)
    assert np.array_equal(res[False][0], res[True][0])
    vf = evo.VoxelFields((64, 64, 64), (64, 64, 64))
    vf.add_field("p", np.random.default_rng(1).random((64, 64, 64)).astype(np.float32))
    ref = vf.fields["p"].copy()
    evo.run_allen_cahn_solver(vf, "p", backend="torch", device="cuda", jit=True, frames=1, max_iters=20,
                              time_increment=0.05, verbose=False)
    a = vf.fields["p"].copy()
    vf.set_field("p", ref)
    evo.run_allen_cahn_solver(vf, "p", backend="torch", device="cuda", jit=False, frames=1, max_iters=20,
                              time_increment=0.05, verbose=False)
    assert np.array_equal(a, vf.fields["p"])


def test_rhs_convergence_order(cuda_device):
    """Spatial order 2 of CahnHilliard.rhs / TwoPhaseAllenCahn.rhs in float64, the check of
    the reference's tests/test_rhs.py:10-37 (MMS on the unit cube)."""
    import sympy as sp
    import sympy.vector as spv
    CS = spv.CoordSys3D("CS")
    cases = [(CahnHilliard, dict(eps=3.0, D=1.0, A=0.25), 0.4 + 0.1 * sp.sin(2 * sp.pi * CS.x)),
             (TwoPhaseAllenCahn, dict(eps=3.0, curvature=0.5, force=1),
              0.5 + 0.3 * sp.cos(4 * sp.pi * CS.x) * sp.cos(2 * sp.pi * CS.y) * (CS.z ** 2 / 2 - CS.z ** 3 / 3))]
    for cls, kw, fun in cases:
        dx, err = [], []
        for p in (3, 4, 5, 6):
            n = 2 ** p
            vf = evo.VoxelFields((n, n, n), (1, 1, 1))
            vf.precision = "float64"
            vg = VoxelGridTorch(vf.grid_info(), precision="float64", device="cuda")
            grid = vf.meshgrid()
            u = vg.init_scalar_field(sp.lambdify((CS.x, CS.y, CS.z), fun, "numpy")(*grid))
            prob = cls(vg, **kw)
            num = prob.rhs(0, u)[0].cpu().numpy()
            exact = sp.lambdify((CS.x, CS.y, CS.z), prob.rhs_analytic(0, fun), "numpy")(*grid)
            dx.append(vf.spacing[0])
            err.append(np.linalg.norm(num - exact) / np.linalg.norm(exact))
        slope = np.polyfit(np.log(dx), np.log(err), 1)[0]
        assert abs(slope - 2) < 0.1, (cls.__name__, slope)


def test_full_size_properties_512(cuda_device):
    """BASELINE config 2 (512^3 fp32 periodic): size-independent properties - mass
    conservation, translation equivariance of the step, determinism."""
    n = 512
    gen = torch.Generator(device="cuda").manual_seed(0)
    u = (0.5 + 0.1 * torch.rand((1, n, n, n), device="cuda", generator=gen))
    vf, vg = make_grid((n, n, n), (1.0, 1.0, 1.0))
    ts = PseudoSpectralIMEX(CahnHilliard(vg), 0.1)
    v = ts.step(0, u)
    assert abs(float(v.double().mean()) - float(u.double().mean())) <= 2e-7
    assert torch.equal(v, ts.step(0, u))
    shift = (5, 17, 64)
    vs = ts.step(0, torch.roll(u, shift, (1, 2, 3)))
    d = (torch.roll(v, shift, (1, 2, 3)) - vs)
    assert float(d.norm() / v.norm()) <= 1e-6
    r = CahnHilliard(vg).rhs(0, u)
    assert abs(float(r.double().sum())) <= 1e-3 * float(r.double().abs().sum())


def test_ch_step_512_vs_oracle(cuda_device):
    """One step at 512^3 against the CPU oracle (needs ~6 GB host RAM, ~10 s)."""
    n = 512
    u = O.noise_field((n, n, n), seed=0)
    ref = O.CHOracle((n, n, n), (1.0, 1.0, 1.0), 0.1).step(u)
    vf, vg = make_grid((n, n, n), (1.0, 1.0, 1.0))
    got = PseudoSpectralIMEX(CahnHilliard(vg), 0.1).step(0, u.cuda()).cpu()
    assert rel_l2((got - u).numpy(), (ref - u).numpy()) <= 2e-5
    assert rel_l2(got.numpy(), ref.numpy()) <= 1e-5


# ---------------------------------------------------------------------------------------
# CoupledReactionDiffusion / ReactionDiffusion + ExponentialEuler
# ---------------------------------------------------------------------------------------
def _two_species(vg, u0):
    return torch.cat([vg.init_scalar_field(u0[0]), vg.init_scalar_field(u0[1])], 0)


@pytest.mark.parametrize("name", ["crd_default", "crd_pow2", "crd_f64"])
def test_crd_golden(cuda_device, name):
    """Fused two-species rhs, IMEX and exponential-Euler steps against the reference's
    outputs; crd_pow2 runs the native FFT (EVX_FILTER_ETD1 inside the x pass), the others
    the cuFFT back end with the stand-alone weight kernel."""
    g = load_golden(name)
    prec = "float64" if g["u0"].dtype == np.float64 else "float32"
    vf, vg = make_grid(g["u0"].shape[1:], g["spacing"], prec)
    u = _two_species(vg, g["u0"])
    prob = CoupledReactionDiffusion(vg, D_A=g["D_A"], D_B=g["D_B"], feed=g["feed"], kill=g["kill"])
    assert rel_l2(prob.rhs(0.0, u).cpu().numpy(), g["rhs"]) <= RHS_TOL[u.dtype]
    for cls, key in ((ExponentialEuler, "etd1"), (PseudoSpectralIMEX, "imex")):
        ts = cls(prob, g["dt"])
        v = u
        for i in range(1, g["nsteps"] + 1):
            v = ts.step(0.0, v)
            if i == 1:
                assert rel_l2(v.cpu().numpy(), g[key + "_step1"]) <= STEP_TOL[u.dtype], key
        assert rel_l2(v.cpu().numpy(), g[key + "_stepn"]) <= 3 * STEP_TOL[u.dtype], key
    ts = ExponentialEuler(prob, g["dt"])
    assert rel_l2(ts.phi_1_k_squared.cpu().numpy(), g["phi1"]) <= 1e-6
    if name == "crd_pow2":
        assert ts._plan(u.shape[1:], u.dtype, u.device).backend_name == "native"


def test_crd_custom_interaction_and_live_oracle(cuda_device):
    shape, sp = (32, 24, 40), (0.5, 1.0, 0.8)
    vf, vg = make_grid(shape, sp)
    rng = np.random.default_rng(3)
    u0 = np.stack([rng.random(shape), 0.5 * rng.random(shape)]).astype(np.float32)
    u = _two_species(vg, u0)
    fn = lambda v, lib=None: v[0] ** 2 * v[1] + 0.1          # noqa: E731
    for inter in (None, fn):
        prob = CoupledReactionDiffusion(vg, D_A=0.7, D_B=1.1, feed=0.04, kill=0.1, interaction=inter)
        with torch.device("cpu"):
            want = O.crd_rhs(torch.from_numpy(u0), sp, 0.7, 1.1, 0.04, 0.1, interaction=inter)
        assert rel_l2(prob.rhs(0.0, u).cpu().numpy(), want.numpy()) <= RHS_TOL[torch.float32]


@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 32, 256)])
def test_crd_etd1_native_vs_live_oracle(cuda_device, shape):
    sp = (1.0, 0.5, 2.0)
    vf, vg = make_grid(shape, sp)
    rng = np.random.default_rng(4)
    u0 = np.stack([rng.random(shape), 0.5 * rng.random(shape)]).astype(np.float32)
    u = _two_species(vg, u0)
    prob = CoupledReactionDiffusion(vg, D_A=1.0, D_B=0.5)
    out = ExponentialEuler(prob, 0.5, fft_backend="native").step(0.0, u)
    with torch.device("cpu"):
        t = torch.from_numpy(u0)
        sym = O.crd_symbol(shape, sp, 1.0, 0.5)
        want = O.etd1_step(t, O.crd_rhs(t, sp, 1.0, 0.5), sym, 0.5)
    assert rel_l2(out.cpu().numpy(), want.numpy()) <= STEP_TOL[torch.float32]
    assert rel_l2((out - u).cpu().numpy(), (want - t).numpy()) <= 2e-5


@pytest.mark.parametrize("name", ["rd_etd1_periodic", "rd_etd1_neumann_x"])
def test_rd_etd1_golden(cuda_device, name):
    g = load_golden(name)
    vf, vg = make_grid(g["u0"].shape, g["spacing"])
    u = vg.init_scalar_field(g["u0"])
    prob = quiet(ReactionDiffusion, vg, D=g["D"], f=lambda t, c, lib=None: c * (1 - c), bc=g["bc"])
    assert rel_l2(prob.rhs(0.0, u)[0].cpu().numpy(), g["rhs"]) <= RHS_TOL[u.dtype]
    ts = ExponentialEuler(prob, g["dt"])
    v = ts.step(0.0, u)
    assert rel_l2(v[0].cpu().numpy(), g["etd1_step1"]) <= STEP_TOL[u.dtype]
    v = ts.step(0.0, v)
    assert rel_l2(v[0].cpu().numpy(), g["etd1_stepn"]) <= 2 * STEP_TOL[u.dtype]


# ---------------------------------------------------------------------------------------
# asynchronous frame export
# ---------------------------------------------------------------------------------------
def test_async_frame_export_commits_every_frame_and_aborts_on_nan(cuda_device):
    frames_seen = []

    class Spy(TimeDependentSolver):
        def _commit_frame(self, host, has_nan, frame, time):
            frames_seen.append((frame, round(float(time), 6), float(host.numpy()[0, 0, 0, 0])))
            super()._commit_frame(host, has_nan, frame, time)

    vf = evo.VoxelFields((8, 8, 8))
    vf.add_field("a", np.zeros(vf.shape, dtype=np.float32))
    Spy(vf, "a", backend="torch", step_fn=lambda t, u: u + 1, device="cuda").solve(
        time_increment=0.5, frames=4, max_iters=8, verbose=False, jit=False)
    # frames at iterations 0, 2, 4, 6 and the final state, each with the value of its own time
    assert frames_seen == [(0, 0.0, 0.0), (1, 1.0, 2.0), (2, 2.0, 4.0), (3, 3.0, 6.0), (4, 4.0, 8.0)]
    assert np.all(vf.fields["a"] == 8)

    vf = evo.VoxelFields((8, 8, 8))
    vf.add_field("a", np.zeros(vf.shape, dtype=np.float32))
    bad = TimeDependentSolver(vf, "a", backend="torch", device="cuda",
                              step_fn=lambda t, u: u + float("nan"))
    with pytest.raises(SystemExit):
        bad.solve(time_increment=0.5, frames=2, max_iters=4, verbose=False, jit=False)


def test_time_dependent_source_runs_eagerly_under_jit(cuda_device):
    """jit=True must not freeze t (round-1 advisor finding): a ReactionDiffusion source f(t, u)
    makes the problem non-autonomous, the solver then skips graph capture and the result equals
    the eager loop - and differs from what a frozen t = 0 would give."""
    res = {}
    for jit in (False, True):
        vf = evo.VoxelFields((16, 16, 16), (16, 16, 16))
        vf.add_field("c", np.zeros(vf.shape, dtype=np.float32))
        s = TimeDependentSolver(vf, "c", backend="torch", problem_cls=ReactionDiffusion,
                                timestepper_cls=ForwardEuler, device="cuda")
        s.solve(time_increment=0.1, frames=2, max_iters=20, verbose=False, jit=jit,
                problem_kwargs=dict(D=0.1, f=lambda t, c, lib: t + 0 * c))
        assert s._graph_ok is False
        res[jit] = vf.fields["c"].copy()
    assert np.array_equal(res[False], res[True])
    # du/dt = t  ->  forward Euler sum_{i<20} 0.1 * (0.1 i) = 1.9
    assert np.allclose(res[True], 1.9, atol=1e-5)


def test_kernel_paths_refuse_to_drop_gradients(cuda_device):
    _, vg = make_grid((8, 8, 8), (1, 1, 1))
    u = torch.rand((1, 8, 8, 8), device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        ForwardEuler(TwoPhaseAllenCahn(vg), 0.05).step(0.0, u)
    with pytest.raises(NotImplementedError):
        RungeKutta4(TwoPhaseAllenCahn(vg), 0.05).step(0.0, u)
    with pytest.raises(NotImplementedError):
        ExponentialEuler(ReactionDiffusion(vg, D=1.0), 0.1).step(0.0, u)
    with pytest.raises(NotImplementedError):
        vg.laplace(vg.pad_periodic(u))
    with torch.no_grad():
        ForwardEuler(TwoPhaseAllenCahn(vg), 0.05).step(0.0, u)
    # the supported differentiable path still works
    out = PseudoSpectralIMEX(CahnHilliard(vg), 0.1).step(0.0, u)
    out.sum().backward()
    assert u.grad is not None and torch.isfinite(u.grad).all()


def test_subclass_overrides_are_honoured(cuda_device):
    """A CahnHilliard subclass with its own rhs must not be routed to the fused kernel."""
    class WithSource(CahnHilliard):
        def rhs(self, t, c):
            return super().rhs(t, c) + 0.25

    _, vg = make_grid((16, 16, 16), (1, 1, 1))
    u = O.noise_field((16, 16, 16), seed=3).to("cuda")
    stock = PseudoSpectralIMEX(CahnHilliard(vg), 0.1).step(0.0, u)
    sub = PseudoSpectralIMEX(WithSource(vg), 0.1).step(0.0, u)
    # a constant source only feeds the k = 0 mode, whose weight is dt
    assert rel_l2(sub.cpu(), (stock + 0.1 * 0.25).cpu()) <= 1e-6


@pytest.mark.parametrize("shape", [(512, 512, 32), (16, 512, 64), (512, 32, 16)])
@pytest.mark.parametrize("kz", ["8", "16"])
def test_tma_tiled_passes_match_the_cp_async_passes_bit_for_bit(cuda_device, monkeypatch, shape, kz):
    """512-point strided passes: TMA-tiled kernels (fft_line.cu; 64- and 128-byte tile rows)
    against the cp.async kernels they replace - same arithmetic, so the update must be
    bit-identical - for the IMEX and the exponential-Euler weight, plus the oracle step."""
    gen = torch.Generator(device="cuda").manual_seed(2)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    outs = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("EVX_FFT_TMA", tma)
        monkeypatch.setenv("EVX_FFT_TMA_KZ", kz)
        res = []
        for power in (2, 1 | _native.FILTER_ETD1):
            out = torch.full_like(u, float("nan"))
            plan.apply(u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, power)
            res.append(out)
        out = torch.full_like(u, float("nan"))
        plan.ch_step(u, out, (1.0, 1.0, 1.0), 0.1, 3.0, 1.0, 0.25)
        res.append(out)
        torch.cuda.synchronize()
        outs[tma] = res
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.isfinite(b).all()
        assert torch.equal(a, b)
    ref = O.CHOracle(shape, (1.0, 1.0, 1.0), 0.1).step(u.cpu()[None])[0]
    assert rel_l2(outs["1"][2].cpu(), ref) <= 1e-5


@pytest.mark.parametrize("shape", [(32, 512, 512), (8, 512, 512), (512, 512, 512)])
def test_chained_zy_passes_match_one_kernel_per_pass_bit_for_bit(cuda_device, monkeypatch, shape):
    """fft_chain.cu: z lines and y tiles of a plane run inside one persistent kernel per
    direction (the plane's half spectrum goes through L2).  Same arithmetic as the stand-alone
    passes, so update and fused CH step must be bit-identical - also when the planes are fewer
    than the schedule's lag - and repeatable (no dependence on how the blocks interleave).  (The
    one-kernel-per-pass pipeline is the one checked against the oracle, test_ch_step_512_vs_oracle.)"""
    gen = torch.Generator(device="cuda").manual_seed(5)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    outs = {}
    for chain in ("0", "1", "1"):
        monkeypatch.setenv("EVX_FFT_CHAIN", chain)
        res = []
        out = torch.full_like(u, float("nan"))
        plan.apply(u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        res.append(out)
        out = torch.full_like(u, float("nan"))
        plan.apply(None, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        res.append(out)
        out = torch.full_like(u, float("nan"))
        plan.ch_step(u, out, (1.0, 1.0, 1.0), 0.1, 3.0, 1.0, 0.25)
        res.append(out)
        torch.cuda.synchronize()
        outs.setdefault(chain, []).append(res)
    base = outs["0"][0]
    for run in outs["1"]:
        for a, b in zip(base, run):
            assert torch.isfinite(b).all()
            assert torch.equal(a, b)


@pytest.mark.parametrize("ws", ["0", "3", "4", "5"])
def test_line_pass_forms_agree_bit_for_bit(cuda_device, monkeypatch, ws):
    """TMA-tiled strided passes: the two-blocks-per-SM form (tile hand-over by the compute
    threads) and the warp-specialised form with 3 / 4 / 5 tile buffers (loader + retirer warps,
    roots in registers) against the cp.async passes - separate passes, so the y passes go
    through the line kernels as well."""
    shape = (512, 512, 64)
    gen = torch.Generator(device="cuda").manual_seed(8)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    monkeypatch.setenv("EVX_FFT_CHAIN", "0")
    outs = []
    for tma, form in (("0", "0"), ("1", ws)):
        monkeypatch.setenv("EVX_FFT_TMA", tma)
        monkeypatch.setenv("EVX_FFT_LINE_WS", form)
        out = torch.full_like(u, float("nan"))
        plan.apply(u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, 2)
        etd = torch.full_like(u, float("nan"))
        plan.apply(u, r, etd, (1.0, 0.5, 2.0), 0.5, 1.0, 1 | _native.FILTER_ETD1)
        torch.cuda.synchronize()
        outs.append((out, etd))
    for a, b in zip(*outs):
        assert torch.isfinite(b).all() and torch.equal(a, b)


@pytest.mark.parametrize("form", ["default", "8", "16"])
@pytest.mark.parametrize("shape", [(1024, 64, 32), (32, 1024, 64), (1024, 1024, 16), (1024, 8, 16)])
def test_four_stage_tma_passes_match_the_cp_async_passes_bit_for_bit(cuda_device, monkeypatch, shape, form):
    """1024-point strided passes: the TMA-tiled kernel (fft_line4_ws_kernel, 4-D tensor maps) in its
    eight-point (StridedLine4) and sixteen-point (StridedLine16) forms against the cp.async kernels -
    y forward, x forward*weight*inverse (IMEX and exponential-Euler weight; the latter always runs
    the eight-point form), y inverse; the update must be bit-identical."""
    if form != "default":
        monkeypatch.setenv("EVX_LINE4_FORM", form)
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = 0.5 + 0.1 * torch.rand(shape, device="cuda", generator=gen)
    r = torch.randn(shape, device="cuda", generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, "cuda", _native.FFT_NATIVE)
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("EVX_FFT_LINE4", flag)
        res = []
        for power in (2, 1 | _native.FILTER_ETD1):
            out = torch.full_like(u, float("nan"))
            plan.apply(u, r, out, (1.0, 0.5, 2.0), 0.1, 1.5, power)
            res.append(out)
        torch.cuda.synchronize()
        outs[flag] = res
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.isfinite(b).all()
        assert torch.equal(a, b)


def test_cuda_graph_replay_of_the_chained_step(cuda_device, recwarn):
    """The 512-point path (cooperative chained kernels, warp-specialised x pass, two memsets per
    step) inside the solver's CUDA-graph replay: same fields as the eager loop, and the capture
    must really have happened (no fall-back warning)."""
    shape = (16, 512, 512)
    res = {}
    for jit in (False, True):
        vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
        vf.add_field("c", 0.5 + 0.1 * np.random.default_rng(3).random(shape).astype(np.float32))
        evo.run_cahn_hilliard_solver(vf, "c", backend="torch", device="cuda", jit=jit, frames=2,
                                     max_iters=20, time_increment=0.1, verbose=False)
        res[jit] = vf.fields["c"].copy()
    assert np.isfinite(res[True]).all() and np.array_equal(res[False], res[True])
    assert not [w for w in recwarn.list if "CUDA-graph capture" in str(w.message)]


_PER = (("periodic", None),) * 3
_NEU = (("neumann", None),) * 3
_MIX = (("dirichlet", (0.2, 0.6)), ("neumann", None), ("periodic", None))
_MIX2 = (("periodic", None), ("dirichlet", (0.1, 0.9)), ("dirichlet", (0.3, 0.5)))


@pytest.mark.parametrize("cfg", ["0", "1", "2", "3"])
@pytest.mark.parametrize("shape", [(64, 64, 64), (5, 16, 256), (100, 100, 100), (48, 40, 132),
                                   (20, 18, 260), (130, 34, 512), (2, 4, 64)])
def test_ch_rhs_warp_specialised_form_matches_cp_async_form(cuda_device, monkeypatch, shape, cfg):
    """ch_rhs_tma.cu (loader thread + TMA boxes, two rows per thread; tile heights 8 / 12 / 16,
    named-barrier and mbarrier mu exchange) against the cp.async kernel of ch_rhs_core.h on the
    same fields: partial tiles in y and z, one-group tiles, periodic wrap through the ring boxes,
    Neumann / Dirichlet ghosts on every axis.  Both are checked against the oracle elsewhere; the
    two forms differ by the reassociated potential polynomial only."""
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = -0.1 + 1.2 * torch.rand(shape, device="cuda", generator=gen)
    for bc in (_PER, _NEU, _MIX, _MIX2):
        outs = []
        for tma in ("0", "1"):
            monkeypatch.setenv("EVX_CH_TMA", tma)
            monkeypatch.setenv("EVX_CH_TMA_CFG", cfg)
            out = torch.full_like(u, float("nan"))
            _native.ch_rhs(u, out, (1.0, 0.5, 2.0), 3.0, 1.0, bc)
            torch.cuda.synchronize()
            outs.append(out)
        assert torch.isfinite(outs[1]).all()
        assert rel_l2(outs[1].cpu().numpy(), outs[0].cpu().numpy()) <= 5e-7, bc


@pytest.mark.parametrize("bc", [_PER, _NEU])
def test_ch_rhs_warp_specialised_form_with_x_halos(cuda_device, monkeypatch, bc):
    """x-slab halos (the multi-GPU form): three slabs of one field with their neighbours' planes
    handed in as halo_lo / halo_hi reproduce the rhs of the whole field."""
    gen = torch.Generator(device="cuda").manual_seed(2)
    full = -0.1 + 1.2 * torch.rand((24, 32, 128), device="cuda", generator=gen)
    monkeypatch.setenv("EVX_CH_TMA", "0")
    ref = torch.empty_like(full)
    _native.ch_rhs(full, ref, (1.0, 0.5, 2.0), 3.0, 1.0, bc)
    monkeypatch.setenv("EVX_CH_TMA", "1")
    for a, b in ((0, 8), (8, 16), (16, 24)):
        sl = full[a:b].contiguous()
        if bc is _PER:
            lo = torch.stack([full[(a - 2) % 24], full[(a - 1) % 24]]).contiguous()
            hi = torch.stack([full[b % 24], full[(b + 1) % 24]]).contiguous()
        else:
            lo = full[a - 2:a].contiguous() if a >= 2 else None
            hi = full[b:b + 2].contiguous() if b + 2 <= 24 else None
        out = torch.full_like(sl, float("nan"))
        _native.ch_rhs(sl, out, (1.0, 0.5, 2.0), 3.0, 1.0, bc, halo_lo=lo, halo_hi=hi)
        torch.cuda.synchronize()
        assert rel_l2(out.cpu().numpy(), ref[a:b].cpu().numpy()) <= 5e-7


@pytest.mark.parametrize("shape,dtype", [((64, 32, 64), torch.float32), ((16, 8, 16), torch.float32),
                                         ((512, 16, 32), torch.float32), ((1024, 8, 16), torch.float32),
                                         ((50, 12, 14), torch.float32), ((9, 7, 5), torch.float32),
                                         ((30, 10, 12), torch.float64)])
@pytest.mark.parametrize("flag", ["even", "odd"])
def test_mirrored_x_pass_equals_the_extended_transform(cuda_device, shape, dtype, flag):
    """Non-periodic x (reference boundary_conditions.py:65-71: rfftn of the field concatenated with
    its flipped / negated copy): the x pass that synthesises the mirror image of every line in
    shared memory (EVX_FILTER_MIRROR_EVEN / _ODD on the un-extended arrays) against the same plan
    applied to the explicitly extended 2 Nx field - radix-8 passes (powers of two, up to 2 Nx =
    2048), mixed-radix passes (incl. odd extents and float64), IMEX and exponential-Euler weight."""
    nx, ny, nz = shape
    gen = torch.Generator(device="cuda").manual_seed(12)
    u = torch.rand(shape, device="cuda", generator=gen, dtype=dtype)
    r = torch.randn(shape, device="cuda", generator=gen, dtype=dtype)
    sign = 1.0 if flag == "even" else -1.0
    mflag = _native.FILTER_MIRROR_EVEN if flag == "even" else _native.FILTER_MIRROR_ODD
    sp = (1.0, 0.5, 2.0)
    plan = _native.ImexPlan(shape, dtype, "cuda", _native.FFT_AUTO)
    ext_shape = (2 * nx, ny, nz)
    plan2 = _native.ImexPlan(ext_shape, dtype, "cuda", _native.FFT_AUTO)
    if plan.backend_name == "cufft":
        pytest.skip("extents beyond the native passes")
    r_ext = torch.cat([r, sign * torch.flip(r, [0])], 0).contiguous()
    for power in (2, 1 | _native.FILTER_ETD1):
        got = torch.full_like(u, float("nan"))
        plan.apply(u, r, got, sp, 0.1, 1.5, power | mflag)
        upd = torch.empty_like(r_ext)
        plan2.apply(None, r_ext, upd, sp, 0.1, 1.5, power)
        ref = u + upd[:nx]
        torch.cuda.synchronize()
        tol = 1e-12 if dtype == torch.float64 else 2e-6
        scale = float(upd.abs().max())
        assert float((got - ref).abs().max()) <= tol * max(scale, 1.0), (power, float((got - ref).abs().max()), scale)
