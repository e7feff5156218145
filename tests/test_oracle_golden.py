"""The oracle (oracle/evx_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  This is what pins the oracle."""
import ast

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import evx_oracle as O

CH_CASES = ["ch_readme16", "ch_odd_aniso", "ch_line16", "ch_pow2_small", "ch_neumann_x",
            "ch_dirichlet_x", "ch_mixed_rhs_only", "ch_neumann3_rhs_only",
            "ch_zdirichlet_rhs_only", "ch_odd_aniso_f64"]
AC_CASES = ["ac_default_neumann", "ac_curv_force", "ac_periodic", "ac_mixed", "ac_line16",
            "ac_flat_bulk"]


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))[None]


@pytest.mark.parametrize("name", CH_CASES)
def test_ch_rhs_and_steps_match_reference(name):
    g = load_golden(name)
    u = T(g["u0"])
    tol = 1e-12 if u.dtype == torch.float64 else 1e-6
    rhs = O.ch_rhs(u, g["spacing"], g["eps"], g["D"], g["bc"])[0].numpy()
    assert rel_l2(rhs, g["rhs"]) <= tol
    if "padded" in g:
        assert np.array_equal(O.ghost_pad(torch.clip(u, 0, 1) * 0 + u, g["bc"])[0].numpy(), g["padded"])
    if "step1" not in g:
        return
    orc = O.CHOracle(u.shape[1:], g["spacing"], g["dt"], g["eps"], g["D"], g["A"], g["bc"])
    assert rel_l2(orc.prefac.numpy(), g["prefac"]) <= 1e-7
    v = u
    n = g.get("nsteps", 1)
    for i in range(1, n + 1):
        v = orc.step(v)
        if f"step{i}" in g:
            assert rel_l2(v[0].numpy(), g[f"step{i}"]) <= tol


def test_ch_custom_mu_hom():
    g = load_golden("ch_custom_mu")

    def mu_log(c):
        cc = torch.clip(c, 1e-4, 1 - 1e-4)
        return torch.log(cc / (1 - cc)) + 2.5 * (1 - 2 * c)

    u = T(g["u0"])
    rhs = O.ch_rhs(u, g["spacing"], g["eps"], g["D"], g["bc"], mu_log)[0].numpy()
    assert rel_l2(rhs, g["rhs"]) <= 1e-6
    step = O.ch_imex_step(u, g["spacing"], g["dt"], g["eps"], g["D"], g["A"], g["bc"], mu_log)
    assert rel_l2(step[0].numpy(), g["step1"]) <= 1e-6


@pytest.mark.parametrize("name", AC_CASES)
def test_ac_rhs_euler_rk4_match_reference(name):
    g = load_golden(name)
    u = T(g["u0"])
    orc = O.ACOracle(u.shape[1:], g["spacing"], g["dt"], g["eps"], g["gab"], g["M"], g["force"],
                     g["curvature"], g["bc"])
    assert rel_l2(orc.rhs(u)[0].numpy(), g["rhs"]) <= 1e-6
    assert rel_l2(orc.step(u)[0].numpy(), g["euler1"]) <= 1e-6
    orc.scheme = "rk4"
    assert rel_l2(orc.step(u)[0].numpy(), g["rk4_1"]) <= 1e-6


def test_ghost_rules_and_padded_stencils():
    g = load_golden("ghost_and_stencils")
    f = T(g["f0"])
    for i in range(g["n"]):
        bc = ast.literal_eval(str(g[f"bc{i}"]))
        assert np.array_equal(O.ghost_pad(f, bc)[0].numpy(), g[f"pad{i}"]), bc
    pad = O.ghost_pad(f, ("neumann",) * 3)
    assert rel_l2(O.laplace7(pad, g["spacing"])[0].numpy(), g["laplace"]) <= 1e-6
    assert rel_l2(O.normal_laplace19(pad, g["spacing"])[0].numpy(), g["normal_laplace"]) <= 1e-6
    assert rel_l2(O.k_squared(f.shape[1:], g["spacing"]).numpy(), g["k2"]) <= 1e-7
    assert rel_l2(O.k_squared(f.shape[1:], g["spacing"], True).numpy(), g["k2_mirror"]) <= 1e-7


def test_reference_test_suite_known_answer_padding():
    """The explicit expected array of the reference's own test
    (tests/test_solvers.py:105-126): Dirichlet x / Neumann y / periodic z on a 2x2x2 field."""
    a = np.arange(1, 9, dtype=np.float32).reshape(2, 2, 2)
    expected = np.pad(a, 1, mode="wrap")
    expected[0] = 2.0 * 10.0 - expected[1]
    expected[-1] = 2.0 * 20.0 - expected[-2]
    expected[:, 0] = expected[:, 1]
    expected[:, -1] = expected[:, -2]
    got = O.ghost_pad(T(a), (("dirichlet", (10.0, 20.0)), "neumann", "periodic"))[0].numpy()
    assert np.allclose(got, expected)


def test_gradients_match_reference_autograd():
    g = load_golden("ch_grad_f64")
    D = torch.tensor(g["D"], dtype=torch.float64, requires_grad=True)
    eps = torch.tensor(g["eps"], dtype=torch.float64, requires_grad=True)
    u = T(g["u0"]).requires_grad_(True)
    v = u
    for _ in range(g["nsteps"]):
        v = O.ch_imex_step(v, g["spacing"], g["dt"], eps, D, g["A"])
    loss = ((v - T(g["target"])) ** 2).sum()
    gu, gD, ge = torch.autograd.grad(loss, (u, D, eps))
    assert abs(float(loss) - g["loss"]) <= 1e-12 * abs(g["loss"])
    assert rel_l2(gu[0].numpy(), g["grad_u0"]) <= 1e-10
    assert abs(float(gD) - g["grad_D"]) <= 2e-6 * abs(g["grad_D"])   # symbol path is float32 upstream
    assert abs(float(ge) - g["grad_eps"]) <= 2e-6 * abs(g["grad_eps"])


def test_readme_config_first_steps():
    """README.md:98-115 configuration (100^3, dt=0.1): steps 1 and 10 against the stored
    sub-sample / statistics (the 1000-step record is checked in the gpu suite)."""
    g = load_golden("ch_readme100_1000steps")
    u = O.noise_field((100, 100, 100), seed=0)
    orc = O.CHOracle((100, 100, 100), g["spacing"], g["dt"], g["eps"], g["D"], g["A"])
    v = u
    for i in range(1, 11):
        v = orc.step(v)
        if i in (1, 10):
            a = v[0].numpy()
            assert rel_l2(a[::4, ::4, ::4], g[f"sub{i}"]) <= 1e-6
            assert abs(a.astype(np.float64).mean() - g[f"mean{i}"]) <= 1e-7
