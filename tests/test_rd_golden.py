"""SURVEY 8(f) row 4 - CoupledReactionDiffusion / ReactionDiffusion with ExponentialEuler:
the oracle restatement against fixtures made by the UNMODIFIED reference
(tests/golden/make_golden_rd.py).  CPU only; pins the oracle for these rows."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import evx_oracle as O

CRD_CASES = ["crd_default", "crd_pow2", "crd_f64"]


@pytest.mark.parametrize("name", CRD_CASES)
def test_crd_rhs_and_steps_match_reference(name):
    g = load_golden(name)
    u = torch.from_numpy(g["u0"])
    kw = dict(D_A=g["D_A"], D_B=g["D_B"], feed=g["feed"], kill=g["kill"])
    assert np.array_equal(O.crd_rhs(u, g["spacing"], **kw).numpy(), g["rhs"])
    sym = O.crd_symbol(u.shape[1:], g["spacing"], g["D_A"], g["D_B"])
    assert sym.dtype == torch.float32          # wavenumbers stay float32 even for fp64 fields
    assert np.array_equal(O.phi1(g["dt"] * sym).numpy(), g["phi1"])
    ve = vi = u
    pref = O.imex_prefactor(sym, g["dt"])
    for i in range(1, g["nsteps"] + 1):
        ve = O.etd1_step(ve, O.crd_rhs(ve, g["spacing"], **kw), sym, g["dt"])
        vi = O.imex_step(vi, O.crd_rhs(vi, g["spacing"], **kw), pref)
        if i == 1:
            assert np.array_equal(ve.numpy(), g["etd1_step1"])
            assert np.array_equal(vi.numpy(), g["imex_step1"])
    assert np.array_equal(ve.numpy(), g["etd1_stepn"])
    assert np.array_equal(vi.numpy(), g["imex_stepn"])


@pytest.mark.parametrize("name", ["rd_etd1_periodic", "rd_etd1_neumann_x"])
def test_rd_etd1_matches_reference(name):
    g = load_golden(name)
    u = torch.from_numpy(g["u0"])[None]
    bc = O.normalize_bc(g["bc"])
    x_kind = bc[0][0]

    def rhs(v):
        return g["D"] * O.laplace7(O.ghost_pad(v, bc), g["spacing"]) + v * (1 - v)

    assert rel_l2(rhs(u)[0].numpy(), g["rhs"]) <= 1e-7
    sym = O.rd_symbol(u.shape[1:], g["spacing"], g["D"], g["A"], mirrored_x=x_kind != "periodic")
    assert rel_l2(O.phi1(g["dt"] * sym).numpy(), g["phi1"]) <= 1e-7
    v = u
    for i in range(1, g["nsteps"] + 1):
        v = O.etd1_step(v, rhs(v), sym, g["dt"], x_kind)
        if i == 1:
            assert rel_l2(v[0].numpy(), g["etd1_step1"]) <= 1e-7
    assert rel_l2(v[0].numpy(), g["etd1_stepn"]) <= 1e-7


def test_phi1_branches():
    z = torch.tensor([0.0, -1e-6, -0.49, 0.49, -0.5, -3.0, -40.0], dtype=torch.float64)
    p = O.phi1(z)
    exact = torch.where(z == 0, torch.ones_like(z), torch.expm1(z) / torch.where(z == 0, torch.ones_like(z), z))
    assert torch.allclose(p, exact, rtol=1e-12, atol=0)
