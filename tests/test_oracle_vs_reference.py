"""Oracle vs the live, unmodified reference (only where /root/reference exists, i.e. in the
build container - skipped on the GPU box)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import evx_oracle as O
from oracle.ref_shim import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference sources not mounted")

BCS = [("periodic",) * 3, ("neumann",) * 3, ("neumann", "periodic", "periodic"),
       (("dirichlet", (0.3, 0.7)), "periodic", "periodic"),
       (("dirichlet", (0.3, 0.7)), "neumann", "periodic"),
       ("periodic", "neumann", ("dirichlet", (0.1, 0.2)))]


@pytest.fixture(scope="module")
def ref():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return load_reference()


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("shape,dom", [((12, 9, 7), (6.0, 9.0, 3.5)), ((16, 1, 1), (16, 1, 1))])
def test_bitwise_equal_to_reference(ref, prec, shape, dom):
    vf = ref.voxelfields.VoxelFields(shape, dom)
    vf.precision = prec
    vg = ref.voxelgrid.VoxelGridTorch(vf.grid_info(), prec, "cpu")
    u = vg.init_scalar_field(-0.2 + 1.4 * np.random.default_rng(1).random(shape))
    for bc in BCS:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ch = ref.problem_definition.CahnHilliard(vg, eps=2.5, D=1.3, bc=bc)
            ac = ref.problem_definition.TwoPhaseAllenCahn(vg, eps=3.0, curvature=0.5, force=1.0, bc=bc)
        assert torch.equal(O.ch_rhs(u, vf.spacing, 2.5, 1.3, bc), ch.rhs(0, u))
        assert torch.equal(O.ac_rhs(u, vf.spacing, 3.0, 1.0, 1.0, 1.0, 0.5, bc), ac.rhs(0, u))
        assert torch.equal(O.ghost_pad(u, bc), ch.pad_bc(u))
        kinds = ch.bc_type
        if all(k == "periodic" for k in kinds[1:]):
            ts = ref.timesteppers.PseudoSpectralIMEX(ch, 0.1)
            assert torch.equal(O.ch_imex_step(u, vf.spacing, 0.1, 2.5, 1.3, 0.25, bc), ts.step(0, u))
        eu = ref.timesteppers.ForwardEuler(ac, 0.05).step(0, u)
        orc = O.ACOracle(shape, vf.spacing, 0.05, 3.0, 1.0, 1.0, 1.0, 0.5, bc)
        assert torch.equal(orc.step(u), eu)
