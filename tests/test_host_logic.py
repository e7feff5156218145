"""Host-side behaviour that needs no GPU: the C-ABI library loads and exports what the
header declares, boundary-condition normalisation/validation follows the reference
(tests/test_solvers.py:85-102,165-171), the data model follows tests/test_fields.py, and
the product never touches the oracle or a CPU compute path."""
import os
import re
import warnings

import numpy as np
import pytest
import torch

from conftest import ROOT
import evoxels_b200 as evo
from evoxels_b200 import _native
from evoxels_b200.problem_definition import (CahnHilliard, ReactionDiffusion, TwoPhaseAllenCahn,
                                             normalize_bc)
from evoxels_b200.timesteppers import ForwardEuler, PseudoSpectralIMEX, RungeKutta4
from evoxels_b200.voxelgrid import VoxelGridTorch


def host_grid(shape=(4, 4, 4)):
    return VoxelGridTorch(evo.VoxelFields(shape).grid_info(), device="cpu")


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "evoxels_b200.h")).read()
    declared = set(re.findall(r"\b(evx_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_native.SIGNATURES) == declared
    assert lib.evx_version() == 100
    assert b"invalid argument" in lib.evx_strerror(-1)


def test_argument_errors_are_reported_without_a_gpu():
    lib = _native.load_library()
    assert lib.evx_ch_rhs_f32(None, None, None, 4, 4, 4, None, 3.0, 1.0, None, None, None, None, None) == -1
    # entry points of the chunked / hybrid distributed passes: null plan, null tables, empty batches
    assert lib.evx_dist_middle_chunk_p2p_f32(None, None, None, 0, 1, None, 0.1, 1.0, 2, None) == -1
    assert lib.evx_dist_forward_chunk_p2p_f32(None, None, None, None, 0, 1, 3, None) == -1
    assert lib.evx_copy_batch_async(None, 0, None, 0, 16, 1, 1, None) == -1
    import ctypes
    one = (ctypes.c_void_p * 1)(1)
    assert lib.evx_copy_batch_async(one, 16, one, 16, 16, 1, 9, ctypes.c_void_p(1)) == -1      # more than 8 regions
    assert lib.evx_copy_batch_async(one, 16, one, 16, 0, 1, 1, ctypes.c_void_p(1)) == -1       # empty region
    assert lib.evx_copy_batch_async(one, 16, one, 16, 16, 1, 1, None) == -1                    # legacy stream
    with pytest.raises(_native.NativeLibraryError):
        _native.check(-2, "probe")


def test_no_cpu_path():
    vg = host_grid()
    u = torch.rand(1, 4, 4, 4)
    for prob in (CahnHilliard(vg), TwoPhaseAllenCahn(vg)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            prob.rhs(0.0, u)
    with pytest.raises(RuntimeError, match="no CPU path"):
        PseudoSpectralIMEX(CahnHilliard(vg), 0.1).step(0.0, u)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vg.pad_periodic(u)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vg.laplace(torch.rand(1, 6, 6, 6))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            VoxelGridTorch(evo.VoxelFields((4, 4, 4)).grid_info(), device="cuda")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "evoxels_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(base, f)).read()
                for needle in ("import oracle", "from oracle", "evx_oracle", "oracle/", "ref_shim"):
                    assert needle not in text, f"{f} reaches for the oracle ({needle})"
                assert "/root/reference" not in text


def test_bc_normalisation_matches_reference_rules():
    with pytest.warns(UserWarning, match="reduces the spatial order of convergence to 0.5"):
        bc = normalize_bc((("dirichlet", (1, -1)), "periodic", "periodic"))
    assert bc == (("dirichlet", (1, -1)), ("periodic", None), ("periodic", None))
    assert normalize_bc(None) == (("periodic", None),) * 3
    assert normalize_bc("fully_periodic") == (("periodic", None),) * 3
    for bad, msg in [(("periodic", "periodic"), "exactly three"),
                     (("dirichlet", "periodic", "periodic"), "require explicit values"),
                     (("robin", "periodic", "periodic"), "Unsupported BC type"),
                     ((("dirichlet", (1,)), "periodic", "periodic"), "two boundary values"),
                     ((("neumann", (1, 2)), "periodic", "periodic"), "do not accept"),
                     ((("periodic", None, 3), "periodic", "periodic"), "either a string or")]:
        with pytest.raises(ValueError, match=msg):
            normalize_bc(bad)
    vg = host_grid()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob = ReactionDiffusion(vg, D=1.0, bc=(("dirichlet", (1, -1)), "periodic", "periodic"))
    assert prob.bc_type == ("dirichlet", "periodic", "periodic")
    assert CahnHilliard(vg).bc_type == ("periodic",) * 3
    assert TwoPhaseAllenCahn(vg).bc_type == ("neumann",) * 3


def test_fft_steppers_reject_unsupported_bc_layouts():
    vg = host_grid()
    with pytest.raises(ValueError, match="support at most one non-periodic axis"):
        PseudoSpectralIMEX(TwoPhaseAllenCahn(vg), 0.1)
    with pytest.raises(NotImplementedError, match="only implement the single non-periodic axis"):
        PseudoSpectralIMEX(CahnHilliard(vg, bc=("periodic", "neumann", "periodic")), 0.1)
    ts = PseudoSpectralIMEX(CahnHilliard(vg, bc=("neumann", "periodic", "periodic")), 0.1)
    assert ts.order == 1 and ForwardEuler(None, 0.1).order == 1 and RungeKutta4(None, 0.1).order == 4


def test_spectral_forms_and_symbols():
    vg = VoxelGridTorch(evo.VoxelFields((6, 5, 4), (3.0, 5.0, 2.0)).grid_info(), device="cpu")
    ch = CahnHilliard(vg, eps=2.0, D=1.5, A=0.25)
    coef, power = ch.spectral_form()
    assert (coef, power) == (2 * 2.0 * 1.5 * 0.25, 2)
    k2 = vg.rfft_k_squared()
    assert k2.dtype == torch.float32 and tuple(k2.shape) == (6, 5, 3)
    assert torch.allclose(ch.fourier_symbol, -coef * k2 ** 2)
    ac = TwoPhaseAllenCahn(vg, gab=0.5, M=2.0, bc=("periodic",) * 3)
    assert torch.allclose(ac.fourier_symbol, -1.0 * k2)
    assert tuple(vg.rfft_k_squared_nonperiodic().shape) == (12, 5, 3)


def test_voxelfields_data_model():
    vf = evo.VoxelFields((10, 5, 7), (10, 5, 7))
    assert vf.shape == (10, 5, 7) and (vf.Nx, vf.Ny, vf.Nz) == (10, 5, 7)
    assert vf.spacing == (1, 1, 1) and vf.precision == "float32"
    x, y, z = vf.meshgrid()
    assert x[-1, 0, 0] == 10 - vf.spacing[0] / 2
    vs = evo.VoxelFields((11, 5, 7), (10, 5, 7), convention="staggered_x")
    xs, ys, _ = vs.meshgrid()
    assert xs[-1, 0, 0] == 10 and ys[0, -1, 0] == 5 - vs.spacing[1] / 2
    vf.add_field("c", 0.123 * np.ones(vf.shape))
    assert vf.fields["c"][1, 2, 3] == 0.123
    with pytest.raises(ValueError):
        vf.add_field("bad", np.ones((2, 2, 2)))
    with pytest.raises(TypeError):
        vf.set_field("bad", [1, 2, 3])
    with pytest.raises(ValueError):
        evo.VoxelFields((1, 2))
    with pytest.raises(ValueError):
        evo.VoxelFields((2, 2, 2), convention="nope")
    with pytest.raises(ValueError):
        evo.VoxelFields((3, 3, 3)).export_to_vtk("bad_name")
    sp = evo.VoxelFields((6, 5, 5), convention="staggered_x")
    sp.add_field("sphere")
    sp.set_voxel_sphere("sphere", center=(0.5, 0.5, 0.5), radius=0.31, label=1)
    assert np.count_nonzero(sp.fields["sphere"] == 1) == 20
    assert sp.average("sphere") == 0.16
    g = vf.grid_info()
    assert (g.shape, g.convention) == ((10, 5, 7), "cell_center")


def test_solver_rejects_other_backends():
    from evoxels_b200.solvers import TimeDependentSolver
    vf = evo.VoxelFields((4, 4, 4))
    vf.add_field("a")
    with pytest.raises(ValueError, match="Unsupported backend"):
        TimeDependentSolver(vf, "a", backend="jax", step_fn=lambda t, u: u, device="cpu")


def test_kernel_wrappers_refuse_inputs_that_require_grad():
    """The ctypes kernels record no autograd node: with grad mode on they must raise instead
    of returning a detached result (round-1 advisor finding)."""
    x = torch.zeros(2, requires_grad=True)
    with pytest.raises(NotImplementedError, match="without a backward pass"):
        _native.refuse_grad("probe", None, x)
    with torch.no_grad():
        _native.refuse_grad("probe", x)
    _native.refuse_grad("probe", torch.zeros(2), None)


def test_graph_replay_and_fused_paths_are_gated_on_the_stock_classes():
    from evoxels_b200.problem_definition import CoupledReactionDiffusion, is_stock
    from evoxels_b200.timesteppers import _symbol_matches_form
    vg = host_grid()
    assert CahnHilliard(vg).autonomous and TwoPhaseAllenCahn(vg).autonomous
    assert CoupledReactionDiffusion(vg).autonomous
    assert ReactionDiffusion(vg, D=1.0).autonomous
    assert not ReactionDiffusion(vg, D=1.0, f=lambda t, c, lib: t * c).autonomous

    class WithSource(CahnHilliard):
        def rhs(self, t, c):
            return super().rhs(t, c) + 1.0

    class OtherSymbol(CahnHilliard):
        @property
        def fourier_symbol(self):
            return -self.k_squared()

    assert is_stock(CahnHilliard(vg), CahnHilliard)
    assert not is_stock(WithSource(vg), CahnHilliard)
    assert not is_stock(OtherSymbol(vg), CahnHilliard)
    assert _symbol_matches_form(CahnHilliard(vg)) and _symbol_matches_form(WithSource(vg))
    assert not _symbol_matches_form(OtherSymbol(vg))


def test_etd1_and_imex_weights_do_not_share_a_cache():
    from evoxels_b200.timesteppers import ExponentialEuler
    vg = host_grid((8, 8, 8))
    ts = ExponentialEuler(ReactionDiffusion(vg, D=1.0), 0.5)
    a = ts.phi_1_k_squared
    b = ts._fft_prefac
    assert not torch.equal(a, b)
    assert torch.equal(ts.phi_1_k_squared, a) and torch.equal(ts._fft_prefac, b)


def test_pipeline_slice_bounds(monkeypatch):
    """Slice bounds of the pipelined distributed stages (CudaOps._split): equal parts by default,
    weights from the environment (any positive numbers), always a partition of the range with
    non-empty slices; malformed or over-fine specifications fall back to equal parts."""
    from evoxels_b200.distributed import CudaOps
    monkeypatch.delenv("EVX_T_SPLIT", raising=False)
    assert CudaOps._split(512, 4, "EVX_T_SPLIT") == [0, 128, 256, 384, 512]
    assert CudaOps._split(512, 4, "EVX_T_SPLIT", "0.12,0.38,0.38,0.12") == [0, 61, 256, 451, 512]
    monkeypatch.setenv("EVX_T_SPLIT", "1,2,1")
    assert CudaOps._split(128, 4, "EVX_T_SPLIT") == [0, 32, 96, 128]
    for spec, n in (("3,1", 100), ("1,1,1,1,1", 47), ("0.5,0.25,0.25", 128)):
        monkeypatch.setenv("EVX_T_SPLIT", spec)
        b = CudaOps._split(n, 4, "EVX_T_SPLIT")
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:]))
    monkeypatch.setenv("EVX_T_SPLIT", "1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1")       # 17 slices of 64 planes: too fine
    assert CudaOps._split(64, 4, "EVX_T_SPLIT") == [0, 16, 32, 48, 64]
    monkeypatch.setenv("EVX_T_SPLIT", "1,-1")
    assert CudaOps._split(64, 2, "EVX_T_SPLIT") == [0, 32, 64]
