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


class _FakeClock:
    now = 0.0


class _FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = _FakeClock.now

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return other.t - self.t


def _fake_plan(monkeypatch, cost, wrong=()):
    """An ImexPlan whose ch_step is a stand-in: it costs cost(schedule) fake milliseconds and
    writes a schedule-independent field, except for the schedules in `wrong`."""
    plan = object.__new__(_native.ImexPlan)
    plan._handle = None
    plan.shape = (512, 512, 512)
    plan.backend = _native.FFT_NATIVE
    plan.device = torch.device("cpu")
    plan.tuned = False
    plan.tune_report = None
    plan.current = (0, 1, 0)
    plan.log = []

    def set_schedule(chunk_planes=0, streams=1, flags=0):
        plan.current = (chunk_planes, streams, flags)
        plan.log.append(plan.current)

    def ch_step(u, out, spacing, dt, eps, D, A, hom=None):
        _FakeClock.now += cost(plan.current)
        out.copy_(u * 2 + (1 if plan.current in wrong else 0))
        return out

    plan.set_schedule, plan.ch_step, plan.schedule = set_schedule, ch_step, lambda: plan.current
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None: (1 << 40, 1 << 40))

    class Props:
        multi_processor_count = 148
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda dev=None: Props)
    return plan


def test_schedule_tuner_picks_the_fastest_bit_identical_candidate(monkeypatch):
    u = torch.rand(4, 4, 4)
    R, C = _native.SCHED_RING_INV, _native.SCHED_CHUNK_RHS
    fast, faster_but_wrong = (17, 2, 0), (26, 2, R)
    cost = lambda s: {fast: 1.2, (17, 2, R): 1.3, faster_but_wrong: 0.9, (0, 1, 0): 1.8}.get(s, 1.5 + 0.01 * s[0])
    plan = _fake_plan(monkeypatch, cost, wrong={faster_but_wrong})
    assert plan.schedule_sizes() == [8, 9, 11, 12, 16, 17, 18, 23, 24, 26, 27, 32]
    best, report = plan.tune_ch_step(u, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)
    # phase 1 finds size 17 through its ring variant, phase 2 then finds the faster no-ring form
    assert best == fast and plan.current == fast and plan.tuned
    assert report["chosen"] == fast and abs(report["baseline_ms"] - 1.8) < 1e-9
    rejected = [c for c in report["candidates"] if not c["bit_identical"]]
    assert [c["schedule"] for c in rejected] == [faster_but_wrong]
    tried = {c["schedule"] for c in report["candidates"]}
    assert (17, 1, R | C) in tried and (17, 3, R | C) in tried and (26, 2, 0) not in tried   # phase 2 skips rejected sizes
    # a gain below the threshold keeps the one-launch-per-pass schedule
    plan = _fake_plan(monkeypatch, lambda s: 1.0 if s == (0, 1, 0) else 0.99)
    best, _ = plan.tune_ch_step(u, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)
    assert best == (0, 1, 0) and plan.current == (0, 1, 0)
    # large grids: only chunks whose spectrum fits L2 are candidates; too small grids: none
    plan = _fake_plan(monkeypatch, lambda s: 120.0 if s == (0, 1, 0) else 100.0)
    plan.shape = (2048, 2048, 2048)
    assert plan.schedule_sizes() == [2, 3]
    best, report = plan.tune_ch_step(u, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)
    assert best[0] in (2, 3) and len(report["candidates"]) <= 16
    plan.shape = (1024, 1024, 1024)
    assert plan.schedule_sizes() == [8, 9, 11, 12]
    plan.shape = (4, 64, 64)
    plan.current = (0, 1, 0)
    assert plan.schedule_sizes() == []
    best, report = plan.tune_ch_step(u, (1, 1, 1), 0.1, 3.0, 1.0, 0.25)
    assert best == (0, 1, 0) and report["skipped"] == "grid too small"


def test_stepper_schedule_policy(monkeypatch):
    """Small grids and EVX_TUNE=0 keep the baseline without measuring, EVX_SCHEDULE forces a
    schedule, and a failing tuner falls back to the baseline with a warning."""
    ts = PseudoSpectralIMEX(CahnHilliard(host_grid()), 0.1)
    calls = []

    def tune(*a, **k):
        calls.append(a)
        return (8, 2, 1), {"chosen": (8, 2, 1)}
    plan = _fake_plan(monkeypatch, lambda s: 1.0)
    plan.tune_ch_step = tune
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    small, large = torch.rand(8, 8, 8), torch.empty(0).new_empty((1 << 24,))
    ts._choose_schedule(plan, small, (1, 1, 1))
    assert plan.tuned and not calls and not plan.log
    plan.tuned = False
    monkeypatch.setenv("EVX_TUNE", "0")
    ts._choose_schedule(plan, large, (1, 1, 1))
    assert plan.tuned and not calls
    monkeypatch.delenv("EVX_TUNE")
    plan.tuned = False
    monkeypatch.setenv("EVX_SCHEDULE", "16,2,3")
    ts._choose_schedule(plan, small, (1, 1, 1))
    assert plan.current == (16, 2, 3) and plan.tuned and not calls
    monkeypatch.delenv("EVX_SCHEDULE")
    plan.tuned = False
    ts._choose_schedule(plan, large, (1, 1, 1))
    assert len(calls) == 1 and plan.tune_report == {"chosen": (8, 2, 1)}

    def broken(*a, **k):
        raise RuntimeError("boom")
    plan.tune_ch_step = broken
    plan.tuned = False
    with pytest.warns(UserWarning, match="schedule tuning failed"):
        ts._choose_schedule(plan, large, (1, 1, 1))
    assert plan.tuned and plan.current == (0, 1, 0)
