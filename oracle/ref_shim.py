"""Import the *unmodified* reference (daubners/evoxels, mounted at /root/reference)
on a box that has neither matplotlib nor jax/diffrax.

TEST INFRASTRUCTURE ONLY.  Used by `tests/golden/make_golden.py` (fixture generation,
in the build container) and by `tests/test_oracle_vs_reference.py` (skipped when
/root/reference is absent, e.g. on the GPU box).  Nothing in the product imports this.

Why a shim is needed (reference file:line):
  * evoxels/voxelfields.py:21-22 and evoxels/solvers.py:8 import matplotlib at module top
  * evoxels/inversion.py:51 evaluates `dfx.ForwardMode()` as a default argument at import
    time, so `evoxels/__init__.py:6` raises without diffrax.
The shim registers stub matplotlib modules and a synthetic `evoxels` package whose
__path__ points at the reference sources, so `evoxels/__init__.py` is never executed.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("EVOXELS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "evoxels"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    """Return a namespace with the reference submodules needed for the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.widgets = _stub("matplotlib.widgets", Slider=object)
        mpl.patches = _stub("matplotlib.patches", Patch=object)
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=object)
    if "evoxels" not in sys.modules or not hasattr(sys.modules["evoxels"], "__evx_shim__"):
        pkg = types.ModuleType("evoxels")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "evoxels")]
        pkg.__evx_shim__ = True
        sys.modules["evoxels"] = pkg
    ns = types.SimpleNamespace()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for sub in ("voxelgrid", "voxelfields", "problem_definition", "timesteppers",
                    "solvers", "precompiled_solvers.cahn_hilliard",
                    "precompiled_solvers.allen_cahn"):
            mod = importlib.import_module("evoxels." + sub)
            setattr(ns, sub.split(".")[-1], mod)
    pkg = sys.modules["evoxels"]
    pkg.VoxelFields = ns.voxelfields.VoxelFields
    pkg.run_cahn_hilliard_solver = ns.cahn_hilliard.run_cahn_hilliard_solver
    pkg.run_allen_cahn_solver = ns.allen_cahn.run_allen_cahn_solver
    return ns
