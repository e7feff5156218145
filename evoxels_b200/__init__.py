"""evoxels_b200 - B200-native drop-in for the per-timestep hot path of daubners/evoxels.

Public names follow evoxels/__init__.py:3-13 (InversionModel is JAX-only upstream; the
differentiable step lives in `evoxels_b200.autograd`).
"""
from .voxelfields import VoxelFields
from .precompiled_solvers.cahn_hilliard import run_cahn_hilliard_solver
from .precompiled_solvers.allen_cahn import run_allen_cahn_solver

__all__ = ["VoxelFields", "run_cahn_hilliard_solver", "run_allen_cahn_solver"]
__version__ = "0.1.0"
