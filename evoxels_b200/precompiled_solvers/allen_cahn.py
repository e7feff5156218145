"""`run_allen_cahn_solver` with the reference's signature
(evoxels/precompiled_solvers/allen_cahn.py:6-50)."""
from typing import Callable

from ..problem_definition import TwoPhaseAllenCahn
from ..solvers import TimeDependentSolver
from ..timesteppers import RungeKutta4


def run_allen_cahn_solver(voxelfields, fieldnames, backend: str = "torch", jit: bool = True,
                          device: str = "cuda", time_increment: float = 0.5, frames: int = 10,
                          max_iters: int = 100, eps: float = 2.0, gab: float = 1.0,
                          M: float = 1.0, force: float = 0.0, curvature: float = 0.01,
                          potential: Callable | None = None, vtk_out: bool = False,
                          verbose: bool = True, plot_bounds=None):
    """Two-phase Allen-Cahn with classical RK4 (fused stage kernels) on CUDA."""
    solver = TimeDependentSolver(voxelfields, fieldnames, backend,
                                 problem_cls=TwoPhaseAllenCahn, timestepper_cls=RungeKutta4,
                                 device=device)
    solver.solve(time_increment=time_increment, frames=frames, max_iters=max_iters,
                 problem_kwargs=dict(eps=eps, gab=gab, M=M, force=force, curvature=curvature,
                                     potential=potential),
                 jit=jit, verbose=verbose, vtk_out=vtk_out, plot_bounds=plot_bounds)
    return solver
