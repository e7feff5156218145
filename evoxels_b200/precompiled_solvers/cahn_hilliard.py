"""`run_cahn_hilliard_solver` with the reference's signature
(evoxels/precompiled_solvers/cahn_hilliard.py:6-42)."""
from typing import Callable

from ..problem_definition import CahnHilliard
from ..solvers import TimeDependentSolver
from ..timesteppers import PseudoSpectralIMEX


def run_cahn_hilliard_solver(voxelfields, fieldnames, backend: str = "torch", jit: bool = True,
                             device: str = "cuda", time_increment: float = 0.1,
                             frames: int = 10, max_iters: int = 100, eps: float = 3.0,
                             diffusivity: float = 1.0, mu_hom: Callable | None = None,
                             vtk_out: bool = False, verbose: bool = True, plot_bounds=None):
    """Cahn-Hilliard with the semi-implicit pseudo-spectral stepper on CUDA."""
    solver = TimeDependentSolver(voxelfields, fieldnames, backend, problem_cls=CahnHilliard,
                                 timestepper_cls=PseudoSpectralIMEX, device=device)
    solver.solve(time_increment=time_increment, frames=frames, max_iters=max_iters,
                 problem_kwargs=dict(eps=eps, D=diffusivity, mu_hom=mu_hom), jit=jit,
                 verbose=verbose, vtk_out=vtk_out, plot_bounds=plot_bounds)
    return solver
