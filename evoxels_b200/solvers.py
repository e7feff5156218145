"""Time loop and field export around `step(t, u)`.

Interface parity with evoxels/solvers.py: `BaseSolver` (:12-187) and
`TimeDependentSolver` (:189-208) keep the dataclass fields
`(vf, fieldnames, backend, problem_cls, timestepper_cls, step_fn, device)` and
`solve(time_increment, frames, max_iters, problem_kwargs, jit, verbose, vtk_out,
plot_bounds, colormap)`.  `backend` must be 'torch' (no JAX dispatch).  `jit=True` (the
reference's torch.compile switch, solvers.py:64-70) here means: capture the launches of a
few consecutive steps in a CUDA graph and replay it, which removes the per-launch host cost
that dominates small grids (100^3: ~10 launches of a few microseconds each per step).
Live plotting (`verbose='plot'`) is host visualisation and not provided.
"""
from __future__ import annotations

import sys
import warnings
from abc import ABC, abstractmethod
from dataclasses import dataclass
from timeit import default_timer as timer
from typing import Any, Callable, Type

import torch

from .problem_definition import ODE
from .timesteppers import TimeStepper
from .voxelgrid import VoxelGridTorch


@dataclass
class BaseSolver(ABC):
    vf: Any
    fieldnames: str | list[str]
    backend: str
    problem_cls: Type[ODE] | None = None
    timestepper_cls: Type[TimeStepper] | None = None
    step_fn: Callable | None = None
    device: str = "cuda"
    # multi-GPU: under an initialised torch.distributed group the grid is x-slab decomposed and
    # every rank solves its slab (VoxelGridTorch); False keeps each rank on the whole grid
    distributed: Any = "auto"

    def __post_init__(self):
        if self.backend != "torch":
            raise ValueError(f"Unsupported backend: {self.backend} "
                             "(evoxels_b200 drives CUDA through torch only)")
        self.vg = VoxelGridTorch(self.vf.grid_info(), precision=self.vf.precision,
                                 device=self.device, distributed=self.distributed)
        self.fieldnames = [self.fieldnames] if isinstance(self.fieldnames, str) \
            else list(self.fieldnames)
        self.problem = None
        self.computation_time = None

    # ---- set-up ----------------------------------------------------------------------
    def _init_fields(self):
        fields = [self.vg.init_scalar_field(self.vf.fields[n]) for n in self.fieldnames]
        return self.vg.bc.trim_boundary_nodes(self.vg.concatenate(fields, 0))

    def _init_stepper(self, time_increment, problem_kwargs, jit):
        if self.step_fn is not None:
            self.problem = None
            return self.step_fn
        if self.problem_cls is None or self.timestepper_cls is None:
            raise ValueError("Either provide step_fn or both problem_cls and timestepper_cls")
        self.problem = self.problem_cls(self.vg, **(problem_kwargs or {}))
        # A captured graph replays step(0.0, u): legal only if rhs ignores t (ODE.autonomous).
        # Everything else - ReactionDiffusion with a source f(t, u), user problem classes,
        # user step_fn - runs the eager loop with the real time, like the reference.
        # (the multi-GPU step runs on several streams with cross-rank barriers: eager only)
        self._graph_ok = (bool(jit) and self.vg.device.type == "cuda"
                          and getattr(self.problem, "autonomous", False) is True
                          and getattr(self.vg, "slab", None) is None)
        return self.timestepper_cls(self.problem, time_increment).step

    @abstractmethod
    def _run_loop(self, u, step, time_increment, frames, max_iters, vtk_out, verbose,
                  plot_bounds, colormap):
        raise NotImplementedError

    def solve(self, time_increment=0.1, frames=10, max_iters=100, problem_kwargs=None,
              jit=True, verbose=True, vtk_out=False, plot_bounds=None, colormap="viridis"):
        u = self._init_fields()
        step = self._init_stepper(time_increment, problem_kwargs, jit)
        if verbose == "plot":
            warnings.warn("verbose='plot' is not available in evoxels_b200; printing stats only")
        cuda = self.vg.device.type == "cuda"
        if cuda:
            torch.cuda.reset_peak_memory_stats(self.vg.device)
            torch.cuda.synchronize(self.vg.device)
        start = timer()
        u = self._run_loop(u, step, time_increment, frames, max_iters, vtk_out, verbose,
                           plot_bounds, colormap)
        if cuda:
            torch.cuda.synchronize(self.vg.device)
        end = timer()
        self.computation_time = end - start
        if verbose:
            per = self.computation_time / max(max_iters, 1)
            print(f"Wall time: {self.computation_time:.4f} s after {max_iters} iterations "
                  f"({per:.6f} s/iter)")
            if cuda:
                mb = 1024 ** 2
                print(f"GPU-RAM (torch) current: {torch.cuda.memory_allocated(self.vg.device)/mb:.2f} MB "
                      f"({torch.cuda.max_memory_allocated(self.vg.device)/mb:.2f} MB max)")

    # ---- per-frame output --------------------------------------------------------------
    def _export_fields(self, u_out):
        for i, name in enumerate(self.fieldnames):
            self.vf.set_field(name, self.vg.export_scalar_field_to_numpy(u_out[i:i + 1]))

    def _handle_outputs(self, u, frame, time, vtk_out, verbose, plot_bounds, colormap):
        """Frame export (reference solvers.py:152-187).  The reference pads and trims here
        (:154-157), which is the identity on a cell-centred grid, copies the state to the
        host synchronously and checks for NaN on the host side.  Here the copy and a
        device-side NaN reduction run on a side stream into pinned memory while the solver
        keeps stepping; the frame is committed to `vf.fields` - and a
        NaN aborts the run - when the next frame is submitted, at the end of the run, or
        immediately if files are written.  Consequence: while solve() runs, `vf.fields` lags
        one frame behind, and a run that went NaN keeps stepping for up to `max_iters//frames`
        more iterations before it exits."""
        if u.device.type != "cuda":
            raise RuntimeError("evoxels_b200 has no CPU path: the state must be a CUDA tensor")
        if getattr(self, "_exporter", None) is None:
            self._exporter = _FrameExporter(u.device)
        if getattr(self.vg, "slab", None) is not None:
            u = self.vg.gather_slabs(u)      # x-slab run: every rank exports the global field
        self._exporter.submit(u, frame, time, self._commit_frame)
        if vtk_out:
            self._exporter.finish()
            prefix = self.problem_cls.__name__ if self.problem_cls else "custom"
            self.vf.export_to_vtk(filename=f"{prefix}_{self.fieldnames[0]}_{frame:03d}.vtk",
                                  field_names=self.fieldnames)

    def _commit_frame(self, host, has_nan, frame, time):
        # through _export_fields, so that subclasses overriding it (the reference's
        # MultiPhaseSolver pattern) see every frame; `host` is the pinned staging buffer,
        # which is reused - hand out a pageable copy (to_numpy() of a CPU tensor is a view)
        self._export_fields(torch.empty_like(host, pin_memory=False).copy_(host))
        if has_nan:
            print(f"NaN detected in frame {frame} at time {time}. Aborting simulation.")
            sys.exit(1)

    def _finish_outputs(self):
        if getattr(self, "_exporter", None) is not None:
            self._exporter.finish()


class _FrameExporter:
    """One frame in flight: device -> pinned host copy plus isnan().any() on a side stream."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.host = None
        self.flag = torch.zeros(1, dtype=torch.bool, device="cpu", pin_memory=True)
        self.pending = None

    def submit(self, u, frame, time, commit):
        self.finish()                                   # the pinned buffer is free again
        if self.host is None or self.host.shape != u.shape or self.host.dtype != u.dtype:
            self.host = torch.empty(tuple(u.shape), dtype=u.dtype, device="cpu", pin_memory=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.host.copy_(u, non_blocking=True)
            # amin propagates NaN and needs no full-size temporary (isnan(u) would allocate
            # a bool field: 1 GB at 1024^3)
            self.flag.copy_(torch.isnan(torch.amin(u)).reshape(1), non_blocking=True)
            done = self.stream.record_event()
        u.record_stream(self.stream)                    # keep u's memory until the copy is done
        self.pending = (done, frame, time, commit)

    def finish(self):
        if self.pending is None:
            return
        done, frame, time, commit = self.pending
        self.pending = None
        done.synchronize()
        commit(self.host, bool(self.flag[0]), frame, time)


class _GraphedSteps:
    """`k` consecutive calls of step(t, u) captured in one CUDA graph (t is not used by the
    autonomous problems of this package).  replay(u) -> state after k steps."""

    def __init__(self, step, u, k):
        self.k = k
        self.static_in = u.clone()
        side = torch.cuda.Stream(device=u.device)
        side.wait_stream(torch.cuda.current_stream(u.device))
        with torch.cuda.stream(side):                  # warm-up outside capture (plans, cuFFT)
            v = self.static_in
            for _ in range(2):
                v = step(0.0, v)
        torch.cuda.current_stream(u.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            v = self.static_in
            for _ in range(k):
                v = step(0.0, v)
            self.static_out = v

    def replay(self, u):
        self.static_in.copy_(u)
        self.graph.replay()
        return self.static_out.clone()


def _largest_divisor_at_most(n, cap):
    for k in range(min(n, cap), 0, -1):
        if n % k == 0:
            return k
    return 1


@dataclass
class TimeDependentSolver(BaseSolver):
    def _run_loop(self, u, step, time_increment, frames, max_iters, vtk_out, verbose,
                  plot_bounds, colormap):
        every = max_iters // frames          # ZeroDivisionError if frames > max_iters, as upstream
        frame = 0
        graphed = None
        k = _largest_divisor_at_most(every, 25)
        if getattr(self, "_graph_ok", False) and k > 1 and max_iters % every == 0:
            try:
                graphed = _GraphedSteps(step, u, k)
            except Exception as exc:     # capture not possible (e.g. a user callback syncs): run eagerly
                warnings.warn(f"CUDA-graph capture of the step failed ({exc}); running eagerly")
                graphed = None
        i = 0
        while i < max_iters:
            now = i * time_increment
            if i % every == 0:
                self._handle_outputs(u, frame, now, vtk_out, verbose, plot_bounds, colormap)
                frame += 1
            if graphed is not None:
                u = graphed.replay(u)
                i += k
            else:
                u = step(now, u)
                i += 1
        self._handle_outputs(u, frame, max_iters * time_increment, vtk_out, verbose,
                             plot_bounds, colormap)
        self._finish_outputs()
        return u
