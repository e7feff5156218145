"""Grid description and the torch/CUDA backend adapter.

Mirrors the reference's `Grid`, `VoxelGrid` and `VoxelGridTorch`
(evoxels/voxelgrid.py:9-245) - same attributes (`shape, origin, spacing, convention, lib,
div_dx, div_dx2, bc, device, precision`) and method names - so problem definitions and
time steppers written against evoxels run unchanged.  Differences, all deliberate:

* CUDA only.  The reference silently falls back to CPU when CUDA is missing
  (voxelgrid.py:172-177); here a CUDA device that is not available raises, and every
  hot-path operator refuses CPU tensors.  A grid *object* can still be built with
  device='cpu' for host-side bookkeeping (wavenumber tables, conversions).
* No JAX backend (`VoxelGridJax`, voxelgrid.py:248-321) - out of scope.
* Stencils and ghost padding run as hand-written kernels (libevx_b200.so) instead of
  strided-slice tensor arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Tuple

import numpy as np
import torch

from . import _native
from .boundary_conditions import CellCenteredBCs, StaggeredXBCs
from .fd_stencils import FDStencils


@dataclass
class Grid:
    shape: Tuple[int, int, int]
    origin: Tuple[float, float, float]
    spacing: Tuple[float, float, float]
    convention: str


class VoxelGrid(FDStencils):
    """Backend-independent part: geometry, wavenumbers, field import/export."""

    def __init__(self, grid: Grid, lib):
        self.shape = tuple(grid.shape)
        self.origin = tuple(grid.origin)
        self.spacing = tuple(grid.spacing)
        self.convention = grid.convention
        self.lib = lib
        h = self.to_backend(np.array(self.spacing))
        self.div_dx = 1 / h
        self.div_dx2 = 1 / h ** 2
        self.bc = {"cell_center": CellCenteredBCs, "staggered_x": StaggeredXBCs}[self.convention](self)

    # coordinates --------------------------------------------------------------------
    def axes(self):
        return tuple(self.lib.arange(0, n) * h + o
                     for n, h, o in zip(self.shape, self.spacing, self.origin))

    def meshgrid(self):
        return tuple(self.lib.meshgrid(*self.axes(), indexing="ij"))

    # wavenumbers: float32 regardless of `precision`, like the reference (no dtype is
    # passed to fftfreq in voxelgrid.py:84-90)
    def fft_axes(self):
        return tuple(2 * self.lib.pi * self.lib.fft.fftfreq(n, h)
                     for n, h in zip(self.shape, self.spacing))

    def rfft_axes(self):
        return tuple(2 * self.lib.pi * self.lib.fft.rfftfreq(n, h)
                     for n, h in zip(self.shape, self.spacing))

    def fft_mesh(self):
        return tuple(self.lib.meshgrid(*self.fft_axes(), indexing="ij"))

    @staticmethod
    def _sum_of_squares(lib, kx, ky, kz):
        KX, KY, KZ = lib.meshgrid(kx, ky, kz, indexing="ij")
        return KX ** 2 + KY ** 2 + KZ ** 2

    def fft_k_squared(self):
        return self._sum_of_squares(self.lib, *self.fft_axes())

    def rfft_k_squared(self):
        kx, ky, _ = self.fft_axes()
        return self._sum_of_squares(self.lib, kx, ky, self.rfft_axes()[2])

    def rfft_k_squared_nonperiodic(self):
        """|k|^2 for the x-mirrored extension used with Neumann/Dirichlet in x
        (reference voxelgrid.py:116-124)."""
        n_ext = 2 * self.shape[0] if self.convention == "cell_center" else 2 * self.shape[0] - 2
        kx = 2 * self.lib.pi * self.lib.fft.fftfreq(n_ext, d=self.spacing[0])
        return self._sum_of_squares(self.lib, kx, self.fft_axes()[1], self.rfft_axes()[2])

    # fields -------------------------------------------------------------------------
    def init_scalar_field(self, array):
        return self.expand_dim(self.to_backend(array), 0)

    def export_scalar_field_to_numpy(self, field):
        return self.to_numpy(self.squeeze(field, 0))

    def average(self, field):
        if tuple(field.shape[1:]) != self.shape:
            raise ValueError(f"The provided field must have the shape {self.shape}.")
        if self.convention == "cell_center":
            return self.lib.mean(field, (1, 2, 3))
        inner = self.lib.sum(field[:, 1:-1], (1, 2, 3))
        ends = 0.5 * (self.lib.sum(field[:, 0], (1, 2)) + self.lib.sum(field[:, -1], (1, 2)))
        return (inner + ends) / ((self.shape[0] - 1) * self.shape[1] * self.shape[2])


class VoxelGridTorch(VoxelGrid):
    """`distributed`: under an initialised `torch.distributed` process group with more than
    one rank the grid is x-slab decomposed: `init_scalar_field` hands out this
    rank's slab of the (replicated) host array, `export_scalar_field_to_numpy` gathers the
    global field on every rank, `shape` stays the GLOBAL shape, and the stock steppers route
    `step(t, u_slab)` to the distributed kernels (timesteppers.py).  "auto" (default) does
    that whenever a group exists and its size divides nx and ny; False keeps every rank on
    its own full copy; True insists (raises where the decomposition does not fit)."""

    def __init__(self, grid: Grid, precision="float32", device: str = "cuda", distributed="auto"):
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type == "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError(
                    "evoxels_b200 needs a CUDA device: there is no CPU fallback "
                    "(the reference would warn and continue on CPU, voxelgrid.py:172-177).")
            _native.load_library()      # fail now, not at the first step
        if precision not in ("float32", "float64"):
            raise ValueError(f"precision must be 'float32' or 'float64', got {precision!r}")
        self.precision = getattr(torch, precision)
        # the reference sets the global default device (voxelgrid.py:178); keep that so
        # user callbacks creating tensors (custom mu_hom, forcing terms) land on the GPU
        torch.set_default_device(self.device)
        super().__init__(grid, torch)
        self.slab = None
        self.group = None
        if distributed not in (False, None):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                from .distributed import Slab
                try:
                    self.slab = Slab(self.shape, dist.get_world_size(), dist.get_rank())
                except ValueError:
                    if distributed is True:
                        raise
                    import warnings
                    warnings.warn(f"grid {self.shape} does not decompose over {dist.get_world_size()} "
                                  "ranks; every rank keeps the whole grid")
            elif distributed is True:
                raise RuntimeError("distributed=True needs an initialised process group with > 1 rank")

    # fields: slab of a replicated host array in, gathered global field out ------------------
    def init_scalar_field(self, array):
        if self.slab is not None:
            array = np.ascontiguousarray(self.slab.take(np.asarray(array)))
        return super().init_scalar_field(array)

    def gather_slabs(self, field):
        """[C, nx/W, ny, nz] on every rank -> the global [C, nx, ny, nz] on every rank."""
        if self.slab is None:
            return field
        import torch.distributed as dist
        field = field.contiguous()
        parts = [torch.empty_like(field) for _ in range(self.slab.world)]
        dist.all_gather(parts, field)
        return torch.cat(parts, dim=1)

    def export_scalar_field_to_numpy(self, field):
        if self.slab is not None and field.shape[1] == self.slab.nxl and self.slab.nxl != self.shape[0]:
            field = self.gather_slabs(field)
        return super().export_scalar_field_to_numpy(field)

    # conversions ----------------------------------------------------------------------
    def to_backend(self, np_arr):
        return torch.tensor(np_arr, dtype=self.precision, device=self.device)

    def to_numpy(self, field):
        return field.detach().cpu().numpy()

    # ghost layers (kernels) -------------------------------------------------------------
    _PERIODIC3 = (("periodic", None),) * 3

    def pad_with_rules(self, field, bc):
        """[C,Nx,Ny,Nz] -> [C,Nx+2,Ny+2,Nz+2] with ghost rules `bc` (one kernel/channel)."""
        _native.require_cuda(field)
        field = field.contiguous()
        return torch.stack([_native.pad_ghost(ch, bc) for ch in field], 0)

    def pad_periodic(self, field):
        return self.pad_with_rules(field, self._PERIODIC3)

    def pad_zeros(self, field):
        _native.require_cuda(field)
        return torch.nn.functional.pad(field, (1, 1, 1, 1, 1, 1), mode="constant", value=0)

    # transforms (cuFFT through torch; the IMEX stepper uses an ImexPlan instead) ---------
    def fftn(self, field, shape):
        _native.require_cuda(field)
        return torch.fft.fftn(field, s=shape)

    def rfftn(self, field, shape):
        _native.require_cuda(field)
        return torch.fft.rfftn(field, s=shape)

    def irfftn(self, field, shape):
        _native.require_cuda(field)
        return torch.fft.irfftn(field, s=shape)

    def real_of_ifftn(self, field, shape):
        _native.require_cuda(field)
        return torch.real(torch.fft.ifftn(field, s=shape))

    # small tensor helpers kept for API parity ---------------------------------------------
    def expand_dim(self, field, dim):
        return field.unsqueeze(dim)

    def squeeze(self, field, dim):
        return torch.squeeze(field, dim)

    def concatenate(self, fieldlist, dim):
        return torch.cat(fieldlist, dim=dim)

    def argmax(self, field, dim=None, keepdim=False):
        return torch.argmax(field, dim=dim, keepdim=keepdim)

    def mean(self, field, dim=None, keepdim=False):
        return torch.mean(field, dim=dim, keepdim=keepdim)

    def sum(self, field, dim=None, keepdim=False):
        return torch.sum(field, dim=dim, keepdim=keepdim)

    def cumsum(self, field, dim=None):
        return torch.cumsum(field, dim=dim)

    def sort(self, field, dim=0, descending=False):
        return torch.sort(field, dim=dim, descending=descending)[0]

    def arange(self, start, stop):
        return torch.arange(start, stop, dtype=self.precision, device=self.device)

    def take_along_dim(self, field, idx, dim=0):
        return torch.take_along_dim(field, idx, dim=dim)

    def set(self, field, index, value):
        field[index] = value
        return field
