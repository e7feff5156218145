"""Host-side NumPy container of named 3-D fields on a uniform voxel grid.

Mirrors the data-model part of the reference's `VoxelFields`
(evoxels/voxelfields.py:65-194: constructor, spacing/origin rules, `grid_info`,
`set_field`/`add_field`, `set_voxel_sphere`, `average`, `axes`/`meshgrid`) so that user
scripts written against evoxels keep working.  Plotting (voxelfields.py:216-325) is host
visualisation outside the hot path and is not provided; VTK export is kept because
`solve(vtk_out=True)` calls it (it needs pyvista, imported lazily).
"""
from __future__ import annotations

import warnings
from typing import Tuple

import numpy as np

from .voxelgrid import Grid

_CONVENTIONS = ("cell_center", "staggered_x")


class VoxelFields:
    def __init__(self, shape: Tuple[int, int, int], domain_size=(1, 1, 1),
                 convention: str = "cell_center"):
        ok_shape = (isinstance(shape, (list, tuple)) and len(shape) == 3
                    and all(isinstance(n, (int, np.integer)) for n in shape))
        if not ok_shape:
            raise ValueError("shape must be a tuple of three integers")
        if not isinstance(domain_size, (list, tuple)) or len(domain_size) != 3:
            raise ValueError("domain_size must be a list or tuple with three elements (dx, dy, dz)")
        if not all(isinstance(length, (int, float)) for length in domain_size):
            raise ValueError("All elements in domain_size must be integers or floats")
        if convention not in _CONVENTIONS:
            raise ValueError("Chosen convention must be cell_center or staggered_x.")

        self.shape = tuple(int(n) for n in shape)
        self.domain_size = domain_size
        self.convention = convention
        self.precision = "float32"
        # cell-centred: N cells of width L/N, first centre at h/2.
        # staggered_x: nodes on both x boundaries, so N-1 intervals and origin 0 in x.
        cells = list(self.shape)
        if convention == "staggered_x":
            cells[0] -= 1
        self.spacing = tuple(length / n for length, n in zip(domain_size, cells))
        self.origin = tuple(0 if (convention == "staggered_x" and a == 0) else h / 2
                            for a, h in enumerate(self.spacing))
        if max(self.spacing) / min(self.spacing) > 10:
            warnings.warn("Simulations become very questionable for largely different "
                          "spacings e.g. dz >> dx.")
        self.grid = None
        self.fields = {}

    Nx = property(lambda self: self.shape[0])
    Ny = property(lambda self: self.shape[1])
    Nz = property(lambda self: self.shape[2])

    def __str__(self):
        return (f"Domain with size {self.domain_size} and {self.shape} grid points on "
                f"{self.convention} position.")

    def grid_info(self) -> Grid:
        return Grid(self.shape, self.origin, self.spacing, self.convention)

    # ---- fields --------------------------------------------------------------------
    def set_field(self, name: str, array: np.ndarray):
        if not isinstance(array, np.ndarray):
            raise TypeError("The provided array must be a numpy array.")
        if array.shape != self.shape:
            raise ValueError(f"The provided array must have the shape {self.shape}.")
        self.fields[name] = array

    def add_field(self, name: str, array=None):
        self.set_field(name, np.zeros(self.shape) if array is None else array)

    def set_voxel_sphere(self, name: str, center, radius, label: int | float = 1):
        """Label all voxels whose centre lies within `radius` of `center`."""
        idx = np.ogrid[:self.Nx, :self.Ny, :self.Nz]
        dist2 = sum((i * h + o - c) ** 2
                    for i, h, o, c in zip(idx, self.spacing, self.origin, center))
        self.fields[name][dist2 <= radius ** 2] = label

    def average(self, name: str):
        f = self.fields[name]
        if self.convention == "cell_center":
            return np.mean(f)
        # staggered_x: boundary planes are half cells
        total = np.sum(f[1:-1]) + 0.5 * np.sum(f[0]) + 0.5 * np.sum(f[-1])
        return total / ((self.Nx - 1) * self.Ny * self.Nz)

    # ---- coordinates ---------------------------------------------------------------
    def axes(self):
        return tuple(np.arange(0, n, dtype=self.precision) * h + o
                     for n, h, o in zip(self.shape, self.spacing, self.origin))

    def meshgrid(self):
        return tuple(np.meshgrid(*self.axes(), indexing="ij"))

    # ---- output --------------------------------------------------------------------
    def export_to_vtk(self, filename="output.vtk", field_names=None):
        """Write the fields as cell data of a VTK image (needs pyvista)."""
        if not str(filename).endswith((".vtk", ".vti")):
            raise ValueError(f"Unsupported VTK file name: {filename!r} (use .vtk or .vti)")
        import pyvista as pv
        image = pv.ImageData()
        image.spacing = self.spacing
        image.dimensions = tuple(n + 1 for n in self.shape)
        image.origin = tuple(o - h / 2 for o, h in zip(self.origin, self.spacing))
        for name in (field_names or list(self.fields)):
            image.cell_data[name] = self.fields[name].flatten(order="F")
        image.save(filename)
