"""Ghost-layer rules for cell-centred grids.

Same entry points as the reference's `CellCenteredBCs`
(evoxels/boundary_conditions.py:5-82).  Every `pad_*` is one launch of the
`evx_pad_ghost_*` kernel, which evaluates the reference's rule set (circular pad, then
Neumann `ghost = inner` / Dirichlet `ghost = 2 v - inner` applied axis by axis in the
order x, y, z over the whole padded plane) by index arithmetic.  The fused right-hand-side
kernels never call these - they exist for user code and for output handling.

`StaggeredXBCs` (boundary_conditions.py:85-156) belongs to the `staggered_x` convention,
which no configuration of the accelerated path uses; it is declared so that grids can be
described, and raises on use.
"""
from __future__ import annotations

import torch

_P = ("periodic", None)


class CellCenteredBCs:
    def __init__(self, vg):
        self.vg = vg

    # ghost padding ----------------------------------------------------------------------
    def pad_bc(self, field, bc):
        for kind, _ in bc:
            if kind not in ("periodic", "neumann", "dirichlet"):
                raise ValueError(f"Unsupported BC type: {kind}")
        return self.vg.pad_with_rules(field, tuple(bc))

    def pad_periodic(self, field):
        return self.vg.pad_with_rules(field, (_P, _P, _P))

    def pad_dirichlet_periodic(self, field, bc0=0, bc1=0):
        return self.vg.pad_with_rules(field, (("dirichlet", (bc0, bc1)), _P, _P))

    def pad_zero_flux_periodic(self, field):
        return self.vg.pad_with_rules(field, (("neumann", None), _P, _P))

    # mirror extensions in x for FFT-based steppers ------------------------------------------
    def pad_fft_periodic(self, field):
        return field

    def pad_fft_dirichlet_periodic(self, field):
        return torch.cat((field, -torch.flip(field, [1])), 1)

    def pad_fft_zero_flux_periodic(self, field):
        return torch.cat((field, torch.flip(field, [1])), 1)

    # trimming -------------------------------------------------------------------------------
    def trim_boundary_nodes(self, field):
        return field

    def trim_ghost_nodes(self, field):
        inner = field[:, 1:-1, 1:-1, 1:-1]
        if tuple(inner.shape[1:]) != tuple(self.vg.shape):
            raise ValueError(f"The provided field has the wrong shape {self.vg.shape}.")
        return inner


class StaggeredXBCs:
    def __init__(self, vg):
        self.vg = vg

    def _unsupported(self, *_, **__):
        raise NotImplementedError(
            "the staggered_x convention is outside the accelerated hot path of evoxels_b200")

    pad_bc = pad_periodic = pad_dirichlet_periodic = pad_zero_flux_periodic = _unsupported
    pad_fft_periodic = pad_fft_dirichlet_periodic = pad_fft_zero_flux_periodic = _unsupported
    trim_ghost_nodes = _unsupported

    def trim_boundary_nodes(self, field):
        if field.shape[1] != self.vg.shape[0]:
            raise ValueError(f"The provided field must have the shape {self.vg.shape}.")
        return field[:, 1:-1]
