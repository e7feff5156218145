"""Finite-difference operators on ghost-padded fields (mixin of `VoxelGrid`).

Same method names and semantics as the reference's `FDStencils`
(evoxels/fd_stencils.py:13-103).  On the fused hot path (CahnHilliard.rhs,
TwoPhaseAllenCahn.rhs) none of these are called - the kernels do the whole right-hand side
in one pass without a padded copy.  They exist for user-defined problems:

* `laplace`, `normal_laplace`, `gradient_norm_squared` run the hand-written
  `evx_padded_stencil_*` kernel (one launch instead of ~8 / ~60 tensor ops);
* the one-line face/centre differences are thin tensor expressions on CUDA tensors.
"""
from __future__ import annotations

import torch

from . import _native


def _pair(field, axis):
    """(upper, lower) neighbours along spatial `axis` of a [C,X,Y,Z] tensor."""
    dim = axis + 1
    n = field.shape[dim]
    return field.narrow(dim, 1, n - 1), field.narrow(dim, 0, n - 1)


def _interior_pair(field, axis):
    """(plus, minus) neighbours of the interior cells along `axis` of a padded field."""
    core = field[:, 1:-1, 1:-1, 1:-1]
    nx, ny, nz = core.shape[1:]
    lo = [1, 1, 1]
    lo[axis] = 0
    hi = [1, 1, 1]
    hi[axis] = 2
    cut = lambda o: field[:, o[0]:o[0] + nx, o[1]:o[1] + ny, o[2]:o[2] + nz]  # noqa: E731
    return cut(hi), cut(lo)


class FDStencils:
    # ---- cell <-> face ---------------------------------------------------------------
    def _to_face(self, field, axis):
        _native.require_cuda(field)
        up, lo = _pair(field, axis)
        return 0.5 * (up + lo)

    def _grad_face(self, field, axis):
        _native.require_cuda(field)
        up, lo = _pair(field, axis)
        return (up - lo) * self.div_dx[axis]

    def _grad_center(self, field, axis):
        _native.require_cuda(field)
        plus, minus = _interior_pair(field, axis)
        return 0.5 * (plus - minus) * self.div_dx[axis]

    def to_x_face(self, field):
        return self._to_face(field, 0)

    def to_y_face(self, field):
        return self._to_face(field, 1)

    def to_z_face(self, field):
        return self._to_face(field, 2)

    def grad_x_face(self, field):
        return self._grad_face(field, 0)

    def grad_y_face(self, field):
        return self._grad_face(field, 1)

    def grad_z_face(self, field):
        return self._grad_face(field, 2)

    def grad_x_center(self, field):
        return self._grad_center(field, 0)

    def grad_y_center(self, field):
        return self._grad_center(field, 1)

    def grad_z_center(self, field):
        return self._grad_center(field, 2)

    # ---- kernels on padded fields ----------------------------------------------------
    def _padded_op(self, field, op):
        _native.require_cuda(field)
        field = field.contiguous()
        return torch.stack([_native.padded_stencil(ch, self.spacing, op) for ch in field], 0)

    def laplace(self, field):
        """7-point Laplacian of a ghost-padded field -> interior."""
        return self._padded_op(field, 0)

    def normal_laplace(self, field):
        """Second derivative along grad(field)/|grad(field)| (19-point) -> interior."""
        return self._padded_op(field, 1)

    def gradient_norm_squared(self, field):
        return self._padded_op(field, 2)
