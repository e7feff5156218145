"""ctypes binding of libevx_b200.so (the C ABI declared in include/evoxels_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
Tensors are passed as raw device pointers together with the current CUDA stream of the
tensor's device; nothing here allocates except through torch's caching allocator.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libevx_b200.so")

BC_CODES = {"periodic": 0, "neumann": 1, "dirichlet": 2}
FFT_AUTO, FFT_CUFFT, FFT_NATIVE, FFT_NATIVE_MIXED = 0, 1, 2, 3

_c_void_p, _c_int, _c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_dptr = ctypes.POINTER(ctypes.c_double)
_iptr = ctypes.POINTER(ctypes.c_int)

# name -> argtypes; every function returns int unless listed in _RESTYPES
_STENCIL_ARGS = [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _dptr, _c_double,
                 _c_double, _iptr, _dptr, _c_void_p, _c_void_p, _c_void_p]
_AC_ARGS = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_double, _c_void_p,
            _c_void_p, _c_double, _c_int, _c_int, _c_int, _dptr, _c_double, _c_double,
            _c_double, _c_double, _c_double, _iptr, _dptr, _c_void_p, _c_void_p, _c_void_p]
_PAD_ARGS = [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _iptr, _dptr, _c_void_p]
_PST_ARGS = [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _dptr, _c_int, _c_void_p]
_APPLY_ARGS = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _dptr, _c_double,
               _c_double, _c_int, _c_void_p]
_STEP_ARGS = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _dptr, _c_double,
              _c_double, _c_double, _c_double, _c_void_p]
_RD2_ARGS = [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _dptr, _c_double, _c_double,
             _c_double, _c_double, _c_void_p]
_FILTER_ARGS = [_c_void_p, _c_int, _c_int, _c_int, _dptr, _c_double, _c_double, _c_int,
                _c_double, _c_void_p]

FILTER_ETD1 = 0x100      # EVX_FILTER_ETD1: OR into `power` for the exponential-Euler weight
FILTER_MIRROR_EVEN = 0x200   # EVX_FILTER_MIRROR_EVEN / _ODD: non-periodic x axis, the mirror image of
FILTER_MIRROR_ODD = 0x400    # every x line is synthesised inside the x pass (un-extended arrays)
ERR_UNSUPPORTED = -2

SIGNATURES = {
    "evx_version": [],
    "evx_strerror": [_c_int],
    "evx_launch_count": [],
    "evx_ch_rhs_f32": _STENCIL_ARGS, "evx_ch_rhs_f64": _STENCIL_ARGS,
    "evx_ac_stage_f32": _AC_ARGS, "evx_ac_stage_f64": _AC_ARGS,
    "evx_pad_ghost_f32": _PAD_ARGS, "evx_pad_ghost_f64": _PAD_ARGS,
    "evx_padded_stencil_f32": _PST_ARGS, "evx_padded_stencil_f64": _PST_ARGS,
    "evx_rd2_rhs_f32": _RD2_ARGS, "evx_rd2_rhs_f64": _RD2_ARGS,
    "evx_imex_plan_create": [ctypes.POINTER(_c_void_p), _c_int, _c_int, _c_int, _c_int, _c_int],
    "evx_imex_plan_destroy": [_c_void_p],
    "evx_imex_plan_backend": [_c_void_p],
    "evx_imex_plan_workspace_bytes": [_c_void_p, ctypes.POINTER(ctypes.c_size_t)],
    "evx_imex_apply_f32": _APPLY_ARGS, "evx_imex_apply_f64": _APPLY_ARGS,
    "evx_ch_imex_step_f32": _STEP_ARGS, "evx_ch_imex_step_f64": _STEP_ARGS,
    "evx_imex_native_pass_f32": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _dptr,
                                 _c_double, _c_double, _c_int, _c_void_p],
    "evx_spectral_filter_c64": _FILTER_ARGS, "evx_spectral_filter_c128": _FILTER_ARGS,
    "evx_ch_mu_f32": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_mu_f64": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_adjoint_flux_f32": [_c_void_p] * 5 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_adjoint_flux_f64": [_c_void_p] * 5 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_adjoint_combine_f32": [_c_void_p] * 6 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_adjoint_combine_f64": [_c_void_p] * 6 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_void_p],
    "evx_ch_adjoint_combine_range_f32": [_c_void_p] * 6 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_int, _c_int, _c_void_p],
    "evx_ch_adjoint_combine_range_f64": [_c_void_p] * 6 + [_c_int, _c_int, _c_int, _dptr, _c_double, _c_int, _c_int, _c_void_p],
    "evx_dist_plan_create": [ctypes.POINTER(_c_void_p), _c_int, _c_int, _c_int, _c_int, _c_int],
    "evx_dist_plan_destroy": [_c_void_p],
    "evx_dist_plan_set_p2p_ctas": [_c_void_p, _c_int],
    "evx_dist_plan_sizes": [_c_void_p, ctypes.POINTER(ctypes.c_size_t), _iptr],
    "evx_dist_forward_f32": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p],
    "evx_dist_middle_f32": [_c_void_p, _c_void_p, _dptr, _c_double, _c_double, _c_int, _c_void_p],
    "evx_dist_forward_chunk_f32": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                   _c_void_p],
    "evx_dist_middle_chunk_f32": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _dptr, _c_double,
                                  _c_double, _c_int, _c_void_p],
    "evx_peer_scatter": [ctypes.POINTER(_c_void_p), ctypes.POINTER(_c_void_p), _c_int, ctypes.c_size_t,
                         ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _c_int, _c_void_p],
    "evx_copy_async": [_c_void_p, _c_void_p, ctypes.c_size_t, _c_void_p],
    "evx_copy2d_async": [_c_void_p, ctypes.c_size_t, _c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                         ctypes.c_size_t, _c_void_p],
    "evx_copy_batch_async": [ctypes.POINTER(_c_void_p), ctypes.c_size_t, ctypes.POINTER(_c_void_p), ctypes.c_size_t,
                             ctypes.c_size_t, ctypes.c_size_t, _c_int, _c_void_p],
    "evx_dist_forward_p2p_f32": [_c_void_p, _c_void_p, _c_void_p, ctypes.POINTER(_c_void_p), _c_void_p],
    "evx_dist_forward_chunk_p2p_f32": [_c_void_p, _c_void_p, _c_void_p, ctypes.POINTER(_c_void_p), _c_int,
                                       _c_int, _c_int, _c_void_p],
    "evx_dist_middle_p2p_f32": [_c_void_p, _c_void_p, ctypes.POINTER(_c_void_p), _dptr, _c_double,
                                _c_double, _c_int, _c_void_p],
    "evx_dist_middle_chunk_p2p_f32": [_c_void_p, _c_void_p, ctypes.POINTER(_c_void_p), _c_int, _c_int, _dptr,
                                      _c_double, _c_double, _c_int, _c_void_p],
    "evx_dist_backward_f32": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p],
}
_RESTYPES = {"evx_strerror": ctypes.c_char_p, "evx_launch_count": ctypes.c_ulonglong}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load_library():
    """Load libevx_b200.so and bind every symbol of the header.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found. Build it with `python -m evoxels_b200.build` "
            "(needs nvcc); evoxels_b200 has no CPU or pure-PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so is stale
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _c_int)
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load_library().evx_strerror(code).decode()
        err = NativeLibraryError(f"{what} failed with code {code}: {msg}")
        err.code = code
        raise err


def launch_count() -> int:
    return int(load_library().evx_launch_count())


# --------------------------------------------------------------------------------------
# argument marshalling
# --------------------------------------------------------------------------------------
def _suffix(t: torch.Tensor) -> str:
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.float64:
        return "f64"
    raise TypeError(f"evoxels_b200 kernels take float32/float64 fields, got {t.dtype}")


def require_cuda(*tensors):
    """The product has no CPU path - refuse anything that is not a CUDA tensor."""
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(
                "evoxels_b200 has no CPU path: hot-path operators accept CUDA tensors only "
                f"(got {type(t).__name__} on "
                f"{getattr(t, 'device', 'host')}).")


def refuse_grad(what: str, *tensors):
    """The kernels behind these wrappers record no autograd node.  Called with grad mode on
    and an input that requires grad, they would return a detached result and backprop would
    silently treat it as a constant (the reference's torch ops are differentiable) - raise
    instead.  Inside `torch.autograd.Function.forward/backward` (autograd.py) grad mode is
    off, so the hand-written adjoint path is not affected."""
    if not torch.is_grad_enabled():
        return
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.requires_grad:
            raise NotImplementedError(
                f"{what} runs a CUDA kernel without a backward pass; an input requires grad. "
                "Differentiable on this path: CahnHilliard (fully periodic, default mu_hom) "
                "through PseudoSpectralIMEX.step / CahnHilliard.rhs. Wrap the call in "
                "torch.no_grad() if no gradient is wanted.")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _h3(spacing: Sequence[float]):
    return (ctypes.c_double * 3)(*[float(s) for s in spacing])


def bc_to_c(bc):
    """((kind, values),)*3 -> (int[3], double[6])."""
    kinds = (ctypes.c_int * 3)(*[BC_CODES[k] for k, _ in bc])
    vals = []
    for _, v in bc:
        vals += [float(v[0]), float(v[1])] if v is not None else [0.0, 0.0]
    return kinds, (ctypes.c_double * 6)(*vals)


def _field3(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 3 or not t.is_contiguous():
        raise ValueError("expected a contiguous [nx,ny,nz] tensor")
    return t


# --------------------------------------------------------------------------------------
# thin typed wrappers (one per C entry point)
# --------------------------------------------------------------------------------------
def ch_rhs(c, out, spacing, eps, D, bc, hom=None, halo_lo=None, halo_hi=None):
    require_cuda(c, out, hom, halo_lo, halo_hi)
    refuse_grad("evx_ch_rhs", c, hom, halo_lo, halo_hi)
    lib = load_library()
    nx, ny, nz = _field3(c).shape
    kinds, vals = bc_to_c(bc)
    with torch.cuda.device(c.device):
        check(getattr(lib, "evx_ch_rhs_" + _suffix(c))(
            _ptr(c), _ptr(hom), _ptr(_field3(out)), nx, ny, nz, _h3(spacing), float(eps),
            float(D), kinds, vals, _ptr(halo_lo), _ptr(halo_hi), _stream(c)), "evx_ch_rhs")
    return out


def ac_stage(phi, spacing, eps, gab, M, force, curvature, bc, *, pot=None, k_out=None,
             base=None, y_out=None, alpha=0.0, acc_in=None, acc_out=None, beta=0.0,
             halo_lo=None, halo_hi=None):
    require_cuda(phi, pot, k_out, base, y_out, acc_in, acc_out, halo_lo, halo_hi)
    refuse_grad("evx_ac_stage", phi, pot, base, acc_in, halo_lo, halo_hi)
    lib = load_library()
    nx, ny, nz = _field3(phi).shape
    kinds, vals = bc_to_c(bc)
    with torch.cuda.device(phi.device):
        check(getattr(lib, "evx_ac_stage_" + _suffix(phi))(
            _ptr(phi), _ptr(pot), _ptr(k_out), _ptr(base), _ptr(y_out), float(alpha),
            _ptr(acc_in), _ptr(acc_out), float(beta), nx, ny, nz, _h3(spacing), float(eps),
            float(gab), float(M), float(force), float(curvature), kinds, vals,
            _ptr(halo_lo), _ptr(halo_hi), _stream(phi)), "evx_ac_stage")


def pad_ghost(field, bc):
    require_cuda(field)
    refuse_grad("evx_pad_ghost", field)
    lib = load_library()
    nx, ny, nz = _field3(field).shape
    out = torch.empty((nx + 2, ny + 2, nz + 2), dtype=field.dtype, device=field.device)
    kinds, vals = bc_to_c(bc)
    with torch.cuda.device(field.device):
        check(getattr(lib, "evx_pad_ghost_" + _suffix(field))(
            _ptr(field), _ptr(out), nx, ny, nz, kinds, vals, _stream(field)), "evx_pad_ghost")
    return out


def padded_stencil(padded, spacing, op):
    require_cuda(padded)
    refuse_grad("evx_padded_stencil", padded)
    lib = load_library()
    px, py, pz = _field3(padded).shape
    nx, ny, nz = px - 2, py - 2, pz - 2
    out = torch.empty((nx, ny, nz), dtype=padded.dtype, device=padded.device)
    with torch.cuda.device(padded.device):
        check(getattr(lib, "evx_padded_stencil_" + _suffix(padded))(
            _ptr(padded), _ptr(out), nx, ny, nz, _h3(spacing), int(op), _stream(padded)),
            "evx_padded_stencil")
    return out


def rd2_rhs(u, spacing, D_A, D_B, feed, kill, interaction=None):
    """Two-species reaction-diffusion rhs of u [2,nx,ny,nz] (fully periodic)."""
    require_cuda(u, interaction)
    refuse_grad("evx_rd2_rhs", u, interaction)
    lib = load_library()
    assert u.dim() == 4 and u.shape[0] == 2 and u.is_contiguous()
    out = torch.empty_like(u)
    _, nx, ny, nz = u.shape
    with torch.cuda.device(u.device):
        check(getattr(lib, "evx_rd2_rhs_" + _suffix(u))(
            _ptr(u), _ptr(interaction), _ptr(out), nx, ny, nz, _h3(spacing), float(D_A),
            float(D_B), float(feed), float(kill), _stream(u)), "evx_rd2_rhs")
    return out


def peer_scatter(src_ptrs, dst_ptrs, row_bytes, rows, src_pitch, dst_pitch, ctas_per_region, stream):
    """One kernel copying len(src_ptrs) pitched regions src[i] -> dst[i] (raw device pointers,
    local or mapped peer memory) on `stream`."""
    n = len(src_ptrs)
    arr = ctypes.c_void_p * n
    check(load_library().evx_peer_scatter(arr(*[int(p) for p in src_ptrs]), arr(*[int(p) for p in dst_ptrs]),
                                          n, int(row_bytes), int(rows), int(src_pitch), int(dst_pitch),
                                          int(ctas_per_region), _c_void_p(stream.cuda_stream)),
          "evx_peer_scatter")


def copy_async(dst_ptr, src_ptr, nbytes, stream):
    """cudaMemcpyAsync between raw device pointers (local or mapped peer memory) on `stream`."""
    check(load_library().evx_copy_async(_c_void_p(int(dst_ptr)), _c_void_p(int(src_ptr)), int(nbytes),
                                        _c_void_p(stream.cuda_stream)), "evx_copy_async")


def copy_batch_async(dst_ptrs, dpitch, src_ptrs, spitch, width_bytes, height, stream):
    """The copies of one chunk to all peers as ONE unordered batch (the driver may spread them over
    its copy engines); height 1 = contiguous regions."""
    n = len(dst_ptrs)
    d = (_c_void_p * n)(*[int(p) for p in dst_ptrs])
    s = (_c_void_p * n)(*[int(p) for p in src_ptrs])
    with torch.cuda.device(stream.device):
        check(load_library().evx_copy_batch_async(d, int(dpitch), s, int(spitch), int(width_bytes), int(height),
                                                  n, _c_void_p(stream.cuda_stream)), "evx_copy_batch_async")


def copy2d_async(dst_ptr, dpitch, src_ptr, spitch, width_bytes, height, stream):
    check(load_library().evx_copy2d_async(_c_void_p(int(dst_ptr)), int(dpitch), _c_void_p(int(src_ptr)),
                                          int(spitch), int(width_bytes), int(height),
                                          _c_void_p(stream.cuda_stream)), "evx_copy2d_async")


def spectral_filter(spec, shape, spacing, dt, coef, power, scale=1.0):
    """In-place P(k) multiply of a cuFFT-layout half spectrum (complex64/128 tensor)."""
    require_cuda(spec)
    refuse_grad("evx_spectral_filter", spec)
    lib = load_library()
    fn = lib.evx_spectral_filter_c64 if spec.dtype == torch.complex64 else lib.evx_spectral_filter_c128
    nx, ny, nz = shape
    with torch.cuda.device(spec.device):
        check(fn(_ptr(spec), nx, ny, nz, _h3(spacing), float(dt), float(coef), int(power),
                 float(scale), _stream(spec)), "evx_spectral_filter")
    return spec


class ImexPlan:
    """Owns an evx_imex_plan and its torch-allocated scratch buffer."""

    def __init__(self, shape, dtype=torch.float32, device="cuda", backend=FFT_AUTO):
        self._handle = None
        lib = load_library()
        self.shape = tuple(int(n) for n in shape)
        self.dtype = dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("evoxels_b200 has no CPU path: ImexPlan needs a CUDA device")
        handle = _c_void_p()
        with torch.cuda.device(self.device):
            check(lib.evx_imex_plan_create(ctypes.byref(handle), *self.shape,
                                           1 if dtype == torch.float64 else 0, int(backend)),
                  "evx_imex_plan_create")
            self._handle = handle
            nbytes = ctypes.c_size_t()
            check(lib.evx_imex_plan_workspace_bytes(handle, ctypes.byref(nbytes)),
                  "evx_imex_plan_workspace_bytes")
            self.workspace = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8,
                                         device=self.device)
        self.backend = int(lib.evx_imex_plan_backend(handle))

    @property
    def backend_name(self):
        return {FFT_CUFFT: "cufft", FFT_NATIVE: "native", FFT_NATIVE_MIXED: "native-mixed"}[self.backend]

    def apply(self, u, r, out, spacing, dt, coef, power):
        """out = u + irfftn(P * rfftn(r)); u may be None (out = update only)."""
        require_cuda(u, r, out)
        refuse_grad("evx_imex_apply", u, r)
        lib = load_library()
        assert tuple(r.shape) == self.shape and r.dtype == self.dtype
        with torch.cuda.device(self.device):
            check(getattr(lib, "evx_imex_apply_" + _suffix(r))(
                self._handle, _ptr(u), _ptr(_field3(r)), _ptr(_field3(out)),
                _ptr(self.workspace), _h3(spacing), float(dt), float(coef), int(power),
                _stream(r)), "evx_imex_apply")
        return out

    def native_pass(self, which, u, r, out, spacing, dt, coef, power):
        """One pass of the native pipeline (0 z fwd, 1 y fwd, 2 x fwd*filter*inv, 3 y inv, 4 z inv)."""
        require_cuda(u, r, out)
        with torch.cuda.device(self.device):
            check(load_library().evx_imex_native_pass_f32(
                self._handle, int(which), _ptr(u), _ptr(r), _ptr(out), _ptr(self.workspace),
                _h3(spacing), float(dt), float(coef), int(power), _stream(r)), "evx_imex_native_pass")

    def ch_step(self, u, out, spacing, dt, eps, D, A, hom=None):
        require_cuda(u, out, hom)
        refuse_grad("evx_ch_imex_step", u, hom)
        lib = load_library()
        assert tuple(u.shape) == self.shape and u.dtype == self.dtype
        with torch.cuda.device(self.device):
            check(getattr(lib, "evx_ch_imex_step_" + _suffix(u))(
                self._handle, _ptr(_field3(u)), _ptr(hom), _ptr(_field3(out)),
                _ptr(self.workspace), _h3(spacing), float(dt), float(eps), float(D), float(A),
                _stream(u)), "evx_ch_imex_step")
        return out

    def close(self):
        if self._handle is not None and _lib is not None:
            _lib.evx_imex_plan_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ch_rhs_vjp(u, w, spacing, eps, D, lam_in=None, deps_planes=None):
    """Vector-Jacobian product of the periodic CH rhs at state `u` with cotangent `w`.
    Returns (dL/du [+ lam_in], dL/deps as a 0-dim float64 CUDA tensor).  Three kernels.
    deps_planes = (x_lo, x_hi): only those planes enter dL/deps (x-slab extended by halos)."""
    require_cuda(u, w, lam_in)
    lib = load_library()
    sfx = _suffix(u)
    nx, ny, nz = _field3(u).shape
    mu, z, m, lam = (torch.empty_like(u) for _ in range(4))
    deps = torch.zeros((), dtype=torch.float64, device=u.device)
    h = _h3(spacing)
    with torch.cuda.device(u.device):
        st = _stream(u)
        check(getattr(lib, "evx_ch_mu_" + sfx)(_ptr(u), _ptr(mu), nx, ny, nz, h, float(eps), st),
              "evx_ch_mu")
        check(getattr(lib, "evx_ch_adjoint_flux_" + sfx)(_ptr(u), _ptr(mu), _ptr(_field3(w)), _ptr(z),
                                                         _ptr(m), nx, ny, nz, h, float(D), st),
              "evx_ch_adjoint_flux")
        if deps_planes is None:
            check(getattr(lib, "evx_ch_adjoint_combine_" + sfx)(_ptr(u), _ptr(z), _ptr(m), _ptr(lam_in),
                                                                _ptr(lam), _ptr(deps), nx, ny, nz, h,
                                                                float(eps), st), "evx_ch_adjoint_combine")
        else:
            check(getattr(lib, "evx_ch_adjoint_combine_range_" + sfx)(
                _ptr(u), _ptr(z), _ptr(m), _ptr(lam_in), _ptr(lam), _ptr(deps), nx, ny, nz, h, float(eps),
                int(deps_planes[0]), int(deps_planes[1]), st), "evx_ch_adjoint_combine_range")
    return lam, deps


class DistPlan:
    """x-slab distributed spectral stage of one rank (evx_dist_plan_* in the header)."""

    def __init__(self, global_shape, world, rank, device):
        self._handle = None
        lib = load_library()
        self.shape = tuple(int(n) for n in global_shape)
        self.world, self.rank = int(world), int(rank)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("evoxels_b200 has no CPU path: DistPlan needs a CUDA device")
        handle = _c_void_p()
        with torch.cuda.device(self.device):
            check(lib.evx_dist_plan_create(ctypes.byref(handle), *self.shape, self.world, self.rank),
                  "evx_dist_plan_create")
            self._handle = handle
            nbytes, pitch = ctypes.c_size_t(), ctypes.c_int()
            check(lib.evx_dist_plan_sizes(handle, ctypes.byref(nbytes), ctypes.byref(pitch)),
                  "evx_dist_plan_sizes")
        self.pitch = int(pitch.value)
        nx, ny, _ = self.shape
        self.block_shape = (self.world, nx // self.world, ny // self.world, self.pitch)
        assert nbytes.value == 8 * self.world * self.block_shape[1] * self.block_shape[2] * self.pitch

    def new_buffer(self):
        return torch.empty(self.block_shape, dtype=torch.complex64, device=self.device)

    def set_p2p_ctas(self, n):
        check(load_library().evx_dist_plan_set_p2p_ctas(self._handle, int(n)), "evx_dist_plan_set_p2p_ctas")

    def forward(self, r_local, spec, send):
        require_cuda(r_local, spec, send)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_forward_f32(self._handle, _ptr(_field3(r_local)), _ptr(spec),
                                                      _ptr(send), _stream(r_local)), "evx_dist_forward")

    def middle(self, recv, spacing, dt, coef, power):
        require_cuda(recv)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_middle_f32(self._handle, _ptr(recv), _h3(spacing), float(dt),
                                                     float(coef), int(power), _stream(recv)),
                  "evx_dist_middle")

    def forward_chunk(self, r_local, spec, send, x0, nxc, self_block=None):
        """z + y pass of the local x planes [x0, x0+nxc) into the block layout `send`; block
        `rank` goes to `self_block` instead when given."""
        require_cuda(r_local, spec, send, self_block)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_forward_chunk_f32(
                self._handle, _ptr(_field3(r_local)), _ptr(spec), _ptr(send), _ptr(self_block),
                int(x0), int(nxc), _stream(r_local)), "evx_dist_forward_chunk")

    def middle_chunk(self, recv, yl0, nylc, spacing, dt, coef, power, self_block=None):
        """x pass (forward * weight * inverse) of the local y-pencil rows [yl0, yl0+nylc), in
        place except for the rows of block `rank`, which go to `self_block` when given."""
        require_cuda(recv, self_block)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_middle_chunk_f32(
                self._handle, _ptr(recv), _ptr(self_block), int(yl0), int(nylc), _h3(spacing),
                float(dt), float(coef), int(power), _stream(recv)), "evx_dist_middle_chunk")

    @staticmethod
    def _ptr_array(ptrs):
        return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])

    def forward_p2p(self, r_local, spec, peer_ptrs):
        """ZFwd + y pass whose stores go straight into the peers' block buffers."""
        require_cuda(r_local, spec)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_forward_p2p_f32(self._handle, _ptr(_field3(r_local)), _ptr(spec),
                                                          self._ptr_array(peer_ptrs), _stream(r_local)),
                  "evx_dist_forward_p2p")

    def forward_chunk_p2p(self, r_local, spec, peer_ptrs, x0, nxc, parts=3):
        """forward_p2p for the local x planes [x0, x0+nxc) on the current stream
        (parts: 1 z pass, 2 y pass with peer stores, 3 both)."""
        require_cuda(r_local, spec)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_forward_chunk_p2p_f32(
                self._handle, _ptr(_field3(r_local)), _ptr(spec), self._ptr_array(peer_ptrs),
                int(x0), int(nxc), int(parts), _stream(r_local)), "evx_dist_forward_chunk_p2p")

    def middle_p2p(self, recv, peer_ptrs, spacing, dt, coef, power):
        require_cuda(recv)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_middle_p2p_f32(self._handle, _ptr(recv), self._ptr_array(peer_ptrs),
                                                         _h3(spacing), float(dt), float(coef), int(power),
                                                         _stream(recv)), "evx_dist_middle_p2p")

    def middle_chunk_p2p(self, recv, peer_ptrs, yl0, nylc, spacing, dt, coef, power):
        """middle_p2p for the local y-pencil rows [yl0, yl0+nylc); the table may mix peer and local
        destinations (hybrid transport)."""
        require_cuda(recv)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_middle_chunk_p2p_f32(
                self._handle, _ptr(recv), self._ptr_array(peer_ptrs), int(yl0), int(nylc), _h3(spacing),
                float(dt), float(coef), int(power), _stream(recv)), "evx_dist_middle_chunk_p2p")

    def backward(self, recv, spec, u_local, out_local):
        require_cuda(recv, spec, u_local, out_local)
        with torch.cuda.device(self.device):
            check(load_library().evx_dist_backward_f32(self._handle, _ptr(recv), _ptr(spec),
                                                       _ptr(u_local), _ptr(_field3(out_local)),
                                                       _stream(recv)), "evx_dist_backward")

    def close(self):
        if self._handle is not None and _lib is not None:
            _lib.evx_dist_plan_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
