"""CPU replay of the CUDA thread-block programs (tests/emu/emu_kernels.cpp compiles the
very same *_core.h phase functions with g++) against the oracle.  Catches index / halo /
boundary-rule mistakes without a GPU.  Test infrastructure only - the product never runs
these on the host."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2
from oracle import evx_oracle as O

KIND = {"periodic": 0, "neumann": 1, "dirichlet": 2}
BCS = [("periodic",) * 3, ("neumann",) * 3,
       (("dirichlet", (0.2, 0.6)), "neumann", "periodic"),
       ("periodic", "neumann", ("dirichlet", (0.1, 0.9))),
       (("dirichlet", (1, 2)), ("dirichlet", (3, 4)), ("dirichlet", (5, 6)))]
SHAPES = [((12, 9, 8), (0.5, 1.0, 0.5)), ((7, 20, 68), (1.0, 1.0, 1.0)), ((16, 1, 1), (1, 1, 1)),
          ((3, 33, 12), (1, 2, 1)), ((2, 2, 2), (1, 1, 1)), ((9, 8, 7), (1, 1, 1))]


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "emu_kernels.cpp")
    out = os.path.join(ROOT, "tests", "emu", "libevx_emu.so")
    deps = [src] + [os.path.join(ROOT, "evoxels_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "evoxels_b200", "csrc")) if f.endswith(".h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", out, src])
    return ctypes.CDLL(out)


def _bc(bc):
    bc = O.normalize_bc(bc)
    k = (ctypes.c_int * 3)(*[KIND[b[0]] for b in bc])
    v = (ctypes.c_double * 6)(*sum([list(b[1]) if b[1] else [0.0, 0.0] for b in bc], []))
    return k, v


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu_ch(lib, u, sp, eps, D, bc, xchunk, vec, hom=None, hlo=None, hhi=None):
    f = lib.emu_ch_rhs_f32 if u.dtype == np.float32 else lib.emu_ch_rhs_f64
    out = np.full_like(u, np.nan)
    k, v = _bc(bc)
    rc = f(_p(u), _p(hom), _p(out), *u.shape, (ctypes.c_double * 3)(*sp), ctypes.c_double(eps),
           ctypes.c_double(D), k, v, _p(hlo), _p(hhi), xchunk, vec)
    assert rc == 0
    return out


def emu_ac(lib, u, sp, kw, bc, xchunk, vec, alpha=0.0, beta=0.0, acc=None, hlo=None, hhi=None):
    f = lib.emu_ac_stage_f32 if u.dtype == np.float32 else lib.emu_ac_stage_f64
    k_, y_, a_ = (np.full_like(u, np.nan) for _ in range(3))
    k, v = _bc(bc)
    d = ctypes.c_double
    rc = f(_p(u), None, _p(k_), _p(u), _p(y_), d(alpha), _p(acc), _p(a_), d(beta), *u.shape,
           (ctypes.c_double * 3)(*sp), d(kw["eps"]), d(kw["gab"]), d(kw["M"]), d(kw["force"]),
           d(kw["curvature"]), k, v, _p(hlo), _p(hhi), xchunk, vec)
    assert rc == 0
    return k_, y_, a_


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ch_rhs_program(emu, dtype):
    tol = 5e-12 if dtype == np.float64 else 2e-5
    for shape, sp in SHAPES:
        u = (-0.2 + 1.4 * np.random.default_rng(3).random(shape)).astype(dtype)
        for bc in BCS:
            ref = O.ch_rhs(torch.from_numpy(u)[None], sp, 2.5, 1.3, bc)[0].numpy()
            for xchunk in (shape[0], 3):
                for vec in (0, 1):
                    if vec and shape[2] % (16 // np.dtype(dtype).itemsize):
                        continue
                    got = emu_ch(emu, u, sp, 2.5, 1.3, bc, xchunk, vec)
                    assert rel_l2(got, ref) <= tol, (shape, bc, xchunk, vec)


def test_ch_rhs_program_custom_potential_and_halos(emu):
    u = (0.2 + 0.6 * np.random.default_rng(1).random((6, 10, 8)))
    mu = lambda c: torch.log(c.clip(1e-4, 1 - 1e-4) / (1 - c.clip(1e-4, 1 - 1e-4))) + 2.5 * (1 - 2 * c)  # noqa: E731
    ref = O.ch_rhs(torch.from_numpy(u)[None], (1, 1, 1), 3.0, 1.0, ("periodic",) * 3, mu)[0].numpy()
    hom = mu(torch.from_numpy(u).clip(0, 1)).numpy()
    assert rel_l2(emu_ch(emu, u, (1, 1, 1), 3.0, 1.0, ("periodic",) * 3, 4, 1, hom=hom), ref) <= 1e-12
    # x-slab decomposition with 2-plane halos reproduces the single-domain result
    u = (-0.2 + 1.4 * np.random.default_rng(2).random((12, 9, 8)))
    for bc in [("periodic",) * 3, ("neumann", "periodic", "neumann"),
               (("dirichlet", (0.2, 0.6)), "neumann", "periodic")]:
        ref = O.ch_rhs(torch.from_numpy(u)[None], (1, 1, 1), 3.0, 1.0, bc)[0].numpy()
        per = bc[0] == "periodic"
        parts = []
        for a, b in [(0, 5), (5, 12)]:
            lo = np.ascontiguousarray(np.take(u, [a - 2, a - 1], axis=0, mode="wrap")) if (per or a > 0) else None
            hi = np.ascontiguousarray(np.take(u, [b, b + 1], axis=0, mode="wrap")) if (per or b < 12) else None
            parts.append(emu_ch(emu, np.ascontiguousarray(u[a:b]), (1, 1, 1), 3.0, 1.0, bc, 3, 1, hlo=lo, hhi=hi))
        assert rel_l2(np.concatenate(parts), ref) <= 5e-12, bc
        # every phase of the 6-fold unrolled plane loop and of the cp.async staging cells
        for xchunk in (1, 2, 4, 5, 6, 7, 12):
            got = emu_ch(emu, u, (1, 1, 1), 3.0, 1.0, bc, xchunk, 1)
            assert rel_l2(got, ref) <= 5e-12, (bc, xchunk)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ac_stage_program(emu, dtype):
    tol = 5e-12 if dtype == np.float64 else 3e-5
    kw = dict(eps=3.0, gab=0.8, M=1.5, force=1.0, curvature=0.5)
    for shape, sp in SHAPES:
        u = (-0.2 + 1.4 * np.random.default_rng(3).random(shape)).astype(dtype)
        acc = np.random.default_rng(4).random(shape).astype(dtype)
        for bc in BCS:
            ref = O.ac_rhs(torch.from_numpy(u)[None], sp, bc=bc, **kw)[0].numpy()
            # vec: 0 scalar register-window program, 1 vector register-window program,
            #      2 shared-memory tile program (AcTileProgram - the one the library launches
            #      for aligned shapes)
            for vec in (0, 1, 2):
                if vec and shape[2] % (16 // np.dtype(dtype).itemsize):
                    continue
                k, y, a = emu_ac(emu, u, sp, kw, bc, 3, vec, alpha=0.05, beta=2.0, acc=acc)
                assert rel_l2(k, ref) <= tol, (shape, bc, vec)
                assert rel_l2(y, u + dtype(0.05) * ref) <= tol
                assert rel_l2(a, acc + 2 * ref) <= tol


def test_ac_stage_program_halos(emu):
    kw = dict(eps=3.0, gab=0.8, M=1.5, force=1.0, curvature=0.5)
    u = (-0.2 + 1.4 * np.random.default_rng(2).random((12, 9, 8)))
    for bc in [("periodic",) * 3, ("neumann",) * 3, (("dirichlet", (0.2, 0.6)), "neumann", "periodic")]:
        ref = O.ac_rhs(torch.from_numpy(u)[None], (1, 1, 1), bc=bc, **kw)[0].numpy()
        per = bc[0] == "periodic"
        parts = []
        for a, b in [(0, 5), (5, 12)]:
            lo = np.ascontiguousarray(np.take(u, [a - 1], axis=0, mode="wrap")) if (per or a > 0) else None
            hi = np.ascontiguousarray(np.take(u, [b], axis=0, mode="wrap")) if (per or b < 12) else None
            parts.append(emu_ac(emu, np.ascontiguousarray(u[a:b]), (1, 1, 1), kw, bc, 3, 1, hlo=lo, hhi=hi)[0])
        assert rel_l2(np.concatenate(parts), ref) <= 5e-12, bc
        for xchunk in (1, 2, 3, 5, 7):      # tile program: every prefetch / slot-rotation phase
            parts = []
            for a, b in [(0, 5), (5, 12)]:
                lo = np.ascontiguousarray(np.take(u, [a - 1], axis=0, mode="wrap")) if (per or a > 0) else None
                hi = np.ascontiguousarray(np.take(u, [b], axis=0, mode="wrap")) if (per or b < 12) else None
                parts.append(emu_ac(emu, np.ascontiguousarray(u[a:b]), (1, 1, 1), kw, bc, xchunk, 2,
                                    hlo=lo, hhi=hi)[0])
            assert rel_l2(np.concatenate(parts), ref) <= 5e-12, (bc, xchunk)


def test_pad_and_padded_stencils(emu):
    for dtype in (np.float64, np.float32):
        sfx = "f64" if dtype == np.float64 else "f32"
        for shape in [(5, 4, 3), (2, 2, 2), (3, 1, 4)]:
            u = np.random.default_rng(5).random(shape).astype(dtype)
            for bc in BCS:
                out = np.zeros(tuple(s + 2 for s in shape), dtype)
                k, v = _bc(bc)
                getattr(emu, "emu_pad_ghost_" + sfx)(_p(u), _p(out), *shape, k, v)
                ref = O.ghost_pad(torch.from_numpy(u)[None], bc)[0].numpy()
                assert np.abs(out - ref).max() <= 1e-6
                h = (ctypes.c_double * 3)(0.5, 1.0, 0.25)
                for op, fn in [(0, O.laplace7), (1, O.normal_laplace19)]:
                    o2 = np.zeros(shape, dtype)
                    getattr(emu, "emu_padded_stencil_" + sfx)(_p(ref), _p(o2), *shape, h, op)
                    r2 = fn(torch.from_numpy(ref)[None], (0.5, 1.0, 0.25))[0].numpy()
                    assert rel_l2(o2, r2) <= (1e-12 if dtype == np.float64 else 2e-4)


@pytest.mark.parametrize("pipe_blocks,xkz", [(0, 8), (1, 8), (3, 16), (0, 16)])
@pytest.mark.parametrize("shape", [(8, 8, 16), (16, 32, 64), (32, 16, 128), (8, 8, 256), (64, 8, 32)])
def test_native_fft_pipeline(emu, shape, pipe_blocks, xkz):
    """Five-pass native FFT pipeline (ZFwd, Y, X fwd*filter*inv, Y inv, ZInv+u) replayed on
    the CPU: forward spectrum against numpy's rfftn, full update against the oracle."""
    # pipe_blocks > 0: persistent software-pipelined strided passes with that many blocks
    # xkz: columns per x-pass tile; 16-column tiles straddle y groups when the pitch is 8 mod 16
    emu.emu_set_pipe_blocks(pipe_blocks)
    emu.emu_set_xpass_columns(xkz)
    nx, ny, nz = shape
    rng = np.random.default_rng(0)
    r = rng.standard_normal(shape).astype(np.float32)
    u = rng.random(shape).astype(np.float32)
    M = nz // 2
    P = ((M + 1 + 7) // 8) * 8
    sp = (1.0, 0.5, 2.0)
    h = (ctypes.c_double * 3)(*sp)
    spec = np.zeros((nx, ny, P, 2), np.float32)
    d = ctypes.c_double
    rc = emu.emu_native_apply(_p(u), _p(r), None, _p(spec), nx, ny, nz, h, d(0.1), d(1.5), 2)
    assert rc == P
    got = spec[:, :, :M + 1, 0] + 1j * spec[:, :, :M + 1, 1]
    assert np.linalg.norm(got - np.fft.rfftn(r.astype(np.float64))) / np.linalg.norm(got) < 5e-7
    out = np.zeros(shape, np.float32)
    assert emu.emu_native_apply(_p(u), _p(r), _p(out), None, nx, ny, nz, h, d(0.1), d(1.5), 2) == 0
    pref = O.imex_prefactor(O.ch_symbol(shape, sp, 3.0, 1.0, 0.25), 0.1)
    want = O.imex_step(torch.from_numpy(u)[None], torch.from_numpy(r)[None], pref)[0].numpy()
    emu.emu_set_pipe_blocks(0)
    emu.emu_set_xpass_columns(8)
    assert rel_l2(out - u, want - u) < 2e-6
    assert rel_l2(out, want) < 1e-6


@pytest.mark.parametrize("kz", [8, 16])
@pytest.mark.parametrize("shape,blocks,power", [((8, 512, 16), 3, 2), ((512, 8, 16), 2, 2),
                                                ((512, 512, 16), 7, 2), ((512, 8, 32), 5, 1 | 0x100)])
def test_tma_tiled_strided_passes_are_bit_identical(emu, shape, blocks, power, kz):
    """The TMA-tiled program of the 512-point strided passes (fft_line_core.h: two-line groups,
    swizzled tile image, exchanges through X and through the tile itself) replayed on the CPU
    gives the very same bits as the cp.async passes - y forward, x forward*weight*inverse (IMEX
    and exponential-Euler weight) and y inverse, 64- and 128-byte tile rows, including the
    partly out-of-range last tile of every row (nz/2+1 is not a multiple of the tile width)."""
    nx, ny, nz = shape
    rng = np.random.default_rng(5)
    r = rng.standard_normal(shape).astype(np.float32)
    u = rng.random(shape).astype(np.float32)
    h = (ctypes.c_double * 3)(1.0, 0.5, 2.0)
    d = ctypes.c_double
    want = np.zeros(shape, np.float32)
    emu.emu_set_line_columns(0)
    assert emu.emu_native_apply(_p(u), _p(r), _p(want), None, nx, ny, nz, h, d(0.1), d(1.5), power) == 0
    got = np.full(shape, np.nan, np.float32)
    emu.emu_set_line_columns(kz)
    emu.emu_set_pipe_blocks(blocks)
    try:
        assert emu.emu_native_apply(_p(u), _p(r), _p(got), None, nx, ny, nz, h, d(0.1), d(1.5), power) == 0
    finally:
        emu.emu_set_line_columns(0)
        emu.emu_set_pipe_blocks(0)
    assert np.array_equal(got, want)


def test_tma_tile_image_is_conflict_free():
    """Shared-memory bank check of the TMA-tiled pass on the index formulas of fft_line_core.h:
    every 64-bit access pattern of a half-warp (8 consecutive t x the 2 lines of a group) hits
    16 distinct 8-byte bank pairs - tile image with the 64-/128-byte TMA swizzle (natural order
    and stage-1 output order) and the padded exchange buffer (natural and stage-0 order)."""
    T = 64
    pad = lambda i: i + (i >> 3)
    for kz in (8, 16):
        rowb = 8 * kz
        swz = (lambda r: ((r >> 1) & 3) << 4) if kz == 8 else (lambda r: (r & 7) << 4)
        tile = lambda r, c: r * rowb + ((c * 8) ^ swz(r))
        for g in range(kz // 2):
            for t0 in range(0, T, 8):
                lanes = [(t, 2 * g + c2) for t in range(t0, t0 + 8) for c2 in (0, 1)]
                for e in range(8):
                    nat = {(tile(t + e * T, c) // 8) % 16 for t, c in lanes}
                    st1 = {(tile(64 * (t // 8) + t % 8 + 8 * e, c) // 8) % 16 for t, c in lanes}
                    xn = {(pad(t + e * T) * 2 + c % 2) % 16 for t, c in lanes}
                    xs = {(pad(8 * t + e) * 2 + c % 2) % 16 for t, c in lanes}
                    assert len(nat) == len(st1) == len(xn) == len(xs) == 16, (kz, g, t0, e)


@pytest.mark.parametrize("form", [8, 16])
@pytest.mark.parametrize("shape,blocks,power", [((8, 1024, 16), 3, 2), ((1024, 8, 16), 2, 2),
                                                ((1024, 8, 32), 5, 1 | 0x100)])
def test_four_stage_tma_tiled_passes_are_bit_identical(emu, shape, blocks, power, form):
    """1024-point lines (StridedLine4 in fft_line_core.h: two groups of 256 threads, two line
    pairs per group and tile, exchanges alternating between two padded buffers, tile touched in
    natural order only; form 16 = StridedLine16: sixteen points per thread, radix-2 and first
    radix-8 stage merged in registers, one padded buffer per group + the tile columns) replayed on
    the CPU against the cp.async passes: y forward, x forward*weight*inverse (IMEX and
    exponential-Euler weight), y inverse - same bits, including the partly out-of-range last tile
    of a row."""
    nx, ny, nz = shape
    rng = np.random.default_rng(5)
    r = rng.standard_normal(shape).astype(np.float32)
    u = rng.random(shape).astype(np.float32)
    h = (ctypes.c_double * 3)(1.0, 0.5, 2.0)
    d = ctypes.c_double
    want = np.zeros(shape, np.float32)
    emu.emu_set_line_columns(0)
    assert emu.emu_native_apply(_p(u), _p(r), _p(want), None, nx, ny, nz, h, d(0.1), d(1.5), power) == 0
    got = np.full(shape, np.nan, np.float32)
    emu.emu_set_line_columns(8)
    emu.emu_set_pipe_blocks(blocks)
    emu.emu_set_line4_form(form)
    try:
        assert emu.emu_native_apply(_p(u), _p(r), _p(got), None, nx, ny, nz, h, d(0.1), d(1.5), power) == 0
    finally:
        emu.emu_set_line_columns(0)
        emu.emu_set_pipe_blocks(0)
        emu.emu_set_line4_form(16)
    assert np.array_equal(got, want)


def test_sixteen_point_form_layout_is_conflict_free():
    """Bank check of StridedLine16 (T = 64 threads per line): 64-bit accesses of a half-warp (8
    consecutive q x the 2 lines of a group) hit 16 distinct 8-byte bank pairs - tile rows in natural
    order (q + 64 e) and in stage-2 output order ((q/16) 128 + q%16 + 512 i + 16 r) under the
    64-byte TMA swizzle, the exchange buffer (index i + (i >> 4)) in natural order and in the output
    order of the merged stage (16 q + k); every index is written exactly once."""
    T = 64
    pad16 = lambda i: i + (i >> 4)
    tile = lambda r, c: r * 64 + ((c * 8) ^ (((r >> 1) & 3) << 4))
    for g in range(4):
        for q0 in range(0, T, 8):
            lanes = [(q, 2 * g + c2) for q in range(q0, q0 + 8) for c2 in (0, 1)]
            for e in range(16):
                assert len({(tile(q + 64 * e, c) // 8) % 16 for q, c in lanes}) == 16
                assert len({(pad16(q + 64 * e) * 2 + c % 2) % 16 for q, c in lanes}) == 16
                assert len({(pad16(16 * q + e) * 2 + c % 2) % 16 for q, c in lanes}) == 16
                i, k = e % 2, e // 2
                rows = {(tile((q // 16) * 128 + q % 16 + 512 * i + 16 * k, c) // 8) % 16 for q, c in lanes}
                assert len(rows) == 16
    assert sorted(16 * q + k for q in range(T) for k in range(16)) == list(range(1024))
    assert sorted((q // 16) * 128 + q % 16 + 512 * i + 16 * k for q in range(T) for i in range(2)
                  for k in range(8)) == list(range(1024))
    assert max(pad16(q + 64 * e) for q in range(T) for e in range(16)) * 2 + 1 < (1024 + 64) * 2


def test_four_stage_tile_and_exchange_layout_is_conflict_free():
    """Bank check of StridedLine4: 64-bit accesses of a half-warp (8 consecutive t x the 2 lines of
    a pair) hit 16 distinct 8-byte bank pairs - natural-order tile rows under the 64-byte TMA
    swizzle for both line pairs of a group, natural-order reads and the stage-0/1/2 output orders
    of the padded exchange buffers."""
    T, L = 128, 1024
    pad = lambda i: i + (i >> 3)
    tile = lambda r, c: r * 64 + ((c * 8) ^ (((r >> 1) & 3) << 4))
    ns = [1, 2, 16]                                   # Ns of stages 0, 1, 2 (radix 2, 8, 8)
    radix = [2, 8, 8]

    def out_index(s, t, e):
        q = 8 // radix[s]
        i, r = e % q, e // q
        jv = t + i * T
        return (jv // ns[s]) * ns[s] * radix[s] + jv % ns[s] + r * ns[s]

    for g in range(2):
        for t0 in range(0, T, 8):
            for e in range(8):
                for ps in range(2):
                    lanes = [(t, 2 * g + c2 + 4 * ps) for t in range(t0, t0 + 8) for c2 in (0, 1)]
                    nat = {(tile(t + e * T, c) // 8) % 16 for t, c in lanes}
                    assert len(nat) == 16, ("tile", g, t0, e, ps)
                lanes = [(t, c2) for t in range(t0, t0 + 8) for c2 in (0, 1)]
                assert len({(pad(t + e * T) * 2 + c2) % 16 for t, c2 in lanes}) == 16
                for s in range(3):
                    idx = {(pad(out_index(s, t, e)) * 2 + c2) % 16 for t, c2 in lanes}
                    assert len(idx) == 16, ("x", s, t0, e)
    # every line index is written exactly once per stage
    for s in range(3):
        assert sorted(out_index(s, t, e) for t in range(T) for e in range(8)) == list(range(L))


@pytest.mark.parametrize("shape", [(8, 8, 16), (16, 32, 64), (8, 16, 256)])
def test_native_fft_pipeline_etd1_weight(emu, shape):
    """Same pipeline with the exponential-Euler weight (EVX_FILTER_ETD1) against the oracle's
    etd1_step; dt*coef*k^2 spans both the Pade branch (|z| < 0.5) and the exp branch."""
    nx, ny, nz = shape
    rng = np.random.default_rng(1)
    r = rng.standard_normal(shape).astype(np.float32)
    u = rng.random(shape).astype(np.float32)
    sp = (1.0, 0.5, 2.0)
    h = (ctypes.c_double * 3)(*sp)
    d = ctypes.c_double
    out = np.zeros(shape, np.float32)
    assert emu.emu_native_apply(_p(u), _p(r), _p(out), None, nx, ny, nz, h, d(0.3), d(0.7), 1 | 0x100) == 0
    sym = -0.7 * O.k_squared(shape, sp)
    z = 0.3 * sym
    assert float(z.abs().min()) < 0.5 < float(z.abs().max())
    want = O.etd1_step(torch.from_numpy(u)[None], torch.from_numpy(r)[None], sym, 0.3)[0].numpy()
    assert rel_l2(out - u, want - u) < 2e-6
    assert rel_l2(out, want) < 1e-6


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rd2_rhs_program(emu, dtype):
    """Two-species reaction-diffusion rhs (rd_core.h) against the oracle, scalar and vector
    paths, default and caller-supplied interaction."""
    tol = 5e-13 if dtype == np.float64 else 2e-6
    f = emu.emu_rd2_rhs_f32 if dtype == np.float32 else emu.emu_rd2_rhs_f64
    d = ctypes.c_double
    kw = dict(D_A=0.8, D_B=1.3, feed=0.03, kill=0.06)
    for shape, sp in SHAPES:
        rng = np.random.default_rng(6)
        u = np.stack([rng.random(shape), 0.5 * rng.random(shape)]).astype(dtype)
        for custom in (False, True):
            fn = (lambda v: v[0] ** 2 * v[1] + 0.1) if custom else None
            ref = O.crd_rhs(torch.from_numpy(u), sp, interaction=fn, **kw).numpy()
            inter = fn(u).astype(dtype) if custom else None
            for vec in (0, 1):
                if vec and shape[2] % (16 // np.dtype(dtype).itemsize):
                    continue
                out = np.full_like(u, np.nan)
                rc = f(_p(u), _p(inter), _p(out), *shape, (ctypes.c_double * 3)(*sp), d(kw["D_A"]),
                       d(kw["D_B"]), d(kw["feed"]), d(kw["kill"]), vec)
                assert rc == 0
                assert rel_l2(out, ref) <= tol, (shape, custom, vec)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,W,nthreads", [((10, 12, 14), 8, 32), ((9, 7, 5), 4, 7), ((16, 1, 1), 8, 16),
                                              ((1, 6, 1), 2, 5), ((20, 25, 18), 8, 64), ((3, 4, 35), 1, 3),
                                              ((8, 8, 16), 8, 32), ((21, 2, 49), 8, 33)])
def test_mixed_radix_fft_pipeline(emu, shape, W, nthreads, dtype):
    """The mixed-radix passes (fft_generic_core.h: extents with prime factors <= 7, odd and
    degenerate axes included) replayed on the CPU against the oracle's IMEX and ETD1 steps."""
    nx, ny, nz = shape
    rng = np.random.default_rng(2)
    r = rng.standard_normal(shape).astype(dtype)
    u = rng.random(shape).astype(dtype)
    sp = (1.0, 0.5, 2.0)
    h = (ctypes.c_double * 3)(*sp)
    d = ctypes.c_double
    f = emu.emu_generic_apply_f32 if dtype == np.float32 else emu.emu_generic_apply_f64
    tol = 2e-6 if dtype == np.float32 else 2e-7      # fp64: the weight is float32 upstream
    out = np.zeros(shape, dtype)
    assert f(_p(u), _p(r), _p(out), nx, ny, nz, h, d(0.1), d(1.5), 2, W, nthreads) == 0
    pref = O.imex_prefactor(O.ch_symbol(shape, sp, 3.0, 1.0, 0.25), 0.1)
    want = O.imex_step(torch.from_numpy(u)[None], torch.from_numpy(r)[None], pref)[0].numpy()
    assert rel_l2(out - u, want - u) < tol, rel_l2(out - u, want - u)
    out = np.zeros(shape, dtype)
    assert f(_p(u), _p(r), _p(out), nx, ny, nz, h, d(0.3), d(0.7), 1 | 0x100, W, nthreads) == 0
    sym = -0.7 * O.k_squared(shape, sp)
    want = O.etd1_step(torch.from_numpy(u)[None], torch.from_numpy(r)[None], sym, 0.3)[0].numpy()
    assert rel_l2(out - u, want - u) < tol


def test_ch_rhs_adjoint_against_autograd(emu):
    """The hand-written VJP of the periodic CH rhs (adjoint_core.h) against torch autograd
    through the oracle, float64, values outside [0,1] included (clip mask)."""
    for shape, sp in [((6, 5, 4), (1.0, 0.5, 2.0)), ((8, 8, 8), (1.0, 1.0, 1.0)), ((3, 1, 7), (1, 1, 1))]:
        rng = np.random.default_rng(4)
        u = (-0.2 + 1.4 * rng.random(shape))
        w = rng.standard_normal(shape)
        ut = torch.from_numpy(u)[None].requires_grad_(True)
        D = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
        eps = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
        R = O.ch_rhs(ut, sp, eps, D)
        gu, gD, ge = torch.autograd.grad((R * torch.from_numpy(w)[None]).sum(), (ut, D, eps))
        lam = np.zeros(shape)
        deps = ctypes.c_double(0.0)
        emu.emu_ch_rhs_vjp_f64(_p(u), _p(w), _p(lam), ctypes.byref(deps), *shape,
                               (ctypes.c_double * 3)(*sp), ctypes.c_double(2.5), ctypes.c_double(1.3))
        assert rel_l2(lam, gu[0].numpy()) < 1e-12, shape
        assert abs(deps.value - float(ge)) < 1e-10 * max(1.0, abs(float(ge))), shape
        assert abs(float((R.detach()[0] * torch.from_numpy(w)).sum() / 1.3) - float(gD)) < 1e-10 * max(1.0, abs(float(gD)))


@pytest.mark.parametrize("L,KZ", [(64, 8), (128, 8), (256, 8), (512, 8), (512, 16), (1024, 8), (2048, 4),
                                  (64, 16), (256, 16)])
def test_tile_prefetch_forms_agree(emu, L, KZ):
    """The tabulated-step tile prefetch of the persistent strided passes copies exactly what the
    direct per-chunk address computation copies - for the plain y and x layouts and for the
    all-to-all block layout of the distributed inverse y pass (W = 2, 4, 8 splits)."""
    emu.emu_prefetch_mismatches.restype = ctypes.c_longlong
    for P in (KZ, 3 * KZ, 40 if KZ <= 8 else 48):
        if P % KZ and (3 * P) % KZ:
            continue
        groups = 3 if L <= 512 else 2
        if (groups * P) % KZ == 0:
            assert emu.emu_prefetch_mismatches(L, KZ, 0, groups, P, 1) == 0
            assert emu.emu_prefetch_mismatches(L, KZ, 1, groups, P, 1) == 0
            for W in (2, 4, 8):
                assert emu.emu_prefetch_mismatches(L, KZ, 2, groups, P, W) == 0


@pytest.mark.parametrize("shape,world", [((16, 16, 32), 2), ((32, 16, 64), 4), ((64, 32, 32), 8), ((32, 64, 16), 8)])
@pytest.mark.parametrize("pipe_blocks", [0, 3])
def test_distributed_pipeline_virtual_ranks(emu, shape, world, pipe_blocks):
    """The x-slab pipeline replayed for W virtual ranks with the launch parameters of
    dist_params.h (block-layout y passes, global ky offset in the x pass, peer tables) gives the
    single-domain result bit for bit - for the all-to-all, block-copy and peer-store transports,
    with and without pipeline chunks."""
    emu.emu_set_pipe_blocks(pipe_blocks)
    emu.emu_set_xpass_columns(8)
    nx, ny, nz = shape
    rng = np.random.default_rng(11)
    r = rng.standard_normal(shape).astype(np.float32)
    u = rng.random(shape).astype(np.float32)
    h = (ctypes.c_double * 3)(1.0, 0.5, 2.0)
    d = ctypes.c_double
    want = np.zeros(shape, np.float32)
    assert emu.emu_native_apply(_p(u), _p(r), _p(want), None, nx, ny, nz, h, d(0.1), d(1.5), 2) == 0
    nxl = nx // world
    cases = [(tr, fc, mc) for tr in (0, 1, 2) for fc, mc in ((1, 1), (4, 2)) if fc <= nxl]
    for transport, fwd_chunks, mid_chunks in cases:
        got = np.full(shape, np.nan, np.float32)
        rc = emu.emu_dist_apply(_p(u), _p(r), _p(got), nx, ny, nz, world, transport, fwd_chunks,
                                mid_chunks, h, d(0.1), d(1.5), 2)
        assert rc == 0
        assert np.array_equal(got, want), (transport, fwd_chunks, mid_chunks)
    # update only (u = NULL) with the exponential-Euler weight
    want = np.zeros(shape, np.float32)
    assert emu.emu_native_apply(None, _p(r), _p(want), None, nx, ny, nz, h, d(0.3), d(0.7), 1 | 0x100) == 0
    got = np.full(shape, np.nan, np.float32)
    assert emu.emu_dist_apply(None, _p(r), _p(got), nx, ny, nz, world, 1, 2, 2, h, d(0.3), d(0.7), 1 | 0x100) == 0
    assert np.array_equal(got, want)
    emu.emu_set_pipe_blocks(0)


@pytest.mark.parametrize("nplanes,lag,n0,n1,blocks", [
    (512, 12, 32, 33, 296), (512, 12, 33, 32, 296), (8, 12, 32, 33, 296), (16, 1, 32, 33, 296),
    (64, 3, 32, 33, 7), (512, 1, 32, 33, 296), (33, 5, 2, 3, 4), (9, 9, 4, 1, 50)])
def test_chain_schedule_covers_everything_in_dependency_order(emu, nplanes, lag, n0, n1, blocks):
    """Work list of the chained z/y kernels (fft_chain_core.h): complete, every dependency has a
    smaller item number, blocks walking their items in order never deadlock; with the shipped
    lag a stage-1 item practically never finds its plane unfinished."""
    waits = ctypes.c_longlong(-1)
    rc = emu.emu_chain_schedule_check(nplanes, lag, n0, n1, blocks, ctypes.byref(waits))
    assert rc == 0
    if lag >= 12 and nplanes >= 512:
        assert waits.value == 0
    # the division-free cursor the kernels walk (virtual numbering with empty slots)
    rc = emu.emu_chain_cursor_check(nplanes, lag, n0, n1, blocks, ctypes.byref(waits))
    assert rc == 0
    if lag >= 12 and nplanes >= 512:
        assert waits.value <= 64      # of 33280 items (the empty head slots skew the blocks a little)


def test_two_lines_per_group_z_passes_are_bit_identical(emu):
    """fft_zline_core.h (z lines as used by the chained kernels: rows in natural order at a
    264-complex pitch, two lines per 64-thread group, no address swizzle) against ZPass: forward
    spectrum rows and inverse real rows equal bit for bit, on poisoned scratch."""
    rng = np.random.default_rng(11)
    rows = 48
    r = rng.standard_normal((rows, 512)).astype(np.float32)
    u = rng.standard_normal((rows, 512)).astype(np.float32)
    emu.emu_zgroup_mismatches.restype = ctypes.c_longlong
    assert emu.emu_zgroup_mismatches(_p(r), _p(u), rows) == 0


def test_two_lines_per_group_layout_is_conflict_free():
    """64-bit shared-memory accesses of a half-warp (8 consecutive t x 2 lines) of the z-line
    program hit 16 distinct 8-byte banks, except the in-place stage-1 writes (two-way)."""
    ROWP, G = 264, 2
    pad = lambda i: i + (i >> 3)
    def banks(addr_cf):       # cf index -> 8-byte bank (16 banks of 8 bytes = 128 bytes)
        return addr_cf % 16
    worst = {}
    for t0 in range(0, 32, 8):
        for e in range(8):
            pats = {
                "rows natural": [(c2 * ROWP + t + 32 * e) for t in range(t0, t0 + 8) for c2 in range(G)],
                "rows partner": [(c2 * ROWP + 256 - t - 32 * e) for t in range(t0, t0 + 8) for c2 in range(G)],
                "rows stage1": [(c2 * ROWP + (t // 4) * 32 + t % 4 + 4 * e) for t in range(t0, t0 + 8) for c2 in range(G)],
                "x natural": [(pad(t + 32 * e) * G + c2) for t in range(t0, t0 + 8) for c2 in range(G)],
                "x stage0": [((4 * t + (t >> 1) + (e >> 1) + 144 * (e & 1)) * G + c2) for t in range(t0, t0 + 8) for c2 in range(G)],
            }
            for name, idx in pats.items():
                b = [banks(i) for i in idx]
                worst[name] = max(worst.get(name, 0), max(b.count(x) for x in set(b)))
    assert worst == {"rows natural": 1, "rows partner": 1, "rows stage1": 2, "x natural": 1, "x stage0": 1}
    # the stage-0 index really is the padded Stockham output index
    for t in range(32):
        for e in range(8):
            i, r = e & 1, e >> 1
            assert pad(4 * (t + 32 * i) + r) == 4 * t + (t >> 1) + (e >> 1) + 144 * (e & 1)
