// Pass programs of the native 3-D real FFT (host/device, barrier-free phases).
//
// Spectrum layout: S[nx][ny][P] complex64, P = pitch >= nz/2+1 (multiple of KZ), kz fastest.
//   ZFwd   real line (nz) -> half spectrum line (nz/2+1): half-length complex FFT + untangle
//   Strided<FWD|INV>   in-place complex FFT along y or x (line stride P or ny*P)
//   Strided<XMID>      forward along x, multiply by P(k)/N, inverse along x, one pass
//   ZInv   half spectrum line -> real line, fused  out = u + line
// Together: out = u + irfftn(P * rfftn(r)) with 5 passes over the data instead of the
// 6 transform passes + filter + add of a library FFT (reference timesteppers.py:85-89).
#pragma once
#include "fft_core.h"
#include "spectral_math.h"
#include "packed_f32.h"

namespace evx {

EVX_HD int smem_pad(int i) { return i + (i >> 3); }
constexpr int smem_padded_len(int n) { return n + (n >> 3) + 1; }

// Shared-memory index map of the z passes (lines are contiguous, lanes run along a line).
// For M >= 128 an XOR swizzle of the low four index bits makes every Stockham write pattern
// (strides 2/4/8 and the run-of-Ns patterns of later stages) and the natural-order reads
// conflict-free for 64-bit accesses (searched exhaustively); short lines keep
// the additive padding.
template <int M>
EVX_HD int zline_idx(int i) {
  return M >= 128 ? (i ^ ((i >> 3) & 15) ^ ((i >> 4) & 3)) : i + (i >> 3);
}
template <int M>
constexpr int zline_len() { return M >= 128 ? M + 16 : smem_padded_len(M + 1); }

// COPY: access-pattern probe; XMID_ETD1: XMID with the exponential-Euler weight
enum : int { PASS_FWD = 0, PASS_INV = 1, PASS_XMID = 2, PASS_COPY = 3, PASS_XMID_ETD1 = 4 };
constexpr bool pass_is_xmid(int mode) { return mode == PASS_XMID || mode == PASS_XMID_ETD1; }

// ------------------------------------------------------------------------------------
// strided passes (y and x)
// ------------------------------------------------------------------------------------
// Addressing of one side (input or output) of a strided pass.  Column c = grp * P + kz;
// point idx of its line lives at
//   grp * plane_stride + kz + (idx / split) * split_stride + (idx % split) * line_stride,
// split = 2^split_shift.
// split == L (split_stride unused) is the plain layout; split = L / W writes / reads the
// line in W chunks that are split_stride apart - the block layout of the slab<->pencil
// all-to-all, so that no separate pack / unpack pass exists.
struct StridedIO {
  long long line_stride;
  long long plane_stride;
  long long split_stride;
  int split_shift;         // split = 1 << split_shift (power of two)
};

struct StridedParams {
  const cf* in;
  cf* out;                 // may equal `in` when both sides use the same addressing
  StridedIO src, dst;
  const cf* tw;            // W_L table
  int P;                   // pitch (columns per group)
  int ncols_valid;         // columns kz < ncols_valid carry data (nz/2+1)
  long long ncols_total;   // number of columns incl. pitch padding (groups * P)
  int kother_offset;       // XMID: global index of the first local column group (y-pencils)
  FilterParams filt;       // XMID only; n0 = L (this axis), n1 = the other strided axis, n2 = nz
  long long src_step[8];   // strided_step(src, e * L/8), filled by finalize_strided()
  long long dst_step[8];
  long long src_pf_step[4];  // strided_step(src, it * L/4): row steps of the tile prefetch
  // peer-store mode (x-slab transposes over NVLink): chunk hi = idx >> dst.split_shift of a
  // line is written straight into rank hi's buffer out_peers[hi] at element offset
  // dst_peer_base + grp * plane_stride + kz + lo * line_stride  (no all-to-all afterwards)
  cf* out_peers[8];
  long long dst_peer_base;
  int use_peers;
  int max_ctas;            // > 0: cap of the persistent grid (NVLink-bound launches leave SMs free)
  // XMID on a non-periodic axis: the array holds L/2 rows; row i >= L/2 of the transformed line
  // is mirror * row (L-1-i) on load and is not stored (mirror = +1 even, -1 odd, 0 periodic)
  int mirror = 0;
};

inline StridedIO plain_io(long long line_stride, long long plane_stride, int L) {
  return StridedIO{line_stride, plane_stride, 0, ilog2(L)};
}

EVX_HD long long strided_offset(const StridedIO& io, long long grp, int kz, int idx) {
  const int hi = idx >> io.split_shift, lo = idx & ((1 << io.split_shift) - 1);
  return grp * io.plane_stride + kz + hi * io.split_stride + lo * io.line_stride;
}
// strided_offset(io, grp, kz, t + c) = strided_base(io, grp, kz, t) + strided_step(io, c)
// for c a multiple of T (T and split are powers of two, so one divides the other); the
// step part is the same for all threads of a block.
EVX_HD long long strided_base(const StridedIO& io, long long grp, int kz, int t) {
  return strided_offset(io, grp, kz, t);
}
EVX_HD long long strided_step(const StridedIO& io, int c) {
  const int hi = c >> io.split_shift, lo = c & ((1 << io.split_shift) - 1);
  return hi * io.split_stride + lo * io.line_stride;
}
// host: tabulate the eight per-element steps of a pass over lines of length L
inline void finalize_strided(StridedParams& p, int L) {
  for (int e = 0; e < 8; ++e) {
    p.src_step[e] = strided_step(p.src, e * (L / 8));
    p.dst_step[e] = strided_step(p.dst, e * (L / 8));
  }
  for (int it = 0; it < 4; ++it) p.src_pf_step[it] = strided_step(p.src, it * (L / 4));
}

// weight of the fused x pass applied to the 8 points t + e*T of a line (kother: index along the
// other strided axis, kz: column); shared by every form of the pass so that they agree bit for bit
template <int MODE, int T>
EVX_HD void xmid_apply_filter(cf* v, int t, int kother, int kz, const FilterParams& f) {
  const float k1 = wavenumber(signed_freq(kother, f.n1), f.inv_len1);
  const float k2 = wavenumber(kz, f.inv_len2);
  const float k12 = fma_rn(k1, k1, fmul_rn(k2, k2));
  const float s0 = 6.283185307179586f * f.inv_len0;
  if (MODE == PASS_XMID_ETD1 || f.n0 != 8 * T) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float k0 = fmul_rn(s0, (float)signed_freq(t + e * T, f.n0));
      v[e] = cscale(v[e], xpass_weight<MODE == PASS_XMID_ETD1>(k0, k12, f));
    }
    return;
  }
  // IMEX weight of the points t + e*T of a line of 8 T points, two at a time on the packed
  // FP32 pipe.  The signed frequency of point t + e*T is t + (e < 4 ? e : e - 8) * T, and
  // float(t) + that constant is exact, so the values (and every rounding step: the operations
  // are those of xpass_weight, lane by lane) equal the scalar sequence above bit for bit.
  const f2 ft = f2_splat((float)t), s02 = f2_splat(s0), k122 = f2_splat(k12);
  const f2 dt2 = f2_splat(f.dt), coef2 = f2_splat(f.coef), one2 = f2_splat(1.0f), scale2 = f2_splat(f.scale);
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    const f2 off = f2{(float)((e < 4 ? e : e - 8) * T), (float)((e + 1 < 4 ? e + 1 : e + 1 - 8) * T)};
    const f2 k0 = f2_mul(s02, f2_add(ft, off));
    const f2 kk = f2_fma(k0, k0, k122);
    const f2 kp = f.power == 2 ? f2_mul(kk, kk) : kk;
    const f2 den = f2_fma(dt2, f2_mul(coef2, kp), one2);
    const f2 w = f2_mul(f2{fdiv_fast(f.dt, den.a), fdiv_fast(f.dt, den.b)}, scale2);
    v[e] = cscale(v[e], w.a);
    v[e + 1] = cscale(v[e + 1], w.b);
  }
}

template <int L, int KZ, int MODE>
struct StridedPass {
  static constexpr int T = L / 8;
  static constexpr int NTHREADS = T * KZ;
  static constexpr int S = num_stages(L);
  static constexpr int NPHASES = pass_is_xmid(MODE) ? 2 * S - 1 : (MODE == PASS_COPY ? 1 : S);
  static constexpr int LP = smem_padded_len(L);
  static constexpr size_t SMEM_BYTES =
      (S > 1 && MODE != PASS_COPY ? 2 : 0) * (size_t)LP * KZ * sizeof(cf);

  struct Regs {
    cf v[8];
    cf w[3];               // roots of the next twiddled stage (fetched before the barrier)
    int t, cl;
    bool valid;
    long long grp;
    long long src_base, dst_base;
    long long src_base_m;  // mirror mode: offset of row T-1-t (rows L-1-(t+eT) = (7-e)T + T-1-t)
    int kz, kother;
  };

  EVX_HD static cf* buf(cf* smem, int which) { return smem + (size_t)which * LP * KZ; }

  EVX_HD static void init(Regs& r, const StridedParams& p, int tid, long long block) {
    const long long c = block * KZ + tid % KZ;
    const long long grp = c / p.P;
    init_at(r, p, tid, c, grp, (int)(c - grp * p.P));
  }
  // same with the (group, kz) decomposition of the thread's column c supplied by the caller
  // (the persistent kernels step it from tile to tile without dividing)
  EVX_HD static void init_at(Regs& r, const StridedParams& p, int tid, long long c, long long grp, int kz) {
    r.cl = tid % KZ;
    r.t = tid / KZ;
    r.kz = kz;
    r.kother = (int)grp + p.kother_offset;
    r.grp = grp;
    r.valid = c < p.ncols_total && r.kz < p.ncols_valid;
    r.src_base = strided_base(p.src, grp, r.kz, r.t);
    r.dst_base = strided_base(p.dst, grp, r.kz, r.t);
    r.src_base_m = p.mirror ? strided_base(p.src, grp, r.kz, T - 1 - r.t) : 0;
  }

  template <int DIR>
  EVX_HD static void write_stage(Regs& r, cf* b, int s) {
    cf* base = b + smem_pad(stage_out_base<L>(s, r.t)) * KZ + r.cl;
#pragma unroll
    for (int e = 0; e < 8; ++e) base[smem_pad(stage_out_const<L>(s, e)) * KZ] = r.v[e];
  }
  EVX_HD static void read_natural(Regs& r, const cf* b) {
    const cf* base = b + smem_pad(r.t) * KZ + r.cl;
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = base[smem_pad(e * T) * KZ];
  }
  EVX_HD static void load_global(Regs& r, const StridedParams& p) {
    const cf* base = p.in + r.src_base;
    if (p.mirror) {          // lower half from memory, upper half = +- the lower half reversed
      const cf* mbase = p.in + r.src_base_m;
      const float sg = (float)p.mirror;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r.v[e] = r.valid ? base[p.src_step[e]] : cf{0.f, 0.f};
        const cf m = r.valid ? mbase[p.src_step[3 - e]] : cf{0.f, 0.f};
        r.v[4 + e] = cf{sg * m.x, sg * m.y};
      }
      return;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
      r.v[e] = r.valid ? base[p.src_step[e]] : cf{0.f, 0.f};
  }
  EVX_HD static void store_global(Regs& r, const StridedParams& p) {
    if (!r.valid) return;
    if (p.mirror) {          // the mirror image of the result is implied, only rows < L/2 exist
      cf* base = p.out + r.dst_base;
#pragma unroll
      for (int e = 0; e < 4; ++e) base[p.dst_step[e]] = r.v[e];
      return;
    }
    if (p.use_peers) {
      const long long within = p.dst_peer_base + r.grp * p.dst.plane_stride + r.kz;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = r.t + e * T;
        const int hi = idx >> p.dst.split_shift, lo = idx & ((1 << p.dst.split_shift) - 1);
        p.out_peers[hi][within + lo * p.dst.line_stride] = r.v[e];
      }
      return;
    }
    cf* base = p.out + r.dst_base;
#pragma unroll
    for (int e = 0; e < 8; ++e) base[p.dst_step[e]] = r.v[e];
  }
  EVX_HD static void apply_filter(Regs& r, const StridedParams& p) {
    xmid_apply_filter<MODE, T>(r.v, r.t, r.kother, r.kz, p.filt);
  }

  // transform step q of the pass: FWD/INV: stage q; XMID: q < S forward stage q, else inverse q-S
  template <int DIRSEL>
  EVX_HD static void compute(Regs& r, const StridedParams& p, int stage) {
    line_stage_compute_pre<L, DIRSEL>(stage, r.v, r.t, r.w);
  }
  // roots for `stage` (no-op for stage 0 and past the last stage)
  EVX_HD static void fetch_tw(Regs& r, const StridedParams& p, int stage) {
    stage_twiddles<L>(stage, r.t, p.tw, r.w);
  }

  EVX_HD static void phase(int k, Regs& r, cf* smem, const StridedParams& p) {
    if (MODE == PASS_COPY) {
      load_global(r, p);
      store_global(r, p);
    } else if (MODE == PASS_FWD || MODE == PASS_INV) {
      if (k == 0) load_global(r, p); else read_natural(r, buf(smem, (k - 1) & 1));
      if (MODE == PASS_FWD) compute<-1>(r, p, k); else compute<+1>(r, p, k);
      if (k == S - 1) store_global(r, p);
      else if (MODE == PASS_FWD) write_stage<-1>(r, buf(smem, k & 1), k);
      else write_stage<+1>(r, buf(smem, k & 1), k);
      fetch_tw(r, p, k + 1);
    } else {
      // XMID: phases 0..S-2 forward stages, phase S-1: last forward + filter + first inverse,
      // phases S..2S-2 remaining inverse stages
      if (k == 0) load_global(r, p); else read_natural(r, buf(smem, (k - 1) & 1));
      if (k < S - 1) {
        compute<-1>(r, p, k);
        write_stage<-1>(r, buf(smem, k & 1), k);
        fetch_tw(r, p, k + 1);
      } else if (k == S - 1) {
        compute<-1>(r, p, S - 1);
        apply_filter(r, p);
        compute<+1>(r, p, 0);
        if (S == 1) store_global(r, p); else write_stage<+1>(r, buf(smem, k & 1), 0);
        fetch_tw(r, p, 1);
      } else {
        const int s = k - (S - 1);
        compute<+1>(r, p, s);
        if (s == S - 1) store_global(r, p); else write_stage<+1>(r, buf(smem, k & 1), s);
        fetch_tw(r, p, s + 1);
      }
    }
  }
};

// ------------------------------------------------------------------------------------
// Pipelined (persistent) form of the strided passes.
//
// A block walks tiles b, b + gridDim, ... and keeps THREE padded line buffers: while it
// transforms tile i out of buffer A it has the asynchronous copy (cp.async / LDGSTS) of
// tile i+1 in flight into buffer B; the stage exchanges alternate between the third buffer
// C and A (A is free once every thread has picked up its stage-0 inputs).  Next tile: A and
// B swap.  Memory latency is thereby hidden inside one block instead of relying on other
// resident blocks, and two blocks (2 x 111 KB for L = 512) still fit an SM.
// ------------------------------------------------------------------------------------
template <int L, int KZ, int MODE>
struct StridedPipe {
  using Base = StridedPass<L, KZ, MODE>;
  using Regs = typename Base::Regs;
  static constexpr int T = Base::T, S = Base::S, NTHREADS = Base::NTHREADS, NPHASES = Base::NPHASES;
  static constexpr int LP = Base::LP;
  static constexpr int BUF = LP * KZ;                       // cf elements per buffer
  static constexpr size_t SMEM_BYTES = 3 * (size_t)BUF * sizeof(cf);
  static constexpr int CHUNKS_PER_ROW = KZ * (int)sizeof(cf) / 16;
  static constexpr int CHUNKS = L * CHUNKS_PER_ROW;
  static_assert(KZ * sizeof(cf) % 16 == 0, "rows must be 16-byte multiples");

  EVX_HD static long long num_tiles(const StridedParams& p) { return (p.ncols_total + KZ - 1) / KZ; }

  // Column position of a block's current tile, stepped by gridDim.x tiles per iteration with
  // an add-and-carry instead of the two 64-bit divisions per tile a direct decode costs
  // (they sat right behind the tile barrier, where every warp waits for them).
  struct Cursor {
    int grp, kz;             // own column of the current tile: c = grp*P + kz
    int bgrp, bkz;           // first column of the NEXT tile to prefetch, same decomposition
    int step_grp, step_kz;   // gridDim.x * KZ columns, decomposed
  };
  EVX_HD static void cursor_init(Cursor& q, const StridedParams& p, int tid, long long tile,
                                 long long nblocks) {
    const long long step = nblocks * KZ, c = tile * KZ + tid % KZ, bc = tile * KZ;
    q.step_grp = (int)(step / p.P);
    q.step_kz = (int)(step - (long long)q.step_grp * p.P);
    q.grp = (int)(c / p.P);
    q.kz = (int)(c - (long long)q.grp * p.P);
    q.bgrp = (int)(bc / p.P);
    q.bkz = (int)(bc - (long long)q.bgrp * p.P);
  }
  EVX_HD static long long cursor_column(const Cursor& q, const StridedParams& p) {
    return (long long)q.grp * p.P + q.kz;
  }
  EVX_HD static void cursor_step_own(Cursor& q, const StridedParams& p) {
    q.grp += q.step_grp; q.kz += q.step_kz;
    if (q.kz >= p.P) { q.kz -= p.P; ++q.grp; }
  }
  EVX_HD static void cursor_step_base(Cursor& q, const StridedParams& p) {
    q.bgrp += q.step_grp; q.bkz += q.step_kz;
    if (q.bkz >= p.P) { q.bkz -= p.P; ++q.bgrp; }
  }

  // all threads: enqueue the copy of tile `tile` (dense [L][KZ] rows of KZ*8 bytes) into dst
  EVX_HD static void prefetch(int tid, const StridedParams& p, long long tile, cf* dst) {
    const long long c0 = tile * KZ;
    const long long grp = c0 / p.P;
    prefetch_at(tid, p, grp, (int)(c0 - grp * p.P), dst);
  }
  // Every thread moves PF_ITERS = 4 16-byte chunks of the tile: chunk q = tid + it * NTHREADS is
  // (row0 + it * L/4, part) with row0 = tid / CHUNKS_PER_ROW < L/4.  Because L/4 and the split
  // of the addressing are powers of two, the row part of the offset is additive,
  //   offset(row0 + it * L/4) = offset(row0) + strided_step(src, it * L/4)
  // (same argument as strided_base / strided_step), so one 64-bit address is computed per tile
  // and the four copies use host-tabulated steps - the per-chunk division and the two 64-bit
  // multiplies of the direct form cost 19 instructions per copy in an issue-bound loop.
  static constexpr int PF_ITERS = CHUNKS / NTHREADS;
  static constexpr int PF_ROWS = NTHREADS / CHUNKS_PER_ROW;
  static_assert(PF_ITERS == 4 && CHUNKS % NTHREADS == 0 && NTHREADS % CHUNKS_PER_ROW == 0 &&
                PF_ROWS * 4 == L, "tile prefetch: four chunks per thread, L/4 rows apart");
  EVX_HD static void prefetch_at(int tid, const StridedParams& p, long long grp, int kz0, cf* dst) {
    constexpr int CF16 = 16 / (int)sizeof(cf);
    const int row0 = tid / CHUNKS_PER_ROW, part = tid - row0 * CHUNKS_PER_ROW;
    const cf* src = p.in + strided_offset(p.src, grp, kz0, row0) + part * CF16;
    cf* d = dst + (size_t)row0 * KZ + part * CF16;
#pragma unroll
    for (int it = 0; it < PF_ITERS; ++it)
      async_copy16(d + (size_t)it * PF_ROWS * KZ, src + p.src_pf_step[it]);
  }
  // direct form of the same copies (one address computation per chunk): kept for the CPU
  // replay, which checks that both forms touch the same (destination, source) pairs
  EVX_HD static void prefetch_at_direct(int tid, const StridedParams& p, long long grp, int kz0, cf* dst) {
    for (int q = tid; q < CHUNKS; q += NTHREADS) {
      const int row = q / CHUNKS_PER_ROW, part = q - row * CHUNKS_PER_ROW;
      const cf* src = p.in + strided_offset(p.src, grp, kz0, row) + part * (16 / (int)sizeof(cf));
      async_copy16(dst + (size_t)row * KZ + part * (16 / (int)sizeof(cf)), src);
    }
  }

  EVX_HD static void read_tile(Regs& r, const cf* a) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = a[(size_t)(r.t + e * T) * KZ + r.cl];
  }

  // phase 0 is split around the prefetch of the next tile: `pick` then `phase(0, ..)`
  EVX_HD static void phase(int k, Regs& r, cf* a, cf* c, const StridedParams& p) {
    // exchange buffers alternate C, A, C, ...: phase k writes (k even ? C : A)
    cf* wr = (k & 1) ? a : c;
    const cf* rd = (k & 1) ? c : a;        // written by phase k-1
    if (k > 0) Base::read_natural(r, rd);
    if (MODE == PASS_FWD || MODE == PASS_INV) {
      if (MODE == PASS_FWD) Base::template compute<-1>(r, p, k); else Base::template compute<+1>(r, p, k);
      if (k == S - 1) Base::store_global(r, p);
      else if (MODE == PASS_FWD) Base::template write_stage<-1>(r, wr, k);
      else Base::template write_stage<+1>(r, wr, k);
      Base::fetch_tw(r, p, k + 1);
    } else {
      if (k < S - 1) {
        Base::template compute<-1>(r, p, k);
        Base::template write_stage<-1>(r, wr, k);
        Base::fetch_tw(r, p, k + 1);
      } else if (k == S - 1) {
        Base::template compute<-1>(r, p, S - 1);
        Base::apply_filter(r, p);
        Base::template compute<+1>(r, p, 0);
        if (S == 1) Base::store_global(r, p); else Base::template write_stage<+1>(r, wr, 0);
        Base::fetch_tw(r, p, 1);
      } else {
        const int s = k - (S - 1);
        Base::template compute<+1>(r, p, s);
        if (s == S - 1) Base::store_global(r, p); else Base::template write_stage<+1>(r, wr, s);
        Base::fetch_tw(r, p, s + 1);
      }
    }
  }
};

// ------------------------------------------------------------------------------------
// z passes (contiguous axis, real <-> half spectrum)
// ------------------------------------------------------------------------------------
struct ZParams {
  const float* real_in;    // ZFwd: r [rows][nz]            ZInv: u (may be null: out = update)
  float* real_out;         // ZInv: out [rows][nz]
  cf* spec;                // [rows][P]
  const cf* tw;            // W_M table, M = nz/2
  const cf* twr;           // W_nz[k], k = 0..M (untangle roots)
  long long rows;          // nx*ny
  int nz, P;
  int pf_blocks;           // > 0: L2 prefetch distance of the z kernels, in blocks (device only)
};

template <int M, int NL, bool INVERSE>
struct ZPass {
  static constexpr int T = M / 8;
  static constexpr int NTHREADS = T * NL;
  static constexpr int S = num_stages(M);
  static constexpr int NPHASES = S + 1;
  static constexpr int LP = zline_len<M>();
  static constexpr size_t SMEM_BYTES = 2 * (size_t)LP * NL * sizeof(cf);

  struct Regs {
    cf v[8];
    cf w[3];                 // roots of the next twiddled stage
    cf u[INVERSE ? 8 : 1];   // ZInv: the u values added at the end, fetched up front
    int t, l;
    bool valid;
    long long row;
  };

  EVX_HD static cf* buf(cf* smem, int which, int l) { return smem + ((size_t)which * NL + l) * LP; }

  EVX_HD static void init(Regs& r, const ZParams& p, int tid, long long block) {
    r.t = tid % T;
    r.l = tid / T;
    r.row = block * NL + r.l;
    r.valid = r.row < p.rows;
  }

  EVX_HD static void write_stage(Regs& r, cf* b, int s) {
#pragma unroll
    for (int e = 0; e < 8; ++e) b[zline_idx<M>(line_stage_out_index<M>(s, r.t, e))] = r.v[e];
  }
  EVX_HD static void read_natural(Regs& r, const cf* b) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = b[zline_idx<M>(r.t + e * T)];
  }

  // X[k] from Z[k], Z[M-k]  (forward untangle)
  EVX_HD static cf untangle_fwd(cf zk, cf zmk, cf w) {
    const cf a = cadd(zk, cconj(zmk));              // 2 E[k]
    const cf b = csub(zk, cconj(zmk));              // 2 i O[k]
    const cf o = cf{b.y, -b.x};                     // -i * b = 2 O[k]
    return cscale(cadd(a, cmul(w, o)), 0.5f);
  }
  // Z'[k] from X[k], X[M-k]  (inverse; unnormalised: ifft_M(Z') = N x)
  EVX_HD static cf untangle_inv(cf xk, cf xmk, cf w) {
    const cf a = cadd(xk, cconj(xmk));              // 2 E[k]
    const cf b = csub(xk, cconj(xmk));
    const cf wo = cmul(cconj(w), b);                // 2 O[k]
    return cadd(a, cf{-wo.y, wo.x});                // a + i * wo
  }

  EVX_HD static void phase(int k, Regs& r, cf* smem, const ZParams& p) {
    if (!INVERSE) {
      if (k < S) {
        if (k == 0) {
          const cf* line = reinterpret_cast<const cf*>(p.real_in + r.row * p.nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) r.v[e] = r.valid ? line[r.t + e * T] : cf{0.f, 0.f};
        } else {
          read_natural(r, buf(smem, (k - 1) & 1, r.l));
        }
        line_stage_compute_pre<M, -1>(k, r.v, r.t, r.w);
        write_stage(r, buf(smem, k & 1, r.l), k);     // last stage lands in natural order
        stage_twiddles<M>(k + 1, r.t, p.tw, r.w);
      } else {
        const cf* z = buf(smem, (S - 1) & 1, r.l);
        if (!r.valid) return;
        cf* out = p.spec + r.row * p.P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = r.t + e * T;
          const cf zk = z[zline_idx<M>(kk)];
          const cf zmk = z[zline_idx<M>(kk == 0 ? 0 : M - kk)];
          out[kk] = untangle_fwd(zk, zmk, p.twr[kk]);
        }
        if (r.t == 0) {
          const cf z0 = z[zline_idx<M>(0)];
          out[M] = cf{z0.x - z0.y, 0.f};
        }
      }
    } else {
      if (k == 0) {
        cf* x = buf(smem, 1, r.l);
        const cf* in = p.spec + r.row * p.P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = r.t + e * T;
          x[zline_idx<M>(kk)] = r.valid ? in[kk] : cf{0.f, 0.f};
        }
        if (r.t == 0) x[zline_idx<M>(M)] = r.valid ? in[M] : cf{0.f, 0.f};
        if (p.real_in && r.valid) {
          const cf* u = reinterpret_cast<const cf*>(p.real_in + r.row * p.nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = u[r.t + e * T];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = cf{0.f, 0.f};
        }
      } else {
        const int s = k - 1;
        if (s == 0) {
          const cf* x = buf(smem, 1, r.l);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kk = r.t + e * T;
            r.v[e] = untangle_inv(x[zline_idx<M>(kk)], x[zline_idx<M>(M - kk)], p.twr[kk]);
          }
        } else {
          read_natural(r, buf(smem, (s - 1) & 1, r.l));
        }
        line_stage_compute_pre<M, +1>(s, r.v, r.t, r.w);
        if (s < S - 1) {
          write_stage(r, buf(smem, s & 1, r.l), s);
          stage_twiddles<M>(s + 1, r.t, p.tw, r.w);
        } else if (r.valid) {
          cf* out = reinterpret_cast<cf*>(p.real_out + r.row * p.nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) out[r.t + e * T] = cadd(r.v[e], r.u[e]);
        }
      }
    }
  }
};

}  // namespace evx
