// z lines (512 reals <-> 257 complex) in the "two lines per 64-thread group" form used by the
// chained kernels (fft_chain.cu).  Host/device, barrier-free phases.
//
// Same arithmetic as ZPass<256, .> in fft_pass_core.h - radix 4, 8, 8 Stockham stages of the
// half-length complex transform with the same roots, the same untangle formulas - hence
// bit-identical results, but organised like the TMA-tiled strided passes (fft_line_core.h):
//
//  * The rows of an item sit in shared memory in natural order at a pitch of 264 complex
//    (2112 bytes = 64 mod 128): a half-warp is 8 consecutive t x the 2 lines of a group, so
//    natural-order, reversed-order (untangle partner) and stage-1 accesses of both lines are
//    64-bit conflict-free (stage 1: two-way) WITHOUT an address swizzle - ZPass spends 37 % of
//    its instructions on the XOR swizzle of its one-line-per-warp layout.
//  * Every shared-memory access is one per-thread base plus an immediate.
//  * The stage-0 exchange goes through a padded per-group buffer X (index i + i/8, lines
//    interleaved), stage 1 writes in place into the rows, so one group barrier per exchange.
//  * Inputs arrive by bulk copy (the caller), outputs leave with 64-bit stores.
#pragma once
#include "fft_pass_core.h"

namespace evx {

struct ZGroupParams {
  const cf* tw;            // W_256
  const cf* twr;           // W_512[k], k = 0..256
  int nz, P;               // 512, pitch of the spectrum rows (264)
};

template <bool INVERSE>
struct ZGroupLine {
  static constexpr int M = 256, T = 32, G = 2, GT = T * G, S = 3;
  static constexpr int ROWP = 264;                         // cf per row of the row buffer
  static constexpr int XG = (M - 1 + ((M - 1) >> 3) + 1) * G;   // cf per group in X (574)
  static constexpr int NPHASES = INVERSE ? 3 : 4;
  static_assert(num_stages(M) == 3 && first_radix(M) == 4, "radix 4, 8, 8");

  struct Regs {
    cf v[8];
    cf w[3];
    cf u[8];               // inverse: the u values added at the end; forward: the untangle roots
    int t, c2, g;
    int rb;                // cf index in the row buffer of element t of the own row
    int pb;                // ... of element M - t (untangle partner of k = t)
    int s1;                // ... of the stage-1 output base (t/4)*32 + t%4
    int xn, xs;            // cf index in X: natural-order base, stage-0 output base
  };

  EVX_HD static void init(Regs& r, int tid) {
    r.g = tid / GT;
    const int tg = tid - r.g * GT;
    r.c2 = tg % G;
    r.t = tg / G;
    const int row = (r.g * G + r.c2) * ROWP;
    r.rb = row + r.t;
    r.pb = row + (M - r.t);
    r.s1 = row + (r.t / 4) * 32 + r.t % 4;
    r.xn = smem_pad(r.t) * G + r.c2;
    r.xs = (4 * r.t + (r.t >> 1)) * G + r.c2;       // smem_pad(4 t): (4t + r) >> 3 == t >> 1 for r < 4
  }

  // stage-0 output (radix 4, two butterflies per thread): v[i + 2r] belongs at 4 (t + 32 i) + r,
  // padded: 4t + (t >> 1) + r + 144 i
  EVX_HD static void write_x_stage0(const Regs& r, cf* xg) {
#pragma unroll
    for (int e = 0; e < 8; ++e) xg[r.xs + ((e >> 1) + 144 * (e & 1)) * G] = r.v[e];
  }
  EVX_HD static void read_x_natural(Regs& r, const cf* xg) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = xg[r.xn + smem_pad(e * T) * G];
  }
  EVX_HD static void read_rows_natural(Regs& r, const cf* rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = rows[r.rb + e * T];
  }
  EVX_HD static void write_rows_natural(const Regs& r, cf* rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) rows[r.rb + e * T] = r.v[e];
  }
  // stage-1 output (radix 8, Ns = 4): v[e] belongs at (t/4)*32 + t%4 + 4 e
  EVX_HD static void write_rows_stage1(const Regs& r, cf* rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) rows[r.s1 + 4 * e] = r.v[e];
  }

  // The same phases with all roots supplied by the caller (fetched once per kernel instead of
  // once per item): w1 / w2 = stage_twiddles<M>(1 / 2, t, tw, .), root[e] = twr[t + 32 e].
  template <bool ROOTS_IN_REGS>
  EVX_HD static void phase_tw(int k, Regs& r, cf* rows, cf* xg, const cf* w1, const cf* w2, const cf* root,
                              const cf* twr, int nz, int P, long long grow, const float* real_in,
                              float* real_out, cf* spec) {
    if (!INVERSE) {
      if (k == 0) {
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, -1>(0, r.v, r.t, w1);
        write_x_stage0(r, xg);
      } else if (k == 1) {
        read_x_natural(r, xg);
        line_stage_compute_pre<M, -1>(1, r.v, r.t, w1);
        write_rows_stage1(r, rows);
      } else if (k == 2) {
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, -1>(2, r.v, r.t, w2);
        write_rows_natural(r, rows);
        if (r.t == 0) rows[r.rb + M] = r.v[0];
        if (!ROOTS_IN_REGS) {
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = twr[r.t + e * T];
        }
      } else {
        cf* out = spec + grow * P;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          out[r.t + e * T] = ZPass<M, 1, false>::untangle_fwd(r.v[e], rows[r.pb - e * T],
                                                              ROOTS_IN_REGS ? root[e] : r.u[e]);
        if (r.t == 0) out[M] = cf{r.v[0].x - r.v[0].y, 0.f};
      }
    } else {
      if (k == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          r.v[e] = ZPass<M, 1, true>::untangle_inv(rows[r.rb + e * T], rows[r.pb - e * T],
                                                   ROOTS_IN_REGS ? root[e] : twr[r.t + e * T]);
        line_stage_compute_pre<M, +1>(0, r.v, r.t, w1);
        write_x_stage0(r, xg);
      } else if (k == 1) {
        read_x_natural(r, xg);
        line_stage_compute_pre<M, +1>(1, r.v, r.t, w1);
        write_rows_stage1(r, rows);
      } else {
        if (real_in) {
          const cf* u = reinterpret_cast<const cf*>(real_in + grow * nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = u[r.t + e * T];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = cf{0.f, 0.f};
        }
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, +1>(2, r.v, r.t, w2);
        cf* out = reinterpret_cast<cf*>(real_out + grow * nz);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[r.t + e * T] = cadd(r.v[e], r.u[e]);
      }
    }
  }

  // `rows`: the item's row buffer (input rows already there), `xg`: the group's exchange buffer,
  // `grow`: global row index of the thread's line.  Caller synchronises the GROUP between phases.
  EVX_HD static void phase(int k, Regs& r, cf* rows, cf* xg, const ZGroupParams& p, long long grow,
                           const float* real_in, float* real_out, cf* spec) {
    if (!INVERSE) {
      if (k == 0) {
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, -1>(0, r.v, r.t, r.w);
        write_x_stage0(r, xg);
        stage_twiddles<M>(1, r.t, p.tw, r.w);
      } else if (k == 1) {
        read_x_natural(r, xg);
        line_stage_compute_pre<M, -1>(1, r.v, r.t, r.w);
        write_rows_stage1(r, rows);
        stage_twiddles<M>(2, r.t, p.tw, r.w);
      } else if (k == 2) {
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, -1>(2, r.v, r.t, r.w);
        write_rows_natural(r, rows);               // own slots: no barrier needed before
        if (r.t == 0) rows[r.rb + M] = r.v[0];     // Z[M] := Z[0], the partner of k = 0
        // roots of the untangle step, fetched before the barrier (and before the stores below,
        // which the compiler must assume to alias the table)
#pragma unroll
        for (int e = 0; e < 8; ++e) r.u[e] = p.twr[r.t + e * T];
      } else {
        cf* out = spec + grow * p.P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = r.t + e * T;
          out[kk] = ZPass<M, 1, false>::untangle_fwd(r.v[e], rows[r.pb - e * T], r.u[e]);
        }
        if (r.t == 0) out[M] = cf{r.v[0].x - r.v[0].y, 0.f};
      }
    } else {
      if (k == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = r.t + e * T;
          r.v[e] = ZPass<M, 1, true>::untangle_inv(rows[r.rb + e * T], rows[r.pb - e * T], p.twr[kk]);
        }
        line_stage_compute_pre<M, +1>(0, r.v, r.t, r.w);
        write_x_stage0(r, xg);
        stage_twiddles<M>(1, r.t, p.tw, r.w);
      } else if (k == 1) {
        read_x_natural(r, xg);
        line_stage_compute_pre<M, +1>(1, r.v, r.t, r.w);
        write_rows_stage1(r, rows);
        stage_twiddles<M>(2, r.t, p.tw, r.w);
      } else {
        // the u row comes from L2 (the loader prefetches it when it issues the item); fetched
        // here, first thing of the last phase, its eight values are live for this phase only
        if (real_in) {
          const cf* u = reinterpret_cast<const cf*>(real_in + grow * p.nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = u[r.t + e * T];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = cf{0.f, 0.f};
        }
        read_rows_natural(r, rows);
        line_stage_compute_pre<M, +1>(2, r.v, r.t, r.w);
        cf* out = reinterpret_cast<cf*>(real_out + grow * p.nz);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[r.t + e * T] = cadd(r.v[e], r.u[e]);
      }
    }
  }
};

}  // namespace evx
