// Mixed-radix Stockham passes for extents that are not powers of two (host/device).
//
// The radix-8 programs of fft_pass_core.h cover 2^k extents - the bench configurations.  The
// reference's own example grid is 100^3 (README.md:98-115), its tests use odd anisotropic
// shapes; those used to fall back to cuFFT.  This file gives them hand-written passes as
// well: any extent whose prime factors are <= 7, float32 and float64, including degenerate
// (size-1) axes.  Same five-pass structure and spectrum layout S[nx][ny][P] as the radix-8
// path (z forward, y forward, x forward * weight * x inverse, y inverse, z inverse + u);
// correctness first - these grids are small, the passes are not tuned.
//
// One block transforms W interleaved lines held in shared memory ([i][w] layout, two
// buffers, autosort => natural order after the last stage).  A stage of radix r does
// N/r butterflies per line; butterfly j multiplies input k by W_N^(m k), m = (j mod Ns) *
// N/(Ns r), takes a direct r-point DFT (roots looked up in the same W_N table) and writes
// output q to (j / Ns) Ns r + j mod Ns + q Ns.  The z passes transform the real line as a
// full complex line (the half-spectrum trick needs even nz; odd nz must work too).
#pragma once
#include "evx_hd.h"
#include "spectral_math.h"

namespace evx {

template <typename R>
struct gcplx {
  R x, y;
};
template <typename R>
EVX_HD gcplx<R> gmul(gcplx<R> a, gcplx<R> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

constexpr int kMaxStages = 12;
struct LineDesc {
  int N;                    // line length
  int nstages;
  int radix[kMaxStages];    // product = N; each in {2,3,4,5,7,8}
};

// host: factor n into radices (largest first); false if a prime factor > 7 remains
inline bool factor_line(int n, LineDesc& d) {
  d.N = n;
  d.nstages = 0;
  const int cand[6] = {8, 4, 2, 3, 5, 7};
  int m = n;
  for (int c = 0; c < 6; ++c)
    while (m % cand[c] == 0 && m > 1) {
      if (d.nstages >= kMaxStages) return false;
      d.radix[d.nstages++] = cand[c];
      m /= cand[c];
    }
  return m == 1;
}

enum : int { GEN_Z_FWD = 0, GEN_Z_INV = 1, GEN_FWD = 2, GEN_INV = 3, GEN_XMID = 4 };

template <typename R>
struct GenericParams {
  int mode;
  LineDesc line;
  const gcplx<R>* tw;       // W_N[m] = exp(-2 pi i m / N), m < N
  int W;                    // interleaved lines per block
  // z passes: rows of the real field <-> rows of the spectrum
  const R* real_in;         // GEN_Z_FWD: r;  GEN_Z_INV: u (may be null)
  R* real_out;              // GEN_Z_INV
  long long rows;           // nx * ny
  int nz, P;                // real row length, spectrum row pitch
  gcplx<R>* spec;
  // strided passes: column c = group * P + kz; element i of it at
  //   spec[group * group_stride + kz + i * line_stride]
  long long line_stride, group_stride, ncols_total;
  int ncols_valid;          // kz < ncols_valid is data, the rest of a row is padding
  FilterParams filt;        // GEN_XMID: n0 = this axis, n1 = the group axis, n2 = nz
};

template <typename R>
struct GenericProgram {
  using C = gcplx<R>;
  using P = GenericParams<R>;

  // phases: 0 load, 1..S stages (forward), [XMID: S+1 weight, S+2..2S+1 inverse stages], last store
  EVX_HD static int nphases(const P& p) {
    const int S = p.line.nstages;
    return p.mode == GEN_XMID ? 2 * S + 3 : S + 2;
  }
  EVX_HD static size_t smem_elems(const P& p) { return 2 * (size_t)p.line.N * p.W; }

  EVX_HD static C root(const P& p, long long idx, int dir) {
    const C w = p.tw[idx % p.line.N];
    return dir < 0 ? w : C{w.x, -w.y};
  }

  // one Stockham stage for all W lines of the block; in/out are [N][W]
  EVX_HD static void stage(const P& p, const C* in, C* out, int s, int dir, int tid, int nthreads) {
    const int N = p.line.N, W = p.W, r = p.line.radix[s];
    int Ns = 1;
    for (int q = 0; q < s; ++q) Ns *= p.line.radix[q];
    const int nb = N / r;
    for (int item = tid; item < nb * W; item += nthreads) {
      const int c = item % W, j = item / W;
      const long long m = (long long)(j % Ns) * (N / (Ns * r));
      C x[8];
      for (int k = 0; k < r; ++k) x[k] = gmul(in[(size_t)(j + k * nb) * W + c], root(p, m * k, dir));
      const int base = (j / Ns) * Ns * r + j % Ns;
      for (int q = 0; q < r; ++q) {
        C acc = x[0];
        for (int k = 1; k < r; ++k) {
          const C t = gmul(x[k], root(p, (long long)k * q * nb, dir));
          acc.x += t.x;
          acc.y += t.y;
        }
        out[(size_t)(base + q * Ns) * W + c] = acc;
      }
    }
  }

  // where the data of a block sits after `n` stages (buffers alternate, loads go to buffer 0)
  EVX_HD static C* bufn(C* smem, const P& p, int n) { return smem + (size_t)(n & 1) * p.line.N * p.W; }

  EVX_HD static void load(const P& p, C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    for (int item = tid; item < N * W; item += nthreads) {
      C v{R(0), R(0)};
      if (p.mode == GEN_Z_FWD || p.mode == GEN_Z_INV) {
        const int i = item % N, c = item / N;          // i fastest: contiguous global reads
        const long long row = block * W + c;
        if (row < p.rows) {
          if (p.mode == GEN_Z_FWD) {
            v.x = p.real_in[row * p.nz + i];
          } else {                                       // Hermitian completion of the half row
            const int h = N / 2;
            if (i <= h) {
              v = p.spec[row * p.P + i];
            } else {
              const C w = p.spec[row * p.P + (N - i)];
              v = C{w.x, -w.y};
            }
          }
        }
        b[(size_t)i * W + c] = v;
      } else {
        const int c = item % W, i = item / W;           // columns fastest
        const long long col = block * W + c;
        const long long g = col / p.P;
        const int kz = (int)(col - g * p.P);
        if (col < p.ncols_total && kz < p.ncols_valid)
          v = p.spec[g * p.group_stride + kz + (long long)i * p.line_stride];
        b[(size_t)i * W + c] = v;
      }
    }
  }

  EVX_HD static void store(const P& p, const C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    if (p.mode == GEN_Z_FWD) {
      const int h = N / 2 + 1;
      for (int item = tid; item < h * W; item += nthreads) {
        const int i = item % h, c = item / h;
        const long long row = block * W + c;
        if (row < p.rows) p.spec[row * p.P + i] = b[(size_t)i * W + c];
      }
    } else if (p.mode == GEN_Z_INV) {
      for (int item = tid; item < N * W; item += nthreads) {
        const int i = item % N, c = item / N;
        const long long row = block * W + c;
        if (row < p.rows) {
          const R v = b[(size_t)i * W + c].x;
          p.real_out[row * p.nz + i] = p.real_in ? p.real_in[row * p.nz + i] + v : v;
        }
      }
    } else {
      for (int item = tid; item < N * W; item += nthreads) {
        const int c = item % W, i = item / W;
        const long long col = block * W + c;
        const long long g = col / p.P;
        const int kz = (int)(col - g * p.P);
        if (col < p.ncols_total && kz < p.ncols_valid)
          p.spec[g * p.group_stride + kz + (long long)i * p.line_stride] = b[(size_t)i * W + c];
      }
    }
  }

  // multiply the transformed lines by weight(k) * scale (natural order along the line)
  EVX_HD static void weight(const P& p, C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    const FilterParams& f = p.filt;
    for (int item = tid; item < N * W; item += nthreads) {
      const int c = item % W, i = item / W;
      const long long col = block * W + c;
      const long long g = col / p.P;
      const int kz = (int)(col - g * p.P);
      const float k0 = wavenumber(signed_freq(i, f.n0), f.inv_len0);
      const float k1 = wavenumber(signed_freq((int)g, f.n1), f.inv_len1);
      const float k2 = wavenumber(kz, f.inv_len2);
      const float ksq = fadd_rn(fadd_rn(fmul_rn(k0, k0), fmul_rn(k1, k1)), fmul_rn(k2, k2));
      const R w = (R)spectral_weight(ksq, f) * (sizeof(R) == 8 ? (R)f.scale_d : (R)f.scale);
      C& v = b[(size_t)i * W + c];
      v.x *= w;
      v.y *= w;
    }
  }

  EVX_HD static void phase(int k, const P& p, C* smem, long long block, int tid, int nthreads) {
    const int S = p.line.nstages;
    if (k == 0) {
      load(p, bufn(smem, p, 0), block, tid, nthreads);
      return;
    }
    const int dir1 = (p.mode == GEN_Z_INV || p.mode == GEN_INV) ? +1 : -1;
    if (k <= S) {
      stage(p, bufn(smem, p, k - 1), bufn(smem, p, k), k - 1, dir1, tid, nthreads);
      return;
    }
    if (p.mode != GEN_XMID) {
      store(p, bufn(smem, p, S), block, tid, nthreads);
      return;
    }
    if (k == S + 1) {
      weight(p, bufn(smem, p, S), block, tid, nthreads);
    } else if (k <= 2 * S + 1) {
      const int s = k - (S + 2);
      stage(p, bufn(smem, p, S + s), bufn(smem, p, S + s + 1), s, +1, tid, nthreads);
    } else {
      store(p, bufn(smem, p, 2 * S), block, tid, nthreads);
    }
  }
};

}  // namespace evx
