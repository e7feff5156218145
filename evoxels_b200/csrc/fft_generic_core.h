// Mixed-radix Stockham passes for extents that are not powers of two (host/device).
//
// The radix-8 programs of fft_pass_core.h cover 2^k extents - the bench configurations.  The
// reference's own example grid is 100^3 (README.md:98-115), its tests use odd anisotropic
// shapes; those used to fall back to cuFFT.  This file gives them hand-written passes as
// well: any extent whose prime factors are <= 7, float32 and float64, including degenerate
// (size-1) axes.  Same five-pass structure and spectrum layout S[nx][ny][P] as the radix-8
// path (z forward, y forward, x forward * weight * x inverse, y inverse, z inverse + u);
// radices are compile-time (2, 4, 8 as register butterflies, 3 / 5 / 7 as direct sums with the
// r roots in registers), twiddle indices need no reduction, all index arithmetic is 32-bit.
//
// One block transforms W interleaved lines held in shared memory ([i][w] layout, two
// buffers, autosort => natural order after the last stage).  A stage of radix r does
// N/r butterflies per line; butterfly j multiplies input k by W_N^(m k), m = (j mod Ns) *
// N/(Ns r), takes a direct r-point DFT (roots looked up in the same W_N table) and writes
// output q to (j / Ns) Ns r + j mod Ns + q Ns.  The z passes transform TWO real rows as one
// complex line (row A in the real part, row B in the imaginary part; the spectra are separated
// / recombined with the Hermitian symmetry while they are stored / loaded), which works for odd
// nz as well.
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

// unsigned division by a run-time constant: q = (x * m) >> 32 with m = floor(2^32 / d) + 1,
// exact while x * d < 2^32 (the operands here are shared-memory indices, < 2^18, and extents
// <= 4096).  The kernels are integer-bound (ncu: 60 % of the executed instructions were index
// arithmetic, a third of that the I2F / MUFU / F2I sequences of `/` and `%`).
struct FastDiv {
  unsigned d, m;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.d = d < 1 ? 1 : d;
  f.m = f.d == 1 ? 0u : (unsigned)(0x100000000ull / f.d) + 1u;
  return f;
}
EVX_HD unsigned fdiv(unsigned x, FastDiv f) {
  return f.d == 1 ? x : (unsigned)(((unsigned long long)x * f.m) >> 32);
}

constexpr int kMaxStages = 12;
struct LineDesc {
  int N;                    // line length
  int nstages;
  int radix[kMaxStages];    // product = N; each in {2,3,4,5,7,8}
  FastDiv dN, dH;           // divisors N and N/2 + 1 (z passes: row / element of an item)
  FastDiv dNs[kMaxStages];  // product of the radices before stage s
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
  d.dN = make_fastdiv((unsigned)n);
  d.dH = make_fastdiv((unsigned)(n / 2 + 1));
  unsigned ns = 1;
  for (int s = 0; s < d.nstages; ++s) {
    d.dNs[s] = make_fastdiv(ns);
    ns *= (unsigned)d.radix[s];
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
  // GEN_XMID on a non-periodic axis: the array holds N/2 rows, row i >= N/2 of the line is
  // mirror * row (N-1-i) and is not stored (+1 even, -1 odd, 0 periodic)
  int mirror;
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
  // complex lines a block holds in shared memory: W columns, or W / 2 pairs of real rows
  EVX_HD static bool paired(const P& p) {
    return (p.mode == GEN_Z_FWD || p.mode == GEN_Z_INV) && p.W % 2 == 0;
  }
  EVX_HD static int nlines(const P& p) { return paired(p) ? p.W / 2 : p.W; }
  // item -> (item / L, item % L) for the L interleaved lines (a power of two on the device)
  EVX_HD static void split_lines(int item, int L, int& q, int& c) {
    if ((L & (L - 1)) == 0) {
      int sh = 0;
      while ((1 << sh) < L) ++sh;
      q = item >> sh;
      c = item & (L - 1);
    } else {
      q = item / L;
      c = item - q * L;
    }
  }

  // ---- r-point DFTs in registers (dir < 0: forward, exp(-i...); dir > 0: inverse) --------
  EVX_HD static C cadd(C a, C b) { return {a.x + b.x, a.y + b.y}; }
  EVX_HD static C csub(C a, C b) { return {a.x - b.x, a.y - b.y}; }
  // a * (-i) for the forward transform, a * (+i) for the inverse one
  EVX_HD static C rot(C a, int dir) { return dir < 0 ? C{a.y, -a.x} : C{-a.y, a.x}; }
  EVX_HD static void dft2(C* x) {
    const C a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
  EVX_HD static void dft4(C& x0, C& x1, C& x2, C& x3, int dir) {
    const C a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = rot(csub(x1, x3), dir);
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = cadd(b, d);
    x3 = csub(b, d);
  }
  EVX_HD static void dft8(C* x, int dir) {
    dft4(x[0], x[2], x[4], x[6], dir);      // even inputs -> E[0..3] in x[0], x[2], x[4], x[6]
    dft4(x[1], x[3], x[5], x[7], dir);      // odd inputs  -> O[0..3] in x[1], x[3], x[5], x[7]
    const R h = R(0.70710678118654752440);
    const C o0 = x[1];
    const C t1 = cadd(x[3], rot(x[3], dir));              // O1 (1 -+ i)
    const C o1 = C{t1.x * h, t1.y * h};
    const C o2 = rot(x[5], dir);
    const C t3 = csub(rot(x[7], dir), x[7]);              // O3 (-1 -+ i)
    const C o3 = C{t3.x * h, t3.y * h};
    const C e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1); x[5] = csub(e1, o1);
    x[2] = cadd(e2, o2); x[6] = csub(e2, o2);
    x[3] = cadd(e3, o3); x[7] = csub(e3, o3);
  }
  EVX_HD static C cscale(C a, R s) { return {a.x * s, a.y * s}; }
  EVX_HD static C cfma(C a, R s, C b) { return {a.x * s + b.x, a.y * s + b.y}; }   // a s + b
  EVX_HD static void dft3(C* x, int dir) {
    const C t = cadd(x[1], x[2]);
    const C m = cfma(t, R(-0.5), x[0]);
    const C s = rot(cscale(csub(x[1], x[2]), R(0.86602540378443864676)), dir);
    x[0] = cadd(x[0], t);
    x[1] = cadd(m, s);
    x[2] = csub(m, s);
  }
  EVX_HD static void dft5(C* x, int dir) {
    const R c1 = R(0.30901699437494742410), c2 = R(-0.80901699437494742410);
    const R s1 = R(0.95105651629515357212), s2 = R(0.58778525229247312917);
    const C a1 = cadd(x[1], x[4]), a2 = cadd(x[2], x[3]);
    const C b1 = csub(x[1], x[4]), b2 = csub(x[2], x[3]);
    const C p1 = cfma(a2, c2, cfma(a1, c1, x[0]));
    const C p2 = cfma(a2, c1, cfma(a1, c2, x[0]));
    const C q1 = rot(cfma(b2, s2, cscale(b1, s1)), dir);
    const C q2 = rot(cfma(b2, -s1, cscale(b1, s2)), dir);
    x[0] = cadd(x[0], cadd(a1, a2));
    x[1] = cadd(p1, q1);
    x[4] = csub(p1, q1);
    x[2] = cadd(p2, q2);
    x[3] = csub(p2, q2);
  }
  // radix 7: direct sum with the r roots W_r^t held in registers
  template <int r>
  EVX_HD static void dft_odd(C* x, const C* wr) {
    C y[r];
#pragma unroll
    for (int q = 0; q < r; ++q) {
      C acc = x[0];
#pragma unroll
      for (int k = 1; k < r; ++k) {
        const C t = gmul(x[k], wr[(k * q) % r]);
        acc.x += t.x;
        acc.y += t.y;
      }
      y[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < r; ++q) x[q] = y[q];
  }

  // one Stockham stage of compile-time radix for all W lines of the block; in/out are [N][W].
  // Butterfly j multiplies input k by W_N^(m k), m = (j mod Ns) N / (Ns r) - m k < N, so the
  // table index needs no reduction - and takes the r-point DFT in registers.
  template <int r>
  EVX_HD static void stage_r(const P& p, const C* in, C* out, int Ns, FastDiv dNs, int dir, int tid,
                             int nthreads) {
    const int N = p.line.N, L = nlines(p);
    const int nb = N / r, step = N / (Ns * r);
    C wr[r];
    if (r == 7) {
#pragma unroll
      for (int t = 0; t < r; ++t) {
        const C w = p.tw[t * nb];
        wr[t] = dir < 0 ? w : C{w.x, -w.y};
      }
    }
    const int total = nb * L;
    for (int item = tid; item < total; item += nthreads) {
      int j, c;
      split_lines(item, L, j, c);
      const int jq = (int)fdiv((unsigned)j, dNs), jm = j - jq * Ns;
      const int m = jm * step;
      const C* src = in + j * L + c;
      C x[r];
      x[0] = src[0];
#pragma unroll
      for (int k = 1; k < r; ++k) {
        const C w = p.tw[m * k];
        x[k] = gmul(src[k * nb * L], dir < 0 ? w : C{w.x, -w.y});
      }
      if (r == 2) dft2(x);
      else if (r == 3) dft3(x, dir);
      else if (r == 4) dft4(x[0], x[1], x[2], x[3], dir);
      else if (r == 5) dft5(x, dir);
      else if (r == 8) dft8(x, dir);
      else dft_odd<r>(x, wr);
      C* dst = out + (jq * Ns * r + jm) * L + c;
#pragma unroll
      for (int q = 0; q < r; ++q) dst[q * Ns * L] = x[q];
    }
  }

  EVX_HD static void stage(const P& p, const C* in, C* out, int s, int dir, int tid, int nthreads) {
    const FastDiv dNs = p.line.dNs[s];
    const int Ns = (int)dNs.d;
    switch (p.line.radix[s]) {
      case 2: stage_r<2>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
      case 3: stage_r<3>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
      case 4: stage_r<4>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
      case 5: stage_r<5>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
      case 7: stage_r<7>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
      default: stage_r<8>(p, in, out, Ns, dNs, dir, tid, nthreads); break;
    }
  }

  // where the data of a block sits after `n` stages (buffers alternate, loads go to buffer 0)
  EVX_HD static C* bufn(C* smem, const P& p, int n) { return smem + (size_t)(n & 1) * p.line.N * p.W; }

  // (row of the pair, element) of a z-pass item; i fastest: contiguous global accesses
  EVX_HD static void split_z(int item, FastDiv dn, int& c, int& i) {
    c = (int)fdiv((unsigned)item, dn);
    i = item - c * (int)dn.d;
  }

  // column of a strided pass -> pointer to its element 0, or null for padding / out of range
  EVX_HD static C* column(const P& p, long long block, int c) {
    const int col = (int)block * p.W + c;             // (column counts fit 32 bits: n <= 4096)
    const int g = col / p.P, kz = col - g * p.P;
    if (col >= p.ncols_total || kz >= p.ncols_valid) return nullptr;
    return p.spec + (g * p.group_stride + kz);
  }

  // element i of a strided line (mirror mode: the upper half is the reversed lower half)
  EVX_HD static C line_element(const P& p, const C* src, int i) {
    if (!src) return C{R(0), R(0)};
    if (p.mirror && 2 * i >= p.line.N) {
      const C v = src[(long long)(p.line.N - 1 - i) * p.line_stride];
      return p.mirror > 0 ? v : C{-v.x, -v.y};
    }
    return src[(long long)i * p.line_stride];
  }

  EVX_HD static void load(const P& p, C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    if (p.mode == GEN_Z_FWD || p.mode == GEN_Z_INV) {
      const bool pair = paired(p);
      const int L = nlines(p), h = N / 2;
      for (int item = tid; item < N * L; item += nthreads) {
        int c, i;
        split_z(item, p.line.dN, c, i);
        const long long ra = block * W + (pair ? 2 * c : c), rb = ra + 1;
        C v{R(0), R(0)};
        if (p.mode == GEN_Z_FWD) {
          if (ra < p.rows) v.x = p.real_in[ra * p.nz + i];
          if (pair && rb < p.rows) v.y = p.real_in[rb * p.nz + i];
        } else {
          // Hermitian completion of the half rows; like a C2R transform the imaginary parts of
          // the k = 0 and k = N/2 entries are ignored.  z = A + i B transforms both rows at once.
          const int k = i <= h ? i : N - i;
          const bool self_conj = i == 0 || 2 * i == N;
          C A{R(0), R(0)}, B{R(0), R(0)};
          if (ra < p.rows) A = p.spec[ra * p.P + k];
          if (pair && rb < p.rows) B = p.spec[rb * p.P + k];
          if (i > h) { A.y = -A.y; B.y = -B.y; }
          if (self_conj) { A.y = R(0); B.y = R(0); }
          v = C{A.x - B.y, A.y + B.x};
        }
        b[i * L + c] = v;
      }
      return;
    }
    if (nthreads % W == 0) {            // the launch configuration: one column per thread
      const int c = tid % W;
      const C* src = column(p, block, c);
      for (int i = tid / W; i < N; i += nthreads / W) b[i * W + c] = line_element(p, src, i);
      return;
    }
    for (int item = tid; item < N * W; item += nthreads) {
      const int i = item / W, c = item - i * W;       // columns fastest
      b[i * W + c] = line_element(p, column(p, block, c), i);
    }
  }

  EVX_HD static void store(const P& p, const C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    if (p.mode == GEN_Z_FWD) {
      const bool pair = paired(p);
      const int L = nlines(p), h = N / 2 + 1;
      for (int item = tid; item < h * L; item += nthreads) {
        int c, i;
        split_z(item, p.line.dH, c, i);
        const long long ra = block * W + (pair ? 2 * c : c), rb = ra + 1;
        const C z = b[i * L + c];
        if (!pair) {
          if (ra < p.rows) p.spec[ra * p.P + i] = z;
          continue;
        }
        // X_A = (Z_k + conj Z_-k) / 2,  X_B = (Z_k - conj Z_-k) / (2 i)
        const C zn = b[(i == 0 ? 0 : N - i) * L + c];
        if (ra < p.rows) p.spec[ra * p.P + i] = C{R(0.5) * (z.x + zn.x), R(0.5) * (z.y - zn.y)};
        if (rb < p.rows) p.spec[rb * p.P + i] = C{R(0.5) * (z.y + zn.y), R(0.5) * (zn.x - z.x)};
      }
    } else if (p.mode == GEN_Z_INV) {
      const bool pair = paired(p);
      const int L = nlines(p);
      for (int item = tid; item < N * L; item += nthreads) {
        int c, i;
        split_z(item, p.line.dN, c, i);
        const long long ra = block * W + (pair ? 2 * c : c), rb = ra + 1;
        const C z = b[i * L + c];
        if (ra < p.rows) p.real_out[ra * p.nz + i] = p.real_in ? p.real_in[ra * p.nz + i] + z.x : z.x;
        if (pair && rb < p.rows)
          p.real_out[rb * p.nz + i] = p.real_in ? p.real_in[rb * p.nz + i] + z.y : z.y;
      }
    } else if (nthreads % W == 0) {
      const int c = tid % W, nst = p.mirror ? N / 2 : N;
      C* dst = column(p, block, c);
      if (dst)
        for (int i = tid / W; i < nst; i += nthreads / W) dst[(long long)i * p.line_stride] = b[i * W + c];
    } else {
      const int nst = p.mirror ? N / 2 : N;
      for (int item = tid; item < nst * W; item += nthreads) {
        const int i = item / W, c = item - i * W;
        C* dst = column(p, block, c);
        if (dst) dst[(long long)i * p.line_stride] = b[i * W + c];
      }
    }
  }

  // multiply the transformed lines by weight(k) * scale (natural order along the line)
  EVX_HD static void weight_one(const P& p, C* b, long long block, int i, int c) {
    const FilterParams& f = p.filt;
    const int col = (int)block * p.W + c;
    const int g = col / p.P, kz = col - g * p.P;
    const float k0 = wavenumber(signed_freq(i, f.n0), f.inv_len0);
    const float k1 = wavenumber(signed_freq(g, f.n1), f.inv_len1);
    const float k2 = wavenumber(kz, f.inv_len2);
    const float ksq = fadd_rn(fadd_rn(fmul_rn(k0, k0), fmul_rn(k1, k1)), fmul_rn(k2, k2));
    const R w = (R)spectral_weight(ksq, f) * (sizeof(R) == 8 ? (R)f.scale_d : (R)f.scale);
    C& v = b[i * p.W + c];
    v.x *= w;
    v.y *= w;
  }
  EVX_HD static void weight(const P& p, C* b, long long block, int tid, int nthreads) {
    const int N = p.line.N, W = p.W;
    if (nthreads % W == 0) {
      for (int i = tid / W; i < N; i += nthreads / W) weight_one(p, b, block, i, tid % W);
      return;
    }
    for (int item = tid; item < N * W; item += nthreads) weight_one(p, b, block, item / W, item % W);
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
