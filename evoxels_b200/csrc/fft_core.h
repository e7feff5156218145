// Radix-2/4/8 Stockham building blocks for the native FFT passes (host/device).
//
// A line of N = 2^k complex points (8 <= N <= 2048) is transformed by T = N/8 threads,
// each holding 8 points in registers.  Stage s has radix R_s (the odd power of two, 2 or
// 4, goes FIRST where no twiddles are needed; all other stages are radix 8) and
// Ns = prod_{s'<s} R_s'.  With q = 8/R, thread t owns the virtual butterflies
// jv = t + i*T (i < q); element v[i + r*q] is butterfly input/output r.  In every stage
//     inputs   v[e] = in [t + e*T]                                  (e = 0..7)
//     outputs  out[(jv/Ns)*Ns*R + jv%Ns + r*Ns] = v[i + r*q]
// and for the last stage the output index collapses to t + e*T again, so the first load
// and the final store are both "natural order, stride T" and can go straight to global
// memory.  Between stages the data crosses shared memory.
//
// DIR = -1: forward (exp(-2 pi i ..)), DIR = +1: unnormalised inverse.
#pragma once
#include "evx_hd.h"

namespace evx {

struct alignas(8) cf {
  float x, y;
};

// Complex arithmetic.  On sm_100a these are the packed-FP32 instructions of Blackwell
// (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): a complex add is ONE
// instruction, a complex multiply two, and multiplication by +-i is an operand modifier
// (half swap + per-half negate) that ptxas folds into the consuming instruction.  The host
// replay uses the scalar expressions.
#if defined(__CUDA_ARCH__)
typedef unsigned long long cf_bits;
EVX_D cf_bits cf_pack(float lo, float hi) {
  cf_bits r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
EVX_D cf cf_unpack(cf_bits v) {
  cf r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
EVX_D cf cadd(cf a, cf b) {
  cf_bits r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cf_pack(a.x, a.y)), "l"(cf_pack(b.x, b.y)));
  return cf_unpack(r);
}
EVX_D cf csub(cf a, cf b) {
  cf_bits r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cf_pack(a.x, a.y)), "l"(cf_pack(b.x, b.y)));
  return cf_unpack(r);
}
EVX_D cf cmul(cf a, cf b) {
  cf_bits t, r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(t) : "l"(cf_pack(a.y, a.y)), "l"(cf_pack(b.y, b.x)));
  const cf tt = cf_unpack(t);
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(cf_pack(a.x, a.x)), "l"(cf_pack(b.x, b.y)), "l"(cf_pack(-tt.x, tt.y)));
  return cf_unpack(r);
}
EVX_D cf cscale(cf a, float s) {
  cf_bits r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cf_pack(a.x, a.y)), "l"(cf_pack(s, s)));
  return cf_unpack(r);
}
#else
EVX_HD cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
EVX_HD cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
EVX_HD cf cmul(cf a, cf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
EVX_HD cf cscale(cf a, float s) { return {a.x * s, a.y * s}; }
#endif
EVX_HD cf cconj(cf a) { return {a.x, -a.y}; }
// multiply by -i (DIR=-1) or +i (DIR=+1)
template <int DIR>
EVX_HD cf mul_dir_i(cf a) {
  return DIR < 0 ? cf{a.y, -a.x} : cf{-a.y, a.x};
}

template <int DIR>
EVX_HD void dft2(cf& a, cf& b) {
  const cf t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

// natural-order in, natural-order out
template <int DIR>
EVX_HD void dft4(cf& c0, cf& c1, cf& c2, cf& c3) {
  const cf s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3);
  const cf s3 = mul_dir_i<DIR>(csub(c1, c3));
  c0 = cadd(s0, s2);
  c2 = csub(s0, s2);
  c1 = cadd(s1, s3);
  c3 = csub(s1, s3);
}

template <int DIR>
EVX_HD void dft8(cf* v) {
  const float h = 0.70710678118654752440f;
  cf a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
  cf a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
  cf a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
  cf a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
  // b_r *= w8^r, w8 = exp(DIR * 2 pi i / 8) = (1 + DIR*i)/sqrt2:
  //   b w8 = h (b + (DIR*i) b),   b w8^3 = h ((DIR*i) b - b)
  b1 = cscale(cadd(b1, mul_dir_i<DIR>(b1)), h);
  b2 = mul_dir_i<DIR>(b2);
  b3 = cscale(csub(mul_dir_i<DIR>(b3), b3), h);
  dft4<DIR>(a0, a1, a2, a3);
  dft4<DIR>(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// log2 helpers ---------------------------------------------------------------------------
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
constexpr bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
// radix of the first stage: 8, or the odd leftover 2 / 4
constexpr int first_radix(int N) { return ilog2(N) % 3 == 0 ? 8 : (ilog2(N) % 3 == 1 ? 2 : 4); }
constexpr int num_stages(int N) { return (ilog2(N) + 2) / 3; }

// One Stockham stage on the 8 register values of thread t.
//   R      radix of this stage,  Ns  product of the previous radices
//   tw     table W_N[m] = exp(-2 pi i m / N), m < N (forward roots; conjugated for DIR=+1)
// After the call v[] holds the butterfly outputs; out_index(i, r) gives where v[i + r*q]
// belongs in the stage's output array.
template <int N, int R, int DIR>
struct Stage {
  static constexpr int T = N / 8;
  static constexpr int Q = 8 / R;

  EVX_HD static void compute(cf* v, int t, int Ns, const cf* tw) {
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      cf a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = v[i + r * Q];
      if (Ns > 1) {
        const int jv = t + i * T;
        const int m = (jv % Ns) * (N / (Ns * R));   // w_r = W_N[m*r]
        cf w1 = tw[m];
        if (DIR > 0) w1 = cconj(w1);
        if (R == 2) {
          a[1] = cmul(a[1], w1);
        } else {
          cf w2 = tw[2 * m];
          if (DIR > 0) w2 = cconj(w2);
          const cf w3 = cmul(w1, w2);
          a[1] = cmul(a[1], w1);
          a[2] = cmul(a[2], w2);
          a[3] = cmul(a[3], w3);
          if (R == 8) {
            cf w4 = tw[4 * m];
            if (DIR > 0) w4 = cconj(w4);
            a[4] = cmul(a[4], w4);
            a[5] = cmul(a[5], cmul(w4, w1));
            a[6] = cmul(a[6], cmul(w4, w2));
            a[7] = cmul(a[7], cmul(w4, w3));
          }
        }
      }
      if (R == 2) dft2<DIR>(a[0], a[1]);
      if (R == 4) dft4<DIR>(a[0], a[1], a[2], a[3]);
      if (R == 8) dft8<DIR>(a);
#pragma unroll
      for (int r = 0; r < R; ++r) v[i + r * Q] = a[r];
    }
  }

  EVX_HD static int out_index(int t, int Ns, int i, int r) {
    const int jv = t + i * T;
    return (jv / Ns) * Ns * R + jv % Ns + r * Ns;
  }
};

// Whole-line transform of the 8 values per thread, with the inter-stage exchange done by
// a caller-supplied "exchange" functor: xchg(stage_index, v, writer) must (1) store
// v[i + r*q] at out_index, (2) synchronise the block, (3) reload v[e] = buf[t + e*T].
// The programs below unroll this by hand because their phases are split at the barriers;
// this helper documents the sequence and is used by the emulator's single-line test.
template <int N>
struct LinePlan {
  static constexpr int S = num_stages(N);
  static constexpr int R0 = first_radix(N);
  static constexpr int T = N / 8;
  // radix and Ns of stage s (0-based)
  EVX_HD static constexpr int radix(int s) { return s == 0 ? R0 : 8; }
  EVX_HD static constexpr int ns(int s) { return s == 0 ? 1 : R0 * (1 << (3 * (s - 1))); }
};

// compute stage `s` of an N-point line (compile-time dispatch on the radix)
template <int N, int DIR>
EVX_HD void line_stage_compute(int s, cf* v, int t, const cf* tw) {
  using LP = LinePlan<N>;
  if (s == 0) {
    Stage<N, LP::R0, DIR>::compute(v, t, 1, tw);
  } else {
    Stage<N, 8, DIR>::compute(v, t, LP::ns(s), tw);
  }
}

// Twiddle prefetch.  Every twiddled stage is radix 8 with one butterfly per thread, whose
// roots W^m, W^2m, W^4m depend only on (stage, thread).  The pass programs fetch them right
// after the previous stage's shared-memory writes - i.e. BEFORE the barrier - so the table
// lookups (L1/L2 latency) overlap the barrier wait instead of stalling the first multiply.
template <int N>
EVX_HD void stage_twiddles(int s, int t, const cf* tw, cf* w) {
  using LP = LinePlan<N>;
  if (s <= 0 || s >= LP::S) return;
  const int Ns = LP::ns(s);
  const int m = (t % Ns) * (N / (Ns * 8));
  w[0] = tw[m];
  w[1] = tw[2 * m];
  w[2] = tw[4 * m];
}

// stage `s` with the roots already in registers (w from stage_twiddles; unused for s == 0)
template <int N, int DIR>
EVX_HD void line_stage_compute_pre(int s, cf* v, int t, const cf* w) {
  using LP = LinePlan<N>;
  if (s == 0) {
    Stage<N, LP::R0, DIR>::compute(v, t, 1, nullptr);
    return;
  }
  cf w1 = w[0], w2 = w[1], w4 = w[2];
  if (DIR > 0) { w1 = cconj(w1); w2 = cconj(w2); w4 = cconj(w4); }
  const cf w3 = cmul(w1, w2);
  v[1] = cmul(v[1], w1);
  v[2] = cmul(v[2], w2);
  v[3] = cmul(v[3], w3);
  v[4] = cmul(v[4], w4);
  v[5] = cmul(v[5], cmul(w4, w1));
  v[6] = cmul(v[6], cmul(w4, w2));
  v[7] = cmul(v[7], cmul(w4, w3));
  dft8<DIR>(v);
}

template <int N>
EVX_HD int line_stage_out_index(int s, int t, int e) {
  using LP = LinePlan<N>;
  const int R = LP::radix(s);
  const int Q = 8 / R;
  const int i = e % Q, r = e / Q;          // e = i + r*Q
  const int jv = t + i * LP::T;
  const int Ns = LP::ns(s);
  return (jv / Ns) * Ns * R + jv % Ns + r * Ns;
}

// Decomposition of the Stockham output index into a per-thread base and a per-element
// constant:  out_index(s, t, e) = stage_out_base(s, t) + stage_out_const(s, e)
// (valid because Ns <= T for every stage, so jv = t + i*T splits cleanly).  With the
// additive smem padding x + (x >> 3) the two parts never carry into each other's low
// three bits, hence pad(base + c) = pad(base) + pad(c): every shared-memory access of a
// stage is one address computation per thread plus compile-time immediates.
template <int N>
EVX_HD int stage_out_base(int s, int t) {
  using LP = LinePlan<N>;
  const int Ns = LP::ns(s), R = LP::radix(s);
  return (t / Ns) * (Ns * R) + (t % Ns);
}
template <int N>
EVX_HD constexpr int stage_out_const(int s, int e) {
  using LP = LinePlan<N>;
  const int R = LP::radix(s), Q = 8 / R;
  return (e % Q) * LP::T * R + (e / Q) * LP::ns(s);
}

}  // namespace evx
