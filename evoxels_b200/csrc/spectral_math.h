// Wavenumber / IMEX prefactor arithmetic shared by the cuFFT-path filter kernel and the
// native x-pass (host/device).  float32 with the reference's rounding sequence:
//   freq = float(idx) * float(1/(n*d))          torch.fft.fftfreq / rfftfreq
//   k    = float(2*pi) * freq                   evoxels/voxelgrid.py:84-90
//   k2   = (kx*kx + ky*ky) + kz*kz              evoxels/voxelgrid.py:110-114
//   P    = (1 / (1 + dt * (coef * k2^power))) * dt   problem_definition.py:303, timesteppers.py:77
//          (torch evaluates `scalar / tensor` as tensor.reciprocal() * scalar)
// The reference keeps all of this in float32 even for float64 fields.
// On the device the *_rn intrinsics keep nvcc from contracting mul+add into fma, so P is
// bit-identical to the reference's stored prefactor array.
#pragma once
#include "evx_hd.h"

namespace evx {

struct FilterParams {
  int n0, n1, n2;                       // extents (n2 is the halved, contiguous one)
  float inv_len0, inv_len1, inv_len2;   // float(1/(n*h)) per axis
  float dt, coef, scale;
  int power;                            // 1: |k|^2, 2: |k|^4
  int kind;                             // FILTER_IMEX or FILTER_ETD1
  double scale_d;
};

// The `power` argument of the C ABI carries the filter kind in bit 8 (EVX_FILTER_ETD1 in
// include/evoxels_b200.h): weight = dt / (1 + dt c |k|^2p)  or  dt * phi1(-dt c |k|^2p).
// Bits 9 / 10 (EVX_FILTER_MIRROR_EVEN / _ODD): the x axis is not periodic - the x pass
// transforms every line together with its even (zero-flux) or odd (Dirichlet) mirror image,
// synthesised on the fly, as ONE line of 2 nx points (reference boundary_conditions.py:65-71
// concatenates the flipped field in memory before rfftn).
enum : int { FILTER_IMEX = 0, FILTER_ETD1 = 1, FILTER_KIND_BIT = 0x100, FILTER_MIRROR_EVEN = 0x200,
             FILTER_MIRROR_ODD = 0x400, FILTER_FLAG_BITS = 0x700 };
inline bool valid_filter_spec(int power) {
  const int pw = power & ~FILTER_FLAG_BITS;
  if ((power & FILTER_MIRROR_EVEN) && (power & FILTER_MIRROR_ODD)) return false;
  return pw == 1 || pw == 2;
}
// +1 even mirror, -1 odd mirror, 0 periodic
inline int filter_mirror(int power) {
  return (power & FILTER_MIRROR_EVEN) ? 1 : ((power & FILTER_MIRROR_ODD) ? -1 : 0);
}

#if defined(__CUDA_ARCH__)
EVX_HD float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
EVX_HD float fadd_rn(float a, float b) { return __fadd_rn(a, b); }
EVX_HD float frcp_rn(float a) { return __frcp_rn(a); }
#else
EVX_HD float fmul_rn(float a, float b) { volatile float r = a * b; return r; }
EVX_HD float fadd_rn(float a, float b) { volatile float r = a + b; return r; }
EVX_HD float frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
#endif

#if defined(__CUDA_ARCH__)
EVX_HD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
EVX_HD float fdiv_fast(float a, float b) { return __fdividef(a, b); }
#else
EVX_HD float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
EVX_HD float fdiv_fast(float a, float b) { return a / b; }
#endif

EVX_HD float wavenumber(int idx, float inv_len) {
  return fmul_rn(6.283185307179586f, fmul_rn((float)idx, inv_len));
}
EVX_HD int signed_freq(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

// Same prefactor with the hardware reciprocal approximation (<= 1 ulp off the IEEE result):
// used inside the native x pass, where the filter sits between two FFTs in an issue-bound
// kernel.  The stand-alone filter kernel keeps the bit-exact sequence.
EVX_HD float imex_prefactor_fast(float k2, const FilterParams& f) {
  const float kp = f.power == 2 ? k2 * k2 : k2;
  const float den = 1.0f + f.dt * (f.coef * kp);
#if defined(__CUDA_ARCH__)
  return __fdividef(f.dt, den);
#else
  return f.dt / den;
#endif
}

EVX_HD float imex_prefactor(float k2, const FilterParams& f) {
  const float kp = f.power == 2 ? fmul_rn(k2, k2) : k2;
  const float den = fadd_rn(1.0f, fmul_rn(f.dt, fmul_rn(f.coef, kp)));
  return fmul_rn(frcp_rn(den), f.dt);
}

// dt * varphi_1(dt * symbol), symbol = -coef |k|^2p, varphi_1(z) = (exp(z) - 1) / z with the
// reference's (6,6) Pade branch for |z| < 0.5 (ExponentialEuler.phi1 / phiPade,
// evoxels/timesteppers.py:155-192; float32 like the reference's symbol array).
EVX_HD float etd1_weight(float k2, const FilterParams& f) {
  const float kp = f.power == 2 ? fmul_rn(k2, k2) : k2;
  const float z = fmul_rn(f.dt, -fmul_rn(f.coef, kp));
  float phi;
  if (fabsf(z) < 0.5f) {
    const float N[7] = {1.f, (float)(1. / 26), (float)(5. / 156), (float)(1. / 858),
                        (float)(1. / 5720), (float)(1. / 205920), (float)(1. / 8648640)};
    const float D[7] = {1.f, (float)(-6. / 13), (float)(5. / 52), (float)(-5. / 429),
                        (float)(1. / 1144), (float)(-1. / 25740), (float)(1. / 1235520)};
    float num = N[6], den = D[6];
#pragma unroll
    for (int k = 5; k >= 0; --k) {
      num = fadd_rn(fmul_rn(num, z), N[k]);
      den = fadd_rn(fmul_rn(den, z), D[k]);
    }
    phi = num / den;
  } else {
    phi = (expf(z) - 1.0f) / z;
  }
  return fmul_rn(f.dt, phi);
}

// Weight of one coefficient inside the fused x pass, FFT normalisation included: k0 is the
// wavenumber along the transformed axis, k12 = k1^2 + k2^2 of the other two.  Every operation is
// an explicit round-to-nearest intrinsic, so all kernels that inline this (cp.async passes,
// TMA-tiled passes, distributed passes) produce the same bits - nvcc's mul/add contraction
// otherwise depends on the surrounding code.
template <bool ETD1>
EVX_HD float xpass_weight(float k0, float k12, const FilterParams& f) {
  const float kk = fma_rn(k0, k0, k12);
  if (ETD1) return fmul_rn(etd1_weight(kk, f), f.scale);
  const float kp = f.power == 2 ? fmul_rn(kk, kk) : kk;
  const float den = fma_rn(f.dt, fmul_rn(f.coef, kp), 1.0f);
  return fmul_rn(fdiv_fast(f.dt, den), f.scale);
}

// weight applied to one spectral coefficient (without the FFT normalisation)
EVX_HD float spectral_weight(float k2, const FilterParams& f) {
  return f.kind == FILTER_ETD1 ? etd1_weight(k2, f) : imex_prefactor(k2, f);
}
EVX_HD float spectral_weight_fast(float k2, const FilterParams& f) {
  return f.kind == FILTER_ETD1 ? etd1_weight(k2, f) : imex_prefactor_fast(k2, f);
}

inline FilterParams make_filter(const int n[3], const double len_h[3], double dt, double coef,
                                int power, double scale) {
  FilterParams f;
  f.n0 = n[0]; f.n1 = n[1]; f.n2 = n[2];
  f.inv_len0 = (float)(1.0 / (n[0] * len_h[0]));
  f.inv_len1 = (float)(1.0 / (n[1] * len_h[1]));
  f.inv_len2 = (float)(1.0 / (n[2] * len_h[2]));
  f.dt = (float)dt;
  f.coef = (float)coef;
  f.power = power & 0xff;
  f.kind = (power & FILTER_KIND_BIT) ? FILTER_ETD1 : FILTER_IMEX;
  f.scale = (float)scale;
  f.scale_d = scale;
  return f;
}

}  // namespace evx
