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
  double scale_d;
};

#if defined(__CUDA_ARCH__)
EVX_HD float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
EVX_HD float fadd_rn(float a, float b) { return __fadd_rn(a, b); }
EVX_HD float frcp_rn(float a) { return __frcp_rn(a); }
#else
EVX_HD float fmul_rn(float a, float b) { volatile float r = a * b; return r; }
EVX_HD float fadd_rn(float a, float b) { volatile float r = a + b; return r; }
EVX_HD float frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
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

inline FilterParams make_filter(const int n[3], const double len_h[3], double dt, double coef,
                                int power, double scale) {
  FilterParams f;
  f.n0 = n[0]; f.n1 = n[1]; f.n2 = n[2];
  f.inv_len0 = (float)(1.0 / (n[0] * len_h[0]));
  f.inv_len1 = (float)(1.0 / (n[1] * len_h[1]));
  f.inv_len2 = (float)(1.0 / (n[2] * len_h[2]));
  f.dt = (float)dt;
  f.coef = (float)coef;
  f.power = power;
  f.scale = (float)scale;
  f.scale_d = scale;
  return f;
}

}  // namespace evx
