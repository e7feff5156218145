// Lane types shared by the stencil kernels: a "lane" is one value or a pair of neighbouring
// z values; float pairs map to Blackwell's packed-FP32 instructions on the device.
#pragma once
#include "evx_hd.h"
#include "packed_f32.h"

namespace evx {

// ------------------------------------------------------------------------------------
// One value, or a pair of neighbouring z values (float pairs use the packed
// FADD2/FMUL2/FFMA2 path on the device).
// ------------------------------------------------------------------------------------
template <typename T, int LW>
struct AcLane;

template <typename T>
struct AcLane<T, 1> {
  T a;
  EVX_HD static AcLane load(const T* w, int k) { return AcLane{w[k]}; }
  EVX_HD void store(T* w, int k) const { w[k] = a; }
  EVX_HD static AcLane add(AcLane x, AcLane y) { return AcLane{x.a + y.a}; }
  EVX_HD static AcLane sub(AcLane x, AcLane y) { return AcLane{x.a - y.a}; }
  EVX_HD static AcLane mul(AcLane x, AcLane y) { return AcLane{x.a * y.a}; }
  EVX_HD static AcLane fma(AcLane x, AcLane y, AcLane z) { return AcLane{x.a * y.a + z.a}; }
  EVX_HD static AcLane muls(AcLane x, T s) { return AcLane{x.a * s}; }
  EVX_HD static AcLane fmas(AcLane x, T s, AcLane z) { return AcLane{x.a * s + z.a}; }
  EVX_HD static AcLane rsubs(T s, AcLane x) { return AcLane{s - x.a}; }
  EVX_HD static AcLane guarded_div(AcLane n, AcLane d) {
    return AcLane{n.a / (d.a <= T(1e-7) ? T(1) : d.a)};
  }
};

template <typename T>
struct AcLane<T, 2> {
  T a, b;
  EVX_HD static AcLane load(const T* w, int k) { return AcLane{w[k], w[k + 1]}; }
  EVX_HD void store(T* w, int k) const { w[k] = a; w[k + 1] = b; }
  EVX_HD static AcLane add(AcLane x, AcLane y) { return AcLane{x.a + y.a, x.b + y.b}; }
  EVX_HD static AcLane sub(AcLane x, AcLane y) { return AcLane{x.a - y.a, x.b - y.b}; }
  EVX_HD static AcLane mul(AcLane x, AcLane y) { return AcLane{x.a * y.a, x.b * y.b}; }
  EVX_HD static AcLane fma(AcLane x, AcLane y, AcLane z) {
    return AcLane{x.a * y.a + z.a, x.b * y.b + z.b};
  }
  EVX_HD static AcLane muls(AcLane x, T s) { return AcLane{x.a * s, x.b * s}; }
  EVX_HD static AcLane fmas(AcLane x, T s, AcLane z) { return AcLane{x.a * s + z.a, x.b * s + z.b}; }
  EVX_HD static AcLane rsubs(T s, AcLane x) { return AcLane{s - x.a, s - x.b}; }
  EVX_HD static AcLane guarded_div(AcLane n, AcLane d) {
    return AcLane{n.a / (d.a <= T(1e-7) ? T(1) : d.a), n.b / (d.b <= T(1e-7) ? T(1) : d.b)};
  }
};

#if defined(__CUDA_ARCH__)
template <>
struct AcLane<float, 2> {
  f2 v;
  EVX_D static AcLane load(const float* w, int k) { return AcLane{f2{w[k], w[k + 1]}}; }
  EVX_D void store(float* w, int k) const { w[k] = v.a; w[k + 1] = v.b; }
  EVX_D static AcLane add(AcLane x, AcLane y) { return AcLane{f2_add(x.v, y.v)}; }
  EVX_D static AcLane sub(AcLane x, AcLane y) { return AcLane{f2_sub(x.v, y.v)}; }
  EVX_D static AcLane mul(AcLane x, AcLane y) { return AcLane{f2_mul(x.v, y.v)}; }
  EVX_D static AcLane fma(AcLane x, AcLane y, AcLane z) { return AcLane{f2_fma(x.v, y.v, z.v)}; }
  EVX_D static AcLane muls(AcLane x, float s) { return AcLane{f2_mul(x.v, f2_splat(s))}; }
  EVX_D static AcLane fmas(AcLane x, float s, AcLane z) { return AcLane{f2_fma(x.v, f2_splat(s), z.v)}; }
  EVX_D static AcLane rsubs(float s, AcLane x) { return AcLane{f2_sub(f2_splat(s), x.v)}; }
  EVX_D static AcLane guarded_div(AcLane n, AcLane d) {
    // hardware reciprocal (<= 2 ulp); the reference divides exactly, the difference is far
    // below the test tolerances
    const float da = d.v.a <= 1e-7f ? 1.0f : d.v.a, db = d.v.b <= 1e-7f ? 1.0f : d.v.b;
    return AcLane{f2{__fdividef(n.v.a, da), __fdividef(n.v.b, db)}};
  }
};
#endif

}  // namespace evx
