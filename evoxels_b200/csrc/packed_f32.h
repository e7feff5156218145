// Pairs of fp32 values processed by Blackwell's packed-FP32 instructions.
// PTX add/sub/mul/fma .f32x2 (sm_100+) become SASS FADD2 / FMUL2 / FFMA2: two results per
// issued instruction, with broadcast and half-swap operand modifiers folded in by ptxas.
// The stencil kernels are fp32-issue bound on B200 (8 B/voxel against 45-85 flop/voxel), so
// halving the FP instruction count is what moves them towards the HBM roofline.
// Host replay: plain scalar arithmetic on both halves.
#pragma once
#include "evx_hd.h"

namespace evx {

struct f2 {
  float a, b;
};

#if defined(__CUDA_ARCH__)
EVX_D unsigned long long f2_bits(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
EVX_D f2 f2_from(unsigned long long v) {
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.a), "=f"(r.b) : "l"(v));
  return r;
}
EVX_D f2 f2_add(f2 x, f2 y) {
  unsigned long long r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(x.a, x.b)), "l"(f2_bits(y.a, y.b)));
  return f2_from(r);
}
EVX_D f2 f2_sub(f2 x, f2 y) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(x.a, x.b)), "l"(f2_bits(y.a, y.b)));
  return f2_from(r);
}
EVX_D f2 f2_mul(f2 x, f2 y) {
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(x.a, x.b)), "l"(f2_bits(y.a, y.b)));
  return f2_from(r);
}
EVX_D f2 f2_fma(f2 x, f2 y, f2 z) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(f2_bits(x.a, x.b)), "l"(f2_bits(y.a, y.b)), "l"(f2_bits(z.a, z.b)));
  return f2_from(r);
}
#else
EVX_HD f2 f2_add(f2 x, f2 y) { return {x.a + y.a, x.b + y.b}; }
EVX_HD f2 f2_sub(f2 x, f2 y) { return {x.a - y.a, x.b - y.b}; }
EVX_HD f2 f2_mul(f2 x, f2 y) { return {x.a * y.a, x.b * y.b}; }
EVX_HD f2 f2_fma(f2 x, f2 y, f2 z) { return {fmaf(x.a, y.a, z.a), fmaf(x.b, y.b, z.b)}; }   // fused, like the device
#endif
EVX_HD f2 f2_splat(float s) { return {s, s}; }

}  // namespace evx
