// Shared plumbing of the warp-specialised stencil kernels (ch_rhs_tma.cu): mbarrier
// waits for the loader / consumer roles and the TMA tensor maps of a plane tile and its ring.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include "evx_internal.h"
#include "tma_ptx.h"

namespace evx {
namespace stma {

#if defined(__CUDACC__)
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float sat(float a) { return __saturatef(a); }

__device__ __forceinline__ bool mbar_try(unsigned addr, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
// slow path of a wait, out of line: bounded (a pipeline bug must end in a trap, not in a hang)
static __device__ __noinline__ void mbar_wait_slow(unsigned addr, unsigned parity) {
  const long long t0 = clock64();
  for (unsigned spin = 0;; ++spin) {
    if (mbar_try(addr, parity)) return;
    if ((spin & 255u) == 255u && clock64() - t0 > 4000000000LL) __trap();
  }
}
// consumers: the plane has normally landed long ago - one probe, no clock read
__device__ __forceinline__ void wait_full(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  if (!mbar_try(addr, parity)) mbar_wait_slow(addr, parity);
}
// loader: it is normally ahead of the consumers and polls; sleep between probes so that the
// polling does not take issue slots from the compute warps
__device__ __forceinline__ void wait_empty(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  const long long t0 = clock64();
  for (unsigned spin = 0; !mbar_try(addr, parity); ++spin) {
    __nanosleep(100);
    if ((spin & 255u) == 255u && clock64() - t0 > 4000000000LL) __trap();
  }
}

#endif

// tensor maps of one launch: [array: c, halo_lo, halo_hi][box: tile (128 x TY), row pair
// (128 x 2), column (4 x TY), corner (4 x 1)]
struct Maps {
  CUtensorMap m[3][4];
};

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---- tensor maps ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return (EncodeTiledFn)sym;
  }();
  return fn;
}

// the four boxes (tile 128 x ty, ring rows 128 x ring_rows, ring column 4 x ty, corner 4 x 1)
// over one [nplanes, ny, nz] fp32 array; encoding is pure host arithmetic, the
// last few arrays are remembered (a stepper alternates between two or three fields)
inline bool make_maps(CUtensorMap out[4], const float* base, int nplanes, int ny, int nz, int ty,
                      int ring_rows) {
  struct Entry {
    const float* base = nullptr;
    int nplanes = 0, ny = 0, nz = 0, ty = 0, rr = 0;
    CUtensorMap m[4];
  };
  constexpr int NE = 8;
  thread_local Entry cache[NE];
  thread_local int next = 0;
  for (int i = 0; i < NE; ++i) {
    const Entry& e = cache[i];
    if (e.base == base && e.nplanes == nplanes && e.ny == ny && e.nz == nz && e.ty == ty &&
        e.rr == ring_rows) {
      for (int k = 0; k < 4; ++k) out[k] = e.m[k];
      return true;
    }
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nplanes};
  const cuuint64_t strides[2] = {(cuuint64_t)nz * sizeof(float), (cuuint64_t)ny * nz * sizeof(float)};
  const cuuint32_t boxes[4][3] = {{128, (cuuint32_t)ty, 1},
                                 {128, (cuuint32_t)ring_rows, 1},
                                 {4, (cuuint32_t)ty, 1},
                                 {4, 1, 1}};
  const cuuint32_t estr[3] = {1, 1, 1};
  Entry e;
  e.base = base; e.nplanes = nplanes; e.ny = ny; e.nz = nz; e.ty = ty; e.rr = ring_rows;
  for (int k = 0; k < 4; ++k) {
    const CUresult rc = enc(&e.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, boxes[k],
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            k < 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
  }
  cache[next] = e;
  next = (next + 1) % NE;
  for (int k = 0; k < 4; ++k) out[k] = e.m[k];
  return true;
}

}  // namespace stma
}  // namespace evx
