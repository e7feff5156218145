// Internal declarations shared by the .cu translation units of libevx_b200.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/evoxels_b200.h"

namespace evx {

extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Opt-in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute is a per-device setting,
// so it is remembered per (kernel instantiation, device): one `static SmemOptIn` per launcher.
struct SmemOptIn {
  std::atomic<unsigned long long> done{0};   // bit d: made on device d (d < 64)
  template <class K>
  int ensure(K kern, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    const cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    done.fetch_or(bit, std::memory_order_release);
    return 0;
  }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// defined in stencil.cu, used by the fused step in spectral.cu
template <typename T>
int ch_rhs_impl(const T* c, const T* hom, T* rhs, int nx, int ny, int nz, const double* h,
                double eps, double D, const int* bc_kind, const double* bc_val,
                const T* halo_lo, const T* halo_hi, cudaStream_t st);

// ch_rhs_tma.cu: warp-specialised fp32 form; EVX_ERR_UNSUPPORTED when the shape is not covered
template <typename T> struct ChParams;
int ch_rhs_tma_f32(ChParams<float> p, cudaStream_t st);

}  // namespace evx
