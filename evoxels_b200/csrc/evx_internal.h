// Internal declarations shared by the .cu translation units of libevx_b200.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/evoxels_b200.h"

namespace evx {

extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// defined in stencil.cu, used by the fused step in spectral.cu
template <typename T>
int ch_rhs_impl(const T* c, const T* hom, T* rhs, int nx, int ny, int nz, const double* h,
                double eps, double D, const int* bc_kind, const double* bc_val,
                const T* halo_lo, const T* halo_hi, cudaStream_t st);

}  // namespace evx
