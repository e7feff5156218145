// Host/device portability layer.
//
// Every kernel in this library is written as a "program": a struct with per-thread
// register state (Regs), a shared-memory image (Smem) and phase functions that contain
// no barrier.  On the GPU a thin __global__ wrapper runs the phases with __syncthreads()
// between them.  The same phase functions compile as plain C++ so that tests/emu/ can
// replay a thread block on the CPU (phase by phase, thread by thread) and compare the
// index/halo/boundary logic against the oracle without a GPU.  The emulator is test
// infrastructure; nothing in the shipped library executes on the host.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define EVX_HD __host__ __device__ __forceinline__
#define EVX_D __device__ __forceinline__
#else
#define EVX_HD inline
#define EVX_D inline
#endif

namespace evx {

enum : int { BC_PERIODIC = 0, BC_NEUMANN = 1, BC_DIRICHLET = 2 };

// V consecutive elements along the contiguous (z) axis; 16-byte aligned for
// float x4 / double x2 so that loads and stores become single 128-bit accesses.
template <typename T, int V>
struct alignas(sizeof(T) * V >= 16 ? 16 : sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
EVX_HD Vec<T, V> vec_load(const T* p) {
  return *reinterpret_cast<const Vec<T, V>*>(p);
}
template <typename T, int V>
EVX_HD void vec_store(T* p, const Vec<T, V>& x) {
  *reinterpret_cast<Vec<T, V>*>(p) = x;
}
template <typename T, int V>
EVX_HD Vec<T, V> vec_splat(T a) {
  Vec<T, V> r;
#pragma unroll
  for (int k = 0; k < V; ++k) r.v[k] = a;
  return r;
}

// torch.clip(c, 0, 1).  On the device float uses the single-instruction saturate; a NaN
// input becomes 0 there instead of propagating, which is harmless for the time steppers:
// u+ = u + update keeps the NaN of u, so the solver's per-frame NaN abort still fires.
template <typename T>
EVX_HD T clip01(T a) {
  return a < T(0) ? T(0) : (a > T(1) ? T(1) : a);
}
#if defined(__CUDA_ARCH__)
template <>
EVX_HD float clip01<float>(float a) {
  return __saturatef(a);
}
#endif

EVX_HD int wrap_index(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}
EVX_HD int clamp_index(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

// Asynchronous 16-byte global -> shared copy (cp.async / LDGSTS): the data never passes
// through a register, so no scoreboard of the issuing warp is tied up while it is in flight.
EVX_HD void async_copy16(void* smem_dst, const void* gmem_src) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#else
  // host replay: the copy lands immediately (callers never touch the destination between
  // issue and wait)
  const unsigned char* s = reinterpret_cast<const unsigned char*>(gmem_src);
  unsigned char* d = reinterpret_cast<unsigned char*>(smem_dst);
  for (int i = 0; i < 16; ++i) d[i] = s[i];
#endif
}
// same for a whole Vec<T, V> (4, 8 or 16 bytes)
template <int BYTES>
EVX_HD void async_copy_bytes(void* smem_dst, const void* gmem_src) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async moves 4, 8 or 16 bytes");
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
  else if (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
#else
  const unsigned char* s = reinterpret_cast<const unsigned char*>(gmem_src);
  unsigned char* d = reinterpret_cast<unsigned char*>(smem_dst);
  for (int i = 0; i < BYTES; ++i) d[i] = s[i];
#endif
}
EVX_HD void async_copy_commit_and_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}
EVX_HD void async_copy_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
EVX_HD void async_copy_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

}  // namespace evx
