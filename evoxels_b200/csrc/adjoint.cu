// CUDA wrappers + C ABI of the Cahn-Hilliard rhs adjoint (see adjoint_core.h).
#include <cuda_runtime.h>
#include "evx_internal.h"
#include "adjoint_core.h"

namespace evx {

template <typename T, int OP>
__global__ void __launch_bounds__(256) ch_adjoint_kernel(const AdjParams<T> p) {
  const long long rows = (long long)p.nx * p.ny;
  double acc = 0.0;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int x = (int)(row / p.ny), y = (int)(row - (long long)x * p.ny);
    for (int z = threadIdx.x; z < p.nz; z += blockDim.x) {
      const long long i = row * p.nz + z;
      if (OP == 0) {
        p.out0[i] = adj_mu(p, x, y, z);
      } else if (OP == 1) {
        T zo, mo;
        adj_flux(p, x, y, z, zo, mo);
        p.out0[i] = zo;
        p.out1[i] = mo;
      } else {
        double term;
        p.out0[i] = adj_combine(p, x, y, z, term);
        if (x >= p.red_x0 && x < p.red_x1) acc += term;
      }
    }
  }
  if (OP == 2 && p.red) {
    __shared__ double part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += part[k];
      atomicAdd(p.red, s);
    }
  }
}

template <typename T>
static int launch_adjoint(int op, const T* u, const T* mu, const T* w, const T* z, const T* m,
                          const T* lam_in, T* out0, T* out1, double* red, int nx, int ny, int nz,
                          const double* h, double eps, double D, cudaStream_t st, int red_x0 = 0,
                          int red_x1 = 2147483647) {
  if (!u || !out0 || !h || nx < 1 || ny < 1 || nz < 1) return EVX_ERR_ARG;
  AdjParams<T> p;
  p.red_x0 = red_x0; p.red_x1 = red_x1;
  p.u = u; p.mu = mu; p.w = w; p.z = z; p.m = m; p.lam_in = lam_in; p.out0 = out0; p.out1 = out1;
  p.red = red; p.nx = nx; p.ny = ny; p.nz = nz;
  p.ihx2 = T(1.0 / (h[0] * h[0])); p.ihy2 = T(1.0 / (h[1] * h[1])); p.ihz2 = T(1.0 / (h[2] * h[2]));
  p.eps = T(eps); p.D = T(D);
  const long long rows = (long long)nx * ny;
  const unsigned grid = (unsigned)(rows < 148 * 16 ? rows : 148 * 16);
  const int threads = nz >= 256 ? 256 : (nz >= 128 ? 128 : (nz >= 64 ? 64 : 32));
  if (op == 0) {
    ch_adjoint_kernel<T, 0><<<grid, threads, 0, st>>>(p);
  } else if (op == 1) {
    if (!mu || !w || !out1) return EVX_ERR_ARG;
    ch_adjoint_kernel<T, 1><<<grid, threads, 0, st>>>(p);
  } else if (op == 2) {
    if (!z || !m) return EVX_ERR_ARG;
    ch_adjoint_kernel<T, 2><<<grid, threads, 0, st>>>(p);
  } else {
    return EVX_ERR_ARG;
  }
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace evx

using namespace evx;

extern "C" {

int evx_ch_mu_f32(const float* u, float* mu, int nx, int ny, int nz, const double* h, double eps,
                  void* stream) {
  return launch_adjoint<float>(0, u, nullptr, nullptr, nullptr, nullptr, nullptr, mu, nullptr,
                               nullptr, nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream);
}
int evx_ch_mu_f64(const double* u, double* mu, int nx, int ny, int nz, const double* h, double eps,
                  void* stream) {
  return launch_adjoint<double>(0, u, nullptr, nullptr, nullptr, nullptr, nullptr, mu, nullptr,
                                nullptr, nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream);
}
int evx_ch_adjoint_flux_f32(const float* u, const float* mu, const float* w, float* z, float* m,
                            int nx, int ny, int nz, const double* h, double D, void* stream) {
  return launch_adjoint<float>(1, u, mu, w, nullptr, nullptr, nullptr, z, m, nullptr, nx, ny, nz,
                               h, 1.0, D, (cudaStream_t)stream);
}
int evx_ch_adjoint_flux_f64(const double* u, const double* mu, const double* w, double* z,
                            double* m, int nx, int ny, int nz, const double* h, double D,
                            void* stream) {
  return launch_adjoint<double>(1, u, mu, w, nullptr, nullptr, nullptr, z, m, nullptr, nx, ny, nz,
                                h, 1.0, D, (cudaStream_t)stream);
}
int evx_ch_adjoint_combine_f32(const float* u, const float* z, const float* m,
                               const float* lam_in, float* lam_out, double* deps_acc, int nx,
                               int ny, int nz, const double* h, double eps, void* stream) {
  return launch_adjoint<float>(2, u, nullptr, nullptr, z, m, lam_in, lam_out, nullptr, deps_acc,
                               nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream);
}
int evx_ch_adjoint_combine_f64(const double* u, const double* z, const double* m,
                               const double* lam_in, double* lam_out, double* deps_acc, int nx,
                               int ny, int nz, const double* h, double eps, void* stream) {
  return launch_adjoint<double>(2, u, nullptr, nullptr, z, m, lam_in, lam_out, nullptr, deps_acc,
                                nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream);
}

int evx_ch_adjoint_combine_range_f32(const float* u, const float* z, const float* m,
                                     const float* lam_in, float* lam_out, double* deps_acc, int nx,
                                     int ny, int nz, const double* h, double eps, int x_lo, int x_hi,
                                     void* stream) {
  if (x_lo < 0 || x_hi > nx || x_lo > x_hi) return EVX_ERR_ARG;
  return launch_adjoint<float>(2, u, nullptr, nullptr, z, m, lam_in, lam_out, nullptr, deps_acc,
                               nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream, x_lo, x_hi);
}
int evx_ch_adjoint_combine_range_f64(const double* u, const double* z, const double* m,
                                     const double* lam_in, double* lam_out, double* deps_acc, int nx,
                                     int ny, int nz, const double* h, double eps, int x_lo, int x_hi,
                                     void* stream) {
  if (x_lo < 0 || x_hi > nx || x_lo > x_hi) return EVX_ERR_ARG;
  return launch_adjoint<double>(2, u, nullptr, nullptr, z, m, lam_in, lam_out, nullptr, deps_acc,
                                nx, ny, nz, h, eps, 1.0, (cudaStream_t)stream, x_lo, x_hi);
}

}  // extern "C"
