// Semi-implicit spectral stage:  out = u + irfftn( P(k) * rfftn(r) ),
//   P(k) = dt / (1 + dt * coef * |k|^(2*power))
// (reference evoxels/timesteppers.py:75-89).  This file holds the plan object, the cuFFT
// back end with its fused filter / add kernels, and the C-ABI entry points.  The native
// sm_100a FFT back end lives in fft_native.cu and is reached through the same plan.
#include <cuda_runtime.h>
#include <cufft.h>
#include <new>
#include "evx_internal.h"
#include "spectral_plan.h"
#include "spectral_math.h"

namespace evx {

// cuFFT layout: [n0, n1, n2/2+1] complex, contiguous.  Each block owns ROWS consecutive
// (i0,i1) rows and walks their elements with a flat index, one complex (8/16 B) per access.
template <typename R>
struct Cplx { R x, y; };

template <typename R>
__global__ void __launch_bounds__(256) spectral_filter_kernel(Cplx<R>* __restrict__ spec,
                                                              const FilterParams f,
                                                              int rows_per_block) {
  const unsigned nh = (unsigned)(f.n2 / 2 + 1);
  const long long rows = (long long)f.n0 * f.n1;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  const long long nrow = rows - row0 < rows_per_block ? rows - row0 : rows_per_block;
  const unsigned count = (unsigned)nrow * nh;
  Cplx<R>* base = spec + row0 * nh;
  for (unsigned e = threadIdx.x; e < count; e += blockDim.x) {
    const unsigned r = e / nh, i2 = e - r * nh;
    const long long row = row0 + r;
    const int i0 = (int)(row / f.n1), i1 = (int)(row - (long long)i0 * f.n1);
    const float k0 = wavenumber(signed_freq(i0, f.n0), f.inv_len0);
    const float k1 = wavenumber(signed_freq(i1, f.n1), f.inv_len1);
    const float k2v = wavenumber((int)i2, f.inv_len2);
    const float ksq = __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(k1, k1)), __fmul_rn(k2v, k2v));
    const R w = (R)spectral_weight(ksq, f) * (sizeof(R) == 8 ? (R)f.scale_d : (R)f.scale);
    Cplx<R> v = base[e];
    v.x *= w;
    v.y *= w;
    base[e] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ u,
                                                  const T* __restrict__ upd,
                                                  T* __restrict__ out, long long n) {
  constexpr int V = 16 / (int)sizeof(T);
  const long long nv = n / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned16_dev(u) && aligned16_dev(upd) && aligned16_dev(out)) {
    for (long long i = t0; i < nv; i += stride) {
      Vec<T, V> a = vec_load<T, V>(u + i * V), b = vec_load<T, V>(upd + i * V), c;
#pragma unroll
      for (int k = 0; k < V; ++k) c.v[k] = a.v[k] + b.v[k];
      vec_store<T, V>(out + i * V, c);
    }
    for (long long i = nv * V + t0; i < n; i += stride) out[i] = u[i] + upd[i];
  } else {
    for (long long i = t0; i < n; i += stride) out[i] = u[i] + upd[i];
  }
}

template <typename R>
static int launch_filter(R* spec, const FilterParams& f, cudaStream_t st) {
  const long long rows = (long long)f.n0 * f.n1;
  const int nh = f.n2 / 2 + 1;
  int rpb = (int)((8192 + nh - 1) / nh);          // ~8k complex values per block
  if (rpb < 1) rpb = 1;
  const long long grid = (rows + rpb - 1) / rpb;
  if (grid > 2147483647LL) return EVX_ERR_UNSUPPORTED;
  spectral_filter_kernel<R><<<(unsigned)grid, 256, 0, st>>>((Cplx<R>*)spec, f, rpb);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T>
static int launch_add(const T* u, const T* upd, T* out, long long n, cudaStream_t st) {
  long long blocks = (n / (16 / (int)sizeof(T)) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  add_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(u, upd, out, n);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static int cufft_err(cufftResult r) { return r == CUFFT_SUCCESS ? 0 : EVX_ERR_CUFFT - (int)r; }

int plan_create(evx_imex_plan** out, int nx, int ny, int nz, int is_f64, int backend) {
  if (!out || nx < 1 || ny < 1 || nz < 1) return EVX_ERR_ARG;
  if (backend < EVX_FFT_AUTO || backend > EVX_FFT_NATIVE_MIXED) return EVX_ERR_ARG;
  evx_imex_plan* p = new (std::nothrow) evx_imex_plan();
  if (!p) return EVX_ERR_ARG;
  p->nx = nx; p->ny = ny; p->nz = nz; p->is_f64 = is_f64;
  // squeeze size-1 extents to the front: the memory layout is unchanged and the spectrum
  // layout is internal, so (16,1,1) is transformed as a rank-1 problem of length 16.
  int src[3] = {nx, ny, nz};
  int k = 2;
  for (int a = 0; a < 3; ++a) { p->n[a] = 1; p->axis_of[a] = -1; }
  for (int a = 2; a >= 0; --a)
    if (src[a] > 1) { p->n[k] = src[a]; p->axis_of[k] = a; --k; }
  p->rank = 2 - k;
  if (p->rank == 0) { p->rank = 1; p->axis_of[2] = 2; }
  p->real_elems = (size_t)nx * ny * nz;
  p->spec_elems = (size_t)p->n[0] * p->n[1] * (p->n[2] / 2 + 1);

  // auto: radix-8 passes for power-of-two float32 grids, mixed-radix passes for every other
  // grid whose extents are 7-smooth, cuFFT for the rest (a prime factor > 7 somewhere)
  const bool native_ok = !is_f64 && native_fft_supported(nx, ny, nz);
  const bool mixed_ok = generic_fft_supported(nx, ny, nz);
  if (backend == EVX_FFT_NATIVE && !native_ok) { delete p; return EVX_ERR_UNSUPPORTED; }
  if (backend == EVX_FFT_NATIVE_MIXED && !mixed_ok) { delete p; return EVX_ERR_UNSUPPORTED; }
  if (backend == EVX_FFT_AUTO)
    p->backend = native_ok ? EVX_FFT_NATIVE : (mixed_ok ? EVX_FFT_NATIVE_MIXED : EVX_FFT_CUFFT);
  else
    p->backend = backend;

  const size_t esz = is_f64 ? 8 : 4;
  p->real_bytes = align256(p->real_elems * esz);
  if (p->backend == EVX_FFT_NATIVE) {
    int rc = native_plan_init(p);
    if (rc) { delete p; return rc; }
  } else if (p->backend == EVX_FFT_NATIVE_MIXED) {
    int rc = generic_plan_init(p);
    if (rc) { delete p; return rc; }
  } else {
    p->spec_bytes = align256(p->spec_elems * 2 * esz);
    int* dims = p->n + (3 - p->rank);
    size_t w1 = 0, w2 = 0;
    cufftResult r;
    if ((r = cufftCreate(&p->fwd)) != CUFFT_SUCCESS) { delete p; return cufft_err(r); }
    if ((r = cufftCreate(&p->inv)) != CUFFT_SUCCESS) { cufftDestroy(p->fwd); delete p; return cufft_err(r); }
    p->have_cufft = true;
    cufftSetAutoAllocation(p->fwd, 0);
    cufftSetAutoAllocation(p->inv, 0);
    r = cufftMakePlanMany(p->fwd, p->rank, dims, nullptr, 1, 0, nullptr, 1, 0,
                          is_f64 ? CUFFT_D2Z : CUFFT_R2C, 1, &w1);
    if (r == CUFFT_SUCCESS)
      r = cufftMakePlanMany(p->inv, p->rank, dims, nullptr, 1, 0, nullptr, 1, 0,
                            is_f64 ? CUFFT_Z2D : CUFFT_C2R, 1, &w2);
    if (r != CUFFT_SUCCESS) { plan_destroy(p); return cufft_err(r); }
    p->work_bytes = align256(w1 > w2 ? w1 : w2);
  }
  *out = p;
  return EVX_OK;
}

int plan_destroy(evx_imex_plan* p) {
  if (!p) return EVX_OK;
  if (p->have_cufft) { cufftDestroy(p->fwd); cufftDestroy(p->inv); }
  native_plan_free(p);
  delete p;
  return EVX_OK;
}

template <typename T>
struct CufftExec;
template <>
struct CufftExec<float> {
  static cufftResult fwd(cufftHandle h, float* in, void* out) { return cufftExecR2C(h, in, (cufftComplex*)out); }
  static cufftResult inv(cufftHandle h, void* in, float* out) { return cufftExecC2R(h, (cufftComplex*)in, out); }
};
template <>
struct CufftExec<double> {
  static cufftResult fwd(cufftHandle h, double* in, void* out) { return cufftExecD2Z(h, in, (cufftDoubleComplex*)out); }
  static cufftResult inv(cufftHandle h, void* in, double* out) { return cufftExecZ2D(h, (cufftDoubleComplex*)in, out); }
};

// out = u + irfftn(P * rfftn(r)); `r` may be the plan's own real scratch buffer.
template <typename T>
static int apply_cufft(evx_imex_plan* p, const T* u, const T* r, T* out, void* workspace,
                       const double* h, double dt, double coef, int power, cudaStream_t st) {
  char* ws = (char*)workspace;
  T* real_buf = (T*)ws;
  T* spec = (T*)(ws + p->real_bytes);
  void* work = ws + p->real_bytes + p->spec_bytes;
  cufftResult cr;
  if ((cr = cufftSetStream(p->fwd, st)) != CUFFT_SUCCESS) return cufft_err(cr);
  if ((cr = cufftSetStream(p->inv, st)) != CUFFT_SUCCESS) return cufft_err(cr);
  if (p->work_bytes) {
    if ((cr = cufftSetWorkArea(p->fwd, work)) != CUFFT_SUCCESS) return cufft_err(cr);
    if ((cr = cufftSetWorkArea(p->inv, work)) != CUFFT_SUCCESS) return cufft_err(cr);
  }
  if ((cr = CufftExec<T>::fwd(p->fwd, const_cast<T*>(r), spec)) != CUFFT_SUCCESS) return cufft_err(cr);
  double len_h[3];
  for (int a = 0; a < 3; ++a) len_h[a] = p->axis_of[a] >= 0 ? h[p->axis_of[a]] : 1.0;
  FilterParams f = make_filter(p->n, len_h, dt, coef, power, 1.0 / (double)p->real_elems);
  int rc = launch_filter<T>(spec, f, st);
  if (rc) return rc;
  if (!u) {  // update only
    if ((cr = CufftExec<T>::inv(p->inv, spec, out)) != CUFFT_SUCCESS) return cufft_err(cr);
    return EVX_OK;
  }
  if ((cr = CufftExec<T>::inv(p->inv, spec, real_buf)) != CUFFT_SUCCESS) return cufft_err(cr);
  return launch_add<T>(u, real_buf, out, (long long)p->real_elems, st);
}

template <typename T>
int imex_apply_impl(evx_imex_plan* p, const T* u, const T* r, T* out, void* workspace,
                    const double* h, double dt, double coef, int power, cudaStream_t st) {
  if (!p || !r || !out || !workspace || !h || r == out) return EVX_ERR_ARG;
  if (!valid_filter_spec(power)) return EVX_ERR_ARG;
  if ((sizeof(T) == 8) != (p->is_f64 != 0)) return EVX_ERR_ARG;
  if (!aligned16(workspace)) return EVX_ERR_ALIGN;
  if (p->backend == EVX_FFT_NATIVE)
    return native_apply(p, (const float*)u, (const float*)r, (float*)out, workspace, h, dt, coef,
                        power, st);
  if (p->backend == EVX_FFT_NATIVE_MIXED)
    return generic_apply<T>(p, u, r, out, workspace, h, dt, coef, power, st);
  if (filter_mirror(power)) return EVX_ERR_UNSUPPORTED;   // the caller extends the field itself
  return apply_cufft<T>(p, u, r, out, workspace, h, dt, coef, power, st);
}

template <typename T>
int ch_step_impl(evx_imex_plan* p, const T* u, const T* hom, T* out, void* workspace,
                 const double* h, double dt, double eps, double D, double A, cudaStream_t st) {
  if (!p || !u || !out || !workspace || !h) return EVX_ERR_ARG;
  if (u == out) return EVX_ERR_ARG;
  if ((sizeof(T) == 8) != (p->is_f64 != 0)) return EVX_ERR_ARG;
  if (!aligned16(workspace)) return EVX_ERR_ALIGN;
  const int per[3] = {BC_PERIODIC, BC_PERIODIC, BC_PERIODIC};
  if (p->backend == EVX_FFT_NATIVE)
    return native_ch_step(p, (const float*)u, (const float*)hom, (float*)out, workspace, h, dt,
                          eps, D, A, st);
  T* rhs = (T*)workspace;   // the plan's real scratch buffer
  int rc = ch_rhs_impl<T>(u, hom, rhs, p->nx, p->ny, p->nz, h, eps, D, per, nullptr, nullptr,
                          nullptr, st);
  if (rc) return rc;
  if (p->backend == EVX_FFT_NATIVE_MIXED)
    return generic_apply<T>(p, u, rhs, out, workspace, h, dt, 2.0 * eps * D * A, 2, st);
  return apply_cufft<T>(p, u, rhs, out, workspace, h, dt, 2.0 * eps * D * A, 2, st);
}

template <typename R>
static int filter_entry(void* spec, int nx, int ny, int nz, const double* h, double dt,
                        double coef, int power, double scale, cudaStream_t st) {
  if (!spec || !h || nx < 1 || ny < 1 || nz < 1 || !valid_filter_spec(power)) return EVX_ERR_ARG;
  const int n[3] = {nx, ny, nz};
  FilterParams f = make_filter(n, h, dt, coef, power, scale);
  return launch_filter<R>((R*)spec, f, st);
}

}  // namespace evx

using namespace evx;

extern "C" {

int evx_imex_plan_create(evx_imex_plan** plan, int nx, int ny, int nz, int is_f64, int backend) {
  return plan_create(plan, nx, ny, nz, is_f64, backend);
}
int evx_imex_plan_destroy(evx_imex_plan* plan) { return plan_destroy(plan); }
int evx_imex_plan_backend(const evx_imex_plan* plan) { return plan ? plan->backend : EVX_ERR_ARG; }
int evx_imex_plan_workspace_bytes(const evx_imex_plan* plan, size_t* bytes) {
  if (!plan || !bytes) return EVX_ERR_ARG;
  *bytes = plan->real_bytes + plan->spec_bytes + plan->work_bytes;
  return EVX_OK;
}

int evx_imex_apply_f32(evx_imex_plan* plan, const float* u, const float* r, float* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       void* stream) {
  return imex_apply_impl<float>(plan, u, r, out, workspace, h, dt, coef, power, (cudaStream_t)stream);
}
int evx_imex_apply_f64(evx_imex_plan* plan, const double* u, const double* r, double* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       void* stream) {
  return imex_apply_impl<double>(plan, u, r, out, workspace, h, dt, coef, power, (cudaStream_t)stream);
}
int evx_ch_imex_step_f32(evx_imex_plan* plan, const float* u, const float* hom, float* out,
                         void* workspace, const double* h, double dt, double eps, double D,
                         double A, void* stream) {
  return ch_step_impl<float>(plan, u, hom, out, workspace, h, dt, eps, D, A, (cudaStream_t)stream);
}
int evx_ch_imex_step_f64(evx_imex_plan* plan, const double* u, const double* hom, double* out,
                         void* workspace, const double* h, double dt, double eps, double D,
                         double A, void* stream) {
  return ch_step_impl<double>(plan, u, hom, out, workspace, h, dt, eps, D, A, (cudaStream_t)stream);
}
int evx_imex_native_pass_f32(evx_imex_plan* plan, int which, const float* u, const float* r,
                             float* out, void* workspace, const double* h, double dt, double coef,
                             int power, void* stream) {
  if (!plan || !workspace || !h || !u || !r || !out) return EVX_ERR_ARG;
  if (plan->backend != EVX_FFT_NATIVE) return EVX_ERR_UNSUPPORTED;
  return native_single_pass(plan, which, u, r, out, workspace, h, dt, coef, power,
                            (cudaStream_t)stream);
}
int evx_spectral_filter_c64(void* spec, int nx, int ny, int nz, const double* h, double dt,
                            double coef, int power, double scale, void* stream) {
  return filter_entry<float>(spec, nx, ny, nz, h, dt, coef, power, scale, (cudaStream_t)stream);
}
int evx_spectral_filter_c128(void* spec, int nx, int ny, int nz, const double* h, double dt,
                             double coef, int power, double scale, void* stream) {
  return filter_entry<double>(spec, nx, ny, nz, h, dt, coef, power, scale, (cudaStream_t)stream);
}

}  // extern "C"
