// Launchers of the mixed-radix passes (fft_generic_core.h) behind evx_imex_plan for extents
// that are not powers of two (back end EVX_FFT_NATIVE_MIXED).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "evx_internal.h"
#include "spectral_plan.h"
#include "fft_generic_core.h"

namespace evx {

constexpr int kGenThreads = 256;
constexpr size_t kGenSmemCap = 160 * 1024;

template <typename R>
__global__ void __launch_bounds__(kGenThreads) fft_generic_kernel(const GenericParams<R> p,
                                                                  long long nblocks) {
  extern __shared__ __align__(16) unsigned char gen_smem_raw[];
  gcplx<R>* smem = reinterpret_cast<gcplx<R>*>(gen_smem_raw);
  const int nph = GenericProgram<R>::nphases(p);
  for (long long block = blockIdx.x; block < nblocks; block += gridDim.x) {
    for (int k = 0; k < nph; ++k) {
      __syncthreads();         // also protects the buffers of the previous block
      GenericProgram<R>::phase(k, p, smem, block, threadIdx.x, (int)blockDim.x);
    }
  }
}

bool generic_fft_supported(int nx, int ny, int nz) {
  LineDesc d;
  const int n[3] = {nx, ny, nz};
  for (int a = 0; a < 3; ++a)
    if (n[a] < 1 || n[a] > 4096 || !factor_line(n[a], d)) return false;
  return true;
}

// lines per block / threads per block: 8 lines and 128 threads measured best at 100^3 and 150^3
// (82 / 172 us per CH step; 16 lines and 256 threads: 85 / 178); fewer lines where the two
// [N][W] buffers would exceed ~48 KB and cost resident blocks
static int gen_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int lines_per_block(int N, size_t esz) {
  int W = gen_env("EVX_GEN_W", 8);
  if (W != 1 && W != 2 && W != 4 && W != 8 && W != 16 && W != 32) W = 8;
  while (W > 1 && 2 * (size_t)N * W * 2 * esz > 48 * 1024) W /= 2;
  return W;
}

template <typename R>
static int launch_generic(GenericParams<R> p, long long nblocks, cudaStream_t st) {
  if (nblocks < 1) return EVX_ERR_UNSUPPORTED;
  const size_t smem = GenericProgram<R>::smem_elems(p) * sizeof(gcplx<R>);
  if (smem > kGenSmemCap) return EVX_ERR_UNSUPPORTED;
  auto kern = fft_generic_kernel<R>;
  static SmemOptIn optin;              // per instantiation (float / double)
  if (int rc = optin.ensure(kern, kGenSmemCap)) return rc;
  const long long cap = 148LL * 16;
  const unsigned grid = (unsigned)(nblocks < cap ? nblocks : cap);
  int threads = gen_env("EVX_GEN_THREADS", 128);
  if (threads < 32 || threads > kGenThreads || threads % 32) threads = kGenThreads;
  kern<<<grid, threads, smem, st>>>(p, nblocks);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename R>
static void fill_roots_r(std::vector<gcplx<R>>& w, size_t off, int n) {
  for (int m = 0; m < n; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)n;
    w[off + m] = gcplx<R>{(R)std::cos(a), (R)std::sin(a)};
  }
}

template <typename R>
static int generic_tables(evx_imex_plan* p) {
  // W_nx | W_ny | W_nz | W_2nx (x pass of a mirrored, non-periodic x axis)
  const size_t total = (size_t)p->nx + p->ny + p->nz + 2 * (size_t)p->nx;
  std::vector<gcplx<R>> host(total);
  fill_roots_r<R>(host, 0, p->nx);
  fill_roots_r<R>(host, p->nx, p->ny);
  fill_roots_r<R>(host, (size_t)p->nx + p->ny, p->nz);
  fill_roots_r<R>(host, (size_t)p->nx + p->ny + p->nz, 2 * p->nx);
  cudaError_t e = cudaMalloc(&p->twiddles, total * sizeof(gcplx<R>));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpy(p->twiddles, host.data(), total * sizeof(gcplx<R>), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(p->twiddles); p->twiddles = nullptr; return (int)e; }
  return EVX_OK;
}

int generic_plan_init(evx_imex_plan* p) {
  const size_t csz = p->is_f64 ? 16 : 8;
  p->spec_pitch = ((p->nz / 2 + 1 + 7) / 8) * 8;
  p->spec_bytes = ((size_t)p->nx * p->ny * p->spec_pitch * csz + 255) & ~(size_t)255;
  p->work_bytes = 0;
  return p->is_f64 ? generic_tables<double>(p) : generic_tables<float>(p);
}

template <typename R>
int generic_apply(evx_imex_plan* pl, const R* u, const R* r, R* out, void* workspace,
                  const double* h, double dt, double coef, int power, cudaStream_t st) {
  const int nx = pl->nx, ny = pl->ny, nz = pl->nz, P = pl->spec_pitch;
  gcplx<R>* spec = reinterpret_cast<gcplx<R>*>((char*)workspace + pl->real_bytes);
  const gcplx<R>* twx = (const gcplx<R>*)pl->twiddles;
  const gcplx<R>* twy = twx + nx;
  const gcplx<R>* twz = twy + ny;
  const int mirror = filter_mirror(power);
  const int nxl = mirror ? 2 * nx : nx;       // points of an x line
  const gcplx<R>* twx2 = twz + nz;
  LineDesc lx, ly, lz;
  if (!factor_line(nxl, lx) || !factor_line(ny, ly) || !factor_line(nz, lz)) return EVX_ERR_UNSUPPORTED;
  if (nxl > 4096) return EVX_ERR_UNSUPPORTED;
  const size_t esz = sizeof(R);
  int rc;

  GenericParams<R> p{};
  p.spec = spec; p.nz = nz; p.P = P; p.rows = (long long)nx * ny; p.ncols_valid = nz / 2 + 1;

  // z forward
  p.mode = GEN_Z_FWD; p.line = lz; p.tw = twz; p.W = lines_per_block(nz, esz);
  p.real_in = r; p.real_out = nullptr;
  if ((rc = launch_generic<R>(p, (p.rows + p.W - 1) / p.W, st))) return rc;

  // y forward: columns (x, kz)
  GenericParams<R> y = p;
  y.mode = GEN_FWD; y.line = ly; y.tw = twy; y.W = lines_per_block(ny, esz);
  y.line_stride = P; y.group_stride = (long long)ny * P; y.ncols_total = (long long)nx * P;
  if ((rc = launch_generic<R>(y, (y.ncols_total + y.W - 1) / y.W, st))) return rc;

  // x forward * weight * x inverse: columns (y, kz)
  GenericParams<R> x = p;
  x.mode = GEN_XMID; x.line = lx; x.tw = mirror ? twx2 : twx; x.W = lines_per_block(nxl, esz);
  x.line_stride = (long long)ny * P; x.group_stride = P; x.ncols_total = (long long)ny * P;
  x.mirror = mirror;
  const int n[3] = {nxl, ny, nz};
  x.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)nxl * ny * nz));
  if ((rc = launch_generic<R>(x, (x.ncols_total + x.W - 1) / x.W, st))) return rc;

  // y inverse
  y.mode = GEN_INV;
  if ((rc = launch_generic<R>(y, (y.ncols_total + y.W - 1) / y.W, st))) return rc;

  // z inverse, out = u + update
  p.mode = GEN_Z_INV; p.real_in = u; p.real_out = out;
  return launch_generic<R>(p, (p.rows + p.W - 1) / p.W, st);
}

template int generic_apply<float>(evx_imex_plan*, const float*, const float*, float*, void*,
                                  const double*, double, double, int, cudaStream_t);
template int generic_apply<double>(evx_imex_plan*, const double*, const double*, double*, void*,
                                   const double*, double, double, int, cudaStream_t);

}  // namespace evx
