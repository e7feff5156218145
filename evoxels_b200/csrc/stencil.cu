// CUDA wrappers + C-ABI launchers for the stencil layer (see include/evoxels_b200.h).
// Kernel bodies live in ch_rhs_core.h / ac_core.h as barrier-free phase functions.
#include <cuda_runtime.h>
#include <cstdlib>
#include "evx_internal.h"
#include "evx_params.h"
#include "rd_core.h"

namespace evx {

std::atomic<unsigned long long> g_launches{0};

// ------------------------------------------------------------------------------------
// Cahn-Hilliard rhs
// ------------------------------------------------------------------------------------
template <typename T, int V, int TY, int G, bool HOM, bool GHOSTS, int MINB = 2>
__global__ void __launch_bounds__(ChRhsProgram<T, V, TY, G, HOM, GHOSTS>::NTHREADS,
                                  (sizeof(T) == 4 && TY <= 16 && !HOM) ? MINB : 1)
    ch_rhs_kernel(const ChParams<T> p) {
  using Prog = ChRhsProgram<T, V, TY, G, HOM, GHOSTS>;
  __shared__ typename Prog::Smem s;
  typename Prog::Regs t;
  Prog::init(t, s, p, threadIdx.x, blockIdx.x, blockIdx.y);
  __syncthreads();
#define EVX_CH_PLANE(PAR, ROT, OFF)                          \
  if (pl + (OFF) <= t.xb) { /* uniform across the block */   \
    Prog::template phase_a<PAR, ROT>(t, s, p, pl + (OFF));    \
    __syncthreads();                                          \
    Prog::template phase_b<PAR, ROT>(t, s, p, pl + (OFF));    \
  }
  for (int pl = t.xa - 1; pl <= t.xb; pl += 6) {
    EVX_CH_PLANE(0, 0, 0) EVX_CH_PLANE(1, 1, 1) EVX_CH_PLANE(0, 2, 2)
    EVX_CH_PLANE(1, 0, 3) EVX_CH_PLANE(0, 1, 4) EVX_CH_PLANE(1, 2, 5)
  }
#undef EVX_CH_PLANE
}

static int pick_xchunk(int nx, long long tiles, int min_chunk) {
  // enough CTAs for ~4 waves of 148 SMs x 2-4 resident CTAs, but chunks no shorter than
  // min_chunk planes (each chunk re-reads its 2+2 priming planes)
  long long want = (2400 + tiles - 1) / tiles;
  long long maxc = (nx + min_chunk - 1) / min_chunk;
  long long chunks = want < 1 ? 1 : (want > maxc ? maxc : want);
  if (chunks < 1) chunks = 1;
  return (int)((nx + chunks - 1) / chunks);
}

template <typename T, int V, int TY, int G>
static int launch_ch(ChParams<T> p, cudaStream_t st) {
  using Prog = ChRhsProgram<T, V, TY, G>;
  const long long tiles = (long long)((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  p.xchunk = pick_xchunk(p.nx, tiles, 32);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  if (tiles > 2147483647LL || chunks > 65535) return EVX_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  const bool ghosts = p.bc_kind[0] != BC_PERIODIC || p.bc_kind[1] != BC_PERIODIC ||
                      p.bc_kind[2] != BC_PERIODIC;
  if (p.hom) {   // user potential: rare, keep one (general) instantiation
    ch_rhs_kernel<T, V, TY, G, true, true><<<grid, Prog::NTHREADS, 0, st>>>(p);
  } else if (ghosts) {
    static const int gocc = [] { const char* e = getenv("EVX_CH_GOCC"); return e ? atoi(e) : 3; }();
    if (gocc == 3 && sizeof(T) == 4 && TY <= 16)
      ch_rhs_kernel<T, V, TY, G, false, true, 3><<<grid, Prog::NTHREADS, 0, st>>>(p);
    else
      ch_rhs_kernel<T, V, TY, G, false, true><<<grid, Prog::NTHREADS, 0, st>>>(p);
  } else {
    // three resident CTAs per SM (72 registers, a few spilled words) beat two (94 registers):
    // 0.365 vs 0.408 ms at 512^3 - the kernel is latency-bound, the extra warps pay
    static const int occ = [] { const char* e = getenv("EVX_CH_OCC"); return e ? atoi(e) : 3; }();
    if (occ == 3 && sizeof(T) == 4 && TY <= 16)
      ch_rhs_kernel<T, V, TY, G, false, false, 3><<<grid, Prog::NTHREADS, 0, st>>>(p);
    else
      ch_rhs_kernel<T, V, TY, G, false, false><<<grid, Prog::NTHREADS, 0, st>>>(p);
  }
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T>
int ch_rhs_impl(const T* c, const T* hom, T* rhs, int nx, int ny, int nz, const double* h,
                double eps, double D, const int* bc_kind, const double* bc_val,
                const T* halo_lo, const T* halo_hi, cudaStream_t st) {
  if (!c || !rhs || !h || !bc_kind || nx < 1 || ny < 1 || nz < 1) return EVX_ERR_ARG;
  if (c == rhs) return EVX_ERR_ARG;
  for (int a = 0; a < 3; ++a) {
    if (bc_kind[a] < 0 || bc_kind[a] > 2) return EVX_ERR_ARG;
    if (bc_kind[a] == BC_DIRICHLET && !bc_val) return EVX_ERR_ARG;
  }
  if (hom && (halo_lo || halo_hi)) return EVX_ERR_UNSUPPORTED;
  ChParams<T> p = make_ch_params<T>(c, hom, rhs, nx, ny, nz, h, eps, D, bc_kind, bc_val,
                                    halo_lo, halo_hi, nx);
  constexpr int VW = 16 / (int)sizeof(T);
  const bool vec = nz % VW == 0 && aligned16(c) && aligned16(rhs) && aligned16(hom) &&
                   aligned16(halo_lo) && aligned16(halo_hi);
  if (vec) {
    if constexpr (sizeof(T) == 4) {
      // warp-specialised form (loader warp + bulk copies, ch_rhs_tma.cu); it declines shapes
      // it does not cover
      const int e = ch_rhs_tma_f32(p, st);
      if (e != EVX_ERR_UNSUPPORTED) return e;
    }
    // tile height: 14 rows x 16 groups = 7 interior warps + 2 ring warps = 288 threads
    return launch_ch<T, VW, 14, 16>(p, st);
  }
  return launch_ch<T, 1, 8, 32>(p, st);
}

// ------------------------------------------------------------------------------------
// Allen-Cahn stage
// ------------------------------------------------------------------------------------
template <typename T, int V, int TY, int G, int MINB>
__global__ void __launch_bounds__(TY* G, MINB) ac_stage_kernel(const AcParams<T> p) {
  AcProgram<T, V, TY, G>::run(p, threadIdx.x, blockIdx.x, blockIdx.y);
}

template <typename T, int V, int TY, int G, int MINB = 2>
__global__ void __launch_bounds__(AcTileProgram<T, V, TY, G>::NTHREADS, sizeof(T) == 4 ? MINB : 1)
    ac_tile_kernel(const AcParams<T> p) {
  using Prog = AcTileProgram<T, V, TY, G>;
  __shared__ typename Prog::Smem s;
  typename Prog::Regs t;
  Prog::init(t, s, p, threadIdx.x, blockIdx.x, blockIdx.y);
  __syncthreads();
#define EVX_AC_PLANE(ROT, OFF)                              \
  if (x + (OFF) < t.xb) { /* uniform across the block */    \
    Prog::template phase_a<ROT>(t, s, p, x + (OFF));        \
    __syncthreads();                                        \
    Prog::template phase_b<ROT>(t, s, p, x + (OFF));        \
  }
  for (int x = t.xa; x < t.xb; x += 4) {
    EVX_AC_PLANE(0, 0) EVX_AC_PLANE(1, 1) EVX_AC_PLANE(2, 2) EVX_AC_PLANE(3, 3)
  }
#undef EVX_AC_PLANE
}

template <typename T, int V, int TY, int G>
static int launch_ac_tile(AcParams<T> p, cudaStream_t st) {
  using Prog = AcTileProgram<T, V, TY, G>;
  const long long tiles = (long long)((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  p.xchunk = pick_xchunk(p.nx, tiles, 32);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  if (tiles > 2147483647LL || chunks > 65535) return EVX_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  // 72 registers without spills at three CTAs per SM: 0.510 vs 0.540 ms at 512^3 (four CTAs
  // at 56 registers spill and fall back to 0.534 ms)
  static const int occ = [] { const char* e = getenv("EVX_AC_OCC"); return e ? atoi(e) : 3; }();
  if (occ == 3) ac_tile_kernel<T, V, TY, G, 3><<<grid, Prog::NTHREADS, 0, st>>>(p);
  else ac_tile_kernel<T, V, TY, G, 2><<<grid, Prog::NTHREADS, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T, int V, int TY, int G>
static int launch_ac(AcParams<T> p, cudaStream_t st) {
  using Prog = AcProgram<T, V, TY, G>;
  const long long tiles = (long long)((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  p.xchunk = pick_xchunk(p.nx, tiles, 16);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  if (tiles > 2147483647LL || chunks > 65535) return EVX_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  ac_stage_kernel<T, V, TY, G, 2><<<grid, Prog::NTHREADS, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T>
int ac_stage_impl(const T* phi, const T* pot, T* k_out, const T* base, T* y_out, double alpha,
                  const T* acc_in, T* acc_out, double beta, int nx, int ny, int nz,
                  const double* h, double eps, double gab, double M, double force,
                  double curvature, const int* bc_kind, const double* bc_val, const T* halo_lo,
                  const T* halo_hi, cudaStream_t st) {
  if (!phi || !h || !bc_kind || nx < 1 || ny < 1 || nz < 1) return EVX_ERR_ARG;
  if (!k_out && !y_out && !acc_out) return EVX_ERR_ARG;
  if (y_out && !base) return EVX_ERR_ARG;
  if (k_out == phi || y_out == phi || acc_out == phi) return EVX_ERR_ARG;
  for (int a = 0; a < 3; ++a) {
    if (bc_kind[a] < 0 || bc_kind[a] > 2) return EVX_ERR_ARG;
    if (bc_kind[a] == BC_DIRICHLET && !bc_val) return EVX_ERR_ARG;
  }
  AcParams<T> p = make_ac_params<T>(phi, pot, k_out, base, y_out, alpha, acc_in, acc_out, beta,
                                    nx, ny, nz, h, eps, gab, M, force, curvature, bc_kind,
                                    bc_val, halo_lo, halo_hi, nx);
  constexpr int VW = 16 / (int)sizeof(T);
  const bool vec = nz % VW == 0 && aligned16(phi) && aligned16(pot) && aligned16(k_out) &&
                   aligned16(base) && aligned16(y_out) && aligned16(acc_in) &&
                   aligned16(acc_out) && aligned16(halo_lo) && aligned16(halo_hi);
  if (vec) {
    return launch_ac_tile<T, VW, 14, 16>(p, st);
  }
  return launch_ac<T, 1, 8, 32>(p, st);     // register-window variant: unaligned / odd nz
}

// ------------------------------------------------------------------------------------
// ghost padding and stencils on padded fields (API parity for VoxelGrid/FDStencils)
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pad_ghost_kernel(const PadParams<T> p) {
  const int K = p.nz + 2, J = p.ny + 2;
  const long long rows = (long long)(p.nx + 2) * J;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int i = (int)(row / J), j = (int)(row % J);
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      p.out[row * K + k] = padded_value(p, i, j, k);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) padded_stencil_kernel(const PaddedStencilParams<T> p) {
  const long long rows = (long long)p.nx * p.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int x = (int)(row / p.ny), y = (int)(row % p.ny);
    for (int z = threadIdx.x; z < p.nz; z += blockDim.x)
      p.out[row * p.nz + z] = padded_stencil_value(p, x, y, z);
  }
}

template <typename T>
int pad_ghost_impl(const T* in, T* out, int nx, int ny, int nz, const int* bc_kind,
                   const double* bc_val, cudaStream_t st) {
  if (!in || !out || !bc_kind || nx < 1 || ny < 1 || nz < 1) return EVX_ERR_ARG;
  for (int a = 0; a < 3; ++a) {
    if (bc_kind[a] < 0 || bc_kind[a] > 2) return EVX_ERR_ARG;
    if (bc_kind[a] == BC_DIRICHLET && !bc_val) return EVX_ERR_ARG;
  }
  PadParams<T> p = make_pad_params<T>(in, out, nx, ny, nz, bc_kind, bc_val);
  const long long rows = (long long)(nx + 2) * (ny + 2);
  const unsigned grid = (unsigned)(rows < 148 * 16 ? rows : 148 * 16);
  pad_ghost_kernel<T><<<grid, 256, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T>
int padded_stencil_impl(const T* g, T* out, int nx, int ny, int nz, const double* h, int op,
                        cudaStream_t st) {
  if (!g || !out || !h || nx < 1 || ny < 1 || nz < 1 || op < 0 || op > 2) return EVX_ERR_ARG;
  PaddedStencilParams<T> p = make_padded_stencil_params<T>(g, out, nx, ny, nz, h, op);
  const long long rows = (long long)nx * ny;
  const unsigned grid = (unsigned)(rows < 148 * 16 ? rows : 148 * 16);
  padded_stencil_kernel<T><<<grid, 256, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// two-species reaction-diffusion rhs
// ------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) rd_rhs_kernel(const RdParams<T> p) {
  const long long groups = (long long)p.nx * p.ny * (p.nz / V);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x)
    RdProgram<T, V>::run(p, g);
}

template <typename T>
static int rd_rhs_impl(const T* u, const T* inter, T* out, int nx, int ny, int nz, const double* h,
                       double DA, double DB, double feed, double kill, cudaStream_t st) {
  if (!u || !out || !h || nx < 1 || ny < 1 || nz < 1 || u == out) return EVX_ERR_ARG;
  RdParams<T> p;
  p.u = u; p.inter = inter; p.out = out; p.nx = nx; p.ny = ny; p.nz = nz;
  fill_rd_metric(p, h);
  p.DA = (T)DA; p.DB = (T)DB; p.feed = (T)feed; p.kill = (T)kill;
  constexpr int VW = 16 / (int)sizeof(T);
  const long long n = (long long)nx * ny * nz;
  // both species must be 16-byte aligned for the vector path (n * sizeof(T) offset)
  const bool vec = nz % VW == 0 && aligned16(u) && aligned16(out) && aligned16(inter) &&
                   (n * (long long)sizeof(T)) % 16 == 0;
  const long long groups = vec ? n / VW : n;
  long long blocks = (groups + 255) / 256;
  const long long cap = 148LL * 8 * 8;       // grid-stride beyond ~8 waves of 8 CTAs/SM
  if (blocks > cap) blocks = cap;
  if (vec) rd_rhs_kernel<T, VW><<<(unsigned)blocks, 256, 0, st>>>(p);
  else rd_rhs_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template int ch_rhs_impl<float>(const float*, const float*, float*, int, int, int, const double*,
                                double, double, const int*, const double*, const float*,
                                const float*, cudaStream_t);
template int ch_rhs_impl<double>(const double*, const double*, double*, int, int, int,
                                 const double*, double, double, const int*, const double*,
                                 const double*, const double*, cudaStream_t);

}  // namespace evx

using namespace evx;

extern "C" {

int evx_version(void) { return EVX_VERSION; }

unsigned long long evx_launch_count(void) { return g_launches.load(); }

const char* evx_strerror(int code) {
  if (code == EVX_OK) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case EVX_ERR_ARG: return "evoxels_b200: invalid argument";
    case EVX_ERR_UNSUPPORTED: return "evoxels_b200: configuration not supported on device";
    case EVX_ERR_ALIGN: return "evoxels_b200: pointer alignment";
    default: break;
  }
  if (code <= EVX_ERR_CUFFT) return "evoxels_b200: cuFFT call failed (code = -1000 - cufftResult)";
  return "evoxels_b200: unknown error";
}

int evx_ch_rhs_f32(const float* c, const float* hom, float* rhs, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bc_kind,
                   const double* bc_val, const float* halo_lo, const float* halo_hi,
                   void* stream) {
  return ch_rhs_impl<float>(c, hom, rhs, nx, ny, nz, h, eps, D, bc_kind, bc_val, halo_lo,
                            halo_hi, (cudaStream_t)stream);
}
int evx_ch_rhs_f64(const double* c, const double* hom, double* rhs, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bc_kind,
                   const double* bc_val, const double* halo_lo, const double* halo_hi,
                   void* stream) {
  return ch_rhs_impl<double>(c, hom, rhs, nx, ny, nz, h, eps, D, bc_kind, bc_val, halo_lo,
                             halo_hi, (cudaStream_t)stream);
}

int evx_ac_stage_f32(const float* phi, const float* pot, float* k_out, const float* base,
                     float* y_out, double alpha, const float* acc_in, float* acc_out,
                     double beta, int nx, int ny, int nz, const double* h, double eps,
                     double gab, double M, double force, double curvature,
                     const int* bc_kind, const double* bc_val, const float* halo_lo,
                     const float* halo_hi, void* stream) {
  return ac_stage_impl<float>(phi, pot, k_out, base, y_out, alpha, acc_in, acc_out, beta, nx,
                              ny, nz, h, eps, gab, M, force, curvature, bc_kind, bc_val,
                              halo_lo, halo_hi, (cudaStream_t)stream);
}
int evx_ac_stage_f64(const double* phi, const double* pot, double* k_out, const double* base,
                     double* y_out, double alpha, const double* acc_in, double* acc_out,
                     double beta, int nx, int ny, int nz, const double* h, double eps,
                     double gab, double M, double force, double curvature,
                     const int* bc_kind, const double* bc_val, const double* halo_lo,
                     const double* halo_hi, void* stream) {
  return ac_stage_impl<double>(phi, pot, k_out, base, y_out, alpha, acc_in, acc_out, beta, nx,
                               ny, nz, h, eps, gab, M, force, curvature, bc_kind, bc_val,
                               halo_lo, halo_hi, (cudaStream_t)stream);
}

int evx_pad_ghost_f32(const float* in, float* out, int nx, int ny, int nz, const int* bc_kind,
                      const double* bc_val, void* stream) {
  return pad_ghost_impl<float>(in, out, nx, ny, nz, bc_kind, bc_val, (cudaStream_t)stream);
}
int evx_pad_ghost_f64(const double* in, double* out, int nx, int ny, int nz, const int* bc_kind,
                      const double* bc_val, void* stream) {
  return pad_ghost_impl<double>(in, out, nx, ny, nz, bc_kind, bc_val, (cudaStream_t)stream);
}

int evx_padded_stencil_f32(const float* padded, float* out, int nx, int ny, int nz,
                           const double* h, int op, void* stream) {
  return padded_stencil_impl<float>(padded, out, nx, ny, nz, h, op, (cudaStream_t)stream);
}
int evx_padded_stencil_f64(const double* padded, double* out, int nx, int ny, int nz,
                           const double* h, int op, void* stream) {
  return padded_stencil_impl<double>(padded, out, nx, ny, nz, h, op, (cudaStream_t)stream);
}

int evx_rd2_rhs_f32(const float* u, const float* interaction, float* out, int nx, int ny, int nz,
                    const double* h, double D_A, double D_B, double feed, double kill,
                    void* stream) {
  return rd_rhs_impl<float>(u, interaction, out, nx, ny, nz, h, D_A, D_B, feed, kill,
                            (cudaStream_t)stream);
}
int evx_rd2_rhs_f64(const double* u, const double* interaction, double* out, int nx, int ny,
                    int nz, const double* h, double D_A, double D_B, double feed, double kill,
                    void* stream) {
  return rd_rhs_impl<double>(u, interaction, out, nx, ny, nz, h, D_A, D_B, feed, kill,
                             (cudaStream_t)stream);
}

}  // extern "C"
