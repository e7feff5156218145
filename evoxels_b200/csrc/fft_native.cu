// Native sm_100a FFT back end of the IMEX plan: five hand-written passes
//   ZFwd -> Y fwd -> X fwd*filter*X inv -> Y inv -> ZInv(+u)
// over a pitched half spectrum (see fft_pass_core.h).  Power-of-two extents:
// 8 <= nx, ny <= 2048, 16 <= nz <= 2048.
#include <cmath>
#include <cstdlib>
#include <vector>
#include "evx_internal.h"
#include "spectral_plan.h"
#include "fft_pass_core.h"
#include "fft_line.h"
#include "fft_chain.h"
#include "tma_ptx.h"
#include "dist_params.h"

namespace evx {

constexpr int kPitchAlign = 8;   // spectrum rows are padded to a multiple of 8 complex

// MINB = 0: default residency for the block size
// WARP_LINES: every line of the program is owned by exactly one warp (z passes with M/8 == 32
// threads per line) - the phases then only need warp-level synchronisation, the eight warps of
// a block drift apart and their load, exchange and arithmetic phases overlap
// (512^3: inverse z pass 279 -> 256 us, forward 228 -> 224 us; profiles/r02_exp_passes.json)
// L2 prefetch of the contiguous input rows of the block that will run `pf_blocks` blocks later
// (one bulk prefetch per array and block, issued by one thread): the z passes load straight
// into registers, so their memory pipeline is only as deep as the resident warps - the
// prefetch turns the DRAM latency of a later block into an L2 hit.
template <class Prog>
__device__ __forceinline__ void pass_prefetch(const ZParams& p, long long block) {
  if (p.pf_blocks <= 0 || threadIdx.x != 0) return;
  constexpr int NL = Prog::NTHREADS / Prog::T;
  const long long row = (block + p.pf_blocks) * NL;
  if (row + NL > p.rows) return;
  if (p.real_in) bulk_prefetch_l2(p.real_in + row * p.nz, (unsigned)(NL * p.nz * sizeof(float)));
  if (p.real_out) bulk_prefetch_l2(p.spec + row * p.P, (unsigned)(NL * p.P * sizeof(cf)));
}
template <class Prog>
__device__ __forceinline__ void pass_prefetch(const StridedParams&, long long) {}

template <class Prog, class Params, int MINB = 0, bool WARP_LINES = false>
__global__ void __launch_bounds__(Prog::NTHREADS, MINB ? MINB : (Prog::NTHREADS <= 256 ? 4 : (Prog::NTHREADS <= 512 ? 2 : 1))) fft_pass_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* smem = reinterpret_cast<cf*>(smem_raw);
  typename Prog::Regs r;
  Prog::init(r, p, threadIdx.x, (long long)blockIdx.x);
  pass_prefetch<Prog>(p, (long long)blockIdx.x);
#pragma unroll
  for (int k = 0; k < Prog::NPHASES; ++k) {
    if (k) { if (WARP_LINES) __syncwarp(); else __syncthreads(); }
    Prog::phase(k, r, smem, p);
  }
}

template <class Prog, class Params, int MINB = 0, bool WARP_LINES = false>
static int launch_pass(const Params& p, long long blocks, cudaStream_t st) {
  if (blocks < 1 || blocks > 2147483647LL) return EVX_ERR_UNSUPPORTED;
  auto kern = fft_pass_kernel<Prog, Params, MINB, WARP_LINES>;
  if (Prog::SMEM_BYTES > 48 * 1024) {
    static SmemOptIn optin;           // per instantiation
    if (int rc = optin.ensure(kern, Prog::SMEM_BYTES)) return rc;
  }
  kern<<<(unsigned)blocks, Prog::NTHREADS, Prog::SMEM_BYTES, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

// Columns per tile.  The x pass walks lines whose points are ny*P elements (~1 MB) apart:
// measured with an access-pattern probe (PASS_COPY, 512^3; scripts/exp) a tile of 8
// columns (64-B rows) tops out at 3.3 TB/s there, 16 columns (128-B rows) at 5.3 TB/s,
// while the y pass (2 KB stride) is already at 6.3 TB/s with 8.  x-pass tiles may straddle
// y groups (the column index is the linear offset), so the pitch stays a multiple of 8.
template <int L, int MODE>
struct StridedCfg {
  static constexpr int KZ = L >= 2048 ? 4 : ((pass_is_xmid(MODE) && L <= 512) ? 16 : 8);
};

// persistent, software-pipelined form (see StridedPipe in fft_pass_core.h)
template <class Pipe>
__global__ void __launch_bounds__(Pipe::NTHREADS, Pipe::NTHREADS <= 512 ? 2 : 1)
    fft_pipe_kernel(const StridedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* t0 = reinterpret_cast<cf*>(smem_raw);
  cf* t1 = t0 + Pipe::BUF;
  cf* c = t1 + Pipe::BUF;
  const long long ntiles = Pipe::num_tiles(p);
  long long tile = blockIdx.x;
  typename Pipe::Cursor q;
  Pipe::cursor_init(q, p, threadIdx.x, tile, gridDim.x);
  if (tile < ntiles) Pipe::prefetch_at(threadIdx.x, p, q.bgrp, q.bkz, t0);
  async_copy_commit();
  typename Pipe::Regs r;
  for (int par = 0; tile < ntiles; tile += gridDim.x, par ^= 1) {
    cf* a = par ? t1 : t0;
    cf* b = par ? t0 : t1;
    async_copy_commit_and_wait();
    __syncthreads();
    Pipe::Base::init_at(r, p, threadIdx.x, Pipe::cursor_column(q, p), q.grp, q.kz);
    Pipe::read_tile(r, a);
    Pipe::cursor_step_own(q, p);
    Pipe::cursor_step_base(q, p);
    if (tile + gridDim.x < ntiles) Pipe::prefetch_at(threadIdx.x, p, q.bgrp, q.bkz, b);
    async_copy_commit();
#pragma unroll
    for (int k = 0; k < Pipe::NPHASES; ++k) {
      if (k) __syncthreads();
      Pipe::phase(k, r, a, c, p);
    }
  }
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

static bool use_pipeline() {
  static const bool on = [] { const char* e = getenv("EVX_FFT_PIPE"); return !e || atoi(e) != 0; }();
  return on;
}

template <class Pipe>
static int launch_pipe(const StridedParams& p, cudaStream_t st) {
  const long long ntiles = Pipe::num_tiles(p);
  if (ntiles < 1) return EVX_ERR_UNSUPPORTED;
  auto kern = fft_pipe_kernel<Pipe>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, Pipe::SMEM_BYTES)) return rc;
  long long resident = (long long)sm_count() * (Pipe::NTHREADS <= 512 ? 2 : 1);
  if (p.max_ctas > 0 && p.max_ctas < resident) resident = p.max_ctas;
  const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
  kern<<<grid, Pipe::NTHREADS, Pipe::SMEM_BYTES, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template <int MODE>
static int launch_strided(int L, StridedParams p, cudaStream_t st);
// x forward . weight . x inverse; the weight formula is a compile-time choice so that the
// rarely used exponential-Euler branch costs the IMEX kernel no registers
static int launch_xmid(int L, const StridedParams& p, cudaStream_t st) {
  if (p.filt.kind == FILTER_ETD1) return launch_strided<PASS_XMID_ETD1>(L, p, st);
  return launch_strided<PASS_XMID>(L, p, st);
}

template <int MODE>
static int launch_strided(int L, StridedParams p, cudaStream_t st) {
  finalize_strided(p, L);
#define EVX_CASE(N)                                                                   \
  case N: {                                                                           \
    constexpr int KZ = StridedCfg<N, MODE>::KZ;                                       \
    if (N >= 64 && use_pipeline() && !p.mirror) return launch_pipe<StridedPipe<N, KZ, MODE>>(p, st); \
    return launch_pass<StridedPass<N, KZ, MODE>, StridedParams>(                      \
        p, (p.ncols_total + KZ - 1) / KZ, st);                                        \
  }
  switch (L) {
    EVX_CASE(8) EVX_CASE(16) EVX_CASE(32) EVX_CASE(64) EVX_CASE(128) EVX_CASE(256)
    EVX_CASE(512) EVX_CASE(1024) EVX_CASE(2048)
    default: return EVX_ERR_UNSUPPORTED;
  }
#undef EVX_CASE
}

template <bool INV>
static int launch_z(int M, const ZParams& p, cudaStream_t st) {
  // forward z pass of 512-point lines: 48 registers (no spills) let five CTAs share an SM
  // (227 vs 233 us at 512^3); the inverse pass spills below 64 registers and stays at four
  if (M == 256 && !INV) return launch_pass<ZPass<256, 8, INV>, ZParams, 5, true>(p, (p.rows + 7) / 8, st);
  if (M == 256 && INV) return launch_pass<ZPass<256, 8, INV>, ZParams, 0, true>(p, (p.rows + 7) / 8, st);
#define EVX_CASE(N, NL)                                                               \
  case N: return launch_pass<ZPass<N, NL, INV>, ZParams>(p, (p.rows + NL - 1) / NL, st);
  switch (M) {
    EVX_CASE(8, 32) EVX_CASE(16, 32) EVX_CASE(32, 32) EVX_CASE(64, 32) EVX_CASE(128, 16)
    EVX_CASE(256, 8) EVX_CASE(512, 4) EVX_CASE(1024, 2)
    default: return EVX_ERR_UNSUPPORTED;
  }
#undef EVX_CASE
}

bool native_fft_supported(int nx, int ny, int nz) {
  return is_pow2(nx) && is_pow2(ny) && is_pow2(nz) && nx >= 8 && nx <= 2048 && ny >= 8 &&
         ny <= 2048 && nz >= 16 && nz <= 2048;
}

static void fill_roots(std::vector<cf>& w, size_t off, int n, int count) {
  for (int m = 0; m < count; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)n;
    w[off + m] = cf{(float)std::cos(a), (float)std::sin(a)};
  }
}

int native_plan_init(evx_imex_plan* p) {
  const int M = p->nz / 2;
  p->spec_pitch = ((M + 1 + kPitchAlign - 1) / kPitchAlign) * kPitchAlign;
  p->spec_bytes = ((size_t)p->nx * p->ny * p->spec_pitch * sizeof(cf) + 255) & ~(size_t)255;
  // scratch behind the spectrum: per-plane completion counters of the chained z/y passes
  // (+ a debug area for per-block cycle counters at the very end, EVX_FFT_CHAIN_STATS)
  p->work_bytes = chain_supported(p->nx, p->ny, p->nz)
                      ? (((size_t)p->nx * sizeof(unsigned) + 255) & ~(size_t)255) + kChainStatsBytes : 0;
  // tables: W_nx | W_ny | W_M | W_nz[0..M] | W_2nx (x pass of a mirrored, non-periodic x axis)
  const size_t total = (size_t)p->nx + p->ny + M + (M + 1) + 2 * (size_t)p->nx;
  std::vector<cf> host(total);
  fill_roots(host, 0, p->nx, p->nx);
  fill_roots(host, p->nx, p->ny, p->ny);
  fill_roots(host, (size_t)p->nx + p->ny, M, M);
  fill_roots(host, (size_t)p->nx + p->ny + M, p->nz, M + 1);
  fill_roots(host, (size_t)p->nx + p->ny + M + (M + 1), 2 * p->nx, 2 * p->nx);
  cudaError_t e = cudaMalloc(&p->twiddles, total * sizeof(cf));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpy(p->twiddles, host.data(), total * sizeof(cf), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(p->twiddles); p->twiddles = nullptr; return (int)e; }
  return EVX_OK;
}

void native_plan_free(evx_imex_plan* p) {
  if (!p) return;
  if (p->twiddles) { cudaFree(p->twiddles); p->twiddles = nullptr; }
}

// ---- the five passes on the plan's buffers --------------------------------------------
struct NativeView {
  int nx, ny, nz, M, P;
  cf* spec;
  void* flags;             // work area behind the spectrum (null if the plan has none)
  const cf *twx, *twy, *twz, *twr, *twx2;
};

static NativeView view_of(const evx_imex_plan* p, void* workspace) {
  NativeView v;
  v.nx = p->nx; v.ny = p->ny; v.nz = p->nz; v.M = p->nz / 2; v.P = p->spec_pitch;
  v.spec = reinterpret_cast<cf*>((char*)workspace + p->real_bytes);
  v.flags = p->work_bytes ? (void*)((char*)workspace + p->real_bytes + p->spec_bytes) : nullptr;
  v.twx = (const cf*)p->twiddles;
  v.twy = v.twx + p->nx;
  v.twz = v.twy + p->ny;
  v.twr = v.twz + v.M;
  v.twx2 = v.twr + (v.M + 1);
  return v;
}

// TMA-tiled form of a strided pass (fft_line.cu): 512-point lines on a driver that can encode
// tensor maps.  EVX_FFT_TMA=0 keeps the cp.async passes (A/B measurements), EVX_FFT_TMA_KZ=16
// selects 128-byte tile rows.
// (read per call: a getenv costs nothing next to a launch, and tests switch paths in-process)
static bool use_line_pass(int L) {
  const char* e = getenv("EVX_FFT_TMA");
  return (!e || atoi(e) != 0) && L == 512 && line_pass_available();
}
// 1024-point lines: four-stage TMA-tiled program (StridedLine4); EVX_FFT_LINE4=0 keeps the
// cp.async passes
static bool use_line4(int L) {
  const char* e = getenv("EVX_FFT_TMA");
  const char* e4 = getenv("EVX_FFT_LINE4");
  return (!e || atoi(e) != 0) && (!e4 || atoi(e4) != 0) && L == 1024 && line_pass_available();
}
static int line_kz() {
  const char* e = getenv("EVX_FFT_TMA_KZ");
  return (e && atoi(e) == 16) ? 16 : 8;
}

static int line_l2_ahead() {
  const char* e = getenv("EVX_FFT_TMA_PF");
  const int v = e ? atoi(e) : 1;      // one tile ahead: y passes 0.215 -> 0.208 ms at 512^3
  return v < 0 ? 0 : (v > 8 ? 8 : v);
}

// tensor maps are keyed by the spectrum's address (the caller owns the workspace and may pass
// a different one per call; encoding is a host-side table fill of about a microsecond)
static int line_tmaps(evx_imex_plan* p, const NativeView& v) {
  if (p->tmap_spec == (void*)v.spec && p->tmap_kz == line_kz()) return EVX_OK;
  int rc = line_make_tmap(p->tmap_y, v.spec, v.nx, v.ny, v.P, v.M + 1, 0, line_kz());
  if (!rc) rc = line_make_tmap(p->tmap_x, v.spec, v.nx, v.ny, v.P, v.M + 1, 1, line_kz());
  if (rc) { p->tmap_spec = nullptr; return rc; }
  p->tmap_spec = (void*)v.spec;
  p->tmap_kz = line_kz();
  return EVX_OK;
}

static int z_pf_blocks() {
  const char* e = getenv("EVX_FFT_Z_PF");
  const int v = e ? atoi(e) : 0;
  return v < 0 ? 0 : v;
}

static ZParams z_params(const NativeView& v, const float* real_in, float* real_out) {
  ZParams zp;
  zp.real_in = real_in; zp.real_out = real_out; zp.spec = v.spec; zp.tw = v.twz; zp.twr = v.twr;
  zp.rows = (long long)v.nx * v.ny; zp.nz = v.nz; zp.P = v.P;
  zp.pf_blocks = z_pf_blocks();
  return zp;
}

static StridedParams y_params(const NativeView& v) {
  StridedParams yp;
  yp.in = v.spec; yp.out = v.spec; yp.tw = v.twy;
  yp.src = yp.dst = plain_io(v.P, (long long)v.ny * v.P, v.ny);
  yp.P = v.P; yp.ncols_valid = v.M + 1; yp.ncols_total = (long long)v.nx * v.P;
  yp.kother_offset = 0; yp.use_peers = 0; yp.max_ctas = 0; yp.dst_peer_base = 0;
  for (int i = 0; i < 8; ++i) yp.out_peers[i] = nullptr;
  yp.filt = FilterParams{};
  return yp;
}

static FilterParams filter_of(const NativeView& v, const double* h, double dt, double coef, int power) {
  const int n[3] = {v.nx, v.ny, v.nz};
  return make_filter(n, h, dt, coef, power, 1.0 / ((double)v.nx * v.ny * v.nz));
}

// which: 1 y forward, 2 x forward * weight * x inverse, 3 y inverse
static int strided_pass(evx_imex_plan* p, const NativeView& v, int which, const double* h, double dt,
                        double coef, int power, cudaStream_t st) {
  const bool along_x = which == 2;
  const int L = along_x ? v.nx : v.ny;
  const int mirror = along_x ? filter_mirror(power) : 0;
  if (mirror) {
    // non-periodic x: lines of 2 nx points, the upper half synthesised from the lower one
    if (2 * v.nx > 2048) return EVX_ERR_UNSUPPORTED;
    StridedParams sp = y_params(v);
    sp.tw = v.twx2;
    sp.src = sp.dst = plain_io((long long)v.ny * v.P, v.P, 2 * v.nx);
    sp.ncols_total = (long long)v.ny * v.P;
    sp.mirror = mirror;
    const int n[3] = {2 * v.nx, v.ny, v.nz};
    sp.filt = make_filter(n, h, dt, coef, power, 1.0 / (2.0 * v.nx * v.ny * v.nz));
    return launch_xmid(2 * v.nx, sp, st);
  }
  if (use_line4(L)) {
    // 4-D maps (kz, i_lo, row, i_hi), i = 256 i_hi + i_lo: lines along x have rows y, along y rows x
    alignas(64) unsigned char map[kTensorMapBytes];
    const long long line_stride = along_x ? (long long)v.ny * v.P : v.P;
    const long long row_stride = along_x ? v.P : (long long)v.ny * v.P;
    LineParams lp;
    lp.spec = v.spec; lp.tw = along_x ? v.twx : v.twy;
    lp.nx = v.nx; lp.ny = v.ny; lp.P = v.P; lp.ncols_valid = v.M + 1;
    lp.tiles_per_row = 0; lp.ntiles = 0; lp.along_x = along_x ? 1 : 0;
    lp.l2_ahead = line_l2_ahead();
    lp.nrows = along_x ? v.ny : v.nx; lp.box_rows = 256;
    if (int rc = line_make_tmap4(map, v.spec, v.M + 1, 256, line_stride, lp.nrows, row_stride, 4, 256 * line_stride))
      return rc;
    lp.filt = along_x ? filter_of(v, h, dt, coef, power) : FilterParams{};
    const int mode = which == 1 ? PASS_FWD
                                : (which == 3 ? PASS_INV
                                              : (lp.filt.kind == FILTER_ETD1 ? PASS_XMID_ETD1 : PASS_XMID));
    lp.out_div = 4;
    const void* outs[1] = {map};
    return line4_pass_launch(mode, lp, map, outs, 1, st);
  }
  if (use_line_pass(L)) {
    if (int rc = line_tmaps(p, v)) return rc;
    LineParams lp;
    lp.spec = v.spec; lp.tw = along_x ? v.twx : v.twy;
    lp.nx = v.nx; lp.ny = v.ny; lp.P = v.P; lp.ncols_valid = v.M + 1;
    lp.tiles_per_row = 0; lp.ntiles = 0; lp.along_x = along_x ? 1 : 0;
    lp.l2_ahead = line_l2_ahead();
    lp.filt = along_x ? filter_of(v, h, dt, coef, power) : FilterParams{};
    const int mode = which == 1 ? PASS_FWD
                                : (which == 3 ? PASS_INV
                                              : (lp.filt.kind == FILTER_ETD1 ? PASS_XMID_ETD1 : PASS_XMID));
    return line_pass_launch(mode, line_kz(), lp, along_x ? p->tmap_x : p->tmap_y, st);
  }
  StridedParams sp = y_params(v);
  if (which == 1) return launch_strided<PASS_FWD>(v.ny, sp, st);
  if (which == 3) return launch_strided<PASS_INV>(v.ny, sp, st);
  sp.tw = v.twx;
  sp.src = sp.dst = plain_io((long long)v.ny * v.P, v.P, v.nx);
  sp.ncols_total = (long long)v.ny * v.P;
  sp.filt = filter_of(v, h, dt, coef, power);
  return launch_xmid(v.nx, sp, st);
}

// Chained z/y passes (fft_chain.cu): one persistent kernel per direction whose second stage
// reads the first stage's output from L2.  EVX_FFT_CHAIN=0 keeps one kernel per pass.
static bool use_chain(const evx_imex_plan* p, const NativeView& v) {
  const char* e = getenv("EVX_FFT_CHAIN");
  return (!e || atoi(e) != 0) && v.flags && chain_supported(v.nx, v.ny, v.nz);
}
static int chain_tmap(evx_imex_plan* p, const NativeView& v) {
  if (p->tmap_chain_spec == (void*)v.spec) return EVX_OK;
  const int rc = line_make_tmap(p->tmap_chain, v.spec, v.nx, v.ny, v.P, v.M + 1, 0, 8);
  p->tmap_chain_spec = rc ? nullptr : (void*)v.spec;
  return rc;
}
static int chain_pass(evx_imex_plan* p, const NativeView& v, bool inverse, const float* real_in,
                      float* real_out, cudaStream_t st) {
  if (int rc = chain_tmap(p, v)) return rc;
  ChainArgs a;
  a.nx = v.nx; a.ny = v.ny; a.nz = v.nz; a.P = v.P;
  a.real_in = real_in; a.real_out = real_out; a.spec = v.spec;
  a.twz = v.twz; a.twr = v.twr; a.twy = v.twy; a.flags = v.flags;
  a.stats = (char*)v.flags + p->work_bytes - kChainStatsBytes;
  return chain_launch(inverse, a, p->tmap_chain, st);
}

int native_apply(evx_imex_plan* p, const float* u, const float* r, float* out, void* workspace,
                 const double* h, double dt, double coef, int power, cudaStream_t st) {
  const NativeView v = view_of(p, workspace);
  if (use_chain(p, v)) {
    if (int rc = chain_pass(p, v, false, r, nullptr, st)) return rc;
    if (int rc = strided_pass(p, v, 2, h, dt, coef, power, st)) return rc;
    return chain_pass(p, v, true, u, out, st);
  }
  if (int rc = launch_z<false>(v.M, z_params(v, r, nullptr), st)) return rc;
  for (int which = 1; which <= 3; ++which)
    if (int rc = strided_pass(p, v, which, h, dt, coef, power, st)) return rc;
  return launch_z<true>(v.M, z_params(v, u, out), st);
}

// one pass of the pipeline on the plan's scratch (measurement aid for bench.py's per-kernel
// roofline): which = 0 ZFwd, 1 Y fwd, 2 X fwd*filter*inv, 3 Y inv, 4 ZInv
int native_single_pass(evx_imex_plan* p, int which, const float* u, const float* r, float* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       cudaStream_t st) {
  const NativeView v = view_of(p, workspace);
  if (which == 0) return launch_z<false>(v.M, z_params(v, r, nullptr), st);
  if (which >= 1 && which <= 3) return strided_pass(p, v, which, h, dt, coef, power, st);
  if (which == 4) return launch_z<true>(v.M, z_params(v, u, out), st);
  // 5: chained z + y forward, 6: chained y + z inverse (EVX_ERR_UNSUPPORTED where the plan
  // runs one kernel per pass)
  if (which == 5 || which == 6) {
    if (!use_chain(p, v)) return EVX_ERR_UNSUPPORTED;
    return which == 5 ? chain_pass(p, v, false, r, nullptr, st) : chain_pass(p, v, true, u, out, st);
  }
  return EVX_ERR_ARG;
}

int native_ch_step(evx_imex_plan* p, const float* u, const float* hom, float* out,
                   void* workspace, const double* h, double dt, double eps, double D, double A,
                   cudaStream_t st) {
  float* rhs = (float*)workspace;
  const int per[3] = {BC_PERIODIC, BC_PERIODIC, BC_PERIODIC};
  if (int rc = ch_rhs_impl<float>(u, hom, rhs, p->nx, p->ny, p->nz, h, eps, D, per, nullptr, nullptr,
                                  nullptr, st))
    return rc;
  return native_apply(p, u, rhs, out, workspace, h, dt, 2.0 * eps * D * A, 2, st);
}

// ------------------------------------------------------------------------------------
// x-slab distributed variant: the y passes write / read the all-to-all block layout
// ------------------------------------------------------------------------------------
struct DistPlan : DistDims {
  int p2p_ctas = 0;   // grid cap of the peer-store launches (0: fill the GPU)
  void* twiddles = nullptr;
  void* chain_scratch = nullptr;   // plane counters (+ stats area) of the chained z/y kernels
};

static DistTables tables_of(const DistPlan* p) {
  DistTables t;
  t.twx = (const cf*)p->twiddles;
  t.twy = t.twx + p->nx;
  t.twz = t.twy + p->ny;
  t.twr = t.twz + p->M;
  return t;
}

int dist_plan_create(DistPlan** out, int nx, int ny, int nz, int world, int rank) {
  if (!out || world < 1 || rank < 0 || rank >= world) return EVX_ERR_ARG;
  if (!native_fft_supported(nx, ny, nz) || !is_pow2(world) || nx % world || ny % world)
    return EVX_ERR_UNSUPPORTED;
  DistPlan* p = new DistPlan();
  static_cast<DistDims&>(*p) = make_dist_dims(nx, ny, nz, world, rank, kPitchAlign);
  const size_t total = (size_t)nx + ny + p->M + (p->M + 1);
  std::vector<cf> host(total);
  fill_roots(host, 0, nx, nx);
  fill_roots(host, nx, ny, ny);
  fill_roots(host, (size_t)nx + ny, p->M, p->M);
  fill_roots(host, (size_t)nx + ny + p->M, nz, p->M + 1);
  cudaError_t e = cudaMalloc(&p->twiddles, total * sizeof(cf));
  if (e == cudaSuccess)
    e = cudaMemcpy(p->twiddles, host.data(), total * sizeof(cf), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && chain_supported(p->nxl, ny, nz) && world == 2)
    e = cudaMalloc(&p->chain_scratch, (((size_t)p->nxl * sizeof(unsigned) + 255) & ~(size_t)255) + kChainStatsBytes);
  if (e != cudaSuccess) { if (p->twiddles) cudaFree(p->twiddles); delete p; return (int)e; }
  *out = p;
  return EVX_OK;
}

// ---- chained z/y kernels in the distributed plan (2 ranks, 512-point y and z lines) -------------
// The y tiles of the chained kernels (fft_chain.cu) are [512 x 8]; with ny / W = 256 their two
// 256-row TMA boxes are exactly the two blocks of the all-to-all layout, so the forward pair
// stores its tiles through one tensor map per destination block (local buffers or the peer's
// buffer) and the inverse pair loads them from the two blocks of the receive buffer - the plane's
// spectrum still goes through L2 only.  EVX_ERR_UNSUPPORTED (nothing launched) otherwise.
static bool use_chain_dist(const DistPlan* p, int nxc) {
  const char* e = getenv("EVX_FFT_CHAIN");
  return (!e || atoi(e) != 0) && p->chain_scratch && p->world == 2 && p->nyl == 256 && chain_supported(nxc, p->ny, p->nz);
}
static int dist_chain(DistPlan* p, bool inverse, const float* real_in, float* real_out, cf* spec, cf* blocks,
                      void* const* peers, int x0, int nxc, cudaStream_t st) {
  if (!use_chain_dist(p, nxc)) return EVX_ERR_UNSUPPORTED;
  if (x0 < 0 || x0 + nxc > p->nxl) return EVX_ERR_ARG;
  const long long blk = (long long)p->nxl * p->nyl * p->P, xoff = (long long)x0 * p->nyl * p->P;
  alignas(64) unsigned char map_spec[kTensorMapBytes], map_blk[2][kTensorMapBytes];
  cf* spec_c = spec + (long long)x0 * p->ny * p->P;
  int rc = line_make_tmap(map_spec, spec_c, nxc, p->ny, p->P, p->M + 1, 0, 8);
  for (int w = 0; w < 2 && !rc; ++w) {
    cf* base = inverse ? blocks + w * blk + xoff : (cf*)peers[w] + (long long)p->rank * blk + xoff;
    rc = line_make_tmap(map_blk[w], base, nxc, p->nyl, p->P, p->M + 1, 0, 8);
  }
  if (rc) return rc;
  const DistTables t = tables_of(p);
  const long long roff = (long long)x0 * p->ny * p->nz;
  ChainArgs a;
  a.nx = nxc; a.ny = p->ny; a.nz = p->nz; a.P = p->P;
  a.real_in = real_in ? real_in + roff : nullptr;
  a.real_out = real_out ? real_out + roff : nullptr;
  a.spec = spec_c; a.twz = t.twz; a.twr = t.twr; a.twy = t.twy;
  a.flags = p->chain_scratch;
  a.stats = (char*)p->chain_scratch + (((size_t)p->nxl * sizeof(unsigned) + 255) & ~(size_t)255);
  return chain_launch(inverse, a, map_spec, st, map_blk[0], map_blk[1]);
}

// ---- 1024-point lines of the distributed plan: four-stage TMA-tiled passes (fft_line.cu) ------
// Destination tables follow the peer-store convention of dist_params.h: block `rank` of the
// buffer peers[w] receives what this rank produces for rank w - a peer's symmetric-memory buffer
// (the TMA stores of the pass ARE the slab<->pencil transpose over NVLink) or, with the table of
// local_block_table(), this GPU's own block buffers (copy-engine / NCCL transports).
// EVX_ERR_UNSUPPORTED (nothing launched): the caller uses the cp.async pass.
static int line4_p2p_ctas(const DistPlan* p, bool to_peers, bool mid) {
  if (!to_peers || p->p2p_ctas <= 0) return 0;
  // grid of an NVLink-bound launch in the pipelined plan: the y pass shares the GPU with rhs and
  // z pass of the next chunk; nothing runs next to the x pass
  const char* e = getenv(mid ? "EVX_LINE4_P2P_CTAS_MID" : "EVX_LINE4_P2P_CTAS");
  const int v = e ? atoi(e) : (mid ? 0 : 64);
  return v > 0 ? v : 0;
}

// y pass of the local x planes [x0, x0+nxc).  Forward: plain spectrum -> block `rank` of every
// peers[w]; inverse: block layout `blocks` -> plain spectrum.  One TMA box = the nyl rows of one
// block, hence nyl <= 256.
static int dist_y_line4(DistPlan* p, bool inverse, cf* spec, cf* blocks, void* const* peers, bool remote,
                        int x0, int nxc, cudaStream_t st) {
  if (!use_line4(p->ny) || p->nyl > 256 || p->nyl < 8 || p->world > 8) return EVX_ERR_UNSUPPORTED;
  if (x0 < 0 || nxc < 1 || x0 + nxc > p->nxl) return EVX_ERR_ARG;
  const int box = p->nyl;
  const long long blk = (long long)p->nxl * p->nyl * p->P;
  alignas(64) unsigned char map_plain[kTensorMapBytes], maps[8][kTensorMapBytes];
  int rc = line_make_tmap4(map_plain, spec, p->M + 1, box, p->P, p->nxl, (long long)p->ny * p->P, p->world,
                           (long long)box * p->P);
  const void* outs[8];
  LineParams lp;
  if (inverse) {
    if (!rc) rc = line_make_tmap4(maps[0], blocks, p->M + 1, box, p->P, p->nxl, (long long)p->nyl * p->P, p->world, blk);
    outs[0] = map_plain;
    lp.out_div = p->world;
  } else {
    for (int w = 0; w < p->world && !rc; ++w) {
      rc = line_make_tmap4(maps[w], (cf*)peers[w] + (long long)p->rank * blk, p->M + 1, box, p->P, p->nxl,
                           (long long)p->nyl * p->P, 1, blk);
      outs[w] = maps[w];
    }
    lp.out_div = 1;
  }
  if (rc) return rc;
  lp.spec = spec; lp.tw = tables_of(p).twy;
  lp.nx = p->nxl; lp.ny = p->ny; lp.P = p->P; lp.ncols_valid = p->M + 1;
  lp.tiles_per_row = 0; lp.ntiles = 0; lp.along_x = 0; lp.l2_ahead = line_l2_ahead();
  lp.row0 = x0; lp.nrows = nxc; lp.kother0 = 0; lp.box_rows = box;
  lp.max_ctas = line4_p2p_ctas(p, remote, false);
  lp.filt = FilterParams{};
  return inverse ? line4_pass_launch(PASS_INV, lp, maps[0], outs, 1, st)
                 : line4_pass_launch(PASS_FWD, lp, map_plain, outs, p->world, st);
}

// x pass of the local y-pencil rows [yl0, yl0+nylc): reads recv = [nx][nyl][P], the x range of rank
// w is stored into block `rank` of peers[w].
static int dist_middle_line4(DistPlan* p, cf* recv, void* const* peers, bool remote, const double* h, double dt,
                             double coef, int power, cudaStream_t st, int yl0, int nylc) {
  if (!use_line4(p->nx) || filter_mirror(power) || p->world > 8) return EVX_ERR_UNSUPPORTED;
  const int box = p->nxl < 256 ? p->nxl : 256;
  if (box < 8 || p->nx % box) return EVX_ERR_UNSUPPORTED;
  const long long line_stride = (long long)p->nyl * p->P, blk = (long long)p->nxl * p->nyl * p->P;
  alignas(64) unsigned char map_in[kTensorMapBytes], maps[8][kTensorMapBytes];
  int rc = line_make_tmap4(map_in, recv, p->M + 1, box, line_stride, p->nyl, p->P, p->nx / box, box * line_stride);
  const void* outs[8];
  for (int w = 0; w < p->world && !rc; ++w) {
    rc = line_make_tmap4(maps[w], (cf*)peers[w] + (long long)p->rank * blk, p->M + 1, box, line_stride, p->nyl,
                         p->P, p->nxl / box, box * line_stride);
    outs[w] = maps[w];
  }
  if (rc) return rc;
  LineParams lp;
  lp.spec = recv; lp.tw = tables_of(p).twx;
  lp.nx = p->nx; lp.ny = p->nyl; lp.P = p->P; lp.ncols_valid = p->M + 1;
  lp.tiles_per_row = 0; lp.ntiles = 0; lp.along_x = 1; lp.l2_ahead = line_l2_ahead();
  lp.row0 = yl0; lp.nrows = nylc; lp.kother0 = p->rank * p->nyl; lp.box_rows = box;
  lp.out_div = p->nxl / box;
  lp.max_ctas = line4_p2p_ctas(p, remote, true);
  const int n[3] = {p->nx, p->ny, p->nz};
  lp.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)p->nx * p->ny * p->nz));
  return line4_pass_launch(lp.filt.kind == FILTER_ETD1 ? PASS_XMID_ETD1 : PASS_XMID, lp, map_in, outs, p->world, st);
}

// Parameters: dist_params.h.  [x0, x0+nxc) selects a chunk of local x planes (r_local / spec /
// send still point at the start of the full local arrays), so that several chunks can be
// pipelined on streams: the NVLink-bound y pass of one chunk overlaps the HBM-bound z pass of
// the next.  parts: 1 = z pass, 2 = y pass, 3 = both.  `send`: local block buffer (peers null), or
// `peers`: destination table (see above; `remote` tells whether it points at other GPUs).
int dist_forward(DistPlan* p, const float* r_local, cf* spec, cf* send, void* const* peers,
                 int x0, int nxc, cudaStream_t st, int parts = 3, bool remote = true) {
  if (x0 < 0 || nxc < 1 || x0 + nxc > p->nxl) return EVX_ERR_ARG;
  const DistTables t = tables_of(p);
  void* table[8];
  if (!peers && p->world <= 8) local_block_table(*p, send, send, table);
  if (parts == 3 && p->world <= 8) {
    const int rc = dist_chain(p, false, r_local, nullptr, spec, nullptr, peers ? peers : table, x0, nxc, st);
    if (rc != EVX_ERR_UNSUPPORTED) return rc;
  }
  if (parts & 1) {
    const int rc = launch_z<false>(p->M, dist_zfwd_params(*p, t, r_local, spec, x0, nxc), st);
    if (rc) return rc;
  }
  if (parts & 2) {
    int rc = peers ? dist_y_line4(p, false, spec, nullptr, peers, remote, x0, nxc, st)
                   : (p->world <= 8 ? dist_y_line4(p, false, spec, nullptr, table, false, x0, nxc, st)
                                    : EVX_ERR_UNSUPPORTED);
    if (rc == EVX_ERR_UNSUPPORTED)
      rc = launch_strided<PASS_FWD>(p->ny, dist_yfwd_params(*p, t, spec, send, peers, p->p2p_ctas, x0, nxc), st);
    if (rc) return rc;
  }
  return EVX_OK;
}

// [yl0, yl0+nylc): chunk of the local y-pencil rows (all x, all kz of those rows)
int dist_middle(DistPlan* p, cf* recv, void* const* peers, const double* h, double dt, double coef,
                int power, cudaStream_t st, int yl0 = 0, int nylc = -1, bool remote = true) {
  if (nylc < 0) nylc = p->nyl - yl0;
  if (yl0 < 0 || nylc < 1 || yl0 + nylc > p->nyl) return EVX_ERR_ARG;
  void* table[8];
  if (!peers && p->world <= 8) local_block_table(*p, recv, recv, table);
  const int rc = peers ? dist_middle_line4(p, recv, peers, remote, h, dt, coef, power, st, yl0, nylc)
                       : (p->world <= 8 ? dist_middle_line4(p, recv, table, false, h, dt, coef, power, st, yl0, nylc)
                                        : EVX_ERR_UNSUPPORTED);
  if (rc != EVX_ERR_UNSUPPORTED) return rc;
  return launch_xmid(p->nx, dist_xmid_params(*p, tables_of(p), recv, peers, p->p2p_ctas, h, dt, coef,
                                             power, yl0, nylc), st);
}

int dist_backward(DistPlan* p, const cf* recv, cf* spec, const float* u_local, float* out_local,
                  cudaStream_t st) {
  const DistTables t = tables_of(p);
  int rc = dist_chain(p, true, u_local, out_local, spec, const_cast<cf*>(recv), nullptr, 0, p->nxl, st);
  if (rc != EVX_ERR_UNSUPPORTED) return rc;
  rc = dist_y_line4(p, true, spec, const_cast<cf*>(recv), nullptr, false, 0, p->nxl, st);
  if (rc == EVX_ERR_UNSUPPORTED)
    rc = launch_strided<PASS_INV>(p->ny, dist_yinv_params(*p, t, recv, spec, 0, p->nxl), st);
  if (rc) return rc;
  return launch_z<true>(p->M, dist_zinv_params(*p, t, spec, u_local, out_local, 0, p->nxl), st);
}

// ------------------------------------------------------------------------------------
// Block scatter over peer memory: region i (rows x row_bytes, pitched) is copied from src[i]
// to dst[i] - typically dst[i] lives on another GPU (mapped peer buffer), so the stores are
// NVLink writes of whole 128-byte lines.  A handful of CTAs per peer saturates the links;
// launched on a side stream it runs next to the kernels of the next chunk.  Used where the
// DMA engines' per-copy latency would dominate (many small regions to many peers).
// ------------------------------------------------------------------------------------
struct ScatterParams {
  const uint4* src[8];
  uint4* dst[8];
  long long row_vec, rows, src_pitch_vec, dst_pitch_vec;   // in 16-byte units
};

__global__ void __launch_bounds__(512) peer_scatter_kernel(const ScatterParams p) {
  const uint4* __restrict__ src = p.src[blockIdx.y];
  uint4* __restrict__ dst = p.dst[blockIdx.y];
  const long long total = p.rows * p.row_vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // four independent 16-byte transfers per thread and iteration
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = i + k * stride, r = j / p.row_vec, c = j - r * p.row_vec;
      v[k] = __ldcs(src + r * p.src_pitch_vec + c);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = i + k * stride, r = j / p.row_vec, c = j - r * p.row_vec;
      dst[r * p.dst_pitch_vec + c] = v[k];
    }
  }
  for (; i < total; i += stride) {
    const long long r = i / p.row_vec, c = i - r * p.row_vec;
    dst[r * p.dst_pitch_vec + c] = __ldcs(src + r * p.src_pitch_vec + c);
  }
}

int peer_scatter(const void* const* src, void* const* dst, int n, size_t row_bytes, size_t rows,
                 size_t src_pitch, size_t dst_pitch, int ctas_per_region, cudaStream_t st) {
  if (n < 1 || n > 8 || !row_bytes || !rows) return EVX_ERR_ARG;
  if (row_bytes % 16 || src_pitch % 16 || dst_pitch % 16) return EVX_ERR_UNSUPPORTED;
  ScatterParams p;
  for (int i = 0; i < 8; ++i) {
    p.src[i] = i < n ? (const uint4*)src[i] : nullptr;
    p.dst[i] = i < n ? (uint4*)dst[i] : nullptr;
    if (i < n && (!src[i] || !dst[i] || ((uintptr_t)src[i] | (uintptr_t)dst[i]) % 16)) return EVX_ERR_ARG;
  }
  p.row_vec = (long long)(row_bytes / 16); p.rows = (long long)rows;
  p.src_pitch_vec = (long long)(src_pitch / 16); p.dst_pitch_vec = (long long)(dst_pitch / 16);
  if (ctas_per_region < 1) ctas_per_region = 8;
  dim3 grid((unsigned)ctas_per_region, (unsigned)n);
  peer_scatter_kernel<<<grid, 512, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace evx

using namespace evx;

extern "C" {

int evx_dist_plan_create(evx_dist_plan** plan, int nx, int ny, int nz, int world, int rank) {
  return dist_plan_create((DistPlan**)plan, nx, ny, nz, world, rank);
}
int evx_dist_plan_destroy(evx_dist_plan* plan) {
  DistPlan* p = (DistPlan*)plan;
  if (p) {
    if (p->twiddles) cudaFree(p->twiddles);
    if (p->chain_scratch) cudaFree(p->chain_scratch);
    delete p;
  }
  return EVX_OK;
}
int evx_dist_plan_set_p2p_ctas(evx_dist_plan* plan, int ctas) {
  if (!plan || ctas < 0) return EVX_ERR_ARG;
  ((DistPlan*)plan)->p2p_ctas = ctas;
  return EVX_OK;
}
int evx_dist_plan_sizes(const evx_dist_plan* plan, size_t* spec_bytes, int* pitch) {
  const DistPlan* p = (const DistPlan*)plan;
  if (!p || !spec_bytes) return EVX_ERR_ARG;
  *spec_bytes = (size_t)p->nxl * p->ny * p->P * sizeof(cf);
  if (pitch) *pitch = p->P;
  return EVX_OK;
}
int evx_dist_forward_f32(evx_dist_plan* plan, const float* r_local, void* spec, void* send,
                         void* stream) {
  if (!plan || !r_local || !spec || !send || spec == send) return EVX_ERR_ARG;
  DistPlan* dp = (DistPlan*)plan;
  return dist_forward(dp, r_local, (cf*)spec, (cf*)send, nullptr, 0, dp->nxl, (cudaStream_t)stream);
}
int evx_dist_forward_p2p_f32(evx_dist_plan* plan, const float* r_local, void* spec,
                             void* const* peer_recv, void* stream) {
  if (!plan || !r_local || !spec || !peer_recv) return EVX_ERR_ARG;
  if (((DistPlan*)plan)->world > 8) return EVX_ERR_UNSUPPORTED;
  DistPlan* dp = (DistPlan*)plan;
  return dist_forward(dp, r_local, (cf*)spec, nullptr, peer_recv, 0, dp->nxl, (cudaStream_t)stream);
}
int evx_dist_forward_chunk_p2p_f32(evx_dist_plan* plan, const float* r_local, void* spec,
                                   void* const* peer_recv, int x0, int nxc, int parts, void* stream) {
  if (!plan || !r_local || !spec || !peer_recv || parts < 1 || parts > 3) return EVX_ERR_ARG;
  if (((DistPlan*)plan)->world > 8) return EVX_ERR_UNSUPPORTED;
  return dist_forward((DistPlan*)plan, r_local, (cf*)spec, nullptr, peer_recv, x0, nxc,
                      (cudaStream_t)stream, parts);
}
int evx_dist_forward_chunk_f32(evx_dist_plan* plan, const float* r_local, void* spec, void* send,
                               void* self_block, int x0, int nxc, void* stream) {
  if (!plan || !r_local || !spec || !send || spec == send) return EVX_ERR_ARG;
  DistPlan* dp = (DistPlan*)plan;
  if (!self_block)
    return dist_forward(dp, r_local, (cf*)spec, (cf*)send, nullptr, x0, nxc, (cudaStream_t)stream);
  if (dp->world > 8) return EVX_ERR_UNSUPPORTED;
  void* table[8];
  local_block_table(*dp, (cf*)send, (cf*)self_block, table);
  const int ctas = dp->p2p_ctas;
  dp->p2p_ctas = 0;                       // local stores: fill the GPU
  const int rc = dist_forward(dp, r_local, (cf*)spec, nullptr, table, x0, nxc, (cudaStream_t)stream, 3, false);
  dp->p2p_ctas = ctas;
  return rc;
}
int evx_dist_middle_chunk_f32(evx_dist_plan* plan, void* recv, void* self_block, int yl0, int nylc,
                              const double* h, double dt, double coef, int power, void* stream) {
  if (!plan || !recv || !h || !valid_filter_spec(power)) return EVX_ERR_ARG;
  DistPlan* dp = (DistPlan*)plan;
  if (!self_block)
    return dist_middle(dp, (cf*)recv, nullptr, h, dt, coef, power, (cudaStream_t)stream, yl0, nylc);
  if (dp->world > 8) return EVX_ERR_UNSUPPORTED;
  void* table[8];
  local_block_table(*dp, (cf*)recv, (cf*)self_block, table);
  const int ctas = dp->p2p_ctas;
  dp->p2p_ctas = 0;
  const int rc = dist_middle(dp, (cf*)recv, table, h, dt, coef, power, (cudaStream_t)stream, yl0, nylc, false);
  dp->p2p_ctas = ctas;
  return rc;
}
int evx_peer_scatter(const void* const* src, void* const* dst, int n, size_t row_bytes, size_t rows,
                     size_t src_pitch, size_t dst_pitch, int ctas_per_region, void* stream) {
  if (!src || !dst) return EVX_ERR_ARG;
  return peer_scatter(src, dst, n, row_bytes, rows, src_pitch, dst_pitch, ctas_per_region,
                      (cudaStream_t)stream);
}
int evx_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return EVX_ERR_ARG;
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
}
int evx_copy_batch_async(void* const* dst, size_t dpitch, const void* const* src, size_t spitch,
                         size_t width_bytes, size_t height, int n, void* stream) {
  if (!dst || !src || n < 1 || n > 8 || !width_bytes || !height || !stream) return EVX_ERR_ARG;
  size_t fail = 0;
  if (height == 1) {
    void* d[8]; void* s[8]; size_t sz[8];
    for (int i = 0; i < n; ++i) { d[i] = dst[i]; s[i] = const_cast<void*>(src[i]); sz[i] = width_bytes; }
    cudaMemcpyAttributes attr = {};
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
    size_t idx = 0;
    return (int)cudaMemcpyBatchAsync(d, s, sz, (size_t)n, &attr, &idx, 1, &fail, (cudaStream_t)stream);
  }
  cudaMemcpy3DBatchOp ops[8] = {};
  for (int i = 0; i < n; ++i) {
    ops[i].src.type = cudaMemcpyOperandTypePointer;
    ops[i].src.op.ptr.ptr = const_cast<void*>(src[i]);
    ops[i].src.op.ptr.rowLength = spitch;
    ops[i].src.op.ptr.layerHeight = height;
    ops[i].dst.type = cudaMemcpyOperandTypePointer;
    ops[i].dst.op.ptr.ptr = dst[i];
    ops[i].dst.op.ptr.rowLength = dpitch;
    ops[i].dst.op.ptr.layerHeight = height;
    ops[i].extent = make_cudaExtent(width_bytes, height, 1);
    ops[i].srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    ops[i].flags = cudaMemcpyFlagPreferOverlapWithCompute;
  }
  return (int)cudaMemcpy3DBatchAsync((size_t)n, ops, &fail, 0, (cudaStream_t)stream);
}
int evx_copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes,
                     size_t height, void* stream) {
  if (!dst || !src) return EVX_ERR_ARG;
  return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDefault,
                                (cudaStream_t)stream);
}
int evx_dist_middle_p2p_f32(evx_dist_plan* plan, void* recv, void* const* peer_out, const double* h,
                            double dt, double coef, int power, void* stream) {
  if (!plan || !recv || !peer_out || !h || !valid_filter_spec(power)) return EVX_ERR_ARG;
  if (((DistPlan*)plan)->world > 8) return EVX_ERR_UNSUPPORTED;
  return dist_middle((DistPlan*)plan, (cf*)recv, peer_out, h, dt, coef, power, (cudaStream_t)stream);
}
int evx_dist_middle_chunk_p2p_f32(evx_dist_plan* plan, void* recv, void* const* peer_out, int yl0, int nylc,
                                  const double* h, double dt, double coef, int power, void* stream) {
  if (!plan || !recv || !peer_out || !h || !valid_filter_spec(power)) return EVX_ERR_ARG;
  if (((DistPlan*)plan)->world > 8) return EVX_ERR_UNSUPPORTED;
  return dist_middle((DistPlan*)plan, (cf*)recv, peer_out, h, dt, coef, power, (cudaStream_t)stream, yl0, nylc);
}
int evx_dist_middle_f32(evx_dist_plan* plan, void* recv, const double* h, double dt, double coef,
                        int power, void* stream) {
  if (!plan || !recv || !h || !valid_filter_spec(power)) return EVX_ERR_ARG;
  return dist_middle((DistPlan*)plan, (cf*)recv, nullptr, h, dt, coef, power, (cudaStream_t)stream);
}
int evx_dist_backward_f32(evx_dist_plan* plan, const void* recv, void* spec, const float* u_local,
                          float* out_local, void* stream) {
  if (!plan || !recv || !spec || !out_local || recv == spec) return EVX_ERR_ARG;
  return dist_backward((DistPlan*)plan, (const cf*)recv, (cf*)spec, u_local, out_local,
                       (cudaStream_t)stream);
}

}  // extern "C"
