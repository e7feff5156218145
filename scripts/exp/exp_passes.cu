// Experimental variants of the native FFT passes (measurement aid, NOT part of libevx_b200.so).
// Built by scripts/exp/build.sh into scripts/exp/libevx_exp.so; scripts/exp/run_passes.py times
// the variants against each other with CUDA events and checks that they are bit-identical.
//
// Variants (all for 512-point lines, nz = 512):
//   z pass:   sync = 0 __syncthreads between phases (shipped form), 1 __syncwarp (a line is
//             owned by exactly one warp when M/8 == 32)
//   strided:  kz = 8 | 16 columns per tile, twreg = roots of unity kept in registers over the
//             persistent tile loop instead of being re-fetched per stage, l2 = cp.async L2
//             prefetch size hint (0, 128, 256 bytes)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../evoxels_b200/csrc/fft_pass_core.h"

using namespace evx;

// ------------------------------------------------------------------------------------
template <class Prog, int MINB, bool WARPSYNC>
__global__ void __launch_bounds__(Prog::NTHREADS, MINB) exp_z_kernel(const ZParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* smem = reinterpret_cast<cf*>(smem_raw);
  typename Prog::Regs r;
  Prog::init(r, p, threadIdx.x, (long long)blockIdx.x);
#pragma unroll
  for (int k = 0; k < Prog::NPHASES; ++k) {
    if (k) { if (WARPSYNC) __syncwarp(); else __syncthreads(); }
    Prog::phase(k, r, smem, p);
  }
}

template <int L2>
__device__ __forceinline__ void exp_copy16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (L2 == 128)
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
  else if (L2 == 256)
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <class Pipe, int L2>
__device__ __forceinline__ void exp_prefetch(int tid, const StridedParams& p, long long grp, int kz0, cf* dst) {
  constexpr int CF16 = 16 / (int)sizeof(cf);
  const int row0 = tid / Pipe::CHUNKS_PER_ROW, part = tid - row0 * Pipe::CHUNKS_PER_ROW;
  const cf* src = p.in + strided_offset(p.src, grp, kz0, row0) + part * CF16;
  cf* d = dst + (size_t)row0 * Pipe::Base::NTHREADS / Pipe::Base::T + part * CF16;
#pragma unroll
  for (int it = 0; it < Pipe::PF_ITERS; ++it)
    exp_copy16<L2>(d + (size_t)it * Pipe::PF_ROWS * (Pipe::Base::NTHREADS / Pipe::Base::T), src + p.src_pf_step[it]);
}

// phase k of a pipelined strided pass with the roots of the twiddled stages in registers
template <class Pipe, int MODE>
__device__ __forceinline__ void exp_phase_twreg(int k, typename Pipe::Regs& r, cf* a, cf* c,
                                                const StridedParams& p, const cf (*w)[3]) {
  using Base = typename Pipe::Base;
  constexpr int S = Pipe::S;
  cf* wr = (k & 1) ? a : c;
  const cf* rd = (k & 1) ? c : a;
  if (k > 0) Base::read_natural(r, rd);
  auto setw = [&](int stage) {
    if (stage >= 1 && stage < S) { r.w[0] = w[stage - 1][0]; r.w[1] = w[stage - 1][1]; r.w[2] = w[stage - 1][2]; }
  };
  if (MODE == PASS_FWD || MODE == PASS_INV) {
    setw(k);
    if (MODE == PASS_FWD) Base::template compute<-1>(r, p, k); else Base::template compute<+1>(r, p, k);
    if (k == S - 1) Base::store_global(r, p);
    else if (MODE == PASS_FWD) Base::template write_stage<-1>(r, wr, k);
    else Base::template write_stage<+1>(r, wr, k);
  } else {
    if (k < S - 1) {
      setw(k);
      Base::template compute<-1>(r, p, k);
      Base::template write_stage<-1>(r, wr, k);
    } else if (k == S - 1) {
      setw(S - 1);
      Base::template compute<-1>(r, p, S - 1);
      Base::apply_filter(r, p);
      Base::template compute<+1>(r, p, 0);
      Base::template write_stage<+1>(r, wr, 0);
    } else {
      const int s = k - (S - 1);
      setw(s);
      Base::template compute<+1>(r, p, s);
      if (s == S - 1) Base::store_global(r, p); else Base::template write_stage<+1>(r, wr, s);
    }
  }
}

template <class Pipe, int MODE, bool TWREG, int L2>
__global__ void __launch_bounds__(Pipe::NTHREADS, Pipe::NTHREADS <= 512 ? 2 : 1)
    exp_pipe_kernel(const StridedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* t0 = reinterpret_cast<cf*>(smem_raw);
  cf* t1 = t0 + Pipe::BUF;
  cf* c = t1 + Pipe::BUF;
  const long long ntiles = Pipe::num_tiles(p);
  long long tile = blockIdx.x;
  typename Pipe::Cursor q;
  Pipe::cursor_init(q, p, threadIdx.x, tile, gridDim.x);
  if (tile < ntiles) exp_prefetch<Pipe, L2>(threadIdx.x, p, q.bgrp, q.bkz, t0);
  async_copy_commit();
  typename Pipe::Regs r;
  cf w[Pipe::S > 1 ? Pipe::S - 1 : 1][3];
  if (TWREG) {
#pragma unroll
    for (int s = 1; s < Pipe::S; ++s) {
      cf tmp[3];
      stage_twiddles<Pipe::Base::T * 8>(s, threadIdx.x / (Pipe::NTHREADS / Pipe::T), p.tw, tmp);
      w[s - 1][0] = tmp[0]; w[s - 1][1] = tmp[1]; w[s - 1][2] = tmp[2];
    }
  }
  for (int par = 0; tile < ntiles; tile += gridDim.x, par ^= 1) {
    cf* a = par ? t1 : t0;
    cf* b = par ? t0 : t1;
    async_copy_commit_and_wait();
    __syncthreads();
    Pipe::Base::init_at(r, p, threadIdx.x, Pipe::cursor_column(q, p), q.grp, q.kz);
    Pipe::read_tile(r, a);
    Pipe::cursor_step_own(q, p);
    Pipe::cursor_step_base(q, p);
    if (tile + gridDim.x < ntiles) exp_prefetch<Pipe, L2>(threadIdx.x, p, q.bgrp, q.bkz, b);
    async_copy_commit();
#pragma unroll
    for (int k = 0; k < Pipe::NPHASES; ++k) {
      if (k) __syncthreads();
      if (TWREG) exp_phase_twreg<Pipe, MODE>(k, r, a, c, p, w);
      else Pipe::phase(k, r, a, c, p);
    }
  }
}

static int g_sms = 0;
static int sms() {
  if (!g_sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, d); }
  return g_sms;
}

template <class Pipe, int MODE, bool TWREG, int L2>
static int run_pipe(StridedParams p, cudaStream_t st) {
  finalize_strided(p, Pipe::Base::T * 8);
  auto kern = exp_pipe_kernel<Pipe, MODE, TWREG, L2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pipe::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const long long ntiles = Pipe::num_tiles(p);
  long long resident = (long long)sms() * (Pipe::NTHREADS <= 512 ? 2 : 1);
  const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
  kern<<<grid, Pipe::NTHREADS, Pipe::SMEM_BYTES, st>>>(p);
  return (int)cudaGetLastError();
}

template <int MODE, int KZ>
static int run_pipe_sel(const StridedParams& p, int twreg, int l2, cudaStream_t st) {
  using Pipe = StridedPipe<512, KZ, MODE>;
  if (twreg) {
    if (l2 == 0) return run_pipe<Pipe, MODE, true, 0>(p, st);
    if (l2 == 128) return run_pipe<Pipe, MODE, true, 128>(p, st);
    if (l2 == 256) return run_pipe<Pipe, MODE, true, 256>(p, st);
  } else {
    if (l2 == 0) return run_pipe<Pipe, MODE, false, 0>(p, st);
    if (l2 == 128) return run_pipe<Pipe, MODE, false, 128>(p, st);
    if (l2 == 256) return run_pipe<Pipe, MODE, false, 256>(p, st);
  }
  return -1;
}

extern "C" {

// z pass over `rows` lines of nz = 512 reals. inverse: out = u + irfft(spec)
int exp_z(int inverse, int warpsync, int minb, const float* real_in, float* real_out, void* spec,
          const void* twz, const void* twr, long long rows, int P, void* stream) {
  ZParams p;
  p.real_in = real_in; p.real_out = real_out; p.spec = (cf*)spec; p.tw = (const cf*)twz;
  p.twr = (const cf*)twr; p.rows = rows; p.nz = 512; p.P = P;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
#define EXP_Z(INV, MB, WS)                                                                     \
  {                                                                                            \
    using Prog = ZPass<256, 8, INV>;                                                           \
    auto kern = exp_z_kernel<Prog, MB, WS>;                                                    \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Prog::SMEM_BYTES); \
    kern<<<blocks, Prog::NTHREADS, Prog::SMEM_BYTES, st>>>(p);                                 \
    return (int)cudaGetLastError();                                                            \
  }
  if (!inverse && minb == 5 && !warpsync) EXP_Z(false, 5, false)
  if (!inverse && minb == 5 && warpsync) EXP_Z(false, 5, true)
  if (!inverse && minb == 4 && warpsync) EXP_Z(false, 4, true)
  if (!inverse && minb == 6 && warpsync) EXP_Z(false, 6, true)
  if (inverse && minb == 4 && !warpsync) EXP_Z(true, 4, false)
  if (inverse && minb == 4 && warpsync) EXP_Z(true, 4, true)
  if (inverse && minb == 5 && warpsync) EXP_Z(true, 5, true)
#undef EXP_Z
  return -1;
}

// strided pass over a [nx][ny][P] spectrum (in place). mode: 0 fwd, 1 inv (axis y), 2 x fwd*filter*inv
int exp_strided(int mode, int kz, int twreg, int l2, void* spec, const void* tw, int nx, int ny,
                int P, int ncols_valid, double dt, double coef, void* stream) {
  StridedParams p;
  p.in = (const cf*)spec; p.out = (cf*)spec; p.tw = (const cf*)tw;
  p.P = P; p.ncols_valid = ncols_valid; p.kother_offset = 0; p.use_peers = 0; p.max_ctas = 0;
  p.filt = FilterParams{};
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 2) {
    if (nx != 512) return -2;
    p.src = p.dst = plain_io((long long)ny * P, P, nx);
    p.ncols_total = (long long)ny * P;
    const int n[3] = {nx, ny, 2 * (ncols_valid - 1)};
    const double h[3] = {1.0, 1.0, 1.0};
    p.filt = make_filter(n, h, dt, coef, 2, 1.0 / ((double)nx * ny * n[2]));
    if (kz == 16) return run_pipe_sel<PASS_XMID, 16>(p, twreg, l2, st);
    if (kz == 8) return run_pipe_sel<PASS_XMID, 8>(p, twreg, l2, st);
    return -1;
  }
  if (ny != 512) return -2;
  p.src = p.dst = plain_io(P, (long long)ny * P, ny);
  p.ncols_total = (long long)nx * P;
  if (kz != 8) return -1;
  if (mode == 0) return run_pipe_sel<PASS_FWD, 8>(p, twreg, l2, st);
  if (mode == 1) return run_pipe_sel<PASS_INV, 8>(p, twreg, l2, st);
  return -1;
}

}  // extern "C"
