// CPU replay of the CUDA thread-block programs (TEST INFRASTRUCTURE ONLY).
//
// Each kernel in evoxels_b200/csrc/*_core.h is a set of barrier-free phase functions;
// this file runs them block by block, phase by phase, thread by thread, exactly in the
// order the __global__ wrappers do with __syncthreads() between phases.  It validates
// index / halo / boundary logic against the oracle on a box without a GPU.  It is NOT a
// CPU fallback: the shipped library never links or calls it.
#include <vector>
#include <cstring>
#include "../../evoxels_b200/csrc/evx_params.h"

using namespace evx;

template <typename T, int V, int TY, int G, bool HOM, bool GHOSTS>
static void run_ch_(const ChParams<T>& p) {
  using Prog = ChRhsProgram<T, V, TY, G, HOM, GHOSTS>;
  const int tiles = ((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  typename Prog::Smem* s = new typename Prog::Smem;
  for (int chunk = 0; chunk < chunks; ++chunk)
    for (int tile = 0; tile < tiles; ++tile) {
      std::memset(s, 0, sizeof(*s));
      for (int t = 0; t < Prog::NTHREADS; ++t) Prog::init(regs[t], *s, p, t, tile, chunk);
#define EMU_CH_PLANE(PAR, ROT, OFF)                                                              \
  if (pl + (OFF) <= regs[0].xb) {                                                                 \
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::template phase_a<PAR, ROT>(regs[t], *s, p, pl + (OFF)); \
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::template phase_b<PAR, ROT>(regs[t], *s, p, pl + (OFF)); \
  }
      for (int pl = regs[0].xa - 1; pl <= regs[0].xb; pl += 6) {
        EMU_CH_PLANE(0, 0, 0) EMU_CH_PLANE(1, 1, 1) EMU_CH_PLANE(0, 2, 2)
        EMU_CH_PLANE(1, 0, 3) EMU_CH_PLANE(0, 1, 4) EMU_CH_PLANE(1, 2, 5)
      }
#undef EMU_CH_PLANE
    }
  delete s;
}

template <typename T, int V, int TY, int G>
static void run_ch(const ChParams<T>& p) {
  const bool ghosts = p.bc_kind[0] != BC_PERIODIC || p.bc_kind[1] != BC_PERIODIC ||
                      p.bc_kind[2] != BC_PERIODIC;
  if (p.hom) run_ch_<T, V, TY, G, true, true>(p);
  else if (ghosts) run_ch_<T, V, TY, G, false, true>(p);
  else run_ch_<T, V, TY, G, false, false>(p);
}

template <typename T>
static int emu_ch(const T* c, const T* hom, T* out, int nx, int ny, int nz, const double* h,
                  double eps, double D, const int* bck, const double* bcv, const T* hlo,
                  const T* hhi, int xchunk, int vec) {
  ChParams<T> p = make_ch_params<T>(c, hom, out, nx, ny, nz, h, eps, D, bck, bcv, hlo, hhi, xchunk);
  constexpr int VW = 16 / sizeof(T);
  if (vec) {
    if (nz % VW) return -1;
    if (vec == 16) run_ch<T, VW, 16, 16>(p);
    else if (vec == 30) run_ch<T, VW, 30, 16>(p);
    else run_ch<T, VW, 14, 16>(p);
  } else {
    run_ch<T, 1, 8, 32>(p);
  }
  return 0;
}

template <typename T, int V, int TY, int G>
static void run_ac(const AcParams<T>& p) {
  using Prog = AcProgram<T, V, TY, G>;
  const int tiles = ((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  for (int chunk = 0; chunk < chunks; ++chunk)
    for (int tile = 0; tile < tiles; ++tile)
      for (int t = 0; t < Prog::NTHREADS; ++t) Prog::run(p, t, tile, chunk);
}

template <typename T, int V, int TY, int G>
static void run_ac_tile(const AcParams<T>& p) {
  using Prog = AcTileProgram<T, V, TY, G>;
  const int tiles = ((p.ny + TY - 1) / TY) * ((p.nz + Prog::TZ - 1) / Prog::TZ);
  const int chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  typename Prog::Smem* s = new typename Prog::Smem;
  for (int chunk = 0; chunk < chunks; ++chunk)
    for (int tile = 0; tile < tiles; ++tile) {
      std::memset(s, 0, sizeof(*s));
      for (int t = 0; t < Prog::NTHREADS; ++t) Prog::init(regs[t], *s, p, t, tile, chunk);
#define EMU_AC_PLANE(ROT, OFF)                                                                  \
  if (x + (OFF) < regs[0].xb) {                                                                  \
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::template phase_a<ROT>(regs[t], *s, p, x + (OFF)); \
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::template phase_b<ROT>(regs[t], *s, p, x + (OFF)); \
  }
      for (int x = regs[0].xa; x < regs[0].xb; x += 4) {
        EMU_AC_PLANE(0, 0) EMU_AC_PLANE(1, 1) EMU_AC_PLANE(2, 2) EMU_AC_PLANE(3, 3)
      }
#undef EMU_AC_PLANE
    }
  delete s;
}

template <typename T>
static int emu_ac(const T* phi, const T* pot, T* k_out, const T* base, T* y_out, double alpha,
                  const T* acc_in, T* acc_out, double beta, int nx, int ny, int nz,
                  const double* h, double eps, double gab, double M, double force, double curv,
                  const int* bck, const double* bcv, const T* hlo, const T* hhi, int xchunk,
                  int vec) {
  AcParams<T> p = make_ac_params<T>(phi, pot, k_out, base, y_out, alpha, acc_in, acc_out, beta,
                                    nx, ny, nz, h, eps, gab, M, force, curv, bck, bcv, hlo, hhi,
                                    xchunk);
  constexpr int VW = 16 / sizeof(T);
  if (vec) {
    if (nz % VW) return -1;
    if (vec == 2) run_ac_tile<T, VW, 14, 16>(p);     // shared-memory tile variant
    else run_ac<T, VW, 8, 32>(p);
  } else {
    run_ac<T, 1, 8, 32>(p);
  }
  return 0;
}

template <typename T>
static void emu_pad(const T* in, T* out, int nx, int ny, int nz, const int* bck, const double* bcv) {
  PadParams<T> p = make_pad_params<T>(in, out, nx, ny, nz, bck, bcv);
  for (int i = 0; i < nx + 2; ++i)
    for (int j = 0; j < ny + 2; ++j)
      for (int k = 0; k < nz + 2; ++k)
        out[((long long)i * (ny + 2) + j) * (nz + 2) + k] = padded_value(p, i, j, k);
}

template <typename T>
static void emu_pst(const T* g, T* out, int nx, int ny, int nz, const double* h, int op) {
  PaddedStencilParams<T> p = make_padded_stencil_params<T>(g, out, nx, ny, nz, h, op);
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
        out[((long long)x * ny + y) * nz + z] = padded_stencil_value(p, x, y, z);
}

extern "C" {
#define EMU_AC(SUF, T)                                                                          \
  int emu_ac_stage_##SUF(const T* phi, const T* pot, T* k_out, const T* base, T* y_out,         \
                         double alpha, const T* acc_in, T* acc_out, double beta, int nx, int ny, \
                         int nz, const double* h, double eps, double gab, double M,             \
                         double force, double curv, const int* bck, const double* bcv,          \
                         const T* hlo, const T* hhi, int xchunk, int vec) {                     \
    return emu_ac<T>(phi, pot, k_out, base, y_out, alpha, acc_in, acc_out, beta, nx, ny, nz, h, \
                     eps, gab, M, force, curv, bck, bcv, hlo, hhi, xchunk, vec);                \
  }                                                                                             \
  void emu_pad_ghost_##SUF(const T* in, T* out, int nx, int ny, int nz, const int* bck,         \
                           const double* bcv) {                                                 \
    emu_pad<T>(in, out, nx, ny, nz, bck, bcv);                                                  \
  }                                                                                             \
  void emu_padded_stencil_##SUF(const T* g, T* out, int nx, int ny, int nz, const double* h,    \
                                int op) {                                                       \
    emu_pst<T>(g, out, nx, ny, nz, h, op);                                                      \
  }
EMU_AC(f32, float)
EMU_AC(f64, double)

int emu_ch_rhs_f32(const float* c, const float* hom, float* out, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bck, const double* bcv,
                   const float* hlo, const float* hhi, int xchunk, int vec) {
  return emu_ch<float>(c, hom, out, nx, ny, nz, h, eps, D, bck, bcv, hlo, hhi, xchunk, vec);
}
int emu_ch_rhs_f64(const double* c, const double* hom, double* out, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bck, const double* bcv,
                   const double* hlo, const double* hhi, int xchunk, int vec) {
  return emu_ch<double>(c, hom, out, nx, ny, nz, h, eps, D, bck, bcv, hlo, hhi, xchunk, vec);
}
}

// =====================================================================================
// native FFT passes
// =====================================================================================
#include <cmath>
#include "../../evoxels_b200/csrc/fft_pass_core.h"

static std::vector<cf> make_roots(int n, int count) {   // exp(-2 pi i m / n), m < count
  std::vector<cf> w(count);
  for (int m = 0; m < count; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)n;
    w[m] = cf{(float)std::cos(a), (float)std::sin(a)};
  }
  return w;
}

template <int L, int KZ, int MODE>
static void run_strided(const StridedParams& p) {
  using Prog = StridedPass<L, KZ, MODE>;
  const long long blocks = (p.ncols_total + KZ - 1) / KZ;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  std::vector<cf> smem(Prog::SMEM_BYTES / sizeof(cf) + 1);
  for (long long b = 0; b < blocks; ++b) {
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::init(regs[t], p, t, b);
    for (int k = 0; k < Prog::NPHASES; ++k)
      for (int t = 0; t < Prog::NTHREADS; ++t) Prog::phase(k, regs[t], smem.data(), p);
  }
}

// persistent pipelined variant, replayed with `nblocks` resident blocks
template <int L, int KZ, int MODE>
static void run_pipe(const StridedParams& p, int nblocks) {
  using Pipe = StridedPipe<L, KZ, MODE>;
  const long long ntiles = Pipe::num_tiles(p);
  std::vector<typename Pipe::Regs> regs(Pipe::NTHREADS);
  std::vector<typename Pipe::Cursor> cur(Pipe::NTHREADS);
  std::vector<cf> smem(3 * (size_t)Pipe::BUF);
  for (int blk = 0; blk < nblocks; ++blk) {
    cf* t0 = smem.data(); cf* t1 = t0 + Pipe::BUF; cf* c = t1 + Pipe::BUF;
    long long tile = blk;
    for (int t = 0; t < Pipe::NTHREADS; ++t) {
      Pipe::cursor_init(cur[t], p, t, tile, nblocks);
      if (tile < ntiles) Pipe::prefetch_at(t, p, cur[t].bgrp, cur[t].bkz, t0);
    }
    for (int par = 0; tile < ntiles; tile += nblocks, par ^= 1) {
      cf* a = par ? t1 : t0; cf* b = par ? t0 : t1;
      for (int t = 0; t < Pipe::NTHREADS; ++t) {
        Pipe::Base::init_at(regs[t], p, t, Pipe::cursor_column(cur[t], p), cur[t].grp, cur[t].kz);
        Pipe::read_tile(regs[t], a);
        Pipe::cursor_step_own(cur[t], p);
        Pipe::cursor_step_base(cur[t], p);
      }
      const long long next = tile + nblocks;
      if (next < ntiles) for (int t = 0; t < Pipe::NTHREADS; ++t) Pipe::prefetch_at(t, p, cur[t].bgrp, cur[t].bkz, b);
      for (int k = 0; k < Pipe::NPHASES; ++k)
        for (int t = 0; t < Pipe::NTHREADS; ++t) Pipe::phase(k, regs[t], a, c, p);
    }
  }
}

// TMA-tiled strided pass (fft_line_core.h) replayed with `nblocks` resident blocks: tensor
// copies are restated by host_tile_load / host_tile_store (swizzled tile image, zero fill and
// no write-back outside the valid columns); groups are replayed one after the other through
// all phases of a tile, i.e. with NO ordering between different groups - exactly what the
// kernel guarantees - and two tile buffers alternate like on the device.
#include "../../evoxels_b200/csrc/fft_line_core.h"
template <int KZ, int MODE>
static void run_line(LineParams p, int nblocks) {
  using Prog = StridedLine<512, KZ, MODE>;
  p.tiles_per_row = (p.ncols_valid + KZ - 1) / KZ;
  p.ntiles = (long long)(p.along_x ? p.ny : p.nx) * p.tiles_per_row;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  std::vector<unsigned char> tiles(2 * (size_t)Prog::TILE_BYTES, 0xff);
  std::vector<cf> xall((size_t)Prog::NG * Prog::XG, cf{std::nanf(""), std::nanf("")});
  for (int t = 0; t < Prog::NTHREADS; ++t) Prog::init(regs[t], t);
  for (int blk = 0; blk < nblocks; ++blk) {
    int it = 0;
    for (long long tile = blk; tile < p.ntiles; tile += nblocks, ++it) {
      unsigned char* tb = tiles.data() + (it & 1) * Prog::TILE_BYTES;
      const int row = (int)(tile / p.tiles_per_row), kz0 = (int)(tile % p.tiles_per_row) * KZ;
      Prog::host_tile_load(p, row, kz0, tb);
      for (int g = Prog::NG - 1; g >= 0; --g) {        // any group order must do
        cf* xg = xall.data() + (size_t)g * Prog::XG;
        for (int k = 0; k < Prog::NPHASES; ++k)
          for (int t = g * Prog::GT; t < (g + 1) * Prog::GT; ++t) {
            if (k == 0) Prog::set_tile(regs[t], row, kz0);
            Prog::phase(k, regs[t], tb, xg, p);
          }
      }
      Prog::host_tile_store(p, row, kz0, tb);
    }
  }
}

// 1024-point form (StridedLine4): two line pairs per group and tile, two exchange buffers per
// group; groups (and, inside a group, the pairs) are replayed one after the other.
static int g_emu_line4_form = 16;   // 16: StridedLine16 (sixteen points per thread), 8: StridedLine4
template <class Prog>
static void run_line4_prog(LineParams p, int nblocks) {
  p.tiles_per_row = (p.ncols_valid + 7) / 8;
  if (p.nrows <= 0) p.nrows = (p.along_x ? p.ny : p.nx) - p.row0;
  p.ntiles = (long long)p.nrows * p.tiles_per_row;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  std::vector<typename Prog::Roots> roots(Prog::NTHREADS);
  std::vector<unsigned char> tiles(2 * (size_t)Prog::TILE_BYTES, 0xff);
  std::vector<cf> xall((size_t)Prog::NG * Prog::XSTRIDE, cf{std::nanf(""), std::nanf("")});
  for (int t = 0; t < Prog::NTHREADS; ++t) {
    Prog::init(regs[t], t);
    Prog::load_roots(roots[t], regs[t].t, p.tw);
  }
  for (int blk = 0; blk < nblocks; ++blk) {
    int it = 0;
    for (long long tile = blk; tile < p.ntiles; tile += nblocks, ++it) {
      unsigned char* tb = tiles.data() + (it & 1) * Prog::TILE_BYTES;
      const int row = p.row0 + (int)(tile / p.tiles_per_row), kz0 = (int)(tile % p.tiles_per_row) * 8;
      Prog::host_tile_load(p, row, kz0, tb);
      for (int g = Prog::NG - 1; g >= 0; --g) {
        cf* xg = xall.data() + (size_t)g * Prog::XSTRIDE;
        for (int pass = 0; pass < Prog::NPASS; ++pass)
          for (int k = 0; k < Prog::NPHASES; ++k)
            for (int t = g * Prog::GT; t < (g + 1) * Prog::GT; ++t) {
              if (k == 0) Prog::set_tile(regs[t], row + p.kother0, kz0);
              Prog::phase(pass, k, regs[t], tb, xg, p, roots[t]);
            }
      }
      Prog::host_tile_store(p, row, kz0, tb);
    }
  }
}

template <int MODE>
static void run_line4(LineParams p, int nblocks) {
  if (g_emu_line4_form == 16) run_line4_prog<StridedLine16<1024, 8, MODE>>(p, nblocks);
  else run_line4_prog<StridedLine4<1024, 8, MODE>>(p, nblocks);
}

static int g_emu_line_kz = 0;       // 8 / 16: 512-point strided passes through the TMA-tiled program
static int g_emu_xkz = 8;           // columns per tile of the x pass (8 or 16; 16 may straddle y groups)
static int g_emu_pipe_blocks = 0;   // 0: one block per tile (StridedPass), >0: pipelined with that many blocks

// plain single-domain layouts only (what fft_native.cu routes to the TMA-tiled passes)
template <int MODE>
static bool try_line(int L, const StridedParams& p, int nx, int ny) {
  if (!g_emu_line_kz || (L != 512 && !(L == 1024 && g_emu_line_kz == 8)) || p.use_peers || p.in != p.out)
    return false;
  LineParams lp;
  lp.spec = p.out; lp.tw = p.tw; lp.nx = nx; lp.ny = ny; lp.P = p.P; lp.ncols_valid = p.ncols_valid;
  lp.along_x = pass_is_xmid(MODE) ? 1 : 0; lp.filt = p.filt;
  const int nb = g_emu_pipe_blocks > 0 ? g_emu_pipe_blocks : 5;
  if (L == 1024) { run_line4<MODE>(lp, nb); return true; }
  if (g_emu_line_kz == 16) run_line<16, MODE>(lp, nb); else run_line<8, MODE>(lp, nb);
  return true;
}

template <int KZ, int MODE>
static int dispatch_strided(int L, StridedParams p) {
  finalize_strided(p, L);
  if (g_emu_pipe_blocks > 0) {
    switch (L) {
      case 8: run_pipe<8, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 16: run_pipe<16, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 32: run_pipe<32, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 64: run_pipe<64, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 128: run_pipe<128, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 256: run_pipe<256, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 512: run_pipe<512, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      case 1024: run_pipe<1024, KZ, MODE>(p, g_emu_pipe_blocks); return 0;
      default: return -1;
    }
  }
  switch (L) {
    case 8: run_strided<8, KZ, MODE>(p); return 0;
    case 16: run_strided<16, KZ, MODE>(p); return 0;
    case 32: run_strided<32, KZ, MODE>(p); return 0;
    case 64: run_strided<64, KZ, MODE>(p); return 0;
    case 128: run_strided<128, KZ, MODE>(p); return 0;
    case 256: run_strided<256, KZ, MODE>(p); return 0;
    case 512: run_strided<512, KZ, MODE>(p); return 0;
    case 1024: run_strided<1024, KZ, MODE>(p); return 0;
    default: return -1;
  }
}

template <int M, int NL, bool INV>
static void run_z(const ZParams& p) {
  using Prog = ZPass<M, NL, INV>;
  const long long blocks = (p.rows + NL - 1) / NL;
  std::vector<typename Prog::Regs> regs(Prog::NTHREADS);
  std::vector<cf> smem(Prog::SMEM_BYTES / sizeof(cf) + 1);
  for (long long b = 0; b < blocks; ++b) {
    for (int t = 0; t < Prog::NTHREADS; ++t) Prog::init(regs[t], p, t, b);
    for (int k = 0; k < Prog::NPHASES; ++k)
      for (int t = 0; t < Prog::NTHREADS; ++t) Prog::phase(k, regs[t], smem.data(), p);
  }
}

template <bool INV>
static int dispatch_z(int M, const ZParams& p) {
  switch (M) {
    case 8: run_z<8, 4, INV>(p); return 0;
    case 16: run_z<16, 4, INV>(p); return 0;
    case 32: run_z<32, 4, INV>(p); return 0;
    case 64: run_z<64, 4, INV>(p); return 0;
    case 128: run_z<128, 4, INV>(p); return 0;
    case 256: run_z<256, 8, INV>(p); return 0;
    case 512: run_z<512, 2, INV>(p); return 0;
    case 1024: run_z<1024, 2, INV>(p); return 0;
    default: return -1;
  }
}

extern "C" {

void emu_set_pipe_blocks(int n) { g_emu_pipe_blocks = n; }
void emu_set_line_columns(int kz) { g_emu_line_kz = kz; }
void emu_set_line4_form(int f) { g_emu_line4_form = f; }
void emu_set_xpass_columns(int kz) { g_emu_xkz = kz; }

// out = u + irfftn(P * rfftn(r)) through the five native passes; spec is scratch
// [nx*ny*P] complex with P = roundup(nz/2+1, 8).  mode: 0 full, 1 forward only (spec out),
int emu_native_apply(const float* u, const float* r, float* out, float* spec_out, int nx, int ny,
                     int nz, const double* h, double dt, double coef, int power) {
  constexpr int KZ = 8;
  const int M = nz / 2, P = ((M + 1 + KZ - 1) / KZ) * KZ;
  std::vector<cf> spec((size_t)nx * ny * P, cf{0.f, 0.f});
  auto twx = make_roots(nx, nx), twy = make_roots(ny, ny), twz = make_roots(M, M),
       twr = make_roots(nz, M + 1);
  ZParams zp;
  zp.real_in = r; zp.real_out = nullptr; zp.spec = spec.data(); zp.tw = twz.data();
  zp.twr = twr.data(); zp.rows = (long long)nx * ny; zp.nz = nz; zp.P = P;
  if (dispatch_z<false>(M, zp)) return -1;
  StridedParams sp;
  sp.in = spec.data(); sp.out = spec.data(); sp.P = P; sp.ncols_valid = M + 1; sp.kother_offset = 0;
  sp.use_peers = 0;
  const StridedIO yio = plain_io(P, (long long)ny * P, ny), xio = plain_io((long long)ny * P, P, nx);
  // y pass: columns (x, kz), line stride P, group stride ny*P
  sp.tw = twy.data(); sp.src = sp.dst = yio;
  sp.ncols_total = (long long)nx * P;
  sp.filt = FilterParams{};
  if (!try_line<PASS_FWD>(ny, sp, nx, ny) && dispatch_strided<KZ, PASS_FWD>(ny, sp)) return -2;
  if (spec_out) {   // forward-only check: finish x forward and return the spectrum
    sp.tw = twx.data(); sp.src = sp.dst = xio;
    sp.ncols_total = (long long)ny * P;
    if (dispatch_strided<KZ, PASS_FWD>(nx, sp)) return -3;
    for (size_t i = 0; i < spec.size(); ++i) { spec_out[2 * i] = spec[i].x; spec_out[2 * i + 1] = spec[i].y; }
    return P;
  }
  // x pass: forward, filter, inverse
  const int n[3] = {nx, ny, nz};
  sp.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)nx * ny * nz));
  sp.tw = twx.data(); sp.src = sp.dst = xio;
  sp.ncols_total = (long long)ny * P;
  if (sp.filt.kind == FILTER_ETD1) {
    if (!try_line<PASS_XMID_ETD1>(nx, sp, nx, ny) && dispatch_strided<KZ, PASS_XMID_ETD1>(nx, sp)) return -3;
  } else if (try_line<PASS_XMID>(nx, sp, nx, ny)) {
  } else if (g_emu_xkz == 16 ? dispatch_strided<16, PASS_XMID>(nx, sp) : dispatch_strided<KZ, PASS_XMID>(nx, sp)) return -3;
  sp.tw = twy.data(); sp.src = sp.dst = yio;
  sp.ncols_total = (long long)nx * P;
  sp.filt = FilterParams{};
  if (!try_line<PASS_INV>(ny, sp, nx, ny) && dispatch_strided<KZ, PASS_INV>(ny, sp)) return -4;
  zp.real_in = u; zp.real_out = out;
  if (dispatch_z<true>(M, zp)) return -5;
  return 0;
}
}

// tile prefetch: tabulated-step form against the direct form, same (destination, source) pairs.
// layout 0: plain y pass [nx][L][P]; 1: plain x pass [L][ny][P]; 2: all-to-all block layout
// [W][nxl][L/W][P] read along y (the source of the distributed inverse y pass)
template <int L, int KZ>
static long long prefetch_mismatches(int layout, int groups, int P, int W) {
  using Pipe = StridedPipe<L, KZ, PASS_FWD>;
  StridedParams p{};
  long long elems;
  if (layout == 0) { p.src = plain_io(P, (long long)L * P, L); elems = (long long)groups * L * P; }
  else if (layout == 1) { p.src = plain_io((long long)groups * P, P, L); elems = (long long)L * groups * P; }
  else {
    const int nyl = L / W;
    p.src = StridedIO{P, (long long)nyl * P, (long long)groups * nyl * P, ilog2(nyl)};
    elems = (long long)W * groups * nyl * P;
  }
  p.dst = p.src;
  std::vector<cf> in((size_t)elems);
  for (size_t i = 0; i < in.size(); ++i) in[i] = cf{(float)(i & 0xffff), (float)(i >> 16)};
  p.in = in.data(); p.out = nullptr; p.P = P; p.ncols_valid = P; p.ncols_total = (long long)groups * P;
  finalize_strided(p, L);
  std::vector<cf> a((size_t)Pipe::BUF), b((size_t)Pipe::BUF);
  long long bad = 0;
  for (long long tile = 0; tile < Pipe::num_tiles(p); ++tile) {
    const long long c0 = tile * KZ, grp = c0 / P;
    const int kz0 = (int)(c0 - grp * P);
    std::fill(a.begin(), a.end(), cf{-1.f, -1.f});
    std::fill(b.begin(), b.end(), cf{-2.f, -2.f});
    for (int t = 0; t < Pipe::NTHREADS; ++t) {
      Pipe::prefetch_at(t, p, grp, kz0, a.data());
      Pipe::prefetch_at_direct(t, p, grp, kz0, b.data());
    }
    for (int i = 0; i < L * KZ; ++i) bad += (a[i].x != b[i].x || a[i].y != b[i].y);
  }
  return bad;
}
extern "C" long long emu_prefetch_mismatches(int L, int KZ, int layout, int groups, int P, int W) {
#define EMU_PF(LL, KK) if (L == LL && KZ == KK) return prefetch_mismatches<LL, KK>(layout, groups, P, W);
  EMU_PF(64, 8) EMU_PF(128, 8) EMU_PF(256, 8) EMU_PF(512, 8) EMU_PF(512, 16) EMU_PF(1024, 8)
  EMU_PF(2048, 4) EMU_PF(64, 16) EMU_PF(256, 16)
#undef EMU_PF
  return -1;
}

// -------------------------------------------------------------------------------------
// x-slab distributed pipeline: W virtual ranks through the parameters of dist_params.h
// -------------------------------------------------------------------------------------
#include "../../evoxels_b200/csrc/dist_params.h"
extern "C" {

// out = u + irfftn(P * rfftn(r)) on the global [nx,ny,nz] grid, computed by `world` virtual
// ranks (x slabs).  transport 0: local send buffer + all-to-all (NCCL path), 1: local block
// buffers with the self block written in place + block copies (copy-engine path), 2: stores
// straight into the peers' buffers (peer-store path).  fwd_chunks / mid_chunks: pipeline chunks
// of the host orchestration.
int emu_dist_apply(const float* u, const float* r, float* out, int nx, int ny, int nz, int world,
                   int transport, int fwd_chunks, int mid_chunks, const double* h,
                   double dt, double coef, int power) {
  if (nx % world || ny % world || world > 8) return -1;
  std::vector<DistDims> dims;
  for (int k = 0; k < world; ++k) dims.push_back(make_dist_dims(nx, ny, nz, world, k, 8));
  const DistDims& d0 = dims[0];
  const int M = d0.M, P = d0.P, nxl = d0.nxl, nyl = d0.nyl;
  const long long blk = (long long)nxl * nyl * P, slab_spec = (long long)nxl * ny * P;
  const long long slab_real = (long long)nxl * ny * nz;
  auto twx = make_roots(nx, nx), twy = make_roots(ny, ny), twz = make_roots(M, M),
       twr = make_roots(nz, M + 1);
  const DistTables t{twx.data(), twy.data(), twz.data(), twr.data()};
  const float nanv = std::nanf("");
  std::vector<std::vector<cf>> spec(world), A(world), B(world);
  for (int k = 0; k < world; ++k) {
    spec[k].assign((size_t)slab_spec, cf{nanv, nanv});
    A[k].assign((size_t)world * blk, cf{nanv, nanv});
    B[k].assign((size_t)world * blk, cf{nanv, nanv});
  }
  auto bounds = [](int n, int chunks, int i) { return (int)std::lround((double)i * n / chunks); };
  // ---- forward: z + y passes, blocks into the peers' B -------------------------------------
  for (int k = 0; k < world; ++k) {
    const DistDims& d = dims[k];
    const float* r_local = r + k * slab_real;
    void* table[8];
    void* const* peers = nullptr;
    cf* send = A[k].data();
    if (transport == 1) { local_block_table(d, A[k].data(), B[k].data(), table); peers = table; send = nullptr; }
    if (transport == 2) { for (int j = 0; j < 8; ++j) table[j] = j < world ? B[j].data() : nullptr; peers = table; send = nullptr; }
    for (int i = 0; i < fwd_chunks; ++i) {
      const int x0 = bounds(nxl, fwd_chunks, i), x1 = bounds(nxl, fwd_chunks, i + 1);
      if (x1 <= x0) continue;
      if (dispatch_z<false>(M, dist_zfwd_params(d, t, r_local, spec[k].data(), x0, x1 - x0))) return -2;
      if (dispatch_strided<8, PASS_FWD>(ny, dist_yfwd_params(d, t, spec[k].data(), send, peers, 0, x0, x1 - x0))) return -3;
    }
  }
  if (transport != 2)   // all-to-all / block copies: block j of rank k's A -> block k of rank j's B
    for (int k = 0; k < world; ++k)
      for (int j = 0; j < world; ++j) {
        if (transport == 1 && j == k) continue;          // written in place by the pass
        std::memcpy(B[j].data() + k * blk, A[k].data() + j * blk, (size_t)blk * sizeof(cf));
      }
  // ---- middle: x pass on the y-pencils, blocks into the peers' A ---------------------------
  for (int k = 0; k < world; ++k) std::fill(A[k].begin(), A[k].end(), cf{nanv, nanv});
  for (int k = 0; k < world; ++k) {
    const DistDims& d = dims[k];
    void* table[8];
    void* const* peers = nullptr;
    if (transport == 1) { local_block_table(d, B[k].data(), A[k].data(), table); peers = table; }
    if (transport == 2) { for (int j = 0; j < 8; ++j) table[j] = j < world ? A[j].data() : nullptr; peers = table; }
    const int mc = mid_chunks < 1 ? 1 : (mid_chunks > nyl ? nyl : mid_chunks);
    for (int i = 0; i < mc; ++i) {
      const int y0 = bounds(nyl, mc, i), y1 = bounds(nyl, mc, i + 1);
      if (y1 <= y0) continue;
      const StridedParams xp = dist_xmid_params(d, t, B[k].data(), peers, 0, h, dt, coef, power, y0, y1 - y0);
      if (xp.filt.kind == FILTER_ETD1 ? dispatch_strided<8, PASS_XMID_ETD1>(nx, xp)
                                      : dispatch_strided<8, PASS_XMID>(nx, xp)) return -4;
    }
  }
  if (transport != 2)
    for (int k = 0; k < world; ++k)
      for (int j = 0; j < world; ++j) {
        if (transport == 1 && j == k) continue;
        std::memcpy(A[j].data() + k * blk, B[k].data() + j * blk, (size_t)blk * sizeof(cf));
      }
  // ---- backward: y inverse + z inverse (+u) ------------------------------------------------
  for (int k = 0; k < world; ++k) {
    const DistDims& d = dims[k];
    std::fill(spec[k].begin(), spec[k].end(), cf{nanv, nanv});
    if (dispatch_strided<8, PASS_INV>(ny, dist_yinv_params(d, t, A[k].data(), spec[k].data(), 0, nxl))) return -5;
    if (dispatch_z<true>(M, dist_zinv_params(d, t, spec[k].data(), u ? u + k * slab_real : nullptr,
                                             out + k * slab_real, 0, nxl))) return -6;
  }
  return 0;
}
}

// =====================================================================================
// two-species reaction-diffusion rhs
// =====================================================================================
#include "../../evoxels_b200/csrc/rd_core.h"
template <typename T>
static int emu_rd2(const T* u, const T* inter, T* out, int nx, int ny, int nz, const double* h,
                   double DA, double DB, double feed, double kill, int vec) {
  RdParams<T> p;
  p.u = u; p.inter = inter; p.out = out; p.nx = nx; p.ny = ny; p.nz = nz;
  fill_rd_metric(p, h);
  p.DA = (T)DA; p.DB = (T)DB; p.feed = (T)feed; p.kill = (T)kill;
  constexpr int VW = 16 / (int)sizeof(T);
  const long long n = (long long)nx * ny * nz;
  if (vec) {
    if (nz % VW) return -1;
    for (long long g = 0; g < n / VW; ++g) RdProgram<T, VW>::run(p, g);
  } else {
    for (long long g = 0; g < n; ++g) RdProgram<T, 1>::run(p, g);
  }
  return 0;
}
extern "C" {
int emu_rd2_rhs_f32(const float* u, const float* inter, float* out, int nx, int ny, int nz,
                    const double* h, double DA, double DB, double feed, double kill, int vec) {
  return emu_rd2<float>(u, inter, out, nx, ny, nz, h, DA, DB, feed, kill, vec);
}
int emu_rd2_rhs_f64(const double* u, const double* inter, double* out, int nx, int ny, int nz,
                    const double* h, double DA, double DB, double feed, double kill, int vec) {
  return emu_rd2<double>(u, inter, out, nx, ny, nz, h, DA, DB, feed, kill, vec);
}
}

// =====================================================================================
// adjoint of the CH rhs
// =====================================================================================
#include "../../evoxels_b200/csrc/adjoint_core.h"

extern "C" {
// lam = dR/du^T w  and  *deps = dL/deps  for the periodic CH rhs (float64)
int emu_ch_rhs_vjp_f64(const double* u, const double* w, double* lam, double* deps, int nx, int ny,
                       int nz, const double* h, double eps, double D) {
  const size_t n = (size_t)nx * ny * nz;
  std::vector<double> mu(n), z(n), m(n);
  AdjParams<double> p;
  p.u = u; p.mu = mu.data(); p.w = w; p.z = z.data(); p.m = m.data(); p.lam_in = nullptr;
  p.red = nullptr; p.nx = nx; p.ny = ny; p.nz = nz;
  p.ihx2 = 1.0 / (h[0] * h[0]); p.ihy2 = 1.0 / (h[1] * h[1]); p.ihz2 = 1.0 / (h[2] * h[2]);
  p.eps = eps; p.D = D;
  auto idx = [&](int x, int y, int zz) { return ((size_t)x * ny + y) * nz + zz; };
  for (int x = 0; x < nx; ++x) for (int y = 0; y < ny; ++y) for (int zz = 0; zz < nz; ++zz)
    mu[idx(x, y, zz)] = adj_mu(p, x, y, zz);
  for (int x = 0; x < nx; ++x) for (int y = 0; y < ny; ++y) for (int zz = 0; zz < nz; ++zz)
    adj_flux(p, x, y, zz, z[idx(x, y, zz)], m[idx(x, y, zz)]);
  double acc = 0.0;
  for (int x = 0; x < nx; ++x) for (int y = 0; y < ny; ++y) for (int zz = 0; zz < nz; ++zz) {
    double term;
    lam[idx(x, y, zz)] = adj_combine(p, x, y, zz, term);
    acc += term;
  }
  *deps = acc;
  return 0;
}
}

// =====================================================================================
// mixed-radix passes (non-power-of-two extents)
// =====================================================================================
#include "../../evoxels_b200/csrc/fft_generic_core.h"
template <typename R>
static int emu_generic_pass(const GenericParams<R>& p, long long nblocks, int nthreads) {
  std::vector<gcplx<R>> smem(GenericProgram<R>::smem_elems(p));
  const int nph = GenericProgram<R>::nphases(p);
  for (long long b = 0; b < nblocks; ++b)
    for (int k = 0; k < nph; ++k)
      for (int t = 0; t < nthreads; ++t) GenericProgram<R>::phase(k, p, smem.data(), b, t, nthreads);
  return 0;
}

template <typename R>
static int emu_generic_apply(const R* u, const R* r, R* out, int nx, int ny, int nz, const double* h,
                             double dt, double coef, int power, int W, int nthreads) {
  LineDesc lx, ly, lz;
  if (!factor_line(nx, lx) || !factor_line(ny, ly) || !factor_line(nz, lz)) return -1;
  const int P = ((nz / 2 + 1 + 7) / 8) * 8;
  std::vector<gcplx<R>> spec((size_t)nx * ny * P, gcplx<R>{R(0), R(0)});
  auto roots = [](int n) {
    std::vector<gcplx<R>> w(n);
    for (int m = 0; m < n; ++m) {
      const double a = -2.0 * M_PI * (double)m / (double)n;
      w[m] = gcplx<R>{(R)std::cos(a), (R)std::sin(a)};
    }
    return w;
  };
  auto twx = roots(nx), twy = roots(ny), twz = roots(nz);
  GenericParams<R> p{};
  p.spec = spec.data(); p.nz = nz; p.P = P; p.rows = (long long)nx * ny; p.ncols_valid = nz / 2 + 1;
  p.W = W;
  p.mode = GEN_Z_FWD; p.line = lz; p.tw = twz.data(); p.real_in = r; p.real_out = nullptr;
  emu_generic_pass(p, (p.rows + W - 1) / W, nthreads);
  GenericParams<R> y = p;
  y.mode = GEN_FWD; y.line = ly; y.tw = twy.data();
  y.line_stride = P; y.group_stride = (long long)ny * P; y.ncols_total = (long long)nx * P;
  emu_generic_pass(y, (y.ncols_total + W - 1) / W, nthreads);
  GenericParams<R> x = p;
  x.mode = GEN_XMID; x.line = lx; x.tw = twx.data();
  x.line_stride = (long long)ny * P; x.group_stride = P; x.ncols_total = (long long)ny * P;
  const int n[3] = {nx, ny, nz};
  x.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)nx * ny * nz));
  emu_generic_pass(x, (x.ncols_total + W - 1) / W, nthreads);
  y.mode = GEN_INV;
  emu_generic_pass(y, (y.ncols_total + W - 1) / W, nthreads);
  p.mode = GEN_Z_INV; p.real_in = u; p.real_out = out;
  emu_generic_pass(p, (p.rows + W - 1) / W, nthreads);
  return 0;
}
extern "C" {
int emu_generic_apply_f32(const float* u, const float* r, float* out, int nx, int ny, int nz,
                          const double* h, double dt, double coef, int power, int W, int nthreads) {
  return emu_generic_apply<float>(u, r, out, nx, ny, nz, h, dt, coef, power, W, nthreads);
}
int emu_generic_apply_f64(const double* u, const double* r, double* out, int nx, int ny, int nz,
                          const double* h, double dt, double coef, int power, int W, int nthreads) {
  return emu_generic_apply<double>(u, r, out, nx, ny, nz, h, dt, coef, power, W, nthreads);
}
}

// -------------------------------------------------------------------------------------
// Chained z/y passes: the work list of fft_chain_core.h replayed with `blocks` resident
// blocks.  Checks that every (stage, plane, idx) occurs exactly once, that every dependency
// points to a smaller item number, and that blocks walking their items in order - a stage-1
// item may only start once all stage-0 items of its plane have finished - always make
// progress.  Returns 0 if all holds; *waits = stage-1 items that found their plane unfinished
// when every block advances one item per sweep (how often the kernel would spin).
// -------------------------------------------------------------------------------------
#include "../../evoxels_b200/csrc/fft_chain_core.h"
extern "C" int emu_chain_schedule_check(int nplanes, int lag, int n0, int n1, int blocks,
                                        long long* waits) {
  const ChainSchedule s = make_chain_schedule(nplanes, lag, n0, n1);
  if (s.total != (long long)nplanes * (n0 + n1)) return 1;
  std::vector<int> seen0((size_t)nplanes * n0, 0), seen1((size_t)nplanes * n1, 0);
  std::vector<long long> last0(nplanes, -1);
  for (long long i = 0; i < s.total; ++i) {
    const ChainItem it = chain_decode(s, i);
    if (it.plane < 0 || it.plane >= nplanes) return 2;
    if (it.stage == 0) {
      if (it.idx < 0 || it.idx >= n0 || seen0[(size_t)it.plane * n0 + it.idx]++) return 3;
      last0[it.plane] = i;
    } else if (it.stage == 1) {
      if (it.idx < 0 || it.idx >= n1 || seen1[(size_t)it.plane * n1 + it.idx]++) return 4;
      if (last0[it.plane] < 0) return 5;
      for (int k = 0; k < n0; ++k)
        if (!seen0[(size_t)it.plane * n0 + k]) return 6;      // a dependency with a larger number
    } else {
      return 7;
    }
  }
  for (int v : seen0) if (v != 1) return 8;
  for (int v : seen1) if (v != 1) return 9;
  // progress: sweep over the blocks, each runs its next item if it can
  std::vector<long long> next(blocks);
  std::vector<int> done0(nplanes, 0);
  for (int b = 0; b < blocks; ++b) next[b] = b;
  long long finished = 0, w = 0;
  while (finished < s.total) {
    bool progress = false;
    std::vector<int> inc(nplanes, 0);          // completions become visible after the sweep
    for (int b = 0; b < blocks; ++b) {
      if (next[b] >= s.total) continue;
      const ChainItem it = chain_decode(s, next[b]);
      if (it.stage == 1 && done0[it.plane] < n0) { ++w; continue; }
      if (it.stage == 0) ++inc[it.plane];
      next[b] += blocks;
      ++finished;
      progress = true;
    }
    for (int x = 0; x < nplanes; ++x) done0[x] += inc[x];
    if (!progress) return 10;                  // deadlock
  }
  if (waits) *waits = w;
  return 0;
}

// the same checks for the division-free cursor the kernels use (virtual numbering)
extern "C" int emu_chain_cursor_check(int nplanes, int lag, int n0, int n1, int blocks, long long* waits) {
  const ChainSchedule s = make_chain_schedule(nplanes, lag, n0, n1);
  std::vector<int> seen0((size_t)nplanes * n0, 0), seen1((size_t)nplanes * n1, 0);
  // per block: list of its items in order, with their virtual slot numbers
  struct Ent { long long slot; ChainItem it; };
  std::vector<std::vector<Ent>> lists(blocks);
  for (int b = 0; b < blocks; ++b) {
    ChainCursor c;
    chain_cursor_init(c, s, b, blocks);
    ChainItem it;
    while (chain_cursor_next(c, s, it)) {
      lists[b].push_back({(long long)c.round * (n0 + n1) + c.w, it});
      chain_cursor_step(c, s);
    }
  }
  std::vector<long long> last0(nplanes, -1), first1(nplanes, -1);
  for (int b = 0; b < blocks; ++b) {
    long long prev = -1;
    for (const Ent& e : lists[b]) {
      if (e.slot <= prev || (e.slot - b) % blocks) return 1;
      prev = e.slot;
      const ChainItem& it = e.it;
      if (it.plane < 0 || it.plane >= nplanes) return 2;
      if (it.stage == 0) {
        if (it.idx < 0 || it.idx >= n0 || seen0[(size_t)it.plane * n0 + it.idx]++) return 3;
        if (e.slot > last0[it.plane]) last0[it.plane] = e.slot;
      } else {
        if (it.idx < 0 || it.idx >= n1 || seen1[(size_t)it.plane * n1 + it.idx]++) return 4;
        if (first1[it.plane] < 0 || e.slot < first1[it.plane]) first1[it.plane] = e.slot;
      }
    }
  }
  for (int v : seen0) if (v != 1) return 8;
  for (int v : seen1) if (v != 1) return 9;
  for (int x = 0; x < nplanes; ++x) if (first1[x] <= last0[x]) return 6;   // dependency order
  std::vector<size_t> next(blocks, 0);
  std::vector<int> done0(nplanes, 0);
  long long remaining = (long long)nplanes * (n0 + n1), w = 0;
  while (remaining > 0) {
    bool progress = false;
    std::vector<int> inc(nplanes, 0);
    for (int b = 0; b < blocks; ++b) {
      if (next[b] >= lists[b].size()) continue;
      const ChainItem& it = lists[b][next[b]].it;
      if (it.stage == 1 && done0[it.plane] < n0) { ++w; continue; }
      if (it.stage == 0) ++inc[it.plane];
      ++next[b]; --remaining; progress = true;
    }
    for (int x = 0; x < nplanes; ++x) done0[x] += inc[x];
    if (!progress) return 10;
  }
  if (waits) *waits = w;
  return 0;
}

// -------------------------------------------------------------------------------------
// z lines in the two-lines-per-group form (fft_zline_core.h) against ZPass, bit for bit.
// One item = 16 rows; the bulk copies are restated as row copies into the 264-complex pitch;
// groups are replayed one after the other (no ordering between groups, as on the device), the
// row buffer and the exchange buffers are poisoned before every item.
// returns the number of differing floats (forward spectrum rows + inverse real rows)
// -------------------------------------------------------------------------------------
#include "../../evoxels_b200/csrc/fft_zline_core.h"
extern "C" long long emu_zgroup_mismatches(const float* r, const float* u, int rows_total) {
  constexpr int M = 256, NZ = 512, P = 264, NL = 16, NT = 512;
  if (rows_total % NL) return -1;
  auto twz = make_roots(M, M), twr = make_roots(NZ, M + 1);
  const cf nan2{std::nanf(""), std::nanf("")};
  // reference: ZPass
  std::vector<cf> spec_ref((size_t)rows_total * P, cf{0.f, 0.f}), spec_new((size_t)rows_total * P, cf{0.f, 0.f});
  std::vector<float> out_ref((size_t)rows_total * NZ, 0.f), out_new((size_t)rows_total * NZ, 0.f);
  ZParams zp;
  zp.real_in = r; zp.real_out = nullptr; zp.spec = spec_ref.data(); zp.tw = twz.data(); zp.twr = twr.data();
  zp.rows = rows_total; zp.nz = NZ; zp.P = P; zp.pf_blocks = 0;
  run_z<M, 8, false>(zp);
  zp.real_in = u; zp.real_out = out_ref.data();
  run_z<M, 8, true>(zp);
  // new form
  ZGroupParams gp;
  gp.tw = twz.data(); gp.twr = twr.data(); gp.nz = NZ; gp.P = P;
  using F = ZGroupLine<false>;
  using I = ZGroupLine<true>;
  std::vector<cf> rowsbuf((size_t)NL * F::ROWP + 8), xall((size_t)(NT / F::GT) * F::XG);
  for (int item = 0; item < rows_total / NL; ++item) {
    {  // forward
      std::fill(rowsbuf.begin(), rowsbuf.end(), nan2);
      std::fill(xall.begin(), xall.end(), nan2);
      for (int l = 0; l < NL; ++l)
        std::memcpy(&rowsbuf[(size_t)l * F::ROWP], r + ((size_t)item * NL + l) * NZ, NZ * sizeof(float));
      std::vector<F::Regs> regs(NT);
      for (int t = 0; t < NT; ++t) F::init(regs[t], t);
      for (int g = NT / F::GT - 1; g >= 0; --g)
        for (int k = 0; k < F::NPHASES; ++k)
          for (int t = g * F::GT; t < (g + 1) * F::GT; ++t) {
            const long long grow = (long long)item * NL + regs[t].g * F::G + regs[t].c2;
            F::phase(k, regs[t], rowsbuf.data(), xall.data() + (size_t)g * F::XG, gp, grow, nullptr, nullptr,
                     spec_new.data());
          }
    }
    {  // inverse from the reference spectrum rows
      std::fill(rowsbuf.begin(), rowsbuf.end(), nan2);
      std::fill(xall.begin(), xall.end(), nan2);
      std::memcpy(rowsbuf.data(), &spec_ref[(size_t)item * NL * P], (size_t)NL * P * sizeof(cf));
      std::vector<I::Regs> regs(NT);
      for (int t = 0; t < NT; ++t) I::init(regs[t], t);
      for (int g = 0; g < NT / I::GT; ++g)
        for (int k = 0; k < I::NPHASES; ++k)
          for (int t = g * I::GT; t < (g + 1) * I::GT; ++t) {
            const long long grow = (long long)item * NL + regs[t].g * I::G + regs[t].c2;
            I::phase(k, regs[t], rowsbuf.data(), xall.data() + (size_t)g * I::XG, gp, grow, u, out_new.data(),
                     nullptr);
          }
    }
  }
  long long bad = 0;
  for (long long row = 0; row < rows_total; ++row)
    for (int k = 0; k <= M; ++k) {
      const cf a = spec_ref[(size_t)row * P + k], b = spec_new[(size_t)row * P + k];
      bad += std::memcmp(&a, &b, sizeof(cf)) != 0;
    }
  bad += std::memcmp(out_ref.data(), out_new.data(), out_ref.size() * sizeof(float)) != 0
             ? [&] { long long n = 0; for (size_t i = 0; i < out_ref.size(); ++i) n += std::memcmp(&out_ref[i], &out_new[i], 4) != 0; return n; }()
             : 0;
  return bad;
}
