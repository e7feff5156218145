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
      for (int pl = regs[0].xa - 1; pl <= regs[0].xb; ++pl) {
        for (int t = 0; t < Prog::NTHREADS; ++t) Prog::phase_a(regs[t], *s, p, pl);
        for (int t = 0; t < Prog::NTHREADS; ++t) Prog::phase_b(regs[t], *s, p, pl);
      }
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
    run_ch<T, VW, 16, 16>(p);
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
    run_ac<T, VW, 8, 32>(p);
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
