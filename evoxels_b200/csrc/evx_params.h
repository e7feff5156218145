// Host-side translation of C-ABI arguments into kernel parameter blocks.
// Shared by the CUDA launchers and by the CPU block emulator under tests/emu/.
#pragma once
#include "ch_rhs_core.h"
#include "ac_core.h"

namespace evx {

// bc_kind[3] in {BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET}; bc_val[6] = (x_lo,x_hi,y_lo,...)
template <typename T>
inline void fill_ghost_rules(const int* bc_kind, const double* bc_val, int kind_out[3],
                             T off[3][2], T sgn[3]) {
  for (int a = 0; a < 3; ++a) {
    kind_out[a] = bc_kind[a];
    if (bc_kind[a] == BC_DIRICHLET) {
      sgn[a] = T(-1);
      off[a][0] = T(2.0 * bc_val[2 * a]);
      off[a][1] = T(2.0 * bc_val[2 * a + 1]);
    } else {  // Neumann: ghost = inner.  Periodic never reads these.
      sgn[a] = T(1);
      off[a][0] = off[a][1] = T(0);
    }
  }
}

template <typename T>
inline ChParams<T> make_ch_params(const T* c, const T* hom, T* out, int nx, int ny, int nz,
                                  const double* h, double eps, double D, const int* bc_kind,
                                  const double* bc_val, const T* halo_lo, const T* halo_hi,
                                  int xchunk) {
  ChParams<T> p;
  p.c = c; p.hom = hom; p.out = out; p.halo_lo = halo_lo; p.halo_hi = halo_hi;
  p.nx = nx; p.ny = ny; p.nz = nz; p.xchunk = xchunk;
  // metric factors folded in double, rounded once to T (the kernel reassociates the
  // reference's expression anyway; see ch_rhs_core.h)
  const double ih2[3] = {1.0 / (h[0] * h[0]), 1.0 / (h[1] * h[1]), 1.0 / (h[2] * h[2])};
  p.pot_scale = T(18.0 / eps);
  p.lx = T(-2.0 * eps * ih2[0]); p.ly = T(-2.0 * eps * ih2[1]); p.lz = T(-2.0 * eps * ih2[2]);
  p.l0 = T(4.0 * eps * (ih2[0] + ih2[1] + ih2[2]));
  p.fx = T(0.25 * D * ih2[0]); p.fy = T(0.25 * D * ih2[1]); p.fz = T(0.25 * D * ih2[2]);
  fill_ghost_rules<T>(bc_kind, bc_val, p.bc_kind, p.ghost_off, p.ghost_sgn);
  return p;
}


template <typename T, typename PT>
inline void fill_metric(PT& p, const double* h) {
  const T hx = T(h[0]), hy = T(h[1]), hz = T(h[2]);
  p.ihx = T(1) / hx; p.ihy = T(1) / hy; p.ihz = T(1) / hz;
  p.ihx2 = T(1) / (hx * hx); p.ihy2 = T(1) / (hy * hy); p.ihz2 = T(1) / (hz * hz);
  p.ih2sum = (p.ihx2 + p.ihy2) + p.ihz2;
}

template <typename T>
inline AcParams<T> make_ac_params(const T* phi, const T* pot, T* k_out, const T* base, T* y_out,
                                  double alpha, const T* acc_in, T* acc_out, double beta, int nx,
                                  int ny, int nz, const double* h, double eps, double gab,
                                  double M, double force, double curvature, const int* bc_kind,
                                  const double* bc_val, const T* halo_lo, const T* halo_hi,
                                  int xchunk) {
  AcParams<T> p;
  p.phi = phi; p.pot = pot; p.k_out = k_out; p.base = base; p.y_out = y_out;
  p.acc_in = acc_in; p.acc_out = acc_out; p.alpha = T(alpha); p.beta = T(beta);
  p.halo_lo = halo_lo; p.halo_hi = halo_hi;
  p.nx = nx; p.ny = ny; p.nz = nz; p.xchunk = xchunk;
  fill_metric<T>(p, h);
  p.pot_scale = T(18.0 / eps); p.eps = T(eps); p.gab = T(gab); p.M = T(M); p.force = T(force);
  p.curv = T(curvature); p.omc = T(1.0 - curvature); p.three_over_eps = T(3.0 / eps);
  p.hihx = T(0.5 / h[0]); p.hihy = T(0.5 / h[1]); p.hihz = T(0.5 / h[2]);
  p.hxy = T(0.5 / (h[0] * h[1])); p.hxz = T(0.5 / (h[0] * h[2])); p.hyz = T(0.5 / (h[1] * h[2]));
  p.neg_inv_2eps = T(-0.5 / eps); p.force3 = T(3.0 / eps * force);
  fill_ghost_rules<T>(bc_kind, bc_val, p.bc_kind, p.ghost_off, p.ghost_sgn);
  return p;
}

template <typename T>
inline PadParams<T> make_pad_params(const T* in, T* out, int nx, int ny, int nz,
                                    const int* bc_kind, const double* bc_val) {
  PadParams<T> p;
  p.in = in; p.out = out; p.nx = nx; p.ny = ny; p.nz = nz;
  fill_ghost_rules<T>(bc_kind, bc_val, p.bc_kind, p.ghost_off, p.ghost_sgn);
  return p;
}

template <typename T>
inline PaddedStencilParams<T> make_padded_stencil_params(const T* g, T* out, int nx, int ny,
                                                         int nz, const double* h, int op) {
  PaddedStencilParams<T> p;
  p.g = g; p.out = out; p.nx = nx; p.ny = ny; p.nz = nz; p.op = op;
  fill_metric<T>(p, h);
  return p;
}

}  // namespace evx
