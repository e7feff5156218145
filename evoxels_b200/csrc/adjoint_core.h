// Hand-written adjoint (vector-Jacobian product) of the Cahn-Hilliard right-hand side on a
// fully periodic grid - the backward pass behind evoxels_b200.autograd (the reference gets
// this from JAX/diffrax in evoxels/inversion.py:51-124; on torch it relies on autograd
// through ~60 tensor ops per step).
//
// Forward:  c^ = clip(u,0,1),  mu = g(c^) - 2 eps lap(c^),  g = (18/eps) p(c^),
//           p(c) = c(1-c)(1-2c),  R = D div( M(c_f) grad mu ),  M(c) = c(1-c)
// Given w = dL/dR:
//   z   = D div( M(c_f) grad w )                            (same operator, w in place of mu)
//   m_i = sum_a 1/2 [ t(f+) + t(f-) ],  t(f) = -D (1 - 2 c_f) (d+ w)_f (d+ mu)_f
//   dL/du   = 1[0<=u<=1] * ( g'(c^) z - 2 eps lap(z) + m ),   g' = (18/eps)(1 - 6c + 6c^2)
//   dL/deps = < z, -(18/eps^2) p(c^) - 2 lap(c^) >
//   dL/dD   = < w, R > / D                                   (formed by the caller)
// Three element-wise-with-7-point-footprint kernels: mu, (z, m), combine.  All neighbour
// reads go through L1/L2 with periodic index wrap.
#pragma once
#include "evx_hd.h"

namespace evx {

template <typename T>
struct AdjParams {
  const T* u;      // state (raw)
  const T* mu;     // chemical potential (written by op 0, read by op 1)
  const T* w;      // incoming cotangent of R
  const T* z;      // op 2 input
  const T* m;      // op 2 input
  const T* lam_in; // op 2: optional term added to the result (dL/du+ of the step)
  T* out0;         // op 0: mu      op 1: z      op 2: dL/du
  T* out1;         //               op 1: m
  double* red;     // op 2: device accumulator for dL/deps (atomicAdd of block partials)
  int red_x0, red_x1;   // op 2: planes whose integrand enters `red` (an x-slab extended by its
                        // halo planes passes the slab's own planes; default: all)
  int nx, ny, nz;
  T ihx2, ihy2, ihz2;   // 1/h^2
  T eps, D;
};

template <typename T>
struct Nbr7 {
  long long c, xm, xp, ym, yp, zm, zp;
};

template <typename T>
EVX_HD Nbr7<T> nbr7(const AdjParams<T>& p, int x, int y, int z) {
  const long long sy = p.nz, sx = (long long)p.ny * p.nz;
  const int xm = x == 0 ? p.nx - 1 : x - 1, xp = x == p.nx - 1 ? 0 : x + 1;
  const int ym = y == 0 ? p.ny - 1 : y - 1, yp = y == p.ny - 1 ? 0 : y + 1;
  const int zm = z == 0 ? p.nz - 1 : z - 1, zp = z == p.nz - 1 ? 0 : z + 1;
  Nbr7<T> n;
  n.c = x * sx + y * sy + z;
  n.xm = xm * sx + y * sy + z; n.xp = xp * sx + y * sy + z;
  n.ym = x * sx + ym * sy + z; n.yp = x * sx + yp * sy + z;
  n.zm = x * sx + y * sy + zm; n.zp = x * sx + y * sy + zp;
  return n;
}

template <typename T>
EVX_HD T lap7(const T* f, const Nbr7<T>& n, const AdjParams<T>& p, bool clip) {
  const T c = clip ? clip01(f[n.c]) : f[n.c];
  const T xs = clip ? clip01(f[n.xp]) + clip01(f[n.xm]) : f[n.xp] + f[n.xm];
  const T ys = clip ? clip01(f[n.yp]) + clip01(f[n.ym]) : f[n.yp] + f[n.ym];
  const T zs = clip ? clip01(f[n.zp]) + clip01(f[n.zm]) : f[n.zp] + f[n.zm];
  return (xs - T(2) * c) * p.ihx2 + (ys - T(2) * c) * p.ihy2 + (zs - T(2) * c) * p.ihz2;
}

// op 0: mu = (18/eps) p(c^) - 2 eps lap(c^)
template <typename T>
EVX_HD T adj_mu(const AdjParams<T>& p, int x, int y, int z) {
  const Nbr7<T> n = nbr7(p, x, y, z);
  const T c = clip01(p.u[n.c]);
  return T(18) / p.eps * c * (T(1) - c) * (T(1) - T(2) * c) - T(2) * p.eps * lap7(p.u, n, p, true);
}

// op 1: z and m at one voxel
template <typename T>
EVX_HD void adj_flux(const AdjParams<T>& p, int x, int y, int z, T& zo, T& mo) {
  const Nbr7<T> n = nbr7(p, x, y, z);
  const T c0 = clip01(p.u[n.c]), w0 = p.w[n.c], m0 = p.mu[n.c];
  const long long nb[6] = {n.xp, n.xm, n.yp, n.ym, n.zp, n.zm};
  const T ih2[3] = {p.ihx2, p.ihy2, p.ihz2};
  T zacc = T(0), macc = T(0);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const T cf = T(0.5) * (c0 + clip01(p.u[nb[k]]));
    const T dw = p.w[nb[k]] - w0;            // (neighbour - centre): outward difference
    const T dm = p.mu[nb[k]] - m0;
    const T g = ih2[k / 2];
    zacc += cf * (T(1) - cf) * dw * g;       // sum over the 6 faces of M dw / h^2
    macc += (T(1) - T(2) * cf) * dw * dm * g;
  }
  zo = p.D * zacc;
  mo = -T(0.5) * p.D * macc;
}

// op 2: dL/du (plus optional lam_in) and the dL/deps integrand
template <typename T>
EVX_HD T adj_combine(const AdjParams<T>& p, int x, int y, int z, double& deps_term) {
  const Nbr7<T> n = nbr7(p, x, y, z);
  const T uraw = p.u[n.c];
  const T c = clip01(uraw);
  const T zc = p.z[n.c];
  const T lapz = lap7(p.z, n, p, false);
  const T lapc = lap7(p.u, n, p, true);
  const T pc = c * (T(1) - c) * (T(1) - T(2) * c);
  deps_term = (double)zc * (double)(-T(18) / (p.eps * p.eps) * pc - T(2) * lapc);
  const bool inside = uraw >= T(0) && uraw <= T(1);     // torch.clamp passes gradient on [0,1]
  T g = T(0);
  if (inside)
    g = T(18) / p.eps * (T(1) - T(6) * c + T(6) * c * c) * zc - T(2) * p.eps * lapz + p.m[n.c];
  return p.lam_in ? p.lam_in[n.c] + g : g;
}

}  // namespace evx
