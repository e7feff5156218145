// Two-species reaction-diffusion right-hand side (Gray-Scott form), fused, fully periodic.
//
//   I    = A B^2                                   (or a caller-supplied field)
//   dA   = D_A lap7(A) - I + feed (1 - A)
//   dB   = D_B lap7(B) + I - kill B
//
// Replaces reference evoxels/problem_definition.py:614-633 (CoupledReactionDiffusion.rhs:
// pad_periodic of both channels, fd_stencils.py:62-75 laplace, five elementwise kernels and
// a stack).  The class is always periodic in the reference (it has no `bc` field).
//
// One thread owns V contiguous z values of both species; y/x neighbours are re-read through
// L1/L2 (the 7-point footprint of a 16 B/voxel update keeps every line hot), z neighbours
// come from the own vector plus one scalar each side.  lap7 keeps the reference's
// association: (R+L)/hx^2 + (T+B)/hy^2 + (F+Bk)/hz^2 - 2 C sum(1/h^2).
#pragma once
#include "evx_hd.h"

namespace evx {

template <typename T>
struct RdParams {
  const T* u;        // [2, nx, ny, nz]
  const T* inter;    // optional [nx, ny, nz]: interaction(u) evaluated by the caller
  T* out;            // [2, nx, ny, nz]
  int nx, ny, nz;
  T ihx2, ihy2, ihz2, ih2sum;
  T DA, DB, feed, kill;
};

template <typename T>
inline void fill_rd_metric(RdParams<T>& p, const double* h) {
  const T hx = T(h[0]), hy = T(h[1]), hz = T(h[2]);
  p.ihx2 = T(1) / (hx * hx); p.ihy2 = T(1) / (hy * hy); p.ihz2 = T(1) / (hz * hz);
  p.ih2sum = (p.ihx2 + p.ihy2) + p.ihz2;
}

template <typename T, int V>
struct RdProgram {
  using Vt = Vec<T, V>;
  using P = RdParams<T>;

  // Laplacian of one species at the V values starting at (x, y, z)
  EVX_HD static Vt lap(const P& p, const T* f, int x, int y, int z, const Vt& c) {
    const long long sy = p.nz, sx = (long long)p.ny * p.nz;
    const int xm = x == 0 ? p.nx - 1 : x - 1, xp = x == p.nx - 1 ? 0 : x + 1;
    const int ym = y == 0 ? p.ny - 1 : y - 1, yp = y == p.ny - 1 ? 0 : y + 1;
    const int zm = z == 0 ? p.nz - 1 : z - 1, zp = z + V >= p.nz ? 0 : z + V;
    const Vt r = vec_load<T, V>(f + xp * sx + y * sy + z), l = vec_load<T, V>(f + xm * sx + y * sy + z);
    const Vt t = vec_load<T, V>(f + x * sx + yp * sy + z), b = vec_load<T, V>(f + x * sx + ym * sy + z);
    const T* row = f + x * sx + y * sy;
    const T zl = row[zm], zr = row[zp];
    Vt o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const T fr = k + 1 < V ? c.v[k + 1 < V ? k + 1 : 0] : zr;
      const T bk = k > 0 ? c.v[k > 0 ? k - 1 : 0] : zl;
      o.v[k] = (r.v[k] + l.v[k]) * p.ihx2 + (t.v[k] + b.v[k]) * p.ihy2 + (fr + bk) * p.ihz2 -
               T(2) * c.v[k] * p.ih2sum;
    }
    return o;
  }

  // group index g in [0, nx*ny*nz/V)
  EVX_HD static void run(const P& p, long long g) {
    const int gz = p.nz / V;
    const int z = (int)(g % gz) * V;
    const long long xy = g / gz;
    const int y = (int)(xy % p.ny), x = (int)(xy / p.ny);
    if (x >= p.nx) return;
    const long long n = (long long)p.nx * p.ny * p.nz;
    const long long o = ((long long)x * p.ny + y) * p.nz + z;
    const T* A = p.u;
    const T* B = p.u + n;
    const Vt a = vec_load<T, V>(A + o), b = vec_load<T, V>(B + o);
    const Vt la = lap(p, A, x, y, z, a), lb = lap(p, B, x, y, z, b);
    Vt it;
    if (p.inter) {
      it = vec_load<T, V>(p.inter + o);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) it.v[k] = a.v[k] * (b.v[k] * b.v[k]);
    }
    Vt oa, ob;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      oa.v[k] = p.DA * la.v[k] - it.v[k] + p.feed * (T(1) - a.v[k]);
      ob.v[k] = p.DB * lb.v[k] + it.v[k] - p.kill * b.v[k];
    }
    vec_store<T, V>(p.out + o, oa);
    vec_store<T, V>(p.out + n + o, ob);
  }
};

}  // namespace evx
