// Fused Cahn-Hilliard right-hand side, one pass over the field.
//
//   c^   = clip(c, 0, 1)
//   mu   = g(c^) - 2 eps lap7(c^)              g(c) = 18/eps c (1-c) (1-2c)  (or a caller-
//                                               supplied field "hom" = mu_hom(c^))
//   rhs  = D * sum_a  [ F_a(i+1/2) - F_a(i-1/2) ] / h_a
//   F_a(i+1/2) = cf (1 - cf) (mu_{i+e_a} - mu_i) / h_a ,  cf = (c^_i + c^_{i+e_a}) / 2
//
// Replaces reference evoxels/problem_definition.py:350-371 (CahnHilliard.rhs), which
// itself calls fd_stencils.py:20-42,62-75 and boundary_conditions.py:9-59.  The two
// ghost-padded copies the reference makes (of c^ and of mu) do not exist here: ghost
// values are produced by index arithmetic (periodic wrap) or synthesised from the
// adjacent inner cell (Neumann: ghost = inner, Dirichlet: ghost = 2 v - inner; the
// reference applies the SAME rule and values to mu, problem_definition.py:354).
//
// Layout: field [nx, ny, nz], z contiguous.  A thread block owns a (TY x TZ) tile of the
// y-z plane and marches along x through a chunk of planes.  Each thread owns V contiguous
// z-values ("a group") of one row of the tile *extended by one ring* (the ring is needed
// because rhs at the tile edge needs mu one cell outside the tile, and that mu needs c^
// two cells outside).  x-neighbours live in registers (rolling window of 4 planes of c^,
// 3 of mu); y/z-neighbours come from shared memory (2 plane slots of c^, 2 of mu).
// One barrier per plane.
//
// The footprint of rhs is the 25-point "diamond" |dx|+|dy|+|dz| <= 2, so tile corners of
// the rings are never consumed (they hold don't-care values).
#pragma once
#include "evx_hd.h"

namespace evx {

template <typename T>
struct ChParams {
  const T* c;        // [nx,ny,nz] raw (unclipped) concentration
  const T* hom;      // optional [nx,ny,nz]: mu_hom(clip(c)) evaluated by the caller; or null
  T* out;            // [nx,ny,nz]
  const T* halo_lo;  // optional [2,ny,nz]: raw planes x=-2,-1 (x-slab neighbour); null => BC rule
  const T* halo_hi;  // optional [2,ny,nz]: raw planes x=nx,nx+1
  int nx, ny, nz;
  int xchunk;        // planes per blockIdx.y
  T ihx, ihy, ihz;   // 1/h
  T ihx2, ihy2, ihz2, ih2sum;
  T pot_scale;       // 18/eps
  T two_eps;         // 2*eps
  T D;
  int bc_kind[3];
  T ghost_off[3][2];   // ghost = ghost_off + ghost_sgn * inner   (per axis, lo/hi side)
  T ghost_sgn[3];
};

template <typename T, int V, int TY, int G, bool HOM = false>
struct ChRhsProgram {
  static constexpr int TZ = G * V;
  static constexpr int RZ = V >= 2 ? 1 : 2; // ring groups per side: c^ is needed 2 cells out
  static constexpr int COLS = G + 2 * RZ;   // groups incl. the ring groups
  static constexpr int ROWS_MU = TY + 2;    // rows incl. one ring row each side
  static constexpr int ROWS_C = TY + 4;     // rows incl. two ring rows each side
  static constexpr int NPOS = ROWS_MU * COLS;
  static constexpr int NTHREADS = ((NPOS + 31) / 32) * 32;
  static constexpr int NEXTRA = 2 * G;      // loaders of the two outermost rows of c^
  static_assert(NEXTRA <= NTHREADS, "tile too flat");
  using Vt = Vec<T, V>;
  using P = ChParams<T>;

  struct Smem {
    Vt c[2][ROWS_C][COLS];
    Vt mu[2][ROWS_MU][COLS];
  };

  struct Regs {
    // position of this thread in the extended tile
    int r, g;                 // row in [-1, TY], group in [-RZ, G+RZ-1]
    bool has_pos;             // tid < NPOS
    bool interior;            // produces an output value
    bool gy_lo, gy_hi, gz_lo, gz_hi;   // neighbour in that direction is a non-periodic ghost
    long long off;            // element offset of this position inside a plane
    long long out_off;
    // extra loader (two outermost rows)
    bool has_extra;
    int er, eg;               // smem row / col of the extra element
    long long eoff;
    // rolling windows
    Vt cA, cB, cC, cD;        // c^ at planes p-2, p-1, p, p+1 (own position)
    Vt hC, hD;                // hom at planes p, p+1
    Vt mA, mB;                // mu at planes p-2, p-1
    Vt nxt, hnxt, enxt;       // raw prefetched plane p+2 (own / hom / extra rows)
    Vt sN, sS;                // saved y-neighbours of c^(p-1)
    T sL, sR;                 // saved z-neighbours of c^(p-1)
    int xa, xb;               // chunk [xa, xb)
  };

  EVX_HD static int slot(int q) { return (q + 4) & 1; }

  // pointer to plane q of the raw field (own slab, x-halo, or periodic image); null = ghost
  EVX_HD static const T* plane(const P& p, const T* base, int q, bool use_halo) {
    const long long ps = (long long)p.ny * p.nz;
    if (q >= 0 && q < p.nx) return base + q * ps;
    if (q < 0) {
      if (use_halo && p.halo_lo) return p.halo_lo + (q + 2) * ps;
      if (p.bc_kind[0] == BC_PERIODIC && !p.halo_lo) return base + wrap_index(q, p.nx) * ps;
      return nullptr;
    }
    if (use_halo && p.halo_hi) return p.halo_hi + (q - p.nx) * ps;
    if (p.bc_kind[0] == BC_PERIODIC && !p.halo_hi) return base + wrap_index(q, p.nx) * ps;
    return nullptr;
  }

  EVX_HD static Vt clipv(const Vt& a) {
    Vt r;
#pragma unroll
    for (int k = 0; k < V; ++k) r.v[k] = clip01(a.v[k]);
    return r;
  }
  EVX_HD static Vt ghostv(const Vt& inner, T off, T sgn) {
    Vt r;
#pragma unroll
    for (int k = 0; k < V; ++k) r.v[k] = off + sgn * inner.v[k];
    return r;
  }

  EVX_HD static Vt load_plane(const P& p, const T* base, int q, long long off, bool use_halo) {
    const T* pl = plane(p, base, q, use_halo);
    if (pl) return vec_load<T, V>(pl + off);
    return vec_splat<T, V>(T(0));
  }

  // ---- prologue: decode position, load planes xa-2, xa-1, xa, prefetch xa+1 ----------
  EVX_HD static void init(Regs& t, Smem& s, const P& p, int tid, int tile, int chunk) {
    const int tiles_z = (p.nz + TZ - 1) / TZ;
    const int y0 = (tile / tiles_z) * TY;
    const int z0 = (tile % tiles_z) * TZ;
    t.xa = chunk * p.xchunk;
    t.xb = t.xa + p.xchunk < p.nx ? t.xa + p.xchunk : p.nx;
    t.has_pos = tid < NPOS;
    t.r = tid / COLS - 1;
    t.g = tid % COLS - RZ;
    const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;
    {
      const int y = y0 + t.r, z = z0 + t.g * V;
      const int yi = per_y ? wrap_index(y, p.ny) : clamp_index(y, 0, p.ny - 1);
      const int zi = per_z ? wrap_index(z, p.nz) : clamp_index(z, 0, p.nz - V);
      t.off = (long long)yi * p.nz + zi;
      t.interior = t.has_pos && t.r >= 0 && t.r < TY && t.g >= 0 && t.g < G && y < p.ny &&
                   z + V <= p.nz;
      t.out_off = (long long)y * p.nz + z;
      t.gy_lo = !per_y && y == 0;
      t.gy_hi = !per_y && y == p.ny - 1;
      t.gz_lo = !per_z && z == 0;
      t.gz_hi = !per_z && z + V == p.nz;
    }
    t.has_extra = tid < NEXTRA;
    {
      const int side = tid / G;                 // 0: row -2, 1: row TY+1
      const int eg = tid % G;
      const int rr = side == 0 ? -2 : TY + 1;
      const int y = y0 + rr, z = z0 + eg * V;
      const int yi = per_y ? wrap_index(y, p.ny) : clamp_index(y, 0, p.ny - 1);
      const int zi = per_z ? wrap_index(z, p.nz) : clamp_index(z, 0, p.nz - V);
      t.er = rr + 2;
      t.eg = eg + RZ;
      t.eoff = (long long)yi * p.nz + zi;
    }
    const Vt zero = vec_splat<T, V>(T(0));
    t.cA = t.cB = t.cC = t.cD = zero;
    t.hC = t.hD = zero;
    t.mA = t.mB = zero;
    t.sN = t.sS = zero;
    t.sL = t.sR = T(0);
    t.nxt = t.hnxt = t.enxt = zero;
    if (t.has_pos) {
      t.cB = clipv(load_plane(p, p.c, t.xa - 2, t.off, true));
      t.cC = clipv(load_plane(p, p.c, t.xa - 1, t.off, true));
      t.cD = clipv(load_plane(p, p.c, t.xa, t.off, true));
      t.nxt = load_plane(p, p.c, t.xa + 1, t.off, true);
      if (HOM) {
        t.hC = load_plane(p, p.hom, t.xa - 1, t.off, false);
        t.hD = load_plane(p, p.hom, t.xa, t.off, false);
        t.hnxt = load_plane(p, p.hom, t.xa + 1, t.off, false);
      }
      s.c[slot(t.xa - 1)][t.r + 2][t.g + RZ] = t.cC;
      s.c[slot(t.xa)][t.r + 2][t.g + RZ] = t.cD;
    }
    if (t.has_extra) {
      s.c[slot(t.xa - 1)][t.er][t.eg] = clipv(load_plane(p, p.c, t.xa - 1, t.eoff, true));
      s.c[slot(t.xa)][t.er][t.eg] = clipv(load_plane(p, p.c, t.xa, t.eoff, true));
      t.enxt = load_plane(p, p.c, t.xa + 1, t.eoff, true);
    }
  }

  // ---- phase A of plane p: mu(p) -> smem, then rhs(x = p-1) -> global ------------------
  EVX_HD static void phase_a(Regs& t, Smem& s, const P& p, int pl) {
    if (!t.has_pos) return;
    const int row = t.r + 2, col = t.g + RZ;     // indices into s.c
    const bool real_p = plane(p, p.c, pl, true) != nullptr;
    const bool xm_ghost = plane(p, p.c, pl - 1, true) == nullptr;
    const bool xp_ghost = plane(p, p.c, pl + 1, true) == nullptr;

    // neighbours of c^(pl) in y and z (ghosts synthesised from the own value)
    Vt cN, cS;
    T cL, cR;
    {
      const int sl = slot(pl);
      cS = t.gy_lo ? ghostv(t.cC, p.ghost_off[1][0], p.ghost_sgn[1]) : s.c[sl][row - 1][col];
      cN = t.gy_hi ? ghostv(t.cC, p.ghost_off[1][1], p.ghost_sgn[1]) : s.c[sl][row + 1][col];
      cL = t.gz_lo ? p.ghost_off[2][0] + p.ghost_sgn[2] * t.cC.v[0] : s.c[sl][row][col - 1].v[V - 1];
      cR = t.gz_hi ? p.ghost_off[2][1] + p.ghost_sgn[2] * t.cC.v[V - 1] : s.c[sl][row][col + 1].v[0];
    }
    Vt mC = vec_splat<T, V>(T(0));
    if (real_p && t.g >= -1 && t.g <= G) {
      const Vt cXm = xm_ghost ? ghostv(t.cC, p.ghost_off[0][0], p.ghost_sgn[0]) : t.cB;
      const Vt cXp = xp_ghost ? ghostv(t.cC, p.ghost_off[0][1], p.ghost_sgn[0]) : t.cD;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const T c0 = t.cC.v[k];
        const T zl = k == 0 ? cL : t.cC.v[k - 1];
        const T zr = k == V - 1 ? cR : t.cC.v[k + 1];
        const T lap = (cXp.v[k] + cXm.v[k]) * p.ihx2 + (cN.v[k] + cS.v[k]) * p.ihy2 +
                      (zr + zl) * p.ihz2 - T(2) * c0 * p.ih2sum;
        const T hom = HOM ? t.hC.v[k] : p.pot_scale * c0 * (T(1) - c0) * (T(1) - T(2) * c0);
        mC.v[k] = hom - p.two_eps * lap;
      }
    }
    s.mu[pl & 1][t.r + 1][col] = mC;

    // rhs at plane x = pl-1: c^ window (cA,cB,cC) = (x-1,x,x+1), mu window (mA,mB,mC)
    const int x = pl - 1;
    if (t.interior && x >= t.xa && x < t.xb) {
      const int ms = x & 1;
      const int mrow = t.r + 1;
      const Vt mS = t.gy_lo ? ghostv(t.mB, p.ghost_off[1][0], p.ghost_sgn[1]) : s.mu[ms][mrow - 1][col];
      const Vt mN = t.gy_hi ? ghostv(t.mB, p.ghost_off[1][1], p.ghost_sgn[1]) : s.mu[ms][mrow + 1][col];
      const T mL = t.gz_lo ? p.ghost_off[2][0] + p.ghost_sgn[2] * t.mB.v[0] : s.mu[ms][mrow][col - 1].v[V - 1];
      const T mR = t.gz_hi ? p.ghost_off[2][1] + p.ghost_sgn[2] * t.mB.v[V - 1] : s.mu[ms][mrow][col + 1].v[0];
      const bool gxm = plane(p, p.c, x - 1, true) == nullptr;
      const bool gxp = xp_ghost_of(p, x);
      const Vt cXm = gxm ? ghostv(t.cB, p.ghost_off[0][0], p.ghost_sgn[0]) : t.cA;
      const Vt cXp = gxp ? ghostv(t.cB, p.ghost_off[0][1], p.ghost_sgn[0]) : t.cC;
      const Vt mXm = gxm ? ghostv(t.mB, p.ghost_off[0][0], p.ghost_sgn[0]) : t.mA;
      const Vt mXp = gxp ? ghostv(t.mB, p.ghost_off[0][1], p.ghost_sgn[0]) : mC;
      Vt o;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const T c0 = t.cB.v[k], m0 = t.mB.v[k];
        const T czl = k == 0 ? t.sL : t.cB.v[k - 1];
        const T czr = k == V - 1 ? t.sR : t.cB.v[k + 1];
        const T mzl = k == 0 ? mL : t.mB.v[k - 1];
        const T mzr = k == V - 1 ? mR : t.mB.v[k + 1];
        T div = face_div(c0, m0, cXm.v[k], mXm.v[k], cXp.v[k], mXp.v[k], p.ihx);
        div += face_div(c0, m0, t.sS.v[k], mS.v[k], t.sN.v[k], mN.v[k], p.ihy);
        div += face_div(c0, m0, czl, mzl, czr, mzr, p.ihz);
        o.v[k] = p.D * div;
      }
      vec_store<T, V>(p.out + (long long)x * p.ny * p.nz + t.out_off, o);
    }
    // roll the mu window and remember the y/z neighbours of c^(pl) for the next plane
    t.mA = t.mB;
    t.mB = mC;
    t.sN = cN;
    t.sS = cS;
    t.sL = cL;
    t.sR = cR;
  }

  EVX_HD static bool xp_ghost_of(const P& p, int x) { return plane(p, p.c, x + 1, true) == nullptr; }

  // divergence contribution of one axis: ( F(+1/2) - F(-1/2) ) / h
  EVX_HD static T face_div(T c0, T m0, T cm, T mm, T cp, T mp, T ih) {
    const T fp = T(0.5) * (cp + c0);
    const T fm = T(0.5) * (c0 + cm);
    const T Fp = fp * (T(1) - fp) * ((mp - m0) * ih);
    const T Fm = fm * (T(1) - fm) * ((m0 - mm) * ih);
    return (Fp - Fm) * ih;
  }

  // ---- phase B of plane p (after the barrier): roll c^ window, publish plane p+2 -------
  EVX_HD static void phase_b(Regs& t, Smem& s, const P& p, int pl) {
    const bool more = pl + 3 <= t.xb + 1;     // plane pl+3 is still needed as a centre value
    if (t.has_pos) {
      t.cA = t.cB;
      t.cB = t.cC;
      t.cC = t.cD;
      t.cD = clipv(t.nxt);
      if (HOM) {
        t.hC = t.hD;
        t.hD = t.hnxt;
      }
      s.c[slot(pl + 2)][t.r + 2][t.g + RZ] = t.cD;
      if (more) {
        t.nxt = load_plane(p, p.c, pl + 3, t.off, true);
        if (HOM) t.hnxt = load_plane(p, p.hom, pl + 3, t.off, false);
      }
    }
    if (t.has_extra) {
      s.c[slot(pl + 2)][t.er][t.eg] = clipv(t.enxt);
      if (more) t.enxt = load_plane(p, p.c, pl + 3, t.eoff, true);
    }
  }
};

}  // namespace evx
