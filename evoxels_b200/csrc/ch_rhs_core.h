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
// two cells outside).  x-neighbours live in registers (rolling window of 3 planes of c^,
// mu of the previous plane and the carried x-face term); y/z-neighbours come from shared memory (2 plane slots of c^, 2 of mu).
// One barrier per plane.
//
// Arithmetic: the kernel is FP32-issue bound on B200 (about 45 flop/voxel against
// 8 B/voxel), so the face terms are written as
//      4 cf (1-cf) (mu_b - mu_a) = s (2 - s) (mu_b - mu_a),   s = c_a + c_b
// with all metric factors folded into three constants D/(4 h_a^2); the x-face term is
// carried to the next plane in a register and the z-face terms are shared inside a
// thread's group, so each face is evaluated once per thread.  This reassociates the
// reference's expression; the result differs by fp32 rounding only (tests: rel-L2 <= 5e-6).
//
// The footprint of rhs is the 25-point "diamond" |dx|+|dy|+|dz| <= 2, so tile corners of
// the rings are never consumed (they hold don't-care values).
#pragma once
#include "evx_hd.h"
#include "lanes.h"

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
  T pot_scale;       // 18/eps
  // folded constants
  T lx, ly, lz, l0;  // mu = g + lx (c_x+ + c_x-) + ly (..) + lz (..) + l0 c ; l_a = -2 eps / h_a^2
  T fx, fy, fz;      // rhs = fx dFx + fy dFy + fz dFz ; f_a = D / (4 h_a^2)
  int bc_kind[3];
  T ghost_off[3][2];   // ghost = ghost_off + ghost_sgn * inner   (per axis, lo/hi side)
  T ghost_sgn[3];
};

template <typename T, int V, int TY, int G, bool HOM = false, bool GHOSTS = true>
struct ChRhsProgram {
  static constexpr int TZ = G * V;
  static constexpr int RZ = V >= 2 ? 1 : 2; // ring groups per side: c^ is needed 2 cells out
  static constexpr int COLS = G + 2 * RZ;   // groups incl. the ring groups
  static constexpr int ROWS_MU = TY + 2;    // rows incl. one ring row each side
  static constexpr int ROWS_C = TY + 4;     // rows incl. two ring rows each side
  // thread -> position: interior positions first (whole warps of interior threads run the
  // full body), then the two ring rows (which also load the two outermost rows of c^), then
  // the ring columns; ring warps skip the rhs part as a warp.
  static constexpr int N_INT = TY * G;
  static constexpr int N_RROW = 2 * G;
  static constexpr int N_RCOL = 2 * RZ * (TY + 2);   // incl. the 4 ring corners (c^ only)
  static constexpr int NPOS = N_INT + N_RROW + N_RCOL;
  static constexpr int NTHREADS = ((NPOS + 31) / 32) * 32;
  using Vt = Vec<T, V>;
  using P = ChParams<T>;

  struct Smem {
    Vt c[2][ROWS_C][COLS];
    Vt mu[2][ROWS_MU][COLS];
    // raw planes in flight (cp.async): the own element, the extra outermost-row element of
    // the ring-row threads, and the hom field; cell = parity of the plane that consumes it
    Vt stage[2][NTHREADS];
    Vt stage_e[2][N_RROW];
    Vt stage_h[2][HOM ? NTHREADS : 1];
  };

  struct Regs {
    int row, col, tid;        // smem indices of the own position (row in s.c numbering)
    bool has_pos, want_mu, interior, has_extra;
    bool gy_lo, gy_hi, gz_lo, gz_hi;   // neighbour in that direction is a non-periodic ghost
    bool xlo_ghost, xhi_ghost;         // planes below 0 / above nx-1 are non-periodic ghosts
    long long ps;             // plane stride ny*nz
    const T* pn;              // own element of the next plane to prefetch (null: ghost plane)
    const T* ph;              // same for the hom field
    const T* pe;              // same for the extra (outermost-row) element
    long long off, eoff;      // element offsets inside a plane (own / extra)
    int qn;                   // index of the plane pn/ph/pe point into
    T* po;                    // own output element of the next plane to write
    int er, ec;               // smem row / col of the extra element
    // Rolling windows are indexed with COMPILE-TIME rotation counters (ROT mod 3 for c^,
    // PAR mod 2 for everything else) and the plane loop is unrolled six-fold, so the windows
    // rotate by renaming instead of by ~60 register moves per plane.
    Vt c[3];                  // c^ at planes p-1, p, p+1 live in c[ROT], c[ROT+1], c[ROT+2] (mod 3)
    Vt hC, hD;                // hom at planes p, p+1
    Vt m[2];                  // mu(p-1) = m[PAR^1]; mu(p) is written to m[PAR]
    Vt fx[2];                 // x-face term (p-2,p-1) = fx[PAR^1]; (p-1,p) is written to fx[PAR]
    Vt sN[2], sS[2];          // y-neighbours of c^(p-1) = s*[PAR^1]; those of c^(p) go to s*[PAR]
    T sL[2], sR[2];           // z-neighbours, same convention
    int xa, xb;               // chunk [xa, xb)
  };


  // pointer to plane q of a field (own slab, x-halo, or periodic image); null = ghost plane
  EVX_HD static const T* plane(const P& p, const T* base, int q, bool use_halo) {
    const long long ps = (long long)p.ny * p.nz;
    if (q < 0 || q >= p.nx) {        // rare: chunk ends only
      if (q < 0) {
        if (use_halo && p.halo_lo) return p.halo_lo + (q + 2) * ps;
        if (p.bc_kind[0] != BC_PERIODIC || p.halo_lo) return nullptr;
      } else {
        if (use_halo && p.halo_hi) return p.halo_hi + (q - p.nx) * ps;
        if (p.bc_kind[0] != BC_PERIODIC || p.halo_hi) return nullptr;
      }
      q = wrap_index(q, p.nx);
    }
    return base + q * ps;
  }

  EVX_HD static bool is_ghost_plane(const Regs& t, const P& p, int q) {
    return GHOSTS && ((q < 0 && t.xlo_ghost) || (q >= p.nx && t.xhi_ghost));
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

  static constexpr int LW = (V % 2 == 0) ? 2 : 1;   // lanes: pairs of z values when possible
  using Ln = AcLane<T, LW>;
  EVX_HD static Ln face_l(Ln ca, Ln cb, Ln ma, Ln mb) {
    const Ln s = Ln::add(ca, cb);
    return Ln::mul(Ln::mul(s, Ln::rsubs(T(2), s)), Ln::sub(mb, ma));
  }

  // 4 cf (1 - cf) (mb - ma) with cf = (ca + cb)/2
  EVX_HD static T face(T ca, T cb, T ma, T mb) {
    const T s = ca + cb;
    return s * (T(2) - s) * (mb - ma);
  }

  // ---- prologue: decode position, load planes xa-2, xa-1, xa, prefetch xa+1 ----------
  EVX_HD static void init(Regs& t, Smem& s, const P& p, int tid, int tile, int chunk) {
    const int tiles_z = (p.nz + TZ - 1) / TZ;
    const int y0 = (tile / tiles_z) * TY;
    const int z0 = (tile % tiles_z) * TZ;
    t.xa = chunk * p.xchunk;
    t.xb = t.xa + p.xchunk < p.nx ? t.xa + p.xchunk : p.nx;
    t.has_pos = tid < NPOS;
    int r, g;                            // row in [-1, TY], group in [-RZ, G+RZ-1]
    if (tid < N_INT) {
      r = tid / G;
      g = tid % G;
    } else if (tid < N_INT + N_RROW) {
      const int j = tid - N_INT;
      r = j < G ? -1 : TY;
      g = j % G;
    } else {
      const int j = tid - N_INT - N_RROW;
      const int side = j / (RZ * (TY + 2)), w = j % (RZ * (TY + 2));
      r = w % (TY + 2) - 1;
      g = side == 0 ? -1 - w / (TY + 2) : G + w / (TY + 2);
    }
    t.row = r + 2;
    t.col = g + RZ;
    const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;
    t.xlo_ghost = p.bc_kind[0] != BC_PERIODIC && !p.halo_lo;
    t.xhi_ghost = p.bc_kind[0] != BC_PERIODIC && !p.halo_hi;
    {
      const int y = y0 + r, z = z0 + g * V;
      const int yi = per_y ? wrap_index(y, p.ny) : clamp_index(y, 0, p.ny - 1);
      const int zi = per_z ? wrap_index(z, p.nz) : clamp_index(z, 0, p.nz - V);
      t.off = (long long)yi * p.nz + zi;
      t.want_mu = t.has_pos && g >= -1 && g <= G;
      t.interior = t.has_pos && r >= 0 && r < TY && g >= 0 && g < G && y < p.ny && z + V <= p.nz;
      t.po = p.out + (long long)t.xa * p.ny * p.nz + (long long)y * p.nz + z;
      t.gy_lo = GHOSTS && !per_y && y == 0;
      t.gy_hi = GHOSTS && !per_y && y == p.ny - 1;
      t.gz_lo = GHOSTS && !per_z && z == 0;
      t.gz_hi = GHOSTS && !per_z && z + V == p.nz;
    }
    t.has_extra = tid >= N_INT && tid < N_INT + N_RROW;   // ring-row threads
    {
      const int j = tid - N_INT;
      const int side = j < G ? 0 : 1;           // 0: row -2, 1: row TY+1
      const int eg = (j < 0 ? 0 : j) % G;
      const int rr = side == 0 ? -2 : TY + 1;
      const int y = y0 + rr, z = z0 + eg * V;
      const int yi = per_y ? wrap_index(y, p.ny) : clamp_index(y, 0, p.ny - 1);
      const int zi = per_z ? wrap_index(z, p.nz) : clamp_index(z, 0, p.nz - V);
      t.er = rr + 2;
      t.ec = eg + RZ;
      t.eoff = (long long)yi * p.nz + zi;
    }
    const Vt zero = vec_splat<T, V>(T(0));
    t.c[0] = t.c[1] = t.c[2] = zero;
    t.hC = t.hD = zero;
    t.m[0] = t.m[1] = zero;
    t.fx[0] = t.fx[1] = zero;
    t.sN[0] = t.sN[1] = t.sS[0] = t.sS[1] = zero;
    t.sL[0] = t.sL[1] = t.sR[0] = t.sR[1] = T(0);
    t.tid = tid;
    if (t.has_pos) {
      t.c[0] = clipv(load_plane(p, p.c, t.xa - 2, t.off, true));     // first plane: ROT = 0
      t.c[1] = clipv(load_plane(p, p.c, t.xa - 1, t.off, true));
      t.c[2] = clipv(load_plane(p, p.c, t.xa, t.off, true));
      if (HOM) {
        t.hC = load_plane(p, p.hom, t.xa - 1, t.off, false);
        t.hD = load_plane(p, p.hom, t.xa, t.off, false);
      }
      s.c[0][t.row][t.col] = t.c[1];   // slot = (plane - (xa-1)) & 1
      s.c[1][t.row][t.col] = t.c[2];
    }
    if (t.has_extra) {
      s.c[0][t.er][t.ec] = clipv(load_plane(p, p.c, t.xa - 1, t.eoff, true));
      s.c[1][t.er][t.ec] = clipv(load_plane(p, p.c, t.xa, t.eoff, true));
    }
    t.ps = (long long)p.ny * p.nz;
    t.qn = t.xa + 1;
    resolve_next(t, p);
    fetch_next(t, s, p, 0);            // plane xa+1 -> cell 0 (read by the first phase B)
    advance_next(t, p);
  }

  // start the copies of plane t.qn into staging cell `cell` (ghost planes: nothing to copy,
  // the cell keeps a finite don't-care value)
  EVX_HD static void fetch_next(Regs& t, Smem& s, const P& p, int cell) {
    constexpr int B = (int)sizeof(Vt);
    // fully periodic instantiation: every plane of c exists (own slab, x halo or periodic
    // image), so the pointers are never null and the ghost-plane branches compile away
    constexpr bool NEVER_NULL = !GHOSTS;
    if (t.has_pos) {
      if (NEVER_NULL || t.pn) async_copy_bytes<B>(&s.stage[cell][t.tid], t.pn);
      else s.stage[cell][t.tid] = vec_splat<T, V>(T(0));
      if (HOM) {
        if (t.ph) async_copy_bytes<B>(&s.stage_h[cell][t.tid], t.ph);
        else s.stage_h[cell][HOM ? t.tid : 0] = vec_splat<T, V>(T(0));
      }
    }
    if (t.has_extra) {
      if (NEVER_NULL || t.pe) async_copy_bytes<B>(&s.stage_e[cell][t.tid - N_INT], t.pe);
      else s.stage_e[cell][t.tid - N_INT] = vec_splat<T, V>(T(0));
    }
    async_copy_commit();
  }

  // pointers into plane t.qn (general path: slab ends, halos, periodic images, ghosts)
  EVX_HD static void resolve_next(Regs& t, const P& p) {
    const T* b = plane(p, p.c, t.qn, true);
    t.pn = b ? b + t.off : nullptr;
    t.pe = b ? b + t.eoff : nullptr;
    if (HOM) {
      const T* h = plane(p, p.hom, t.qn, false);
      t.ph = h ? h + t.off : nullptr;
    } else {
      t.ph = nullptr;
    }
  }
  EVX_HD static void advance_next(Regs& t, const P& p) {
    ++t.qn;
    if (t.qn >= 1 && t.qn < p.nx) {      // both planes inside the slab: plain stride
      t.pn += t.ps;
      t.pe += t.ps;
      if (HOM) t.ph += t.ps;
    } else {
      resolve_next(t, p);
    }
  }

  // ---- phase A of plane p: mu(p) -> smem, then rhs(x = p-1) -> global ------------------
  // PAR = (pl - (xa-1)) & 1 is the shared-memory slot of plane pl (static so that all
  // shared-memory offsets are immediates)
  // ROT = (pl - (xa-1)) mod 3 names the registers of the c^ window
  template <int PAR, int ROT>
  EVX_HD static void phase_a(Regs& t, Smem& s, const P& p, int pl) {
    if (!t.has_pos) return;
    const Vt& cB_ = t.c[ROT % 3];
    const Vt& cC_ = t.c[(ROT + 1) % 3];
    const Vt& cD_ = t.c[(ROT + 2) % 3];
    const Vt& mB_ = t.m[PAR ^ 1];
    const Vt& fxm_ = t.fx[PAR ^ 1];
    const Vt& sN_ = t.sN[PAR ^ 1];
    const Vt& sS_ = t.sS[PAR ^ 1];
    const T sL_ = t.sL[PAR ^ 1], sR_ = t.sR[PAR ^ 1];
    const int row = t.row, col = t.col;
    const T oy0 = p.ghost_off[1][0], oy1 = p.ghost_off[1][1], sy = p.ghost_sgn[1];
    const T oz0 = p.ghost_off[2][0], oz1 = p.ghost_off[2][1], sz = p.ghost_sgn[2];
    const T ox0 = p.ghost_off[0][0], ox1 = p.ghost_off[0][1], sx = p.ghost_sgn[0];

    // neighbours of c^(pl) in y and z (ghosts synthesised from the own value)
    Vt cN, cS;
    T cL, cR;
    {
      constexpr int sl = PAR;
      cS = s.c[sl][row - 1][col];
      cN = s.c[sl][row + 1][col];
      cL = s.c[sl][row][col - 1].v[V - 1];
      cR = s.c[sl][row][col + 1].v[0];
      if (GHOSTS) {
        if (t.gy_lo) cS = ghostv(cC_, oy0, sy);
        if (t.gy_hi) cN = ghostv(cC_, oy1, sy);
        if (t.gz_lo) cL = oz0 + sz * cC_.v[0];
        if (t.gz_hi) cR = oz1 + sz * cC_.v[V - 1];
      }
    }
    Vt mC = vec_splat<T, V>(T(0));
    if (t.want_mu && !is_ghost_plane(t, p, pl)) {
      Vt cXm = cB_, cXp = cD_;
      if (GHOSTS) {
        if (is_ghost_plane(t, p, pl - 1)) cXm = ghostv(cC_, ox0, sx);
        if (is_ghost_plane(t, p, pl + 1)) cXp = ghostv(cC_, ox1, sx);
      }
      // window of c^(pl) along z: [cL, c0..c(V-1), cR]
      T cw[V + 2];
      cw[0] = cL;
#pragma unroll
      for (int k = 0; k < V; ++k) cw[k + 1] = cC_.v[k];
      cw[V + 1] = cR;
#pragma unroll
      for (int k = 0; k < V; k += LW) {
        const Ln c0 = Ln::load(cw, k + 1);
        const Ln sx = Ln::add(Ln::load(cXp.v, k), Ln::load(cXm.v, k));
        const Ln sy = Ln::add(Ln::load(cN.v, k), Ln::load(cS.v, k));
        const Ln sz = Ln::add(Ln::load(cw, k + 2), Ln::load(cw, k));
        Ln hom;
        if (HOM) {
          hom = Ln::load(t.hC.v, k);
        } else {
          const Ln cc = Ln::mul(c0, Ln::rsubs(T(1), c0));
          hom = Ln::muls(Ln::mul(cc, Ln::rsubs(T(1), Ln::add(c0, c0))), p.pot_scale);
        }
        Ln::fmas(sx, p.lx, Ln::fmas(sy, p.ly, Ln::fmas(sz, p.lz, Ln::fmas(c0, p.l0, hom))))
            .store(mC.v, k);
      }
    }
    s.mu[PAR][row - 1][col] = mC;

    // x-face term between planes pl-1 and pl (carried to the next plane as "minus" face)
    Vt fxp;
#pragma unroll
    for (int k = 0; k < V; k += LW)
      face_l(Ln::load(cB_.v, k), Ln::load(cC_.v, k), Ln::load(mB_.v, k), Ln::load(mC.v, k))
          .store(fxp.v, k);

    // rhs at plane x = pl-1: c^(x) = cB, mu(x) = mB; the x-faces are fxm (carried) and fxp
    const int x = pl - 1;
    if (t.interior && x >= t.xa) {
      constexpr int ms = PAR ^ 1;
      const int mrow = row - 1;
      Vt mS = s.mu[ms][mrow - 1][col];
      Vt mN = s.mu[ms][mrow + 1][col];
      T mL = s.mu[ms][mrow][col - 1].v[V - 1];
      T mR = s.mu[ms][mrow][col + 1].v[0];
      Vt fxm = fxm_;
      if (GHOSTS) {
        if (t.gy_lo) mS = ghostv(mB_, oy0, sy);
        if (t.gy_hi) mN = ghostv(mB_, oy1, sy);
        if (t.gz_lo) mL = oz0 + sz * mB_.v[0];
        if (t.gz_hi) mR = oz1 + sz * mB_.v[V - 1];
        if (is_ghost_plane(t, p, x - 1)) {
#pragma unroll
          for (int k = 0; k < V; ++k)
            fxm.v[k] = face(ox0 + sx * cB_.v[k], cB_.v[k], ox0 + sx * mB_.v[k], mB_.v[k]);
        }
        if (is_ghost_plane(t, p, x + 1)) {
#pragma unroll
          for (int k = 0; k < V; ++k)
            fxp.v[k] = face(cB_.v[k], ox1 + sx * cB_.v[k], mB_.v[k], ox1 + sx * mB_.v[k]);
        }
      }
      // z windows of c^(x) and mu(x): [left, 0..V-1, right]
      T cw[V + 2], mw[V + 2];
      cw[0] = sL_; mw[0] = mL;
#pragma unroll
      for (int k = 0; k < V; ++k) { cw[k + 1] = cB_.v[k]; mw[k + 1] = mB_.v[k]; }
      cw[V + 1] = sR_; mw[V + 1] = mR;
      // z-face terms: fz[k] is the face between window elements k and k+1
      T fz[V + 2];
#pragma unroll
      for (int k = 0; k + LW <= V + 1; k += LW)
        face_l(Ln::load(cw, k), Ln::load(cw, k + 1), Ln::load(mw, k), Ln::load(mw, k + 1)).store(fz, k);
      if ((V + 1) % LW) fz[V] = face(cw[V], cw[V + 1], mw[V], mw[V + 1]);
      Vt o;
#pragma unroll
      for (int k = 0; k < V; k += LW) {
        const Ln c0 = Ln::load(cB_.v, k), m0 = Ln::load(mB_.v, k);
        const Ln fyp = face_l(c0, Ln::load(sN_.v, k), m0, Ln::load(mN.v, k));
        const Ln fym = face_l(Ln::load(sS_.v, k), c0, Ln::load(mS.v, k), m0);
        const Ln dx = Ln::sub(Ln::load(fxp.v, k), Ln::load(fxm.v, k));
        const Ln dz = Ln::sub(Ln::load(fz, k + 1), Ln::load(fz, k));
        Ln::fmas(dx, p.fx, Ln::fmas(Ln::sub(fyp, fym), p.fy, Ln::muls(dz, p.fz))).store(o.v, k);
      }
      vec_store<T, V>(t.po, o);
      t.po += t.ps;
    }
    // hand the mu window and the y/z neighbours of c^(pl) to the next plane (no moves:
    // the next plane reads them under the flipped parity)
    t.fx[PAR] = fxp;
    t.m[PAR] = mC;
    t.sN[PAR] = cN;
    t.sS[PAR] = cS;
    t.sL[PAR] = cL;
    t.sR[PAR] = cR;
  }

  // ---- phase B of plane p (after the barrier): roll c^ window, publish plane p+2 -------
  template <int PAR, int ROT>
  EVX_HD static void phase_b(Regs& t, Smem& s, const P& p, int pl) {
    const bool more = t.qn <= t.xb + 1;       // plane qn = pl+3 is still needed as a centre value
    async_copy_wait_all();                    // own copies of plane pl+2 have landed
    if (t.has_pos) {
      t.c[ROT % 3] = clipv(s.stage[PAR][t.tid]);     // plane pl+2 replaces plane pl-1
      if (HOM) {
        t.hC = t.hD;
        t.hD = s.stage_h[PAR][HOM ? t.tid : 0];
      }
      s.c[PAR][t.row][t.col] = t.c[ROT % 3];
    }
    if (t.has_extra) s.c[PAR][t.er][t.ec] = clipv(s.stage_e[PAR][t.tid - N_INT]);
    if (more) {
      fetch_next(t, s, p, PAR ^ 1);
      advance_next(t, p);
    }
  }
};

}  // namespace evx
