// Fused two-phase Allen-Cahn right-hand side + explicit-stage update, one pass.
//
//   phi^ = clip(phi, 0, 1)
//   k    = M * [ gab * ( curv * lap7(phi^) + (1-curv) * nlap19(phi^) - g(phi^)/2/eps )
//                + 3/eps * phi^ (1-phi^) * force ]
//   nlap19 = [ sum_a g_a^2 d_aa f + sum_{a<b} 1/2 g_a g_b d_ab f ] / |g|^2 ,  |g|^2<=1e-7 -> 1
//
// Replaces reference evoxels/problem_definition.py:421-447 (TwoPhaseAllenCahn.rhs),
// fd_stencils.py:44-103 (centred gradients, laplace, normal_laplace) and the generic ghost
// rules of boundary_conditions.py:33-59.  Ghost values (including the edge ghosts the
// 19-point stencil touches) are produced on the fly: index map per axis (periodic wrap /
// clamp to the inner cell) and the affine rule ghost = off + sgn * inner composed in the
// reference's axis order x, then y, then z.
//
// A thread owns V contiguous z-values of one (y) row and marches along x keeping a
// 3 planes x 3 rows x (V+2) window in registers; neighbouring rows/columns are re-read
// through L1 (no shared memory, no barrier).
#pragma once
#include "evx_hd.h"
#include "lanes.h"

namespace evx {

template <typename T>
struct AcParams {
  const T* phi;      // [nx,ny,nz] raw
  const T* pot;      // optional potential(clip(phi)) field
  T* k_out;          // optional: k
  const T* base;     // for y_out
  T* y_out;          // optional: base + alpha*k
  const T* acc_in;   // optional
  T* acc_out;        // optional: acc_in + beta*k
  T alpha, beta;
  const T* halo_lo;  // optional [1,ny,nz] raw plane x=-1
  const T* halo_hi;  // optional [1,ny,nz] raw plane x=nx
  int nx, ny, nz, xchunk;
  T ihx, ihy, ihz, ihx2, ihy2, ihz2, ih2sum;
  T pot_scale, eps, gab, M, force, curv, omc, three_over_eps;
  T hihx, hihy, hihz;        // 0.5 / h
  T hxy, hxz, hyz;           // 0.5 / (h_a h_b)
  T neg_inv_2eps, force3;    // -1/(2 eps),  3/eps * force
  int bc_kind[3];
  T ghost_off[3][2];
  T ghost_sgn[3];
};

template <typename T, int V, int TY, int G>
struct AcProgram {
  static constexpr int TZ = G * V;
  static constexpr int NTHREADS = TY * G;
  static constexpr int W = V + 2;
  using P = AcParams<T>;
  using Vt = Vec<T, V>;

  struct PlaneRef {
    const T* ptr;
    int side;   // -1: no x ghost rule, 0/1: apply the lo/hi rule of axis 0
  };

  EVX_HD static PlaneRef plane(const P& p, int q) {
    const long long ps = (long long)p.ny * p.nz;
    PlaneRef r;
    r.side = -1;
    if (q >= 0 && q < p.nx) { r.ptr = p.phi + q * ps; return r; }
    if (q < 0) {
      if (p.halo_lo) { r.ptr = p.halo_lo; return r; }
      if (p.bc_kind[0] == BC_PERIODIC) { r.ptr = p.phi + wrap_index(q, p.nx) * ps; return r; }
      r.ptr = p.phi; r.side = 0; return r;
    }
    if (p.halo_hi) { r.ptr = p.halo_hi; return r; }
    if (p.bc_kind[0] == BC_PERIODIC) { r.ptr = p.phi + wrap_index(q, p.nx) * ps; return r; }
    r.ptr = p.phi + (long long)(p.nx - 1) * ps; r.side = 1; return r;
  }

  struct Pos {
    long long roff[3];   // element offset of rows y-1, y, y+1 (group start) in a plane
    long long loff[3];   // same rows, element z0-1 (mapped)
    long long hoff[3];   // same rows, element z0+V (mapped)
    int yside[3];        // -1 or lo/hi side of the y rule to apply to that row
    int zl_side, zr_side;
  };

  EVX_HD static T rule(const P& p, int axis, int side, T v) {
    return side < 0 ? v : p.ghost_off[axis][side] + p.ghost_sgn[axis] * v;
  }

  // load the (V+2)-wide window of one row of one plane with all ghost rules applied
  EVX_HD static void load_row(const P& p, const PlaneRef& pl, const Pos& ps, int j, T* w) {
    const Vt c = vec_load<T, V>(pl.ptr + ps.roff[j]);
    T l = pl.ptr[ps.loff[j]];
    T h = pl.ptr[ps.hoff[j]];
#pragma unroll
    for (int k = 0; k < V; ++k)
      w[k + 1] = rule(p, 1, ps.yside[j], rule(p, 0, pl.side, clip01(c.v[k])));
    l = rule(p, 1, ps.yside[j], rule(p, 0, pl.side, clip01(l)));
    h = rule(p, 1, ps.yside[j], rule(p, 0, pl.side, clip01(h)));
    w[0] = rule(p, 2, ps.zl_side, l);
    w[V + 1] = rule(p, 2, ps.zr_side, h);
  }

  // ---- lane arithmetic: a "lane" is one value (scalar path) or an aligned/shifted pair of
  // neighbouring z values (packed path; FADD2/FMUL2/FFMA2 on sm_100a for float) -------------
  // one plane of the rolling window: 3 rows x (V+2) values
  struct Plane {
    T w[3][W];
  };

  template <typename L>
  EVX_HD static L at(const T* w, int k) { return L::load(w, k); }

  // the right-hand side for the lanes starting at window index k (element e = k-1)
  template <typename L>
  EVX_HD static L rhs_lanes(const P& p, const Plane& fm, const Plane& fc, const Plane& fp, int k,
                            const T* potv, int e) {
    const L C = at<L>(fc.w[1], k);
    const L R = at<L>(fp.w[1], k), Lf = at<L>(fm.w[1], k);
    const L Tp = at<L>(fc.w[2], k), B = at<L>(fc.w[0], k);
    const L F = at<L>(fc.w[1], k + 1), Bk = at<L>(fc.w[1], k - 1);
    const L two_c = L::add(C, C);
    const L dxx = L::sub(L::add(R, Lf), two_c), dyy = L::sub(L::add(Tp, B), two_c),
            dzz = L::sub(L::add(F, Bk), two_c);
    // lap = dxx/hx^2 + dyy/hy^2 + dzz/hz^2
    const L lap = L::fmas(dxx, p.ihx2, L::fmas(dyy, p.ihy2, L::muls(dzz, p.ihz2)));
    const L gx = L::muls(L::sub(R, Lf), p.hihx), gy = L::muls(L::sub(Tp, B), p.hihy),
            gz = L::muls(L::sub(F, Bk), p.hihz);
    const L mxy = L::sub(L::add(at<L>(fp.w[2], k), at<L>(fm.w[0], k)),
                         L::add(at<L>(fm.w[2], k), at<L>(fp.w[0], k)));
    const L mxz = L::sub(L::add(at<L>(fp.w[1], k + 1), at<L>(fm.w[1], k - 1)),
                         L::add(at<L>(fm.w[1], k + 1), at<L>(fp.w[1], k - 1)));
    const L myz = L::sub(L::add(at<L>(fc.w[2], k + 1), at<L>(fc.w[0], k - 1)),
                         L::add(at<L>(fc.w[0], k + 1), at<L>(fc.w[2], k - 1)));
    const L gxx = L::mul(gx, gx), gyy = L::mul(gy, gy), gzz = L::mul(gz, gz);
    // num = sum_a g_a^2 d_aa / h_a^2 + sum_{a<b} 1/2 g_a g_b d_ab / (h_a h_b)
    L num = L::mul(gxx, L::muls(dxx, p.ihx2));
    num = L::fma(gyy, L::muls(dyy, p.ihy2), num);
    num = L::fma(gzz, L::muls(dzz, p.ihz2), num);
    num = L::fma(L::mul(gx, gy), L::muls(mxy, p.hxy), num);
    num = L::fma(L::mul(gx, gz), L::muls(mxz, p.hxz), num);
    num = L::fma(L::mul(gy, gz), L::muls(myz, p.hyz), num);
    const L n2 = L::add(L::add(gxx, gyy), gzz);
    const L nl = L::guarded_div(num, n2);
    const L one_m_c = L::rsubs(T(1), C);
    const L cc = L::mul(C, one_m_c);                                  // c (1 - c)
    L pot;
    if (p.pot) pot = L::load(potv, e);
    else pot = L::muls(L::mul(cc, L::rsubs(T(1), two_c)), p.pot_scale);
    // df = gab (curv lap + (1-curv) nl - pot/(2 eps)) + 3/eps c(1-c) force
    L inner = L::fmas(lap, p.curv, L::fmas(nl, p.omc, L::muls(pot, p.neg_inv_2eps)));
    const L df = L::fmas(inner, p.gab, L::muls(cc, p.force3));
    return L::muls(df, p.M);
  }

  EVX_HD static PlaneRef plane_of(const P& p, int q) { return plane(p, q); }

  EVX_HD static void load_plane(const P& p, const PlaneRef& pl, const Pos& ps, bool plain, Plane& f) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (plain && pl.side < 0) {        // no ghost rule touches this thread: clip only
        const Vt c = vec_load<T, V>(pl.ptr + ps.roff[j]);
#pragma unroll
        for (int k = 0; k < V; ++k) f.w[j][k + 1] = clip01(c.v[k]);
        f.w[j][0] = clip01(pl.ptr[ps.loff[j]]);
        f.w[j][V + 1] = clip01(pl.ptr[ps.hoff[j]]);
      } else {
        load_row(p, pl, ps, j, f.w[j]);
      }
    }
  }

  // one output plane x: window planes (fm, fc, fp) = (x-1, x, x+1)
  EVX_HD static void emit(const P& p, const Plane& fm, const Plane& fc, const Plane& fp, long long o) {
    Vt potv = vec_splat<T, V>(T(0)), basev = potv, accv = potv;
    if (p.pot) potv = vec_load<T, V>(p.pot + o);
    if (p.y_out) basev = vec_load<T, V>(p.base + o);
    if (p.acc_out && p.acc_in) accv = vec_load<T, V>(p.acc_in + o);
    Vt kv;
    constexpr int LW = (V % 2 == 0) ? 2 : 1;
    using L = AcLane<T, LW>;
#pragma unroll
    for (int e = 0; e < V; e += LW) rhs_lanes<L>(p, fm, fc, fp, e + 1, potv.v, e).store(kv.v, e);
    if (p.k_out) vec_store<T, V>(p.k_out + o, kv);
    if (p.y_out) {
      Vt yv;
#pragma unroll
      for (int e = 0; e < V; ++e) yv.v[e] = basev.v[e] + p.alpha * kv.v[e];
      vec_store<T, V>(p.y_out + o, yv);
    }
    if (p.acc_out) {
      Vt av;
#pragma unroll
      for (int e = 0; e < V; ++e) av.v[e] = accv.v[e] + p.beta * kv.v[e];
      vec_store<T, V>(p.acc_out + o, av);
    }
  }

  EVX_HD static void run(const P& p, int tid, int tile, int chunk) {
    const int tiles_z = (p.nz + TZ - 1) / TZ;
    const int y = (tile / tiles_z) * TY + tid / G;
    const int z = (tile % tiles_z) * TZ + (tid % G) * V;
    if (y >= p.ny || z + V > p.nz) return;
    const int xa = chunk * p.xchunk;
    const int xb = xa + p.xchunk < p.nx ? xa + p.xchunk : p.nx;
    const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;

    Pos ps;
    bool plain = true;
    {
      const int zl = z - 1, zr = z + V;
      const int zli = per_z ? wrap_index(zl, p.nz) : clamp_index(zl, 0, p.nz - 1);
      const int zri = per_z ? wrap_index(zr, p.nz) : clamp_index(zr, 0, p.nz - 1);
      ps.zl_side = (!per_z && zl < 0) ? 0 : -1;
      ps.zr_side = (!per_z && zr >= p.nz) ? 1 : -1;
      plain = ps.zl_side < 0 && ps.zr_side < 0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int yy = y + j - 1;
        const int yi = per_y ? wrap_index(yy, p.ny) : clamp_index(yy, 0, p.ny - 1);
        ps.yside[j] = (!per_y && yy < 0) ? 0 : ((!per_y && yy >= p.ny) ? 1 : -1);
        plain = plain && ps.yside[j] < 0;
        ps.roff[j] = (long long)yi * p.nz + z;
        ps.loff[j] = (long long)yi * p.nz + zli;
        ps.hoff[j] = (long long)yi * p.nz + zri;
      }
    }

    // rolling window of three planes; the loop is unrolled by three so that the planes
    // rotate by renaming instead of by register moves
    Plane f0, f1, f2;
    load_plane(p, plane(p, xa - 1), ps, plain, f0);
    load_plane(p, plane(p, xa), ps, plain, f1);
    const long long plane_sz = (long long)p.ny * p.nz;
    long long o = (long long)xa * plane_sz + (long long)y * p.nz + z;
    for (int x = xa; x < xb; x += 3) {
      load_plane(p, plane(p, x + 1), ps, plain, f2);
      emit(p, f0, f1, f2, o);
      if (x + 1 < xb) {
        load_plane(p, plane(p, x + 2), ps, plain, f0);
        emit(p, f1, f2, f0, o + plane_sz);
      }
      if (x + 2 < xb) {
        load_plane(p, plane(p, x + 3), ps, plain, f1);
        emit(p, f2, f0, f1, o + 2 * plane_sz);
      }
      o += 3 * plane_sz;
    }
  }
};

// ------------------------------------------------------------------------------------
// Shared-memory tile variant of the same update (the fast path for V = 4 floats).
//
// A block owns a (TY x TZ) tile of the y-z plane and marches along x.  Four plane slots of
// the ghost-padded, clipped field live in shared memory ([TY+2][G+2] groups each, ring of
// one cell); ghost values - periodic images, Neumann/Dirichlet rules, x-slab halos, edge and
// corner ghosts in the reference's x,y,z order - are resolved ONCE when a plane is stored,
// so the 19-point evaluation reads plain neighbours.  Per plane: evaluate plane x from slots
// (x-1, x, x+1), barrier, publish plane x+3 (prefetched into a register one plane earlier)
// into the slot of plane x-1.  Ring positions sit in their own warps and only move data.
// Replaces the register-window kernel above where it applies: that one re-reads rows through
// L1 with a 54-register window and is latency-bound at two blocks per SM.
// ------------------------------------------------------------------------------------
template <typename T, int V, int TY, int G>
struct AcTileProgram {
  static constexpr int TZ = G * V;
  static constexpr int COLS = G + 2, ROWS = TY + 2;
  static constexpr int N_INT = TY * G, N_RROW = 2 * G, N_RCOL = 2 * (TY + 2);
  static constexpr int NPOS = N_INT + N_RROW + N_RCOL;
  static constexpr int NTHREADS = ((NPOS + 31) / 32) * 32;
  static_assert(V >= 2, "tile variant needs vector groups");
  using Base = AcProgram<T, V, TY, G>;
  using P = AcParams<T>;
  using Vt = Vec<T, V>;
  using Plane = typename Base::Plane;

  struct Smem {
    Vt c[4][ROWS][COLS];
    Vt stage[2][NTHREADS];   // raw planes in flight (cp.async), one cell per thread
    // z neighbours of every interior group, kept as compact arrays next to the planes: the value
    // left of a group (last element of the group before it) and right of it (first element of
    // the group after it).  Read as 4-byte words straight from c[][][] these sit 16 bytes apart
    // from lane to lane - a 4-way bank conflict on each of the ten neighbour reads of a voxel
    // group, as many shared-memory wavefronts as all its vector reads together; here a warp (two
    // rows of G = 16 groups) reads 32 consecutive words.
    T zl[4][ROWS][G];
    T zr[4][ROWS][G];
  };
  static_assert(sizeof(Vt) == 16, "the staging copy moves 16 bytes per thread");

  struct Regs {
    int row, col, tid;       // smem indices of the own position
    bool has_pos, interior;
    bool plain;              // no y / z ghost rule touches this position
    int yside, zside;        // -1 or lo/hi ghost side of this position (non-periodic only)
    long long off;           // element offset of the (mapped) position inside a plane
    long long o;             // output element index of plane x at the own position
    long long ps;
    const T* pn;             // own element of plane qn, the next one to fetch
    int qn;
    int xa, xb;
  };

  // value stored for this position: clip, then the x / y / z ghost rules in reference order
  EVX_HD static Vt padded(const Regs& t, const P& p, const Vt& raw, int xside) {
    Vt v;
    if (t.plain && xside < 0) {          // steady state: clip only
#pragma unroll
      for (int k = 0; k < V; ++k) v.v[k] = clip01(raw.v[k]);
      return v;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      T a = clip01(raw.v[k]);
      a = Base::rule(p, 0, xside, a);
      a = Base::rule(p, 1, t.yside, a);
      v.v[k] = a;
    }
    if (t.zside == 0) {        // left ghost group: its last element is the ghost of z = 0
      v.v[V - 1] = Base::rule(p, 2, 0, v.v[0]);
    } else if (t.zside == 1) { // right ghost group: its first element is the ghost of z = nz-1
      v.v[0] = Base::rule(p, 2, 1, v.v[V - 1]);
    }
    return v;
  }

  // store the padded group of the own position into plane slot `sl`, and its outer elements
  // into the neighbour arrays of the groups next to it
  EVX_HD static void publish(Smem& s, int sl, const Regs& t, const Vt& v) {
    s.c[sl][t.row][t.col] = v;
    if (t.col <= G - 1) s.zl[sl][t.row][t.col] = v.v[V - 1];      // left neighbour of group col + 1
    if (t.col >= 2) s.zr[sl][t.row][t.col - 2] = v.v[0];          // right neighbour of group col - 1
  }

  EVX_HD static Vt load_plane(const Regs& t, const P& p, int q, int& xside) {
    const typename Base::PlaneRef pl = Base::plane(p, q);
    xside = pl.side;
    return vec_load<T, V>(pl.ptr + t.off);
  }

  // x ghost side of plane q (-1 inside the slab, for halos and for periodic images)
  EVX_HD static int xside_of(const P& p, int q) {
    if (q >= 0 && q < p.nx) return -1;
    return Base::plane(p, q).side;
  }

  // start the copy of plane t.qn into staging cell `cell`, then step to the next plane
  EVX_HD static void fetch_next(Regs& t, Smem& s, const P& p, int cell) {
    async_copy16(&s.stage[cell][t.tid], t.pn);
    async_copy_commit();
    ++t.qn;
    if (t.qn >= 1 && t.qn < p.nx) t.pn += t.ps;             // both planes inside the slab
    else t.pn = Base::plane(p, t.qn).ptr + t.off;
  }

  EVX_HD static void init(Regs& t, Smem& s, const P& p, int tid, int tile, int chunk) {
    const int tiles_z = (p.nz + TZ - 1) / TZ;
    const int y0 = (tile / tiles_z) * TY, z0 = (tile % tiles_z) * TZ;
    t.xa = chunk * p.xchunk;
    t.xb = t.xa + p.xchunk < p.nx ? t.xa + p.xchunk : p.nx;
    t.tid = tid;
    t.has_pos = tid < NPOS;
    int r, g;
    if (tid < N_INT) {
      r = tid / G; g = tid % G;
    } else if (tid < N_INT + N_RROW) {
      const int j = tid - N_INT;
      r = j < G ? -1 : TY; g = j % G;
    } else {
      const int j = tid - N_INT - N_RROW;
      r = j % (TY + 2) - 1; g = j < TY + 2 ? -1 : G;
    }
    t.row = r + 1; t.col = g + 1;
    const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;
    const int y = y0 + r, z = z0 + g * V;
    const int yi = per_y ? wrap_index(y, p.ny) : clamp_index(y, 0, p.ny - 1);
    const int zi = per_z ? wrap_index(z, p.nz) : clamp_index(z, 0, p.nz - V);
    t.off = (long long)yi * p.nz + zi;
    t.yside = (!per_y && y < 0) ? 0 : ((!per_y && y >= p.ny) ? 1 : -1);
    t.zside = (!per_z && z < 0) ? 0 : ((!per_z && z >= p.nz) ? 1 : -1);
    t.plain = t.yside < 0 && t.zside < 0;
    t.interior = tid < N_INT && y < p.ny && z + V <= p.nz;
    t.ps = (long long)p.ny * p.nz;
    t.o = (long long)t.xa * t.ps + (long long)y * p.nz + z;
    t.qn = t.xa + 3;
    t.pn = Base::plane(p, t.qn).ptr + t.off;
    if (t.has_pos) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {     // planes xa-1 .. xa+2 -> slots 0..3
        int xs;
        const Vt raw = load_plane(t, p, t.xa - 1 + i, xs);
        publish(s, i, t, padded(t, p, raw, xs));
      }
      if (t.qn <= t.xb) fetch_next(t, s, p, 0);     // plane xa+3 -> cell 0
    }
  }

  // group `col` of row `row` of slot `sl` as a (V+2)-wide window incl. the neighbouring z values
  EVX_HD static void window(const Smem& s, int sl, int row, int col, T* w) {
    const Vt c = s.c[sl][row][col];
#pragma unroll
    for (int k = 0; k < V; ++k) w[k + 1] = c.v[k];
    w[0] = s.zl[sl][row][col - 1];
    w[V + 1] = s.zr[sl][row][col - 1];
  }
  EVX_HD static void centre_only(const Smem& s, int sl, int row, int col, T* w) {
    const Vt c = s.c[sl][row][col];
#pragma unroll
    for (int k = 0; k < V; ++k) w[k + 1] = c.v[k];
    w[0] = w[V + 1] = T(0);          // (y+-1, z+-1) of the x+-1 planes are not in the stencil
  }

  // ROT = (x - xa) mod 4: planes x-1, x, x+1 live in slots ROT, ROT+1, ROT+2 (mod 4)
  template <int ROT>
  EVX_HD static void phase_a(Regs& t, Smem& s, const P& p, int x) {
    if (!t.interior) return;
    constexpr int sm = ROT % 4, sc = (ROT + 1) % 4, sp = (ROT + 2) % 4;
    Plane fm, fc, fp;
    centre_only(s, sm, t.row - 1, t.col, fm.w[0]);
    window(s, sm, t.row, t.col, fm.w[1]);
    centre_only(s, sm, t.row + 1, t.col, fm.w[2]);
    window(s, sc, t.row - 1, t.col, fc.w[0]);
    window(s, sc, t.row, t.col, fc.w[1]);
    window(s, sc, t.row + 1, t.col, fc.w[2]);
    centre_only(s, sp, t.row - 1, t.col, fp.w[0]);
    window(s, sp, t.row, t.col, fp.w[1]);
    centre_only(s, sp, t.row + 1, t.col, fp.w[2]);
    Base::emit(p, fm, fc, fp, t.o);
    t.o += t.ps;
  }

  // after the barrier: plane x+3 (staged by the previous iteration) replaces plane x-1, and
  // the copy of plane x+4 starts
  template <int ROT>
  EVX_HD static void phase_b(Regs& t, Smem& s, const P& p, int x) {
    if (!t.has_pos) return;
    // planes needed by later iterations: up to xb (as "x+1" of the last plane xb-1)
    if (x + 3 <= t.xb) {
      async_copy_wait_all();
      const Vt raw = s.stage[ROT & 1][t.tid];
      publish(s, ROT % 4, t, padded(t, p, raw, xside_of(p, x + 3)));
      if (x + 4 <= t.xb) fetch_next(t, s, p, (ROT + 1) & 1);
    }
  }
};

// ------------------------------------------------------------------------------------
// Generic ghost-layer fetch used by the pad kernel: value of the padded field at padded
// coordinates (i,j,k) in [0,n+2)^3 (reference boundary_conditions.py:33-59).
// ------------------------------------------------------------------------------------
template <typename T>
struct PadParams {
  const T* in;
  T* out;
  int nx, ny, nz;
  int bc_kind[3];
  T ghost_off[3][2];
  T ghost_sgn[3];
};

template <typename T>
EVX_HD T padded_value(const PadParams<T>& p, int i, int j, int k) {
  const int n[3] = {p.nx, p.ny, p.nz};
  int q[3] = {i - 1, j - 1, k - 1};
  int side[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    side[a] = -1;
    if (q[a] < 0 || q[a] >= n[a]) {
      if (p.bc_kind[a] == BC_PERIODIC) {
        q[a] = wrap_index(q[a], n[a]);
      } else {
        side[a] = q[a] < 0 ? 0 : 1;
        q[a] = q[a] < 0 ? 0 : n[a] - 1;
      }
    }
  }
  T v = p.in[((long long)q[0] * p.ny + q[1]) * p.nz + q[2]];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (side[a] >= 0) v = p.ghost_off[a][side[a]] + p.ghost_sgn[a] * v;
  return v;
}

// Stencils on an already padded field (reference fd_stencils.py:56-103): value at interior
// voxel (x,y,z); g points at the padded array [nx+2,ny+2,nz+2].
template <typename T>
struct PaddedStencilParams {
  const T* g;
  T* out;
  int nx, ny, nz, op;
  T ihx, ihy, ihz, ihx2, ihy2, ihz2, ih2sum;
};

template <typename T>
EVX_HD T padded_stencil_value(const PaddedStencilParams<T>& p, int x, int y, int z) {
  const long long sy = p.nz + 2, sx = (long long)(p.ny + 2) * (p.nz + 2);
  const T* c = p.g + (x + 1) * sx + (y + 1) * sy + (z + 1);
  const T C = c[0], R = c[sx], L = c[-sx], Tp = c[sy], B = c[-sy], F = c[1], Bk = c[-1];
  if (p.op == 0)
    return (R + L) * p.ihx2 + (Tp + B) * p.ihy2 + (F + Bk) * p.ihz2 - T(2) * C * p.ih2sum;
  const T gx = T(0.5) * (R - L) * p.ihx, gy = T(0.5) * (Tp - B) * p.ihy,
          gz = T(0.5) * (F - Bk) * p.ihz;
  T n2 = gx * gx + gy * gy + gz * gz;
  if (p.op == 2) return n2;
  const T mxy = c[sx + sy] + c[-sx - sy] - c[-sx + sy] - c[sx - sy];
  const T mxz = c[sx + 1] + c[-sx - 1] - c[-sx + 1] - c[sx - 1];
  const T myz = c[sy + 1] + c[-sy - 1] - c[-sy + 1] - c[sy - 1];
  const T num = gx * gx * (R - T(2) * C + L) * p.ihx2 + gy * gy * (Tp - T(2) * C + B) * p.ihy2 +
                gz * gz * (F - T(2) * C + Bk) * p.ihz2 + T(0.5) * gx * gy * mxy * p.ihx * p.ihy +
                T(0.5) * gx * gz * mxz * p.ihx * p.ihz + T(0.5) * gy * gz * myz * p.ihy * p.ihz;
  if (n2 <= T(1e-7)) n2 = T(1);
  return num / n2;
}

}  // namespace evx
