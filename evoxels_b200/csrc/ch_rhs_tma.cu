// Fused Cahn-Hilliard right-hand side, warp-specialised form for fp32 fields whose rows are
// whole float4 groups (the headline 512^3 case and every slab of the multi-GPU plan).
//
// Same arithmetic as ch_rhs_core.h (reference evoxels/problem_definition.py:350-371:
// c^ = clip(c), mu = g(c^) - 2 eps lap c^, rhs = D div(c_f (1 - c_f) grad mu)), different
// machine mapping:
//
//   * a block owns a (TY = 2 NWI rows) x (TZ = 128 z) tile of the y-z plane and marches along
//     x through a chunk of planes;
//   * ONE LOADER THREAD brings every plane of the tile, extended by two rows and one float4 per
//     side, into a ring of shared-memory slots with TMA tensor copies (cp.async.bulk.tensor,
//     completion on an mbarrier per slot): the tile itself, the two rows above / below, the
//     float4 columns left / right and the four corner float4 are NINE boxes with separate
//     destinations, so the periodic wrap is nothing but the coordinates of a box and no two
//     boxes ever write the same bytes.  (A first version issued one 1-D bulk copy per row: 20
//     to 60 serialised UBLKCP per plane kept the loader warp busy for ~1100 of the ~1400 cycles
//     a plane may take - ncu showed the compute warps waiting on the full barrier.)
//     Non-periodic ghosts are never copied (the consumers synthesise them from the adjacent
//     inner value, the rule of boundary_conditions.py:9-59).  The compute warps hold no global
//     pointers, issue no loads and carry no plane bookkeeping;
//   * NWI INTERIOR WARPS: a warp covers two adjacent rows x 128 z, a thread 2 rows x 4 z.  The
//     x window of c^ (three planes), mu of the previous plane and the x-face flux live in
//     registers; y neighbours between the two own rows never leave the thread; the other y / z
//     neighbours are read (and clipped) from the raw plane slot, so c^ is never written back;
//   * ONE RING WARP computes mu on the ring of the tile (one row above and below, one cell left
//     and right), which the interior threads need for the flux divergence;
//   * z neighbours (of c^ and of mu) are the adjacent lanes' registers: two shuffles per row and
//     side; only the first / last group of a tile reads the ring;
//   * mu of a plane is exchanged through shared memory as float4 rows; one named barrier per
//     plane among the compute warps (an mbarrier arrive / wait pair with four mu slots, which
//     lets warps drift a plane apart, is kept as a compile-time variant - measured slower).
//
// Per thread and plane (8 voxels): ~250 instructions of which 170 arithmetic (packed FP32
// wherever the operands pair up along z), against 2 x 248 in the cp.async form.  ncu at 512^3:
// FMA pipe 50 % busy (a packed instruction occupies it for two cycles), shared-memory pipe
// 55 %, issue slots 55 % - three resources at half load at once, none saturated.
#include <cuda_runtime.h>
#include <cstdlib>
#include "evx_internal.h"
#include "evx_params.h"
#include "packed_f32.h"
#include "tma_ptx.h"
#include "stencil_tma.h"

namespace evx {
namespace chtma {
using namespace stma;

struct Consts {
  float k2ps, km3ps, kpl;     // g(c) + l0 c = c (kpl + c (km3ps + k2ps c)),  kpl = 18/eps + l0
  float lx, ly, lz, fx, fy, fz;
  float ox0, ox1, sx, oy0, oy1, sy, oz0, oz1, sz;
};

__device__ __forceinline__ Consts make_consts(const ChParams<float>& p) {
  Consts k;
  k.k2ps = 2.0f * p.pot_scale;
  k.km3ps = -3.0f * p.pot_scale;
  k.kpl = p.pot_scale + p.l0;
  k.lx = p.lx; k.ly = p.ly; k.lz = p.lz; k.fx = p.fx; k.fy = p.fy; k.fz = p.fz;
  k.ox0 = p.ghost_off[0][0]; k.ox1 = p.ghost_off[0][1]; k.sx = p.ghost_sgn[0];
  k.oy0 = p.ghost_off[1][0]; k.oy1 = p.ghost_off[1][1]; k.sy = p.ghost_sgn[1];
  k.oz0 = p.ghost_off[2][0]; k.oz1 = p.ghost_off[2][1]; k.sz = p.ghost_sgn[2];
  return k;
}

__device__ __forceinline__ f2 pr(const float* w, int k) { return f2{w[k], w[k + 1]}; }
__device__ __forceinline__ void un(float* w, int k, f2 v) { w[k] = v.a; w[k + 1] = v.b; }

// mu of four consecutive z values of one row.  cC: c^ of the row; xm/xp, yn/ys: c^ of the
// x and y neighbours; cl/cr: c^ left / right of the group.
__device__ __forceinline__ void mu_row(float* m, const float* cC, const float* xm, const float* xp,
                                       const float* yn, const float* ys, float cl, float cr,
                                       const Consts& k) {
  float szv[4];
  szv[0] = cl + cC[1];
  szv[1] = cC[0] + cC[2];
  szv[2] = cC[1] + cC[3];
  szv[3] = cC[2] + cr;
#pragma unroll
  for (int j = 0; j < 4; j += 2) {
    const f2 c0 = pr(cC, j);
    const f2 sx = f2_add(pr(xp, j), pr(xm, j));
    const f2 sy = f2_add(pr(yn, j), pr(ys, j));
    f2 t = f2_fma(c0, f2_splat(k.k2ps), f2_splat(k.km3ps));
    t = f2_fma(c0, t, f2_splat(k.kpl));
    f2 a = f2_mul(c0, t);
    a = f2_fma(pr(szv, j), f2_splat(k.lz), a);
    a = f2_fma(sy, f2_splat(k.ly), a);
    a = f2_fma(sx, f2_splat(k.lx), a);
    un(m, j, a);
  }
}
__device__ __forceinline__ float mu_cell(float c0, float xm, float xp, float yn, float ys, float cl,
                                         float cr, const Consts& k) {
  float t = fmaf(c0, k.k2ps, k.km3ps);
  t = fmaf(c0, t, k.kpl);
  float a = c0 * t;
  a = fmaf(cl + cr, k.lz, a);
  a = fmaf(yn + ys, k.ly, a);
  a = fmaf(xp + xm, k.lx, a);
  return a;
}

// 4 cf (1 - cf) (mb - ma), cf = (ca + cb) / 2
__device__ __forceinline__ f2 face2(f2 ca, f2 cb, f2 ma, f2 mb) {
  const f2 s = f2_add(ca, cb);
  return f2_mul(f2_mul(s, f2_sub(f2_splat(2.0f), s)), f2_sub(mb, ma));
}
__device__ __forceinline__ float face1(float ca, float cb, float ma, float mb) {
  const float s = ca + cb;
  return s * (2.0f - s) * (mb - ma);
}

__device__ __forceinline__ void ld4(float* r, const float* s) {
  const float4 v = *reinterpret_cast<const float4*>(s);
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void ld4sat(float* r, const float* s) {
  const float4 v = *reinterpret_cast<const float4*>(s);
  r[0] = sat(v.x); r[1] = sat(v.y); r[2] = sat(v.z); r[3] = sat(v.w);
}
__device__ __forceinline__ void st4(float* s, const float* r) {
  *reinterpret_cast<float4*>(s) = make_float4(r[0], r[1], r[2], r[3]);
}

template <int NWI, int NS, bool GHOSTS, bool DEC>
struct Prog {
  static constexpr int TY = 2 * NWI, G = 32, TZ = 128;
  static constexpr int NCOMP = (NWI + 1) * 32, NTHREADS = NCOMP + 32;
  // mu slots: a warp may run a whole plane ahead of the slowest one (the mu exchange is an
  // mbarrier arrive after the store and a wait just before the neighbours are read - no
  // block-wide barrier), so four planes of mu are kept
  // (DEC; otherwise one named barrier per plane and two slots)
  static constexpr int MS = DEC ? 4 : 2, MSM = MS - 1, MUSLOT = (TY + 2) * G * 4;
  static_assert(2 * TY <= 32, "ring columns are handled by one warp");
  // one plane slot (float offsets; every piece is the 128-byte aligned image of one TMA box)
  static constexpr int LPAD = (TY * 4 + 31) / 32 * 32;
  static constexpr int OM = 0;                 // [TY][128]  rows y0 .. y0+TY-1, z0 .. z0+127
  static constexpr int OT = OM + TY * TZ;      // [2][128]   rows y0-2, y0-1
  static constexpr int OB = OT + 2 * TZ;       // [2][128]   rows y0+TY, y0+TY+1
  static constexpr int OL = OB + 2 * TZ;       // [TY][4]    z0-4 .. z0-1 of the tile rows
  static constexpr int OR = OL + LPAD;         // [TY][4]    z0+nzt .. z0+nzt+3
  static constexpr int OLT = OR + LPAD;        // [4]        corner float4: row y0-1, left
  static constexpr int OLB = OLT + 32;         //            row y0+TY, left
  static constexpr int ORT = OLB + 32;         //            row y0-1, right
  static constexpr int ORB = ORT + 32;         //            row y0+TY, right
  static constexpr int SLOT = ORB + 32;
  struct Smem {
    float c[NS][SLOT];            // raw planes
    float mu[MS][TY + 2][G * 4];  // rows y0-1 .. y0+TY
    float rl[MS][TY], rr[MS][TY]; // mu of the cells left / right of the tile rows (ring warp)
    unsigned long long full[NS], empty[NS];
    unsigned long long mubar[2];  // "mu of plane i is in shared memory": barrier i & 1, phase i >> 1
  };

  struct Tile {
    int y0, z0, nzt, gv, xa, xb;
    int tv;              // rows of the tile inside the domain (even; < TY in the last tile of y)
    bool xlo_ghost, xhi_ghost;
  };

  __device__ static bool ghost_plane(const Tile& t, const ChParams<float>& p, int q) {
    return GHOSTS && ((q < 0 && t.xlo_ghost) || (q >= p.nx && t.xhi_ghost));
  }

  // ---------------------------------------------------------------------------------
  // loader thread
  // ---------------------------------------------------------------------------------
  // plane q of the slab -> (array: 0 = c, 1 = halo_lo, 2 = halo_hi; plane index inside it);
  // -1 = non-periodic ghost plane (nothing to copy)
  __device__ static int plane_of(const ChParams<float>& p, int q, int& xi) {
    if (q < 0) {
      if (p.halo_lo) { xi = q + 2; return 1; }
      if (p.bc_kind[0] != BC_PERIODIC) return -1;
      xi = wrap_index(q, p.nx);
      return 0;
    }
    if (q >= p.nx) {
      if (p.halo_hi) { xi = q - p.nx; return 2; }
      if (p.bc_kind[0] != BC_PERIODIC) return -1;
      xi = wrap_index(q, p.nx);
      return 0;
    }
    xi = q;
    return 0;
  }

  __device__ static void loader(Smem& S, const ChParams<float>& p, const Maps& maps, const Tile& t) {
    const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;
    // coordinates of the ring pieces (periodic images) or "absent" (non-periodic ghosts)
    int yt2 = t.y0 - 2, yt1 = t.y0 - 1, yb = t.y0 + t.tv;
    bool has_t = true, has_b = true, has_l = true, has_r = true;
    if (yt2 < 0) { if (per_y) { yt2 += p.ny; yt1 += p.ny; } else has_t = false; }
    if (yb >= p.ny) { if (per_y) yb -= p.ny; else has_b = false; }
    int zl = t.z0 - 4, zr = t.z0 + t.nzt;
    if (zl < 0) { if (per_z) zl += p.nz; else has_l = false; }
    if (zr >= p.nz) { if (per_z) zr -= p.nz; else has_r = false; }
    const unsigned bytes = (unsigned)sizeof(float) *
        (TY * TZ + (has_t ? 2 * TZ : 0) + (has_b ? 2 * TZ : 0) + (has_l ? TY * 4 : 0) +
         (has_r ? TY * 4 : 0) + (has_l && has_t ? 4 : 0) + (has_l && has_b ? 4 : 0) +
         (has_r && has_t ? 4 : 0) + (has_r && has_b ? 4 : 0));
    int slot = 0;
    unsigned ph = 0;     // parity to wait for on empty[slot] (first pass through the ring: no wait)
    bool first = true;
    for (int q = t.xa - 2; q <= t.xb + 1; ++q) {
      if (!first) wait_empty(&S.empty[slot], ph);
      int xi = 0;
      const int arr = plane_of(p, q, xi);
      unsigned long long* bar = &S.full[slot];
      if (arr < 0) {
        mbar_arrive(bar);
      } else {
        const CUtensorMap* m = maps.m[arr];
        float* d = &S.c[slot][0];
        mbar_expect_tx(bar, bytes);
        tma_load_3d(d + OM, &m[0], bar, t.z0, t.y0, xi);
        if (has_t) tma_load_3d(d + OT, &m[1], bar, t.z0, yt2, xi);
        if (has_b) tma_load_3d(d + OB, &m[1], bar, t.z0, yb, xi);
        if (has_l) tma_load_3d(d + OL, &m[2], bar, zl, t.y0, xi);
        if (has_r) tma_load_3d(d + OR, &m[2], bar, zr, t.y0, xi);
        if (has_l && has_t) tma_load_3d(d + OLT, &m[3], bar, zl, yt1, xi);
        if (has_l && has_b) tma_load_3d(d + OLB, &m[3], bar, zl, yb, xi);
        if (has_r && has_t) tma_load_3d(d + ORT, &m[3], bar, zr, yt1, xi);
        if (has_r && has_b) tma_load_3d(d + ORB, &m[3], bar, zr, yb, xi);
      }
      if (++slot == NS) {
        slot = 0;
        if (!first) ph ^= 1u;
        first = false;
      }
    }
  }

  // ---------------------------------------------------------------------------------
  // interior warps: thread = rows (ra, ra+1) of the slot x group g
  // ---------------------------------------------------------------------------------
  struct IReg {
    float c[4][2][4];    // c^ of planes pl-1, pl, pl+1 under names (ROT, ROT+1, ROT+2) mod 4
    float m[2][2][4];
    float fx[2][2][4];
    float yS[2][4], yN[2][4];
    float zL[2][2], zR[2][2];
    int o_own, o_S, o_N;     // float offsets inside a plane slot: own group of row a (row b:
    int o_L, o_R;            // +128), row below a / above b, ring float left / right of row a
    int offm;            // float offset of (mu row of ra, group g) inside a mu slot
    int orow;            // tile row of the first own row
    bool edge_l, edge_r; // first / last group of the tile: z neighbours come from the ring
    float* po;           // output element of row a in plane xa
    long long ps;
    bool st_a, st_b;     // rows inside the domain (and group valid)
    bool gvalid;         // group inside the tile's z extent
    bool gy_lo, gy_hi, gz_lo, gz_hi;
  };

  // The plane loop is unrolled FOUR-fold: PAR (mod 2) names the carried mu / flux / neighbour
  // registers, ROT (mod 4) the three live planes of the c^ window, so both rotate by renaming.
  // (Six-fold unrolling with a 3-name window made the kernel 3200 instructions long: 77 %
  // instruction-cache hit rate and `no_instruction` the top stall.)
  template <int PAR, int ROT>
  __device__ static void interior_step(IReg& r, Smem& S, const ChParams<float>& p, const Tile& t,
                                       const Consts& k, int pl, int sp, int sn, unsigned phn, int lane, int it) {
    constexpr int iB = ROT % 4, iC = (ROT + 1) % 4, iD = (ROT + 2) % 4;
    const float* cs_p = &S.c[0][0] + sp * SLOT;
    const float* cs_n = &S.c[0][0] + sn * SLOT;
    // own values of plane pl+1
    wait_full(&S.full[sn], phn);
    ld4sat(r.c[iD][0], cs_n + r.o_own);
    ld4sat(r.c[iD][1], cs_n + r.o_own + TZ);
    // neighbours of plane pl
    float cS[4], cN[4], cL[2], cR[2];
    ld4sat(cS, cs_p + r.o_S);
    ld4sat(cN, cs_p + r.o_N);
    const float* cC0 = r.c[iC][0];
    const float* cC1 = r.c[iC][1];
    // z neighbours: the adjacent lanes hold them (already clipped); the first / last group of
    // the tile reads the ring
    cL[0] = __shfl_up_sync(0xffffffffu, cC0[3], 1);
    cL[1] = __shfl_up_sync(0xffffffffu, cC1[3], 1);
    cR[0] = __shfl_down_sync(0xffffffffu, cC0[0], 1);
    cR[1] = __shfl_down_sync(0xffffffffu, cC1[0], 1);
    if (r.edge_l) { cL[0] = sat(cs_p[r.o_L]); cL[1] = sat(cs_p[r.o_L + 4]); }
    if (r.edge_r) { cR[0] = sat(cs_p[r.o_R]); cR[1] = sat(cs_p[r.o_R + 4]); }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[sp]);
    if (GHOSTS) {
      if (r.gy_lo) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cS[j] = k.oy0 + k.sy * cC0[j];
      }
      if (r.gy_hi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cN[j] = k.oy1 + k.sy * cC1[j];
      }
      if (r.gz_lo) { cL[0] = k.oz0 + k.sz * cC0[0]; cL[1] = k.oz0 + k.sz * cC1[0]; }
      if (r.gz_hi) { cR[0] = k.oz1 + k.sz * cC0[3]; cR[1] = k.oz1 + k.sz * cC1[3]; }
    }
    // mu(pl)
    float mC[2][4];
    if (!ghost_plane(t, p, pl)) {
      float xm[2][4], xp[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xm[a][j] = r.c[iB][a][j];
          xp[a][j] = r.c[iD][a][j];
        }
      if (GHOSTS) {
        if (ghost_plane(t, p, pl - 1))
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) xm[a][j] = k.ox0 + k.sx * r.c[iC][a][j];
        if (ghost_plane(t, p, pl + 1))
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) xp[a][j] = k.ox1 + k.sx * r.c[iC][a][j];
      }
      mu_row(mC[0], cC0, xm[0], xp[0], cC1, cS, cL[0], cR[0], k);
      mu_row(mC[1], cC1, xm[1], xp[1], cN, cC0, cL[1], cR[1], k);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) mC[0][j] = mC[1][j] = 0.0f;
    }
    if (r.st_a) {       // groups / rows beyond a partial tile must not touch the ring's entries
      float* ms = &S.mu[0][0][0] + (it & MSM) * MUSLOT + r.offm;
      st4(ms, mC[0]);
      st4(ms + G * 4, mC[1]);
    }
    if (DEC) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.mubar[it & 1]);
    }
    // x-face term between planes pl-1 and pl
    float fxp[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 4; j += 2)
        un(fxp[a], j, face2(pr(r.c[iB][a], j), pr(r.c[iC][a], j), pr(r.m[PAR ^ 1][a], j), pr(mC[a], j)));

    // rhs at plane x = pl-1
    const int x = pl - 1;
    if (x >= t.xa) {
      const float(*cB)[4] = r.c[iB];
      const float(*mB)[4] = r.m[PAR ^ 1];
      const int mslot = (it - 1) & MSM;
      const float* ms = &S.mu[0][0][0] + mslot * MUSLOT + r.offm;
      float mS[4], mN[4], mL[2], mR[2];
      mL[0] = __shfl_up_sync(0xffffffffu, mB[0][3], 1);
      mL[1] = __shfl_up_sync(0xffffffffu, mB[1][3], 1);
      mR[0] = __shfl_down_sync(0xffffffffu, mB[0][0], 1);
      mR[1] = __shfl_down_sync(0xffffffffu, mB[1][0], 1);
      // mu of plane pl-1 of every warp (and of the ring) is in shared memory
      if (DEC) wait_full(&S.mubar[(it - 1) & 1], (unsigned)((it - 1) >> 1) & 1u);
      ld4(mS, ms - G * 4);
      ld4(mN, ms + 2 * G * 4);
      if (r.edge_l) { mL[0] = S.rl[mslot][r.orow]; mL[1] = S.rl[mslot][r.orow + 1]; }
      if (r.edge_r) { mR[0] = S.rr[mslot][r.orow]; mR[1] = S.rr[mslot][r.orow + 1]; }
      float fxm[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) fxm[a][j] = r.fx[PAR ^ 1][a][j];
      if (GHOSTS) {
        if (r.gy_lo) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mS[j] = k.oy0 + k.sy * mB[0][j];
        }
        if (r.gy_hi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mN[j] = k.oy1 + k.sy * mB[1][j];
        }
        if (r.gz_lo) { mL[0] = k.oz0 + k.sz * mB[0][0]; mL[1] = k.oz0 + k.sz * mB[1][0]; }
        if (r.gz_hi) { mR[0] = k.oz1 + k.sz * mB[0][3]; mR[1] = k.oz1 + k.sz * mB[1][3]; }
        if (ghost_plane(t, p, x - 1)) {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              fxm[a][j] = face1(k.ox0 + k.sx * cB[a][j], cB[a][j], k.ox0 + k.sx * mB[a][j], mB[a][j]);
        }
        if (ghost_plane(t, p, x + 1)) {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              fxp[a][j] = face1(cB[a][j], k.ox1 + k.sx * cB[a][j], mB[a][j], k.ox1 + k.sx * mB[a][j]);
        }
      }
      // y faces: (S | a), (a | b), (b | N)
      float fy[3][4];
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        un(fy[0], j, face2(pr(r.yS[PAR ^ 1], j), pr(cB[0], j), pr(mS, j), pr(mB[0], j)));
        un(fy[1], j, face2(pr(cB[0], j), pr(cB[1], j), pr(mB[0], j), pr(mB[1], j)));
        un(fy[2], j, face2(pr(cB[1], j), pr(r.yN[PAR ^ 1], j), pr(mB[1], j), pr(mN, j)));
      }
      float o[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        // z faces: window [L, 0, 1, 2, 3, R]; sums and differences are scalar (the pairs
        // straddle the register pairs), the products are packed
        const float* cb = cB[a];
        const float* mb = mB[a];
        const float zl = r.zL[PAR ^ 1][a], zr = r.zR[PAR ^ 1][a];
        float s[6], d[6], f[6];
        s[0] = zl + cb[0]; d[0] = mb[0] - mL[a];
        s[1] = cb[0] + cb[1]; d[1] = mb[1] - mb[0];
        s[2] = cb[1] + cb[2]; d[2] = mb[2] - mb[1];
        s[3] = cb[2] + cb[3]; d[3] = mb[3] - mb[2];
        s[4] = cb[3] + zr; d[4] = mR[a] - mb[3];
        s[5] = 0.0f; d[5] = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const f2 sv = pr(s, j);
          un(f, j, f2_mul(f2_mul(sv, f2_sub(f2_splat(2.0f), sv)), pr(d, j)));
        }
        f[4] = s[4] * (2.0f - s[4]) * d[4];
        float dz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) dz[j] = f[j + 1] - f[j];
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const f2 dx = f2_sub(pr(fxp[a], j), pr(fxm[a], j));
          const f2 dy = f2_sub(pr(fy[a + 1], j), pr(fy[a], j));
          f2 acc = f2_mul(pr(dz, j), f2_splat(k.fz));
          acc = f2_fma(dy, f2_splat(k.fy), acc);
          acc = f2_fma(dx, f2_splat(k.fx), acc);
          un(o[a], j, acc);
        }
      }
      if (r.st_a) st4(r.po, o[0]);
      if (r.st_b) st4(r.po + p.nz, o[1]);
      r.po += r.ps;
    }
    // hand over to the next plane (renaming only: the next plane reads under flipped PAR)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r.m[PAR][a][j] = mC[a][j];
        r.fx[PAR][a][j] = fxp[a][j];
      }
      r.zL[PAR][a] = cL[a];
      r.zR[PAR][a] = cR[a];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r.yS[PAR][j] = cS[j];
      r.yN[PAR][j] = cN[j];
    }
  }

  // ---------------------------------------------------------------------------------
  // ring warp: lane g -> mu of the rows above / below the tile (group g); lanes < 2 TY also
  // -> mu of one cell left (side 0) or right (side 1) of the tile
  // ---------------------------------------------------------------------------------
  struct RReg {
    float c[3][2][4];
    float cc[3];          // ring-column cell: c^ window along x
    int oBS, omB;         // last tile row (slot offset); mu row below the tile (offset in a mu slot)
    int oT, oB;           // float offsets inside a plane slot: own group of the row above / below
    int oTL, oTR, oBL, oBR;   // their left / right neighbours
    int oq, oqS, oqN, oqL, oqR;   // ring-column cell and its neighbours
    int offe;             // tile row of the cell; < 0: no cell
    bool edge_l, edge_r;
    bool side;
    bool gy_lo0, gy_hi1;          // ring row above is y = 0 / ring row below is y = ny-1
    bool gz_lo, gz_hi;
    bool cy_lo, cy_hi;            // ring-column cell's row is y = 0 / ny-1
  };

  template <int ROT>
  __device__ static void ring_step(RReg& r, Smem& S, const ChParams<float>& p, const Tile& t,
                                   const Consts& k, int pl, int sp, int sn, unsigned phn, int lane, int it) {
    constexpr int iB = ROT % 3, iC = (ROT + 1) % 3, iD = (ROT + 2) % 3;
    const float* cs_p = &S.c[0][0] + sp * SLOT;
    const float* cs_n = &S.c[0][0] + sn * SLOT;
    wait_full(&S.full[sn], phn);
    ld4sat(r.c[iD][0], cs_n + r.oT);
    ld4sat(r.c[iD][1], cs_n + r.oB);
    const bool cell = r.offe >= 0;
    r.cc[iD] = sat(cs_n[r.oq]);
    float cS[2][4], cN[2][4], cL[2], cR[2];
    ld4sat(cS[0], cs_p + r.oT - TZ);                      // row y0-2
    ld4sat(cN[0], cs_p + OM + 4 * lane);                  // tile row 0
    ld4sat(cS[1], cs_p + r.oBS);                          // last tile row
    ld4sat(cN[1], cs_p + r.oB + TZ);                      // row y0+TY+1
    cL[0] = __shfl_up_sync(0xffffffffu, r.c[iC][0][3], 1);
    cL[1] = __shfl_up_sync(0xffffffffu, r.c[iC][1][3], 1);
    cR[0] = __shfl_down_sync(0xffffffffu, r.c[iC][0][0], 1);
    cR[1] = __shfl_down_sync(0xffffffffu, r.c[iC][1][0], 1);
    if (r.edge_l) { cL[0] = sat(cs_p[r.oTL]); cL[1] = sat(cs_p[r.oBL]); }
    if (r.edge_r) { cR[0] = sat(cs_p[r.oTR]); cR[1] = sat(cs_p[r.oBR]); }
    float qn = sat(cs_p[r.oqN]), qs = sat(cs_p[r.oqS]);
    const float ql = sat(cs_p[r.oqL]), qr = sat(cs_p[r.oqR]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[sp]);
    const bool gp = ghost_plane(t, p, pl);
    float mC[2][4];
    float xm[2][4], xp[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xm[a][j] = r.c[iB][a][j];
        xp[a][j] = r.c[iD][a][j];
      }
    float qm = r.cc[iB], qp = r.cc[iD];
    const float q0 = r.cc[iC];
    if (GHOSTS) {
      const float* c0 = r.c[iC][0];
      const float* c1 = r.c[iC][1];
      if (r.gy_lo0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cS[0][j] = k.oy0 + k.sy * c0[j];
      }
      if (r.gy_hi1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cN[1][j] = k.oy1 + k.sy * c1[j];
      }
      if (r.gz_lo) { cL[0] = k.oz0 + k.sz * c0[0]; cL[1] = k.oz0 + k.sz * c1[0]; }
      if (r.gz_hi) { cR[0] = k.oz1 + k.sz * c0[3]; cR[1] = k.oz1 + k.sz * c1[3]; }
      if (r.cy_lo) qs = k.oy0 + k.sy * q0;
      if (r.cy_hi) qn = k.oy1 + k.sy * q0;
      if (ghost_plane(t, p, pl - 1)) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j) xm[a][j] = k.ox0 + k.sx * r.c[iC][a][j];
        qm = k.ox0 + k.sx * q0;
      }
      if (ghost_plane(t, p, pl + 1)) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j) xp[a][j] = k.ox1 + k.sx * r.c[iC][a][j];
        qp = k.ox1 + k.sx * q0;
      }
    }
    mu_row(mC[0], r.c[iC][0], xm[0], xp[0], cN[0], cS[0], cL[0], cR[0], k);
    mu_row(mC[1], r.c[iC][1], xm[1], xp[1], cN[1], cS[1], cL[1], cR[1], k);
    float mq = mu_cell(q0, qm, qp, qn, qs, ql, qr, k);
    if (gp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mC[0][j] = mC[1][j] = 0.0f;
      mq = 0.0f;
    }
    // the slot was last read for plane pl-4: every warp has passed that point once it has
    // published mu of plane pl-2
    if (DEC && it >= 2) wait_full(&S.mubar[it & 1], (unsigned)((it - 2) >> 1) & 1u);
    const int mslot = it & MSM;
    float* ms = &S.mu[0][0][0] + mslot * MUSLOT + lane * 4;
    st4(ms, mC[0]);
    st4(ms + r.omB, mC[1]);
    if (cell) {
      if (r.side) S.rr[mslot][r.offe] = mq;
      else S.rl[mslot][r.offe] = mq;
    }
    if (DEC) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.mubar[it & 1]);
    }
  }
};

template <int NWI, int NS, bool GHOSTS, int MAXREG, bool DEC>
__global__ void __maxnreg__(MAXREG)
    ch_rhs_tma_kernel(const ChParams<float> p, const __grid_constant__ Maps maps, const int tiles_z) {
  using P = Prog<NWI, NS, GHOSTS, DEC>;
  extern __shared__ unsigned char smem_raw[];
  // TMA destinations want 128-byte aligned shared memory
  typename P::Smem& S = *reinterpret_cast<typename P::Smem*>(
      smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  typename P::Tile t;
  t.y0 = (blockIdx.x / tiles_z) * P::TY;
  t.tv = p.ny - t.y0 < P::TY ? p.ny - t.y0 : P::TY;
  t.z0 = (blockIdx.x % tiles_z) * P::TZ;
  t.nzt = p.nz - t.z0 < P::TZ ? p.nz - t.z0 : P::TZ;
  t.gv = t.nzt >> 2;
  t.xa = blockIdx.y * p.xchunk;
  t.xb = t.xa + p.xchunk < p.nx ? t.xa + p.xchunk : p.nx;
  t.xlo_ghost = p.bc_kind[0] != BC_PERIODIC && !p.halo_lo;
  t.xhi_ghost = p.bc_kind[0] != BC_PERIODIC && !p.halo_hi;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], NWI + 1);
    }
    mbar_init(&S.mubar[0], NWI + 1);
    mbar_init(&S.mubar[1], NWI + 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  if (warp == NWI + 1) {
    if (lane == 0) P::loader(S, p, maps, t);
    return;
  }
  const Consts k = make_consts(p);
  const bool per_y = p.bc_kind[1] == BC_PERIODIC, per_z = p.bc_kind[2] == BC_PERIODIC;
  const int nplanes_it = t.xb - t.xa + 2;      // iterations: pl = xa-1 .. xb
  // slot / phase cursor of plane pl+1 (load index it+2)
  int sp = 1 % NS, sn = 2 % NS;
  unsigned phn = (2 / NS) & 1u;
  auto advance = [&]() {
    sp = sn;
    if (++sn == NS) { sn = 0; phn ^= 1u; }
  };

  if (warp < NWI) {
    typename P::IReg r;
    const int ta = 2 * warp;                           // tile row of the first own row
    r.o_own = P::OM + ta * P::TZ + 4 * lane;
    r.o_S = ta > 0 ? r.o_own - P::TZ : P::OT + P::TZ + 4 * lane;
    r.o_N = ta + 1 < t.tv - 1 ? r.o_own + 2 * P::TZ : P::OB + 4 * lane;
    r.o_L = P::OL + ta * 4 + 3;
    r.o_R = P::OR + ta * 4;
    r.edge_l = lane == 0;
    r.edge_r = lane == t.gv - 1;
    r.orow = ta;
    r.offm = (ta + 1) * P::G * 4 + 4 * lane;
    // keep the offsets in registers (ptxas would otherwise recompute them from the thread index
    // in every plane: ~25 integer instructions per plane)
    asm volatile("" : "+r"(r.o_own), "+r"(r.o_S), "+r"(r.o_N), "+r"(r.offm));
    const int ya = t.y0 + 2 * warp, z = t.z0 + 4 * lane;
    const bool gvalid = lane < t.gv;
    r.gvalid = gvalid;
    r.st_a = gvalid && ya < p.ny;
    r.st_b = gvalid && ya + 1 < p.ny;
    r.ps = (long long)p.ny * p.nz;
    r.po = p.out + (long long)t.xa * r.ps + (long long)ya * p.nz + z;
    r.gy_lo = GHOSTS && !per_y && ya == 0;
    r.gy_hi = GHOSTS && !per_y && ya + 1 == p.ny - 1;
    r.gz_lo = GHOSTS && !per_z && z == 0;
    r.gz_hi = GHOSTS && !per_z && z + 4 == p.nz;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r.m[0][a][j] = r.m[1][a][j] = 0.0f;
        r.fx[0][a][j] = r.fx[1][a][j] = 0.0f;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) r.yS[0][j] = r.yS[1][j] = r.yN[0][j] = r.yN[1][j] = 0.0f;
    r.zL[0][0] = r.zL[0][1] = r.zL[1][0] = r.zL[1][1] = 0.0f;
    r.zR[0][0] = r.zR[0][1] = r.zR[1][0] = r.zR[1][1] = 0.0f;
    // prologue: own values of planes xa-2 (load 0) and xa-1 (load 1)
    {
      const float* c0 = &S.c[0][0] + r.o_own;
      wait_full(&S.full[0], 0);
      ld4sat(r.c[0][0], c0);
      ld4sat(r.c[0][1], c0 + P::TZ);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[0]);
      const float* c1 = &S.c[0][0] + (1 % NS) * P::SLOT + r.o_own;
      wait_full(&S.full[1 % NS], (1 / NS) & 1u);
      ld4sat(r.c[1][0], c1);
      ld4sat(r.c[1][1], c1 + P::TZ);
    }
#define EVX_STEP(PAR, ROT, OFF)                                                          \
  if (it + (OFF) < nplanes_it) {                                                         \
    P::template interior_step<PAR, ROT>(r, S, p, t, k, t.xa - 1 + it + (OFF), sp, sn, phn, lane, it + (OFF)); \
    advance();                                                                           \
    if (!DEC) group_sync(1, P::NCOMP);                                                   \
  }
    for (int it = 0; it < nplanes_it; it += 4) {
      EVX_STEP(0, 0, 0) EVX_STEP(1, 1, 1) EVX_STEP(0, 2, 2) EVX_STEP(1, 3, 3)
    }
#undef EVX_STEP
  } else {
    typename P::RReg r;
    r.oT = P::OT + P::TZ + 4 * lane;
    r.oB = P::OB + 4 * lane;
    r.oTL = P::OLT + 3;
    r.oBL = P::OLB + 3;
    r.oTR = P::ORT;
    r.oBR = P::ORB;
    r.edge_l = lane == 0;
    r.edge_r = lane == t.gv - 1;
    r.gy_lo0 = GHOSTS && !per_y && t.y0 - 1 == 0;
    r.oBS = P::OM + (t.tv - 1) * P::TZ + 4 * lane;
    r.omB = (t.tv + 1) * P::G * 4;
    r.gy_hi1 = GHOSTS && !per_y && t.y0 + t.tv == p.ny - 1;
    const int z = t.z0 + 4 * lane;
    r.gz_lo = GHOSTS && !per_z && z == 0;
    r.gz_hi = GHOSTS && !per_z && z + 4 == p.nz;
    const bool has_cell = lane < 2 * P::TY;
    const int crow = lane % P::TY;
    r.side = lane >= P::TY;
    if (!r.side) {
      r.oq = P::OL + crow * 4 + 3;
      r.oqS = crow > 0 ? r.oq - 4 : P::OLT + 3;
      r.oqN = crow < t.tv - 1 ? r.oq + 4 : P::OLB + 3;
      r.oqL = r.oq - 1;
      r.oqR = P::OM + crow * P::TZ;
    } else {
      r.oq = P::OR + crow * 4;
      r.oqS = crow > 0 ? r.oq - 4 : P::ORT;
      r.oqN = crow < t.tv - 1 ? r.oq + 4 : P::ORB;
      r.oqL = P::OM + crow * P::TZ + t.nzt - 1;
      r.oqR = r.oq + 1;
    }
    r.offe = has_cell && crow < t.tv ? crow : -1;
    r.cy_lo = GHOSTS && !per_y && t.y0 + crow == 0;
    r.cy_hi = GHOSTS && !per_y && t.y0 + crow == p.ny - 1;
    {
      const float* c0 = &S.c[0][0];
      wait_full(&S.full[0], 0);
      ld4sat(r.c[0][0], c0 + r.oT);
      ld4sat(r.c[0][1], c0 + r.oB);
      r.cc[0] = sat(c0[r.oq]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[0]);
      const float* c1 = &S.c[0][0] + (1 % NS) * P::SLOT;
      wait_full(&S.full[1 % NS], (1 / NS) & 1u);
      ld4sat(r.c[1][0], c1 + r.oT);
      ld4sat(r.c[1][1], c1 + r.oB);
      r.cc[1] = sat(c1[r.oq]);
    }
#define EVX_STEP(ROT, OFF)                                                               \
  if (it + (OFF) < nplanes_it) {                                                         \
    P::template ring_step<ROT>(r, S, p, t, k, t.xa - 1 + it + (OFF), sp, sn, phn, lane, it + (OFF)); \
    advance();                                                                           \
    if (!DEC) group_sync(1, P::NCOMP);                                                   \
  }
    for (int it = 0; it < nplanes_it; it += 3) {
      EVX_STEP(0, 0) EVX_STEP(1, 1) EVX_STEP(2, 2)
    }
#undef EVX_STEP
  }
}

template <int NWI, int NS, bool GHOSTS, int MAXREG, bool DEC = false>
static int launch(ChParams<float> p, cudaStream_t st) {
  using P = Prog<NWI, NS, GHOSTS, DEC>;
  static SmemOptIn optin;
  auto kern = ch_rhs_tma_kernel<NWI, NS, GHOSTS, MAXREG, DEC>;
  constexpr int MINB = 65536 / (P::NTHREADS * MAXREG);     // resident blocks per SM
  const size_t smem = sizeof(typename P::Smem) + 128;
  if (int e = optin.ensure(kern, smem)) return e;
  Maps maps;
  if (!make_maps(maps.m[0], p.c, p.nx, p.ny, p.nz, P::TY, 2)) return EVX_ERR_UNSUPPORTED;
  if (p.halo_lo && !make_maps(maps.m[1], p.halo_lo, 2, p.ny, p.nz, P::TY, 2)) return EVX_ERR_UNSUPPORTED;
  if (p.halo_hi && !make_maps(maps.m[2], p.halo_hi, 2, p.ny, p.nz, P::TY, 2)) return EVX_ERR_UNSUPPORTED;
  if (!p.halo_lo) for (int k = 0; k < 4; ++k) maps.m[1][k] = maps.m[0][k];
  if (!p.halo_hi) for (int k = 0; k < 4; ++k) maps.m[2][k] = maps.m[0][k];
  const int tiles_z = (p.nz + P::TZ - 1) / P::TZ;
  const long long tiles = (long long)((p.ny + P::TY - 1) / P::TY) * tiles_z;
  // chunks of x: every chunk re-reads four planes and recomputes mu of two, so they are kept
  // long; their number is chosen so that the blocks fill whole waves of the resident set
  const int want_planes = env_int("EVX_CH_TMA_CHUNK", 64);
  const long long resident = 148LL * MINB;
  int chunks = (p.nx + want_planes - 1) / want_planes;
  if (chunks < 1) chunks = 1;
  {
    // prefer a chunk count whose block count is just below a multiple of the resident set
    double best = -1.0;
    int best_c = chunks;
    for (int c = chunks; (c <= 2 * chunks || c * 8 <= p.nx) && c <= p.nx && c <= 64; ++c) {
      const double waves = (double)(tiles * c) / (double)resident;
      const double eff = waves / (double)(long long)(waves + 0.999999);
      const double over = 1.0 + 4.0 * c / (double)p.nx;      // re-read planes
      const double score = eff / over;
      if (score > best) { best = score; best_c = c; }
    }
    chunks = best_c;
  }
  p.xchunk = (p.nx + chunks - 1) / chunks;
  chunks = (p.nx + p.xchunk - 1) / p.xchunk;
  if (tiles > 2147483647LL || chunks > 65535) return EVX_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  kern<<<grid, P::NTHREADS, smem, st>>>(p, maps, tiles_z);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace chtma

// Entry point used by ch_rhs_impl<float>: returns EVX_ERR_UNSUPPORTED (< 0) when the shape is
// not covered, so that the caller falls back to the cp.async form.
int ch_rhs_tma_f32(ChParams<float> p, cudaStream_t st) {
  const int enabled = chtma::env_int("EVX_CH_TMA", 1);
  if (!enabled || p.hom) return EVX_ERR_UNSUPPORTED;
  if (p.nz % 4 != 0 || p.nz < 64 || p.ny % 2 != 0 || p.ny < 4 || p.nx < 2) return EVX_ERR_UNSUPPORTED;
  if (p.ny > 65535 * 8 || p.nz > (1 << 24)) return EVX_ERR_UNSUPPORTED;
  const bool ghosts = p.bc_kind[0] != BC_PERIODIC || p.bc_kind[1] != BC_PERIODIC ||
                      p.bc_kind[2] != BC_PERIODIC;
  const int cfg = chtma::env_int("EVX_CH_TMA_CFG", 0);
#define EVX_CFG(N, NWI, NS, MAXREG, DEC)                                                \
  if (cfg == N)                                                                         \
    return ghosts ? chtma::launch<NWI, NS, true, MAXREG, DEC>(p, st)                    \
                  : chtma::launch<NWI, NS, false, MAXREG, DEC>(p, st);
  // measured at 512^3 periodic (cp.async form: 0.362 ms): 4 interior warps / 168 registers
  // (two blocks per SM, no spills) 0.300 ms; 6 warps / 128 registers 0.325; 8 warps / 96
  // registers (spills) 0.41; the mbarrier-decoupled mu exchange 0.309 (each try_wait costs ~90
  // cycles of latency that 12 warps per SM cannot hide - the named barrier is cheaper)
  EVX_CFG(1, 6, 4, 128, false)
  EVX_CFG(2, 8, 4, 96, false)
  EVX_CFG(3, 4, 4, 168, true)
  EVX_CFG(0, 4, 4, 168, false)
#undef EVX_CFG
  return EVX_ERR_UNSUPPORTED;
}

}  // namespace evx
