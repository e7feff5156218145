// Strided FFT pass (y or x axis of the half spectrum), TMA-tiled form for 512-point lines.
//
// Same arithmetic as StridedPass / StridedPipe in fft_pass_core.h (same radix-8 Stockham
// stages, same roots, same thread -> butterfly assignment, hence bit-identical results), but
// re-organised around what measurement showed to bound those kernels on B200:
//
//  * Tiles of [512 rows][KZ columns] move between HBM and shared memory with TMA tensor copies
//    (cp.async.bulk.tensor, 64/128-byte swizzle) issued by ONE thread, in both directions.
//    The per-thread cp.async prefetch, the eight strided 64-bit global stores and their
//    64-bit address arithmetic were ~45 % of the instructions of the pipelined y pass.
//  * A line (column) is transformed by 64 threads; TWO neighbouring lines form a "group" of
//    128 threads (4 warps) that synchronises with its own named barrier.  Lines of different
//    groups never exchange data, so there is no block-wide barrier in the tile loop: groups
//    drift apart and their shared-memory bursts and FP32 bursts overlap instead of all 16/32
//    warps of a block running in lockstep.
//  * Stage exchanges alternate between a per-group padded buffer X and the group's own two
//    columns of the tile buffer (in place), so one barrier per exchange suffices and the last
//    stage leaves its natural-order output exactly where the TMA store expects it.
//
// Shared-memory indexing (all accesses are 64-bit, conflict-free per half-warp; a half-warp
// is 8 consecutive t x the 2 lines of a group):
//   tile   element (row r, column c) at byte  r*ROWB + ((c*8) ^ swz(r)),  ROWB = 8*KZ,
//          swz(r) = ((r>>1)&3)<<4 for 64-byte rows (CU_TENSOR_MAP_SWIZZLE_64B),
//                   (r&7)<<4      for 128-byte rows (CU_TENSOR_MAP_SWIZZLE_128B)
//   X      element i of line c2 of the group at cf index  (i + (i>>3)) * 2 + c2
// With r = t + 64 e (natural order) or r = 64 (t/8) + t%8 + 8 e (output of stage 1) the swizzle
// term depends on t only, so every access is one base register plus an immediate.
#pragma once
#include "fft_pass_core.h"

namespace evx {

struct LineParams {
  cf* spec;                 // [nx][ny][P] half spectrum (host replay and address checks)
  const cf* tw;             // W_512 table
  int nx, ny, P;
  int ncols_valid;          // nz/2 + 1
  int tiles_per_row;        // ceil(ncols_valid / KZ)
  long long ntiles;         // (along_x ? ny : nx) * tiles_per_row
  int along_x;              // 0: lines run along y (tile row index = x), 1: along x (row index = y)
  int l2_ahead;             // tiles prefetched into L2 ahead of the shared-memory loads (0 = none)
  FilterParams filt;        // XMID only; n0 = nx
  // 1024-point form (StridedLine4, 4-D tensor maps: kz, line index low part, row, line index
  // high part): first tile row / number of tile rows of this launch, global index of tile row 0
  // along the other strided axis (x-slab y-pencils), rows per TMA box.  Box h of a tile is STORED
  // through output map h / out_div at high coordinate h % out_div (one output map per destination
  // block of the x-slab transposes - a peer's buffer over NVLink, or local; out_div = boxes per
  // tile and a single map for an ordinary pass).  max_ctas > 0 caps the persistent grid (an
  // NVLink-bound launch leaves SMs to the kernels of the next chunk).
  int row0 = 0, nrows = 0, kother0 = 0, box_rows = 256, out_div = 4, max_ctas = 0;
};

template <int L, int KZ, int MODE>
struct StridedLine {
  static_assert(L == 512, "TMA-tiled strided pass: 512-point lines (three radix-8 stages)");
  static_assert(KZ == 8 || KZ == 16, "64- or 128-byte tile rows");
  static constexpr int COLS = KZ;              // columns (lines) per tile
  static constexpr int T = L / 8;              // threads per line
  static constexpr int G = 2;                  // lines per synchronisation group
  static constexpr int GT = T * G;             // threads per group
  static constexpr int NG = KZ / G;            // groups per block
  static constexpr int NTHREADS = T * KZ;
  static constexpr int S = 3;
  static constexpr bool XMID = pass_is_xmid(MODE);
  static constexpr int NPHASES = XMID ? 2 * S - 1 : S;
  static constexpr int ROWB = KZ * (int)sizeof(cf);
  static constexpr int TILE_BYTES = L * ROWB;
  static constexpr int BOX_ROWS = 256;         // TMA box height (boxDim <= 256)
  static constexpr int LP = smem_padded_len(L);
  static constexpr int XG = LP * G;            // cf elements of one group's exchange buffer
  static constexpr int X_BYTES = NG * XG * (int)sizeof(cf);
  // [tile 0 | tile 1 | X | 2 mbarriers | 2 completion counters] + slack to align to 1024 bytes
  static constexpr int CTRL_BYTES = 64;
  static constexpr size_t SMEM_BYTES = 1024 + 2 * (size_t)TILE_BYTES + X_BYTES + CTRL_BYTES;

  struct Regs {
    cf v[8];
    cf w[3];
    int t, c2, g, col;       // position in the line, line of the group, group, column in the tile
    int tb;                  // byte offset of tile element (row t, col)
    int sb;                  // byte offset of tile element (row 64*(t/8) + t%8, col): stage-1 output
    int xn, xs;              // cf index in X: natural order base / stage-0 output base
    int kz, kother;          // XMID: global column and index along the other strided axis
  };

  EVX_HD static int swz(int r) { return KZ == 8 ? ((r >> 1) & 3) << 4 : (r & 7) << 4; }
  EVX_HD static int tile_off(int r, int c) { return r * ROWB + ((c * (int)sizeof(cf)) ^ swz(r)); }

  EVX_HD static void init(Regs& r, int tid) {
    r.g = tid / GT;
    const int tg = tid - r.g * GT;
    r.c2 = tg % G;
    r.t = tg / G;
    r.col = r.g * G + r.c2;
    r.tb = tile_off(r.t, r.col);
    r.sb = tile_off(stage_out_base<L>(1, r.t), r.col);
    r.xn = smem_pad(r.t) * G + r.c2;
    r.xs = smem_pad(stage_out_base<L>(0, r.t)) * G + r.c2;
  }
  // per tile: where the thread's line sits in the spectrum (only the filter needs it)
  EVX_HD static void set_tile(Regs& r, int row, int kz0) {
    r.kz = kz0 + r.col;
    r.kother = row;
  }

  EVX_HD static cf* tile_at(unsigned char* tile, int byte_off) {
    return reinterpret_cast<cf*>(tile + byte_off);
  }
  EVX_HD static void read_tile_natural(Regs& r, unsigned char* tile) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = *tile_at(tile, r.tb + e * T * ROWB);
  }
  EVX_HD static void write_tile_natural(Regs& r, unsigned char* tile) {
#pragma unroll
    for (int e = 0; e < 8; ++e) *tile_at(tile, r.tb + e * T * ROWB) = r.v[e];
  }
  // output of stage 1 (Ns = 8): element e belongs at row 64*(t/8) + t%8 + 8 e
  EVX_HD static void write_tile_stage1(Regs& r, unsigned char* tile) {
#pragma unroll
    for (int e = 0; e < 8; ++e) *tile_at(tile, r.sb + stage_out_const<L>(1, e) * ROWB) = r.v[e];
  }
  EVX_HD static void read_x_natural(Regs& r, const cf* xg) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = xg[r.xn + smem_pad(e * T) * G];
  }
  // output of stage 0 (Ns = 1): element e belongs at index 8 t + e
  EVX_HD static void write_x_stage0(Regs& r, cf* xg) {
#pragma unroll
    for (int e = 0; e < 8; ++e) xg[r.xs + smem_pad(stage_out_const<L>(0, e)) * G] = r.v[e];
  }

  EVX_HD static void apply_filter(Regs& r, const LineParams& p) {
    xmid_apply_filter<MODE, T>(r.v, r.t, r.kother, r.kz, p.filt);
  }

  EVX_HD static void fetch_tw(Regs& r, const LineParams& p, int stage) {
    stage_twiddles<L>(stage, r.t, p.tw, r.w);
  }

  // Phase k of a tile; the caller synchronises the GROUP between phases.  `tile` is the block's
  // current tile buffer, `xg` the group's exchange buffer.
  EVX_HD static void phase(int k, Regs& r, unsigned char* tile, cf* xg, const LineParams& p) {
    constexpr int DIR = MODE == PASS_INV ? +1 : -1;      // direction of the first transform
    if (k == 0) {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, DIR>(0, r.v, r.t, r.w);
      write_x_stage0(r, xg);
      fetch_tw(r, p, 1);
    } else if (k == 1) {
      read_x_natural(r, xg);
      line_stage_compute_pre<L, DIR>(1, r.v, r.t, r.w);
      write_tile_stage1(r, tile);
      fetch_tw(r, p, 2);
    } else if (k == 2) {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, DIR>(2, r.v, r.t, r.w);
      if (!XMID) {
        write_tile_natural(r, tile);
      } else {
        apply_filter(r, p);
        line_stage_compute_pre<L, +1>(0, r.v, r.t, r.w);
        write_x_stage0(r, xg);
        fetch_tw(r, p, 1);
      }
    } else if (k == 3) {
      read_x_natural(r, xg);
      line_stage_compute_pre<L, +1>(1, r.v, r.t, r.w);
      write_tile_stage1(r, tile);
      fetch_tw(r, p, 2);
    } else {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, +1>(2, r.v, r.t, r.w);
      write_tile_natural(r, tile);
    }
  }

  // Same phases with the roots of the two twiddled stages supplied by the caller - a persistent
  // kernel with registers to spare (one block per SM: fft_chain.cu, fft_line_ws_kernel) fetches
  // them once instead of once per stage and tile.  w1 / w2: stage_twiddles<L>(1 / 2, t, tw, .);
  // the inverse stages of the x pass use the same roots (conjugated inside the butterfly).
  EVX_HD static void phase_tw(int k, Regs& r, unsigned char* tile, cf* xg, const LineParams& p,
                              const cf* w1, const cf* w2) {
    constexpr int DIR = MODE == PASS_INV ? +1 : -1;      // direction of the first transform
    if (k == 0) {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, DIR>(0, r.v, r.t, w1);
      write_x_stage0(r, xg);
    } else if (k == 1) {
      read_x_natural(r, xg);
      line_stage_compute_pre<L, DIR>(1, r.v, r.t, w1);
      write_tile_stage1(r, tile);
    } else if (k == 2) {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, DIR>(2, r.v, r.t, w2);
      if (!XMID) {
        write_tile_natural(r, tile);
      } else {
        apply_filter(r, p);
        line_stage_compute_pre<L, +1>(0, r.v, r.t, w1);
        write_x_stage0(r, xg);
      }
    } else if (k == 3) {
      read_x_natural(r, xg);
      line_stage_compute_pre<L, +1>(1, r.v, r.t, w1);
      write_tile_stage1(r, tile);
    } else {
      read_tile_natural(r, tile);
      line_stage_compute_pre<L, +1>(2, r.v, r.t, w2);
      write_tile_natural(r, tile);
    }
  }

  // ---- what the tensor copies do, restated for the host replay (tests/emu) ---------------
  // element (kz, y, x) of the spectrum; tile (row, kz0): rows run along y (along_x = 0, x = row)
  // or along x (along_x = 1, y = row).  Out-of-range columns read as zero and are not written.
  EVX_HD static long long spec_index(const LineParams& p, int row, int kz, int i) {
    return p.along_x ? ((long long)i * p.ny + row) * p.P + kz : ((long long)row * p.ny + i) * p.P + kz;
  }
  static void host_tile_load(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        *tile_at(tile, tile_off(i, c)) = kz < p.ncols_valid ? p.spec[spec_index(p, row, kz, i)] : cf{0.f, 0.f};
      }
  }
  static void host_tile_store(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        if (kz < p.ncols_valid) p.spec[spec_index(p, row, kz, i)] = *tile_at(tile, tile_off(i, c));
      }
  }
};

// ---------------------------------------------------------------------------------------------
// 1024-point lines: four stages (radix 2, 8, 8, 8 - the stages, roots and thread -> butterfly
// assignment of StridedPass<1024>, hence the same bits), T = 128 threads per line.  A tile is
// [1024 rows][8 columns] (64-byte rows, 64 KB); 1024 compute threads would leave no room for the
// loader / retirer warps of the warp-specialised kernel, so a block has 512 compute threads = two
// groups of 256 (two neighbouring lines each, own named barrier) and every group transforms TWO
// line pairs of a tile one after the other (columns 2g, 2g+1, then 4+2g, 4+2g+1: the same tile
// byte offsets with bit 5 flipped).  The tile image is touched in natural order only (first read,
// last write: conflict-free under the 64-byte TMA swizzle); the three (XMID: six) exchanges
// alternate between two padded buffers of the group, one barrier each.  Phase k reads the buffer
// phase k-1 wrote and writes the other one, so the only hazard left - the first write of the next
// line pair - is covered by starting that pair on the buffer the last phase did NOT read (4
// phases: the pairs alternate; 7 phases: always buffer 0).
// ---------------------------------------------------------------------------------------------
template <int L, int KZ, int MODE, int GL = 2>
struct StridedLine4 {
  static_assert(L == 1024 && KZ == 8, "four-stage TMA-tiled pass: 1024-point lines, 64-byte tile rows");
  static_assert(GL == 1 || GL == 2, "lines per synchronisation group");
  static constexpr int LEN = L;
  static constexpr int COLS = KZ;
  static constexpr int T = L / 8;
  static constexpr int G = GL;                 // 2: conflict-free 16-byte pairs, two groups of 8 warps;
                                               // 1: four groups of 4 warps (2-way conflicts on the tile rows)
  static constexpr int GT = T * G;
  static constexpr int NPASS = 2;              // line pairs per group and tile
  static constexpr int NG = KZ / (G * NPASS);
  static constexpr int NTHREADS = GT * NG;
  static constexpr int S = 4;
  static constexpr bool XMID = pass_is_xmid(MODE);
  static constexpr int NPHASES = XMID ? 2 * S - 1 : S;
  static constexpr int ROWB = KZ * (int)sizeof(cf);
  static constexpr int TILE_BYTES = L * ROWB;
  static constexpr int LP = smem_padded_len(L);
  static constexpr int XG = LP * G;            // cf elements of ONE exchange buffer of a group
  static constexpr int XSTRIDE = 2 * XG;       // cf elements of a group's exchange space
  static constexpr int X_BYTES = NG * XSTRIDE * (int)sizeof(cf);

  struct Roots { cf w[3][3]; };                // roots of stages 1, 2, 3 of this thread
  struct Regs {
    cf v[8];
    int t, c2, g, col;       // position in the line, line of the pair, group, column of pair 0
    int tb;                  // byte offset of tile element (row t, col)
    int xn, xs[3];           // cf index in X: natural order base / output base of stages 0, 1, 2
    int kz, kother;          // XMID: global column of pair 0 and index along the other strided axis
  };

  EVX_HD static int swz(int r) { return ((r >> 1) & 3) << 4; }
  EVX_HD static int tile_off(int r, int c) { return r * ROWB + ((c * (int)sizeof(cf)) ^ swz(r)); }

  EVX_HD static void init(Regs& r, int tid) {
    r.g = tid / GT;
    const int tg = tid - r.g * GT;
    r.c2 = tg % G;
    r.t = tg / G;
    r.col = r.g * G + r.c2;
    r.tb = tile_off(r.t, r.col);
    r.xn = smem_pad(r.t) * G + r.c2;
#pragma unroll
    for (int s = 0; s < 3; ++s) r.xs[s] = smem_pad(stage_out_base<L>(s, r.t)) * G + r.c2;
  }
  EVX_HD static void load_roots(Roots& w, int t, const cf* tw) {
#pragma unroll
    for (int s = 1; s < S; ++s) stage_twiddles<L>(s, t, tw, w.w[s - 1]);
  }
  EVX_HD static void set_tile(Regs& r, int kother, int kz0) {
    r.kz = kz0 + r.col;
    r.kother = kother;
  }
  EVX_HD static cf* tile_at(unsigned char* tile, int byte_off) {
    return reinterpret_cast<cf*>(tile + byte_off);
  }
  // pair `pass` of the group: columns + 4 = byte offset bit 5 flipped (both group shapes: pass 0
  // covers columns 0..3, pass 1 columns 4..7)
  EVX_HD static void read_tile(Regs& r, unsigned char* tile, int pass) {
    const int tb = r.tb ^ (pass << 5);
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = *tile_at(tile, tb + e * T * ROWB);
  }
  EVX_HD static void write_tile(Regs& r, unsigned char* tile, int pass) {
    const int tb = r.tb ^ (pass << 5);
#pragma unroll
    for (int e = 0; e < 8; ++e) *tile_at(tile, tb + e * T * ROWB) = r.v[e];
  }
  EVX_HD static void read_x(Regs& r, const cf* xb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = xb[r.xn + smem_pad(e * T) * G];
  }
  template <int STAGE>
  EVX_HD static void write_x(Regs& r, cf* xb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) xb[r.xs[STAGE] + smem_pad(stage_out_const<L>(STAGE, e)) * G] = r.v[e];
  }
  template <int STAGE, int DIR>
  EVX_HD static void stage(Regs& r, const Roots& w) {
    if constexpr (STAGE == 0) line_stage_compute_pre<L, DIR>(0, r.v, r.t, nullptr);
    else line_stage_compute_pre<L, DIR>(STAGE, r.v, r.t, w.w[STAGE - 1]);
  }

  // Phase k of line pair `pass` of a tile; the caller synchronises the GROUP between phases
  // (not between the pairs).  xg: the group's two exchange buffers (XG elements each).
  EVX_HD static void phase(int pass, int k, Regs& r, unsigned char* tile, cf* xg, const LineParams& p,
                           const Roots& w) {
    constexpr int DIR = MODE == PASS_INV ? +1 : -1;      // direction of the first transform
    const int b0 = XMID ? 0 : (pass & 1);
    cf* xw = xg + ((b0 ^ (k & 1)) ? XG : 0);             // written by phase k
    const cf* xr = xg + ((b0 ^ (k & 1)) ? 0 : XG);       // written by phase k - 1
    if (k == 0) {
      read_tile(r, tile, pass);
      stage<0, DIR>(r, w);
      write_x<0>(r, xw);
    } else if (k == 1) {
      read_x(r, xr);
      stage<1, DIR>(r, w);
      write_x<1>(r, xw);
    } else if (k == 2) {
      read_x(r, xr);
      stage<2, DIR>(r, w);
      write_x<2>(r, xw);
    } else if (k == 3) {
      read_x(r, xr);
      stage<3, DIR>(r, w);
      if (!XMID) {
        write_tile(r, tile, pass);
      } else {
        xmid_apply_filter<MODE, T>(r.v, r.t, r.kother, r.kz + pass * (G * NG), p.filt);
        stage<0, +1>(r, w);
        write_x<0>(r, xw);
      }
    } else if (k == 4) {
      read_x(r, xr);
      stage<1, +1>(r, w);
      write_x<1>(r, xw);
    } else if (k == 5) {
      read_x(r, xr);
      stage<2, +1>(r, w);
      write_x<2>(r, xw);
    } else {
      read_x(r, xr);
      stage<3, +1>(r, w);
      write_tile(r, tile, pass);
    }
  }

  // ---- the tensor copies restated for the host replay: plain [nx][ny][P] spectrum ----------
  EVX_HD static long long spec_index(const LineParams& p, int row, int kz, int i) {
    return p.along_x ? ((long long)i * p.ny + row) * p.P + kz : ((long long)row * p.ny + i) * p.P + kz;
  }
  static void host_tile_load(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        *tile_at(tile, tile_off(i, c)) = kz < p.ncols_valid ? p.spec[spec_index(p, row, kz, i)] : cf{0.f, 0.f};
      }
  }
  static void host_tile_store(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        if (kz < p.ncols_valid) p.spec[spec_index(p, row, kz, i)] = *tile_at(tile, tile_off(i, c));
      }
  }
};

// ---------------------------------------------------------------------------------------------
// 1024-point lines, SIXTEEN points per thread (T = 64 threads per line).  The FFT kernels are
// bound by the shared-memory pipe, and StridedLine4 sends every point through
// shared memory 14 times in the x pass.  Here the radix-2 stage and the first radix-8 stage are
// done together in registers - a thread that holds in[q + 64 e'], e' < 16, owns the eight radix-2
// butterflies q + 64 e AND the two radix-8 butterflies 2q, 2q + 1 that consume their outputs - and
// the two remaining radix-8 stages run two butterflies per thread (jv = q, q + 64).  Every output is
// produced by the same operations in the same order as in StridedPass<1024> (same roots from the
// same table), so the results are bit-identical; but a line crosses shared memory only twice per
// transform, the block has the geometry of the 512-point kernel (512 compute threads = four
// two-line groups, all eight lines of a tile in flight), and the exchanges alternate between ONE
// padded buffer per group and the group's own tile columns:
//   k0  tile (natural) -> stages 0+1 -> X at 16 q + k        (index i + (i >> 4), lines interleaved)
//   k1  X (natural)    -> stage 2    -> tile rows (q/16) 128 + q%16 + 512 i + 16 r  (in place)
//   k2  tile (natural) -> stage 3    -> tile (natural)       [x pass: weight, inverse stages 0+1 -> X]
//   k3, k4  the inverse transform of the x pass, same pattern.
// Three (x pass: five) phases per tile instead of eight (fourteen); ten instead of fourteen trips
// through shared memory in the x pass.  All accesses are 64-bit and conflict-free per half-warp
// (8 consecutive q x 2 lines), one base register plus an immediate each.
// ---------------------------------------------------------------------------------------------
template <int L, int KZ, int MODE>
struct StridedLine16 {
  static_assert(L == 1024 && KZ == 8, "sixteen-point form: 1024-point lines, 64-byte tile rows");
  static constexpr int LEN = L;
  static constexpr int COLS = KZ;
  static constexpr int T = L / 16;             // 64 threads per line
  static constexpr int G = 2;
  static constexpr int GT = T * G;             // 128
  static constexpr int NPASS = 1;
  static constexpr int NG = KZ / G;            // 4
  static constexpr int NTHREADS = GT * NG;     // 512
  static constexpr bool XMID = pass_is_xmid(MODE);
  static constexpr int NPHASES = XMID ? 5 : 3;
  static constexpr int ROWB = KZ * (int)sizeof(cf);
  static constexpr int TILE_BYTES = L * ROWB;
  static constexpr int LP = L + (L >> 4);      // padded line length: index i + (i >> 4)
  static constexpr int XG = LP * G;
  static constexpr int XSTRIDE = XG;
  static constexpr int X_BYTES = NG * XSTRIDE * (int)sizeof(cf);

  // roots of this thread's butterflies: stage 2 (both butterflies share them), stage 3 (jv = q and
  // q + 64); the roots of the merged stage are the same for every thread and come from the table
  struct Roots { cf w2[3], w3[2][3]; };
  struct Regs {
    cf v[16];
    int t, c2, g, col;
    int tb;                  // byte offset of tile element (row q, col)
    int sb;                  // byte offset of tile element (row (q/16) 128 + q%16, col): stage-2 output base
    int xn, xs;              // cf index in X: natural-order base / output base of the merged stage
    int kz, kother;
  };

  EVX_HD static int pad16(int i) { return i + (i >> 4); }
  EVX_HD static int swz(int r) { return ((r >> 1) & 3) << 4; }
  EVX_HD static int tile_off(int r, int c) { return r * ROWB + ((c * (int)sizeof(cf)) ^ swz(r)); }

  EVX_HD static void init(Regs& r, int tid) {
    r.g = tid / GT;
    const int tg = tid - r.g * GT;
    r.c2 = tg % G;
    r.t = tg / G;
    r.col = r.g * G + r.c2;
    r.tb = tile_off(r.t, r.col);
    r.sb = tile_off((r.t / 16) * 128 + r.t % 16, r.col);
    r.xn = pad16(r.t) * G + r.c2;
    r.xs = pad16(16 * r.t) * G + r.c2;
  }
  EVX_HD static void load_roots(Roots& w, int t, const cf* tw) {
    stage_twiddles<L>(2, t, tw, w.w2);                 // m = (q % 16) * 8 for jv = q and q + 64
    stage_twiddles<L>(3, t, tw, w.w3[0]);              // m = q
    stage_twiddles<L>(3, t + T, tw, w.w3[1]);          // m = q + 64
  }
  EVX_HD static void set_tile(Regs& r, int kother, int kz0) {
    r.kz = kz0 + r.col;
    r.kother = kother;
  }
  EVX_HD static cf* tile_at(unsigned char* tile, int byte_off) {
    return reinterpret_cast<cf*>(tile + byte_off);
  }
  EVX_HD static void read_tile(Regs& r, unsigned char* tile) {
#pragma unroll
    for (int e = 0; e < 16; ++e) r.v[e] = *tile_at(tile, r.tb + e * T * ROWB);
  }
  EVX_HD static void write_tile(Regs& r, unsigned char* tile) {
#pragma unroll
    for (int e = 0; e < 16; ++e) *tile_at(tile, r.tb + e * T * ROWB) = r.v[e];
  }
  EVX_HD static void read_x(Regs& r, const cf* xg) {
#pragma unroll
    for (int e = 0; e < 16; ++e) r.v[e] = xg[r.xn + pad16(e * T) * G];
  }

  // stages 0 + 1 on v[e'] = in[q + 64 e']; the outputs go to X at 16 q + b + 2 r
  template <int DIR>
  EVX_HD static void merged_to_x(Regs& r, cf* xg, const cf* tw) {
#pragma unroll
    for (int e = 0; e < 8; ++e) dft2<DIR>(r.v[e], r.v[e + 8]);        // butterflies q + 64 e of stage 0
    const cf wa[3] = {tw[0], tw[0], tw[0]};                             // jv = 2 q: m = 0
    line_stage_compute_pre<L, DIR>(1, r.v, 0, wa);
    const cf wb[3] = {tw[L / 16], tw[2 * (L / 16)], tw[4 * (L / 16)]};  // jv = 2 q + 1: m = 64
    line_stage_compute_pre<L, DIR>(1, r.v + 8, 0, wb);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int k = 0; k < 8; ++k) xg[r.xs + (b + 2 * k) * G] = r.v[8 * b + k];
  }
  // stage 2 on v (natural order): butterfly jv = q + 64 i takes v[2 e + i]; outputs to the tile rows
  template <int DIR>
  EVX_HD static void stage2_to_tile(Regs& r, unsigned char* tile, const Roots& w) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      cf a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = r.v[2 * e + i];
      line_stage_compute_pre<L, DIR>(2, a, 0, w.w2);
#pragma unroll
      for (int k = 0; k < 8; ++k) *tile_at(tile, r.sb + (512 * i + 16 * k) * ROWB) = a[k];
    }
  }
  // stage 3 in registers: v (natural order in) -> v (natural order out: jv + 128 r = q + 64 (i + 2 r));
  // FILTER: the weight of the x pass on the way (the points q + 64 i + 128 r are the eight points
  // t + e 128 of "thread" t = q + 64 i of the 128-thread forms - the very same routine)
  template <int DIR, bool FILTER>
  EVX_HD static void stage3(Regs& r, const Roots& w, const LineParams& p) {
    cf o[16];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      cf a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = r.v[2 * e + i];
      line_stage_compute_pre<L, DIR>(3, a, 0, w.w3[i]);
      if (FILTER) xmid_apply_filter<MODE, L / 8>(a, r.t + T * i, r.kother, r.kz, p.filt);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[i + 2 * k] = a[k];
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) r.v[e] = o[e];
  }

  EVX_HD static void phase(int /*pass*/, int k, Regs& r, unsigned char* tile, cf* xg, const LineParams& p,
                           const Roots& w) {
    constexpr int DIR = MODE == PASS_INV ? +1 : -1;      // direction of the first transform
    if (k == 0) {
      read_tile(r, tile);
      merged_to_x<DIR>(r, xg, p.tw);
    } else if (k == 1) {
      read_x(r, xg);
      stage2_to_tile<DIR>(r, tile, w);
    } else if (k == 2) {
      read_tile(r, tile);
      if (!XMID) {
        stage3<DIR, false>(r, w, p);
        write_tile(r, tile);
      } else {
        stage3<DIR, true>(r, w, p);
        merged_to_x<+1>(r, xg, p.tw);
      }
    } else if (k == 3) {
      read_x(r, xg);
      stage2_to_tile<+1>(r, tile, w);
    } else {
      read_tile(r, tile);
      stage3<+1, false>(r, w, p);
      write_tile(r, tile);
    }
  }

  // ---- the tensor copies restated for the host replay: plain [nx][ny][P] spectrum ----------
  EVX_HD static long long spec_index(const LineParams& p, int row, int kz, int i) {
    return p.along_x ? ((long long)i * p.ny + row) * p.P + kz : ((long long)row * p.ny + i) * p.P + kz;
  }
  static void host_tile_load(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        *tile_at(tile, tile_off(i, c)) = kz < p.ncols_valid ? p.spec[spec_index(p, row, kz, i)] : cf{0.f, 0.f};
      }
  }
  static void host_tile_store(const LineParams& p, int row, int kz0, unsigned char* tile) {
    for (int i = 0; i < L; ++i)
      for (int c = 0; c < KZ; ++c) {
        const int kz = kz0 + c;
        if (kz < p.ncols_valid) p.spec[spec_index(p, row, kz, i)] = *tile_at(tile, tile_off(i, c));
      }
  }
};

}  // namespace evx
