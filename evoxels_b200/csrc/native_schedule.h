// Launch schedule of the native IMEX pipeline (host only, no CUDA calls).
//
// The five passes  ZFwd -> Y fwd -> X fwd*filter*inv -> Y inv -> ZInv(+u)  each stream the
// whole field through HBM when launched once per pass.  The z and y passes of one x plane
// touch the same 1 MB of spectrum, so the L2-blocked schedule walks x in chunks of a few
// planes and runs the two passes of a pair back to back on each chunk: the second pass finds
// the chunk the first one wrote still in the 126 MB L2.  Per voxel that removes one spectrum
// write + read per direction (52 -> 36 B, and 28 B with the two ring options below):
//   * EVX_SCHED_RING_INV: the inverse y pass writes a chunk-sized ring slot instead of the
//     spectrum; the slot is overwritten every other chunk and therefore never written back;
//   * EVX_SCHED_CHUNK_RHS: (fused CH step) the rhs kernel runs per chunk as well, into a
//     chunk-sized slot that the z pass consumes at once.
// With two streams the second pass of chunk i runs on a side stream next to the first pass
// of chunk i+1, which fills the tail of every short launch; with three, rhs, z and y of
// consecutive chunks overlap (the issue-bound stencil next to the memory-bound passes).
//
// This header turns (grid, chunk size, options) into a flat list of operations; fft_native.cu
// executes the list with kernel launches and events, the CPU replay in tests/emu executes the
// same list serially, and tests/test_emulated_kernels.py checks on the list itself that every
// pair of operations touching the same memory is ordered by stream order or an event.
#pragma once
#include <vector>
#include "fft_pass_core.h"

namespace evx {

enum : int { SCHED_RING_INV = 1, SCHED_CHUNK_RHS = 2 };   // == EVX_SCHED_* in the C header
constexpr int kMaxChunkPlanes = 32;                      // capacity of a ring slot

enum : int { OP_RHS = 0, OP_ZFWD, OP_YFWD, OP_XMID, OP_YINV, OP_ZINV, OP_RECORD, OP_WAIT };

struct SchedOp {
  int kind;
  int stream;   // 0: the caller's stream, 1 / 2: the plan's side streams
  int x0, nxc;  // x planes [x0, x0+nxc) (compute ops)
  int slot;     // ring slot written (OP_YINV, OP_RHS) / read (OP_ZINV, OP_ZFWD); -1: in place
  int event;    // OP_RECORD / OP_WAIT: event index < kSchedEvents
};

struct NativeDims {
  int nx, ny, nz, M, P;
};

// events: 0 fork, 1 "z forward of a chunk done", 2 join of stream 1 after the forward half,
//         3 "y inverse of a chunk done", 4/5 "z inverse has consumed ring slot 0/1",
//         6/7 "rhs of a chunk is in rhs slot 0/1", 8/9 "z forward has consumed rhs slot 0/1",
//         10 join of stream 2 after the forward half
constexpr int kSchedEvents = 11;
constexpr int kSchedStreams = 3;   // 0: the caller's stream, 1 and 2: owned by the plan

// streams = 1: everything on the caller's stream.  2: the y passes on stream 1.  3 (only with
// per-chunk rhs, else like 2): rhs on the caller's stream, z forward on stream 2, y forward on
// stream 1 - rhs(i+1), z(i) and y(i-1) overlap, the issue-bound stencil kernel next to the
// memory-bound transform passes; the rhs then alternates between two slots.
inline void build_schedule(int nx, int chunk_planes, int streams, int flags, int ring_planes,
                           bool with_rhs, std::vector<SchedOp>& ops) {
  ops.clear();
  const int X = (chunk_planes > 0 && chunk_planes < nx) ? chunk_planes : nx;
  const bool chunked = X < nx;
  const bool two = chunked && streams >= 2;
  const bool ring = chunked && (flags & SCHED_RING_INV) && X <= ring_planes;
  // a chunk's rhs needs the two planes on either side as one contiguous pair
  const bool rhs_chunks = with_rhs && chunked && (flags & SCHED_CHUNK_RHS) && X >= 2 && nx % X != 1;
  const bool three = two && streams == 3 && rhs_chunks && 2 * X <= nx;
  const int side = two ? 1 : 0;
  const int zstream = three ? 2 : 0;
  auto op = [&](int kind, int stream, int x0, int nxc, int slot, int event) {
    ops.push_back(SchedOp{kind, stream, x0, nxc, slot, event});
  };
  auto signal = [&](int from, int to, int event) {     // work enqueued on `to` after this point
    op(OP_RECORD, from, 0, 0, -1, event);              // waits for the work on `from` before it
    op(OP_WAIT, to, 0, 0, -1, event);
  };
  // ---- forward half -------------------------------------------------------------------
  if (with_rhs && !rhs_chunks) op(OP_RHS, 0, 0, nx, -1, -1);
  if (two) signal(0, 1, 0);
  if (three) op(OP_WAIT, 2, 0, 0, -1, 0);
  int i = 0;
  for (int x0 = 0; x0 < nx; x0 += X, ++i) {
    const int nxc = nx - x0 < X ? nx - x0 : X;
    const int rslot = rhs_chunks ? (three ? (i & 1) : 0) : -1;
    if (rhs_chunks) {
      if (three && i >= 2) op(OP_WAIT, 0, 0, 0, -1, 8 + rslot);   // slot free again
      op(OP_RHS, 0, x0, nxc, rslot, -1);
      if (three) signal(0, 2, 6 + rslot);
    }
    op(OP_ZFWD, zstream, x0, nxc, rslot, -1);
    if (three) {
      op(OP_RECORD, 2, 0, 0, -1, 8 + rslot);
      op(OP_WAIT, 1, 0, 0, -1, 8 + rslot);
    } else if (two) {
      signal(0, 1, 1);
    }
    op(OP_YFWD, side, x0, nxc, -1, -1);
  }
  if (three) signal(2, 0, 10);
  if (two) signal(1, 0, 2);
  // ---- x pass -------------------------------------------------------------------------
  op(OP_XMID, 0, 0, nx, -1, -1);
  // ---- inverse half -------------------------------------------------------------------
  if (two) signal(0, 1, 0);
  i = 0;
  for (int x0 = 0; x0 < nx; x0 += X, ++i) {
    const int nxc = nx - x0 < X ? nx - x0 : X;
    const int slot = ring ? (i & 1) : -1;
    if (two && ring && i >= 2) op(OP_WAIT, 1, 0, 0, -1, 4 + (i & 1));
    op(OP_YINV, side, x0, nxc, slot, -1);
    if (two) signal(1, 0, 3);
    op(OP_ZINV, 0, x0, nxc, slot, -1);
    if (two && ring) op(OP_RECORD, 0, 0, 0, -1, 4 + (i & 1));
  }
}

// Buffers of one application.  `r` is the caller's right-hand side (apply) or the plan's
// real scratch buffer (fused step); with chunked rhs the slot is the start of that scratch.
struct NativeBufs {
  const float* u;        // may be null (update only)
  const float* r;
  float* out;
  cf* spec;              // [nx][ny][P]
  cf* ring;              // 2 slots of ring_slot_elems (null: no ring)
  long long ring_slot_elems;
  long long rhs_slot_elems;   // per-chunk rhs: slot s starts at r + s * rhs_slot_elems
  const cf *twx, *twy, *twz, *twr;
};

inline ZParams z_chunk_params(const NativeDims& d, const NativeBufs& b, const SchedOp& o) {
  const long long plane_r = (long long)d.ny * d.nz, plane_s = (long long)d.ny * d.P;
  ZParams zp;
  zp.tw = b.twz; zp.twr = b.twr; zp.nz = d.nz; zp.P = d.P;
  zp.rows = (long long)o.nxc * d.ny;
  if (o.kind == OP_ZFWD) {
    zp.real_in = o.slot >= 0 ? b.r + o.slot * b.rhs_slot_elems : b.r + o.x0 * plane_r;
    zp.real_out = nullptr;
    zp.spec = b.spec + o.x0 * plane_s;
  } else {
    zp.real_in = b.u ? b.u + o.x0 * plane_r : nullptr;
    zp.real_out = b.out + o.x0 * plane_r;
    zp.spec = o.slot >= 0 ? b.ring + o.slot * b.ring_slot_elems : b.spec + o.x0 * plane_s;
  }
  return zp;
}

inline StridedParams y_chunk_params(const NativeDims& d, const NativeBufs& b, const SchedOp& o) {
  const long long plane_s = (long long)d.ny * d.P;
  StridedParams yp;
  yp.in = b.spec + o.x0 * plane_s;
  yp.out = (o.kind == OP_YINV && o.slot >= 0) ? b.ring + o.slot * b.ring_slot_elems
                                              : b.spec + o.x0 * plane_s;
  yp.tw = b.twy;
  yp.src = yp.dst = plain_io(d.P, plane_s, d.ny);
  yp.P = d.P; yp.ncols_valid = d.M + 1; yp.ncols_total = (long long)o.nxc * d.P;
  yp.kother_offset = 0; yp.use_peers = 0; yp.max_ctas = 0; yp.dst_peer_base = 0;
  for (int i = 0; i < 8; ++i) yp.out_peers[i] = nullptr;
  yp.filt = FilterParams{};
  return yp;
}

inline StridedParams x_params(const NativeDims& d, const NativeBufs& b, const double* h, double dt,
                              double coef, int power) {
  SchedOp all{OP_YFWD, 0, 0, d.nx, -1, -1};
  StridedParams xp = y_chunk_params(d, b, all);
  xp.tw = b.twx;
  xp.src = xp.dst = plain_io((long long)d.ny * d.P, d.P, d.nx);
  xp.ncols_total = (long long)d.ny * d.P;
  const int n[3] = {d.nx, d.ny, d.nz};
  xp.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)d.nx * d.ny * d.nz));
  return xp;
}

// rhs of the x planes [x0, x0+nxc) of a periodic field as a slab with halo planes: pointers
// to the slab, its two planes below and its two planes above (periodic images at the ends)
struct RhsChunk {
  const float* c;
  const float* halo_lo;
  const float* halo_hi;
  float* out;
};
inline RhsChunk rhs_chunk(const NativeDims& d, const float* u, float* rhs, long long slot_elems,
                          const SchedOp& o) {
  const long long plane = (long long)d.ny * d.nz;
  RhsChunk k;
  if (o.nxc >= d.nx) { k.c = u; k.halo_lo = k.halo_hi = nullptr; k.out = rhs; return k; }
  const int e = o.x0 + o.nxc;
  k.c = u + o.x0 * plane;
  k.halo_lo = u + (o.x0 >= 2 ? o.x0 - 2 : d.nx - 2) * plane;
  k.halo_hi = u + (e + 2 <= d.nx ? e : 0) * plane;
  k.out = o.slot >= 0 ? rhs + o.slot * slot_elems : rhs + o.x0 * plane;
  return k;
}

}  // namespace evx
