// Pass parameters of the x-slab distributed FFT (host only, no CUDA calls).
//
// Rank r of W owns the x planes [r*nx/W, (r+1)*nx/W).  The y passes write / read the
// all-to-all block layout [W][nxl][nyl][P] directly (block j travels to / came from rank j), so
// there is no pack / unpack kernel around the transposes.  fft_native.cu launches the kernels
// with these parameters; the CPU replay (tests/emu) runs W virtual ranks through the very same
// parameters against the single-domain pipeline.
#pragma once
#include <vector>
#include "fft_pass_core.h"

namespace evx {

struct DistDims {
  int nx, ny, nz, world, rank, nxl, nyl, P, M;
};
struct DistTables {
  const cf *twx, *twy, *twz, *twr;     // W_nx, W_ny, W_M, W_nz[0..M]
};

inline DistDims make_dist_dims(int nx, int ny, int nz, int world, int rank, int pitch_align) {
  DistDims d;
  d.nx = nx; d.ny = ny; d.nz = nz; d.world = world; d.rank = rank;
  d.nxl = nx / world; d.nyl = ny / world; d.M = nz / 2;
  d.P = ((d.M + 1 + pitch_align - 1) / pitch_align) * pitch_align;
  return d;
}

// block layout [world][nxl][nyl][P] seen from local group xl: chunks nxl*nyl*P apart
inline StridedIO block_io(const DistDims& d) {
  return StridedIO{d.P, (long long)d.nyl * d.P, (long long)d.nxl * d.nyl * d.P, ilog2(d.nyl)};
}

inline void set_peers(StridedParams& sp, const DistDims& d, void* const* peers, int p2p_ctas) {
  sp.use_peers = peers ? 1 : 0;
  sp.max_ctas = peers ? p2p_ctas : 0;
  sp.dst_peer_base = (long long)d.rank * d.nxl * d.nyl * d.P;
  for (int i = 0; i < 8; ++i) sp.out_peers[i] = (peers && i < d.world) ? (cf*)peers[i] : nullptr;
}

// Pointer table for "every block into `other`, except block `rank` into `self_buf`" in terms
// of the peer-store addressing (which adds rank * block to the table entry).
inline void local_block_table(const DistDims& d, cf* other, cf* self_buf, void** table) {
  const long long blk = (long long)d.nxl * d.nyl * d.P;
  for (int j = 0; j < 8; ++j)
    table[j] = j >= d.world ? nullptr
                            : (j == d.rank ? (void*)self_buf : (void*)(other + (j - d.rank) * blk));
}

// ---- forward: z pass and y pass of the local x planes [x0, x0+nxc) ------------------------
// (r_local / spec / send point at the start of the full local arrays)
inline ZParams dist_zfwd_params(const DistDims& d, const DistTables& t, const float* r_local,
                                cf* spec, int x0, int nxc) {
  ZParams zp;
  zp.real_in = r_local + (long long)x0 * d.ny * d.nz; zp.real_out = nullptr;
  zp.spec = spec + (long long)x0 * d.ny * d.P; zp.tw = t.twz; zp.twr = t.twr;
  zp.rows = (long long)nxc * d.ny; zp.nz = d.nz; zp.P = d.P; zp.pf_blocks = 0;
  return zp;
}
// peers != null: block `rank` of every peer's buffer is written directly (peer stores)
inline StridedParams dist_yfwd_params(const DistDims& d, const DistTables& t, cf* spec, cf* send,
                                      void* const* peers, int p2p_ctas, int x0, int nxc) {
  StridedParams yp;
  yp.in = spec + (long long)x0 * d.ny * d.P; yp.tw = t.twy;
  yp.src = plain_io(d.P, (long long)d.ny * d.P, d.ny);
  yp.dst = block_io(d);
  yp.out = send ? send + (long long)x0 * yp.dst.plane_stride : nullptr;
  yp.P = d.P; yp.ncols_valid = d.M + 1; yp.ncols_total = (long long)nxc * d.P;
  yp.kother_offset = 0; yp.filt = FilterParams{};
  set_peers(yp, d, peers, p2p_ctas);
  yp.dst_peer_base += (long long)x0 * yp.dst.plane_stride;
  return yp;
}

// ---- middle: x forward * P(k)/N * x inverse on the local y-pencil rows [yl0, yl0+nylc) -----
inline StridedParams dist_xmid_params(const DistDims& d, const DistTables& t, cf* recv,
                                      void* const* peers, int p2p_ctas, const double* h, double dt,
                                      double coef, int power, int yl0, int nylc) {
  recv += (long long)yl0 * d.P;
  StridedParams xp;
  xp.in = recv; xp.out = recv; xp.tw = t.twx;
  xp.src = xp.dst = plain_io((long long)d.nyl * d.P, d.P, d.nx);
  set_peers(xp, d, peers, p2p_ctas);
  if (peers) {   // chunk x / nxl of every x line goes to that rank: [rank j block][xl][yl][kz]
    xp.dst = StridedIO{(long long)d.nyl * d.P, d.P, 0, ilog2(d.nxl)};
    xp.dst_peer_base += (long long)yl0 * d.P;
  }
  xp.P = d.P; xp.ncols_valid = d.M + 1; xp.ncols_total = (long long)nylc * d.P;
  xp.kother_offset = d.rank * d.nyl + yl0;
  const int n[3] = {d.nx, d.ny, d.nz};
  xp.filt = make_filter(n, h, dt, coef, power, 1.0 / ((double)d.nx * d.ny * d.nz));
  return xp;
}

// ---- backward: y inverse (block layout -> plain spectrum at spec_chunk) and z inverse (+u) of
// the local x planes [x0, x0+nxc) -----------------------------------------------------------
inline StridedParams dist_yinv_params(const DistDims& d, const DistTables& t, const cf* recv,
                                      cf* spec_chunk, int x0, int nxc) {
  StridedParams yp;
  yp.src = block_io(d);
  yp.dst = plain_io(d.P, (long long)d.ny * d.P, d.ny);
  yp.in = recv + (long long)x0 * yp.src.plane_stride; yp.out = spec_chunk; yp.tw = t.twy;
  yp.P = d.P; yp.ncols_valid = d.M + 1; yp.ncols_total = (long long)nxc * d.P;
  yp.kother_offset = 0; yp.filt = FilterParams{};
  set_peers(yp, d, nullptr, 0);
  return yp;
}
inline ZParams dist_zinv_params(const DistDims& d, const DistTables& t, cf* spec_chunk,
                                const float* u_local, float* out_local, int x0, int nxc) {
  const long long real_off = (long long)x0 * d.ny * d.nz;
  ZParams zp;
  zp.real_in = u_local ? u_local + real_off : nullptr; zp.real_out = out_local + real_off;
  zp.spec = spec_chunk; zp.tw = t.twz; zp.twr = t.twr;
  zp.rows = (long long)nxc * d.ny; zp.nz = d.nz; zp.P = d.P; zp.pf_blocks = 0;
  return zp;
}

}  // namespace evx
