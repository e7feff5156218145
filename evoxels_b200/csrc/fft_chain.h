// Launcher of the chained z/y passes (fft_chain.cu), used by fft_native.cu.
#pragma once
#include <cuda_runtime.h>

namespace evx {

struct ChainArgs {
  int nx, ny, nz, P;
  const float* real_in;    // forward: r [nx][ny][nz]   inverse: u or null
  float* real_out;         // inverse: out
  void* spec;              // [nx][ny][P] complex64
  const void *twz, *twr, *twy;
  void* flags;             // >= nx unsigned ints of scratch (zeroed by the launcher)
  void* stats;             // kChainStatsBytes of scratch: per-block cycle counters (debug aid)
};

constexpr size_t kChainStatsBytes = 64 * 1024;   // up to 512 blocks x 16 counters

// 512-point y and z lines on a driver that can encode tensor maps
bool chain_supported(int nx, int ny, int nz);
// tmap_y: tensor map of the spectrum with [8 x 256 x 1] boxes (line_make_tmap, along_x = 0, kz = 8)
// tmap_blk0 / tmap_blk1 (both or none): x-slab plan with ny / W = 256 - the y tiles of rows [0, 256) /
// [256, 512) are stored to (forward) or loaded from (inverse) these [8 x 256 x 1]-box maps over the
// blocks of the all-to-all layout instead of the spectrum
int chain_launch(bool inverse, const ChainArgs& a, const void* tmap_y, cudaStream_t st,
                 const void* tmap_blk0 = nullptr, const void* tmap_blk1 = nullptr);

}  // namespace evx
