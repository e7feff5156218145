// Launcher of the TMA-tiled strided passes (fft_line.cu), used by fft_native.cu.
#pragma once
#include <cuda_runtime.h>
#include "fft_line_core.h"

namespace evx {

constexpr size_t kTensorMapBytes = 128;   // sizeof(CUtensorMap), 64-byte aligned

// false when the driver offers no cuTensorMapEncodeTiled (then the cp.async passes are used)
bool line_pass_available();
// tensor map over the valid columns of a [nx][ny][P] half spectrum with a [512 x kz] box along
// y (along_x = 0) or x (along_x = 1); `out` points at kTensorMapBytes bytes, 64-byte aligned
int line_make_tmap(void* out, void* spec, int nx, int ny, int P, int ncols_valid, int along_x, int kz);
// mode: PASS_FWD / PASS_INV / PASS_XMID / PASS_XMID_ETD1 of fft_pass_core.h
int line_pass_launch(int mode, int kz, const LineParams& p, const void* tmap, cudaStream_t st);


// ---- 1024-point lines (StridedLine4, fft_line4_ws_kernel) -------------------------------------
// 4-D tensor map (kz, i_lo, row, i_hi), strides in complex elements: line index i = i_hi *
// box_rows + i_lo, box = [box_rows][8 columns] of one row (box_rows <= 256, divides 1024)
int line_make_tmap4(void* out, void* base, int ncols_valid, int box_rows, long long lo_stride, int nrows,
                    long long row_stride, int nhi, long long hi_stride);
// p: tw, ncols_valid, row0, nrows, kother0, box_rows, out_div, max_ctas, l2_ahead, filt; loads through
// map_in, box h of a tile is stored through maps_out[h / out_div] at high coordinate h % out_div
int line4_pass_launch(int mode, const LineParams& p, const void* map_in, const void* const* maps_out, int nout,
                      cudaStream_t st);

}  // namespace evx
