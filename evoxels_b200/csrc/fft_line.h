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

}  // namespace evx
