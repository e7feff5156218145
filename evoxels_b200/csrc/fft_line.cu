// TMA-tiled strided FFT passes (y / x axis of the half spectrum) - kernel and launcher for the
// program in fft_line_core.h.  One thread per block drives the tensor copies:
//
//   full[b]   mbarrier, completes when the TMA load of the tile in buffer b has landed
//   done[b]   counter of groups that have finished the tile in buffer b; the thread that
//             completes the count issues the TMA store of the tile (bulk async-group) and,
//             once the store has read shared memory, the TMA load of the tile after next
//
// so HBM -> smem, the transform and smem -> HBM of consecutive tiles overlap inside one block
// without a dedicated producer warp (a 17th warp would not fit the 64-register budget at two
// blocks per SM).
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "evx_internal.h"
#include "fft_line_core.h"
#include "fft_line.h"
#include "tma_ptx.h"

namespace evx {

// ---- kernel -------------------------------------------------------------------------
template <class Prog>
__global__ void __launch_bounds__(Prog::NTHREADS, Prog::NTHREADS <= 512 ? 2 : 1)
    fft_line_kernel(const __grid_constant__ CUtensorMap tmap, const LineParams p) {
  extern __shared__ unsigned char smem_raw[];
  // swizzled TMA tiles want 1024-byte aligned shared memory
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = sm;
  cf* xall = reinterpret_cast<cf*>(sm + 2 * Prog::TILE_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + 2 * Prog::TILE_BYTES + Prog::X_BYTES);
  int* done = reinterpret_cast<int*>(full + 2);

  const int tid = threadIdx.x;
  typename Prog::Regs r;
  Prog::init(r, tid);
  cf* xg = xall + r.g * Prog::XG;
  const bool leader = (tid % Prog::GT) == 0;           // one signalling thread per group
  const int tpr = p.tiles_per_row;
  const long long ntiles = p.ntiles;
  const int nblk = gridDim.x;

  // tile -> (row, kz0); rows of a tile are BOX_ROWS-high boxes along y (along_x = 0) or x
  auto load_tile = [&](long long tile, int buf) {
    const int row = (int)(tile / tpr), kz0 = (int)(tile - (long long)row * tpr) * Prog::COLS;
    unsigned char* dst = tiles + buf * Prog::TILE_BYTES;
    mbar_expect_tx(&full[buf], Prog::TILE_BYTES);
#pragma unroll
    for (int h = 0; h < 512 / Prog::BOX_ROWS; ++h) {
      if (p.along_x) tma_load_3d(dst + h * Prog::BOX_ROWS * Prog::ROWB, &tmap, &full[buf], kz0, row, h * Prog::BOX_ROWS);
      else tma_load_3d(dst + h * Prog::BOX_ROWS * Prog::ROWB, &tmap, &full[buf], kz0, h * Prog::BOX_ROWS, row);
    }
  };
  // L2 prefetch of a tile that will be loaded `p.l2_ahead` hand-overs later: deepens the
  // memory pipeline beyond the two shared-memory buffers without costing shared memory
  auto prefetch_tile = [&](long long tile) {
    const int row = (int)(tile / tpr), kz0 = (int)(tile - (long long)row * tpr) * Prog::COLS;
#pragma unroll
    for (int h = 0; h < 512 / Prog::BOX_ROWS; ++h) {
      if (p.along_x) tma_prefetch_3d(&tmap, kz0, row, h * Prog::BOX_ROWS);
      else tma_prefetch_3d(&tmap, kz0, h * Prog::BOX_ROWS, row);
    }
  };
  auto store_tile = [&](long long tile, int buf) {
    const int row = (int)(tile / tpr), kz0 = (int)(tile - (long long)row * tpr) * Prog::COLS;
    const unsigned char* src = tiles + buf * Prog::TILE_BYTES;
#pragma unroll
    for (int h = 0; h < 512 / Prog::BOX_ROWS; ++h) {
      if (p.along_x) tma_store_3d(&tmap, src + h * Prog::BOX_ROWS * Prog::ROWB, kz0, row, h * Prog::BOX_ROWS);
      else tma_store_3d(&tmap, src + h * Prog::BOX_ROWS * Prog::ROWB, kz0, h * Prog::BOX_ROWS, row);
    }
    tma_store_commit();
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    done[0] = 0;
    done[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    if ((long long)blockIdx.x < ntiles) load_tile(blockIdx.x, 0);
    if ((long long)blockIdx.x + nblk < ntiles) load_tile((long long)blockIdx.x + nblk, 1);
    for (int a = 0; a < p.l2_ahead; ++a)
      if ((long long)blockIdx.x + (2LL + a) * nblk < ntiles) prefetch_tile((long long)blockIdx.x + (2LL + a) * nblk);
  }

  // tile cursor: (row, tcol) advanced by nblk tiles per iteration without dividing
  int row = blockIdx.x / tpr, tcol = blockIdx.x - row * tpr;
  const int step_row = nblk / tpr, step_col = nblk - step_row * tpr;
  long long pending_tile = -1;      // leader only: tile to load into pending_buf once the store
  int pending_buf = -1;             // issued by this thread has read shared memory

  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++it) {
    const int buf = it & 1;
    unsigned char* tb = tiles + buf * Prog::TILE_BYTES;
    Prog::set_tile(r, row, tcol * Prog::COLS);
    mbar_wait(&full[buf], (unsigned)(it >> 1) & 1u);
#pragma unroll
    for (int k = 0; k < Prog::NPHASES; ++k) {
      if (k) {
        group_sync(1 + r.g, Prog::GT);
        if (k == 1 && pending_buf >= 0) {       // deferred half of the previous tile's hand-over
          tma_store_wait_read();
          if (pending_tile >= 0) {
            load_tile(pending_tile, pending_buf);
            const long long ahead = pending_tile + (long long)p.l2_ahead * nblk;
            if (p.l2_ahead > 0 && ahead < ntiles) prefetch_tile(ahead);
          }
          pending_buf = -1;
        }
      }
      Prog::phase(k, r, tb, xg, p);
    }
    fence_proxy_async();                        // my tile writes -> visible to the TMA store
    group_sync(1 + r.g, Prog::GT);
    if (leader) {
      __threadfence_block();
      const int old = atomicAdd(&done[buf], 1);
      if (old == Prog::NG - 1) {                // last group: the tile is complete
        done[buf] = 0;
        __threadfence_block();
        fence_proxy_async();
        store_tile(tile, buf);
        pending_buf = buf;
        pending_tile = tile + 2LL * nblk < ntiles ? tile + 2LL * nblk : -1;
      }
    }
    row += step_row; tcol += step_col;
    if (tcol >= tpr) { tcol -= tpr; ++row; }
  }
  tma_store_wait_read();            // shared memory must outlive the stores this thread issued
}

// ---- warp-specialised form --------------------------------------------------------------
// One block per SM: 16 compute warps + a loader warp + one retirer warp per tile buffer (lane 0
// of each acts), NBUF tile buffers.  Compute threads wait for a tile (mbarrier full[b]), run the
// phases with their group barriers and arrive on done[b]; the loader issues the TMA loads up to
// NBUF tiles ahead as buffers come back (empty[b]); retirer b stores the tile in buffer b and
// frees the buffer once the store has read it.  No compute thread ever waits for a store to
// drain, and with the whole register file to itself a thread keeps the roots of its twiddled
// stages in registers for the whole kernel (the per-stage table loads were the top stall).
constexpr int line_ws_threads(int compute, int nbuf) { return compute + 32 + 32 * nbuf; }

__device__ __forceinline__ void mbar_arrive_cta(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class Prog, int NBUF>
__global__ void __launch_bounds__(line_ws_threads(Prog::NTHREADS, NBUF), 1)
    fft_line_ws_kernel(const __grid_constant__ CUtensorMap tmap, const LineParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = sm;
  cf* xall = reinterpret_cast<cf*>(sm + NBUF * Prog::TILE_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + NBUF * Prog::TILE_BYTES + Prog::X_BYTES);
  unsigned long long* done = full + NBUF;
  unsigned long long* empty = full + 2 * NBUF;
  const int tid = threadIdx.x;
  const int tpr = p.tiles_per_row;
  const long long ntiles = p.ntiles;
  const int nblk = gridDim.x;

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&done[b], Prog::NTHREADS);
      mbar_init(&empty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  if (tid == Prog::NTHREADS) {
    // ------------------------------------ loader ------------------------------------------
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
      const int buf = n % NBUF;
      if (n >= NBUF) mbar_wait(&empty[buf], (unsigned)(n / NBUF - 1) & 1u);
      const int row = (int)(tile / tpr), kz0 = (int)(tile - (long long)row * tpr) * Prog::COLS;
      unsigned char* dst = tiles + buf * Prog::TILE_BYTES;
      mbar_expect_tx(&full[buf], Prog::TILE_BYTES);
#pragma unroll
      for (int h = 0; h < 512 / Prog::BOX_ROWS; ++h) {
        if (p.along_x) tma_load_3d(dst + h * Prog::BOX_ROWS * Prog::ROWB, &tmap, &full[buf], kz0, row, h * Prog::BOX_ROWS);
        else tma_load_3d(dst + h * Prog::BOX_ROWS * Prog::ROWB, &tmap, &full[buf], kz0, h * Prog::BOX_ROWS, row);
      }
    }
    return;
  }
  if (tid > Prog::NTHREADS && (tid & 31) == 0) {
    // ------------------------------ retirer of buffer `me` ---------------------------------
    const int me = (tid - Prog::NTHREADS - 32) >> 5;
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
      if (n % NBUF != me) continue;
      mbar_wait(&done[me], (unsigned)(n / NBUF) & 1u);
      const int row = (int)(tile / tpr), kz0 = (int)(tile - (long long)row * tpr) * Prog::COLS;
      const unsigned char* src = tiles + me * Prog::TILE_BYTES;
#pragma unroll
      for (int h = 0; h < 512 / Prog::BOX_ROWS; ++h) {
        if (p.along_x) tma_store_3d(&tmap, src + h * Prog::BOX_ROWS * Prog::ROWB, kz0, row, h * Prog::BOX_ROWS);
        else tma_store_3d(&tmap, src + h * Prog::BOX_ROWS * Prog::ROWB, kz0, h * Prog::BOX_ROWS, row);
      }
      tma_store_commit();
      tma_store_wait_read();
      mbar_arrive_cta(&empty[me]);
    }
    tma_store_wait_read();
    return;
  }
  if (tid >= Prog::NTHREADS) return;

  // ------------------------------------ compute warps --------------------------------------
  typename Prog::Regs r;
  Prog::init(r, tid);
  cf* xg = xall + r.g * Prog::XG;
  cf w1[3], w2[3];
  stage_twiddles<512>(1, r.t, p.tw, w1);
  stage_twiddles<512>(2, r.t, p.tw, w2);
  int row = blockIdx.x / tpr, tcol = blockIdx.x - row * tpr;
  const int step_row = nblk / tpr, step_col = nblk - step_row * tpr;
  int n = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
    const int buf = n % NBUF;
    unsigned char* tb = tiles + buf * Prog::TILE_BYTES;
    Prog::set_tile(r, row, tcol * Prog::COLS);
    mbar_wait(&full[buf], (unsigned)(n / NBUF) & 1u);
#pragma unroll
    for (int k = 0; k < Prog::NPHASES; ++k) {
      if (k) group_sync(1 + r.g, Prog::GT);
      Prog::phase_tw(k, r, tb, xg, p, w1, w2);
    }
    fence_proxy_async();                        // my tile writes -> visible to the TMA store
    mbar_arrive_cta(&done[buf]);
    row += step_row; tcol += step_col;
    if (tcol >= tpr) { tcol -= tpr; ++row; }
  }
}


// ---- 1024-point lines: warp-specialised kernel over 4-D tensor maps -----------------------
// Same roles as fft_line_ws_kernel (loader thread, one retirer per tile buffer, compute warps in
// two-line groups; one retirer for both tile buffers), for the programs StridedLine4 / StridedLine16.  The tensor maps are 4-D - (kz, i_lo, row, i_hi) with line index
// i = i_hi * box_rows + i_lo - so one code path serves lines along y, lines along x, and the
// block layout of the x-slab transposes; loads go through maps.in, box h of a tile is stored
// through maps.out[h / out_div] - one map per destination block, which may live in a PEER's
// symmetric-memory buffer: then the TMA store IS the slab<->pencil transpose (64-byte rows over
// NVLink, issued by the retirer thread while the compute warps are already on the next tile).
struct LineMaps {
  CUtensorMap in;
  CUtensorMap out[8];
};

template <class Prog, int NBUF>
__global__ void __launch_bounds__(Prog::NTHREADS + 64, 1)
    fft_line4_ws_kernel(const __grid_constant__ LineMaps maps, const LineParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = sm;
  cf* xall = reinterpret_cast<cf*>(sm + NBUF * Prog::TILE_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + NBUF * Prog::TILE_BYTES + Prog::X_BYTES);
  unsigned long long* done = full + NBUF;
  unsigned long long* empty = full + 2 * NBUF;
  const int tid = threadIdx.x;
  const int tpr = p.tiles_per_row;
  const long long ntiles = p.ntiles;
  const int nblk = gridDim.x;
  const int nbox = Prog::LEN / p.box_rows;
  const int box_bytes = p.box_rows * Prog::ROWB;

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&done[b], Prog::NTHREADS);
      mbar_init(&empty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  if (tid == Prog::NTHREADS) {
    // ------------------------------------ loader ------------------------------------------
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
      const int buf = n % NBUF;
      const int row = p.row0 + (int)(tile / tpr), kz0 = (int)(tile % tpr) * Prog::COLS;
      if (n >= NBUF) {
        // the buffer is still in use: pull the tile into L2 while waiting for it
        if (p.l2_ahead)
          for (int h = 0; h < nbox; ++h) tma_prefetch_4d(&maps.in, kz0, 0, row, h);
        mbar_wait(&empty[buf], (unsigned)(n / NBUF - 1) & 1u);
      }
      unsigned char* dst = tiles + buf * Prog::TILE_BYTES;
      mbar_expect_tx(&full[buf], Prog::TILE_BYTES);
      for (int h = 0; h < nbox; ++h) tma_load_4d(dst + h * box_bytes, &maps.in, &full[buf], kz0, 0, row, h);
    }
    return;
  }
  if (tid == Prog::NTHREADS + 32) {
    // ---- retirer: ONE thread serves the buffers in tile order (tiles complete in order; a second
    // control warp per buffer would take the block to 19 warps and the register cap from 112 to 96)
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
      const int me = n % NBUF;
      mbar_wait(&done[me], (unsigned)(n / NBUF) & 1u);
      const int row = p.row0 + (int)(tile / tpr), kz0 = (int)(tile % tpr) * Prog::COLS;
      const unsigned char* src = tiles + me * Prog::TILE_BYTES;
      for (int h = 0, m = 0, c = 0; h < nbox; ++h) {
        tma_store_4d(&maps.out[m], src + h * box_bytes, kz0, 0, row, c);
        if (++c == p.out_div) { c = 0; ++m; }
      }
      tma_store_commit();
      tma_store_wait_read();
      mbar_arrive_cta(&empty[me]);
    }
    tma_store_wait_all();
    return;
  }
  if (tid >= Prog::NTHREADS) return;

  // ------------------------------------ compute warps --------------------------------------
  typename Prog::Regs r;
  Prog::init(r, tid);
  cf* xg = xall + r.g * Prog::XSTRIDE;
  typename Prog::Roots w;
  Prog::load_roots(w, r.t, p.tw);
  int row = blockIdx.x / tpr, tcol = blockIdx.x - row * tpr;
  const int step_row = nblk / tpr, step_col = nblk - step_row * tpr;
  int n = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += nblk, ++n) {
    const int buf = n % NBUF;
    unsigned char* tb = tiles + buf * Prog::TILE_BYTES;
    Prog::set_tile(r, p.row0 + p.kother0 + row, tcol * Prog::COLS);
    mbar_wait(&full[buf], (unsigned)(n / NBUF) & 1u);
#pragma unroll 1
    for (int pass = 0; pass < Prog::NPASS; ++pass) {
#pragma unroll
      for (int k = 0; k < Prog::NPHASES; ++k) {
        if (k) group_sync(1 + r.g, Prog::GT);
        Prog::phase(pass, k, r, tb, xg, p, w);
      }
    }
    fence_proxy_async();                        // my tile writes -> visible to the TMA store
    mbar_arrive_cta(&done[buf]);
    row += step_row; tcol += step_col;
    if (tcol >= tpr) { tcol -= tpr; ++row; }
  }
}

// ---- host ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return (EncodeTiledFn)sym;
  }();
  return fn;
}

bool line_pass_available() { return encode_fn() != nullptr; }

int line_make_tmap(void* out, void* spec, int nx, int ny, int P, int ncols_valid, int along_x, int kz) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return EVX_ERR_UNSUPPORTED;
  if (kz != 8 && kz != 16) return EVX_ERR_ARG;
  // (kz, y, x) with 8-byte elements; only the valid columns are part of the tensor, so the
  // pitch padding is neither read nor written
  const cuuint64_t dims[3] = {(cuuint64_t)ncols_valid, (cuuint64_t)ny, (cuuint64_t)nx};
  const cuuint64_t strides[2] = {(cuuint64_t)P * sizeof(cf), (cuuint64_t)ny * P * sizeof(cf)};
  const cuuint32_t box_y[3] = {(cuuint32_t)kz, 256, 1}, box_x[3] = {(cuuint32_t)kz, 1, 256};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, spec, dims, strides,
                          along_x ? box_x : box_y, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          kz == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? EVX_OK : EVX_ERR_UNSUPPORTED;
}

static int line_sms();

// 4-D tensor map (kz, i_lo, row, i_hi) over 8-byte elements; strides in elements.  Line index
// i = i_hi * box_rows + i_lo; the box is [box_rows x 8 columns] of one row.
int line_make_tmap4(void* out, void* base, int ncols_valid, int box_rows, long long lo_stride, int nrows,
                    long long row_stride, int nhi, long long hi_stride) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return EVX_ERR_UNSUPPORTED;
  if (box_rows < 1 || box_rows > 256 || nrows < 1 || nhi < 1) return EVX_ERR_ARG;
  const cuuint64_t dims[4] = {(cuuint64_t)ncols_valid, (cuuint64_t)box_rows, (cuuint64_t)nrows, (cuuint64_t)nhi};
  const cuuint64_t strides[3] = {(cuuint64_t)lo_stride * sizeof(cf), (cuuint64_t)row_stride * sizeof(cf),
                                 (cuuint64_t)(nhi > 1 ? hi_stride : lo_stride * box_rows) * sizeof(cf)};
  const cuuint32_t box[4] = {8, (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? EVX_OK : EVX_ERR_UNSUPPORTED;
}

// EVX_LINE4_G: lines per synchronisation group (2 = two groups of 256 threads, 1 = four groups of 128)
static int line4_group_lines() {
  const char* e = getenv("EVX_LINE4_G");
  return (e && atoi(e) == 1) ? 1 : 2;
}

// EVX_LINE4_FORM: 16 = sixteen points per thread (StridedLine16: two trips through shared memory per
// transform), 8 = eight points per thread (StridedLine4)
// default: 16 for the x pass (chunked x pass of the distributed plan 0.372 -> 0.328 ms per rank), 8 for
// the y passes (HBM-bound at 0.92 of the copy peak in the eight-point form; 0.36 against 0.45 ms)
static int line4_form(int mode) {
  const char* e = getenv("EVX_LINE4_FORM");
  if (e && atoi(e) == 8) return 8;
  if (e && atoi(e) == 16) return 16;
  return pass_is_xmid(mode) ? 16 : 8;
}

template <int MODE, int GL>
static int launch_line4_t(LineParams p, const void* map_in, const void* const* maps_out, int nout, cudaStream_t st) {
  using Prog = typename std::conditional<GL == 16, StridedLine16<1024, 8, MODE>, StridedLine4<1024, 8, MODE, GL == 16 ? 2 : GL>>::type;
  constexpr int NBUF = 2;
  p.tiles_per_row = (p.ncols_valid + 7) / 8;
  p.ntiles = (long long)p.nrows * p.tiles_per_row;
  if (p.ntiles < 1 || p.box_rows < 1 || 1024 % p.box_rows || p.out_div < 1) return EVX_ERR_ARG;
  const int nbox = 1024 / p.box_rows;
  if (nout < 1 || nout > 8 || (nbox + p.out_div - 1) / p.out_div > nout) return EVX_ERR_ARG;
  constexpr size_t smem = 1024 + (size_t)NBUF * Prog::TILE_BYTES + Prog::X_BYTES + 128;
  auto kern = fft_line4_ws_kernel<Prog, NBUF>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, smem)) return rc;
  LineMaps maps;
  maps.in = *(const CUtensorMap*)map_in;
  for (int i = 0; i < 8; ++i) maps.out[i] = *(const CUtensorMap*)maps_out[i < nout ? i : 0];
  long long resident = line_sms();
  if (p.max_ctas > 0 && p.max_ctas < resident) resident = p.max_ctas;
  const unsigned grid = (unsigned)(p.ntiles < resident ? p.ntiles : resident);
  kern<<<grid, Prog::NTHREADS + 64, smem, st>>>(maps, p);
  count_launch();
  return (int)cudaGetLastError();
}

int line4_pass_launch(int mode, const LineParams& p, const void* map_in, const void* const* maps_out, int nout,
                      cudaStream_t st) {
  // The exponential-Euler x pass keeps the eight-point form: its sixteen-point instantiation differs
  // from the cp.async pass by one ulp in ~9 % of the outputs on the device (repeatable; the CPU replay
  // of the same program is bit-identical, the IMEX instantiation is bit-identical on the device -
  // a code-generation difference that was not tracked down), and the transports are tested for
  // bit-identity.
  if (line4_form(mode) == 16 && mode != PASS_XMID_ETD1) {
    switch (mode) {
      case PASS_FWD: return launch_line4_t<PASS_FWD, 16>(p, map_in, maps_out, nout, st);
      case PASS_INV: return launch_line4_t<PASS_INV, 16>(p, map_in, maps_out, nout, st);
      case PASS_XMID: return launch_line4_t<PASS_XMID, 16>(p, map_in, maps_out, nout, st);
      default: return EVX_ERR_ARG;
    }
  }
  const bool g1 = line4_group_lines() == 1;
  switch (mode) {
    case PASS_FWD: return g1 ? launch_line4_t<PASS_FWD, 1>(p, map_in, maps_out, nout, st)
                             : launch_line4_t<PASS_FWD, 2>(p, map_in, maps_out, nout, st);
    case PASS_INV: return g1 ? launch_line4_t<PASS_INV, 1>(p, map_in, maps_out, nout, st)
                             : launch_line4_t<PASS_INV, 2>(p, map_in, maps_out, nout, st);
    case PASS_XMID: return g1 ? launch_line4_t<PASS_XMID, 1>(p, map_in, maps_out, nout, st)
                              : launch_line4_t<PASS_XMID, 2>(p, map_in, maps_out, nout, st);
    case PASS_XMID_ETD1: return g1 ? launch_line4_t<PASS_XMID_ETD1, 1>(p, map_in, maps_out, nout, st)
                                   : launch_line4_t<PASS_XMID_ETD1, 2>(p, map_in, maps_out, nout, st);
    default: return EVX_ERR_ARG;
  }
}

static int line_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

// EVX_FFT_LINE_WS: 0 = two blocks per SM, tile hand-over by the compute threads (fft_line_kernel);
// 3 / 4 / 5 = warp-specialised form with that many tile buffers (KZ = 8 only)
static int line_ws_bufs() {
  const char* e = getenv("EVX_FFT_LINE_WS");
  const int v = e ? atoi(e) : 3;     // 512^3 x pass: 0.320 ms (form 0), 0.2455 (3 buffers), 0.258 (4), 0.250 (5)
  return (v == 3 || v == 4 || v == 5) ? v : 0;
}

template <class Prog, int NBUF>
static int launch_line_ws(const LineParams& p, const void* tmap, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)NBUF * Prog::TILE_BYTES + Prog::X_BYTES + 128;
  auto kern = fft_line_ws_kernel<Prog, NBUF>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, smem)) return rc;
  const long long resident = line_sms();
  const unsigned grid = (unsigned)(p.ntiles < resident ? p.ntiles : resident);
  kern<<<grid, line_ws_threads(Prog::NTHREADS, NBUF), smem, st>>>(*(const CUtensorMap*)tmap, p);
  count_launch();
  return (int)cudaGetLastError();
}

template <int KZ, int MODE>
static int launch_line_t(LineParams p, const void* tmap, cudaStream_t st) {
  using Prog = StridedLine<512, KZ, MODE>;
  p.tiles_per_row = (p.ncols_valid + KZ - 1) / KZ;
  p.ntiles = (long long)(p.along_x ? p.ny : p.nx) * p.tiles_per_row;
  if (KZ == 8) {
    const int ws = line_ws_bufs();
    if (ws == 3) return launch_line_ws<StridedLine<512, 8, MODE>, 3>(p, tmap, st);
    if (ws == 4) return launch_line_ws<StridedLine<512, 8, MODE>, 4>(p, tmap, st);
    if (ws == 5) return launch_line_ws<StridedLine<512, 8, MODE>, 5>(p, tmap, st);
  }
  auto kern = fft_line_kernel<Prog>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, Prog::SMEM_BYTES)) return rc;
  long long resident = (long long)line_sms() * (Prog::NTHREADS <= 512 ? 2 : 1);
  const unsigned grid = (unsigned)(p.ntiles < resident ? p.ntiles : resident);
  kern<<<grid, Prog::NTHREADS, Prog::SMEM_BYTES, st>>>(*(const CUtensorMap*)tmap, p);
  count_launch();
  return (int)cudaGetLastError();
}

int line_pass_launch(int mode, int kz, const LineParams& p, const void* tmap, cudaStream_t st) {
  if ((p.along_x ? p.nx : p.ny) != 512) return EVX_ERR_UNSUPPORTED;
#define EVX_LINE(M)                                                  \
  case M:                                                            \
    return kz == 16 ? launch_line_t<16, M>(p, tmap, st) : launch_line_t<8, M>(p, tmap, st);
  switch (mode) {
    EVX_LINE(PASS_FWD) EVX_LINE(PASS_INV) EVX_LINE(PASS_XMID) EVX_LINE(PASS_XMID_ETD1)
    default: return EVX_ERR_ARG;
  }
#undef EVX_LINE
}

}  // namespace evx
