// Chained z/y passes of the native FFT: ONE persistent kernel per direction that runs the z
// lines and the y tiles of every x plane with the intermediate half spectrum of a plane
// (1 MB at 512^2) staying in L2 - it is written by the first stage, read by the second stage a
// few microseconds later and overwritten in place, so HBM sees 8 B/voxel for the forward pair
// (read r, write S) instead of 16, and the inverse pair reads S once instead of twice.
//
//   forward   z lines  r[x]  -> S[x]  (one warp per 512-point line: fft_pass_core.h ZPass)
//             y tiles  S[x] in place  (fft_line_core.h StridedLine<512, 8, PASS_FWD>, TMA)
//   inverse   y tiles  S[x] in place  (PASS_INV)
//             z lines  S[x], u[x] -> out[x]
//
// Same arithmetic, same thread -> butterfly assignment as the stand-alone passes: results are
// bit-identical (tests/test_gpu_parity.py::test_chained_passes_*).  Work list and dependency
// rule: fft_chain_core.h.  Completion of a plane's first stage is published through one
// counter per plane (release add / acquire load at gpu scope); tiles move by TMA, so the
// generic <-> async proxy hand-overs carry fence.proxy.async on both sides.
#include <cuda.h>
#include <cstdlib>
#include "evx_internal.h"
#include "fft_line_core.h"
#include "fft_chain_core.h"
#include "fft_chain.h"
#include "fft_line.h"
#include "tma_ptx.h"

namespace evx {

struct ChainParams {
  ChainSchedule sched;
  const float* real_in;    // forward: r   inverse: u (may be null: out = update)
  float* real_out;         // inverse: out
  cf* spec;                // [nx][ny][P]
  const cf *twz, *twr, *twy;
  int ny, nz, P;
  unsigned* done0;         // [nplanes] finished stage-0 items per plane; zeroed before the launch
  int ahead;               // 1: the next item's input copy is issued while the current item runs
  unsigned long long* stats;   // optional [gridDim.x][8] cycle counters of thread 0 (EVX_FFT_CHAIN_STATS)
};

constexpr int kChainThreads = 512;
constexpr int kChainZLines = kChainThreads / 32;     // z lines per item: one per warp
// input buffer: a [512 x 8] complex tile (32 KB) or 16 spectrum rows of 264 complex (33 KB)
constexpr int kChainBufBytes = 34 * 1024;
constexpr size_t kChainSmemBytes = 1024 + 2 * (size_t)kChainBufBytes + StridedLine<512, 8, PASS_FWD>::X_BYTES + 64;

// spin until *flag >= target (bounded: a scheduling bug must trap, not hang the GPU)
__device__ __forceinline__ void wait_count(const unsigned* flag, unsigned target) {
  const long long t0 = clock64();
  for (unsigned spin = 0; ld_acquire_gpu(flag) < target; ++spin)
    if ((spin & 255u) == 255u && clock64() - t0 > 4000000000LL) __trap();
}
// thread 0's cycle accounting (only when the launcher passes a stats buffer)
struct ChainStats {
  long long dep = 0, mbar = 0, zitem = 0, yitem = 0, nwait = 0, nearly_fail = 0, pub = 0;
};

template <bool INV, bool STATS>
__global__ void __launch_bounds__(kChainThreads, 2)
    fft_chain_kernel(const __grid_constant__ CUtensorMap tmap, const ChainParams p) {
  using Line = StridedLine<512, 8, INV ? PASS_INV : PASS_FWD>;
  using ZP = ZPass<256, 1, INV>;
  constexpr int M = 256, T = 32, ZLP = ZP::LP;
  constexpr int BUF = kChainBufBytes;
  static_assert(kChainZLines * ZLP * sizeof(cf) <= Line::X_BYTES, "z scratch lives in the exchange area");
  static_assert(Line::TILE_BYTES <= BUF && kChainZLines * 264 * sizeof(cf) <= BUF, "input buffer size");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* bufs = sm;                     // two input buffers: a y tile or the rows of a z item
  cf* xall = reinterpret_cast<cf*>(sm + 2 * BUF);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + 2 * BUF + Line::X_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool leader = tid == 0;
  const ChainSchedule& sc = p.sched;
  const long long total = sc.total, G = gridDim.x;
  typename Line::Regs yr;
  Line::init(yr, tid);
  cf* xg = xall + yr.g * Line::XG;
  cf* zb = xall + warp * ZLP;
  LineParams lp;
  lp.tw = p.twy;
  const unsigned zin_bytes = (unsigned)(kChainZLines * (INV ? p.P * sizeof(cf) : p.nz * sizeof(float)));

  if (leader) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  auto is_y = [](const ChainItem& it) { return INV ? it.stage == 0 : it.stage == 1; };
  int count = 0;                    // items this block has finished (uniform)
  // leader only --------------------------------------------------------------------------
  int issued = 0;                   // items whose input copy has been issued
  int pending_plane = -1;           // inverse: plane of the tile store that is not yet published
  ChainStats cs;
  constexpr bool st_on = STATS;
  const long long t_begin = st_on ? clock64() : 0;
  auto publish_pending = [&]() {
    if (pending_plane < 0) return;
    const long long t0 = st_on ? clock64() : 0;
    tma_store_wait_all();           // the tile is in global memory
    if (st_on) cs.pub += clock64() - t0;
    fence_proxy_async_all();
    __threadfence();
    red_release_gpu_add(p.done0 + pending_plane, 1u);
    pending_plane = -1;
  };
  // Issue the input copy of this block's item number `issued` (item index blockIdx.x + issued*G)
  // into buffer issued & 1; false if it depends on an unfinished plane and !blocking.
  auto issue = [&](bool blocking) -> bool {
    const ChainItem it = chain_decode(sc, (long long)blockIdx.x + (long long)issued * G);
    if (it.stage == 1) {            // second stage of the pair: the plane's first stage must be done
      if (ld_acquire_gpu(p.done0 + it.plane) < (unsigned)sc.n0) {
        if (!blocking) { ++cs.nearly_fail; return false; }
        if (INV) publish_pending(); // never wait while holding back an own tile
        const long long t0 = st_on ? clock64() : 0;
        wait_count(p.done0 + it.plane, (unsigned)sc.n0);
        if (st_on) { cs.dep += clock64() - t0; ++cs.nwait; }
      }
      fence_proxy_async_all();      // other blocks' stores -> this asynchronous copy
    }
    tma_store_wait_read();          // the buffer's previous tile has left shared memory
    const int buf = issued & 1;
    unsigned char* dst = bufs + buf * BUF;
    if (is_y(it)) {
      mbar_expect_tx(&full[buf], Line::TILE_BYTES);
#pragma unroll
      for (int h = 0; h < 512 / Line::BOX_ROWS; ++h)
        tma_load_3d(dst + h * Line::BOX_ROWS * Line::ROWB, &tmap, &full[buf], it.idx * Line::COLS,
                    h * Line::BOX_ROWS, it.plane);
    } else {
      const long long row = (long long)it.plane * p.ny + (long long)it.idx * kChainZLines;
      const void* src = INV ? (const void*)(p.spec + row * p.P) : (const void*)(p.real_in + row * p.nz);
      mbar_expect_tx(&full[buf], zin_bytes);
      bulk_load_1d(dst, src, zin_bytes, &full[buf]);
      // the u rows of an inverse z item are loaded straight into registers: get them into L2
      if (INV && p.real_in) bulk_prefetch_l2(p.real_in + row * p.nz, (unsigned)(kChainZLines * p.nz * sizeof(float)));
    }
    ++issued;
    return true;
  };

  for (long long i = blockIdx.x; i < total; i += G) {
    const ChainItem it = chain_decode(sc, i);
    const bool y_item = is_y(it);
    const long long t_item = (st_on && leader) ? clock64() : 0;
    if (leader) {
      if (issued == count) issue(true);
      // one item ahead: its buffer is free once the previous item's tile store has been read
      if (p.ahead && issued == count + 1 && i + G < total) issue(false);
    }
    const int buf = count & 1;
    unsigned char* tb = bufs + buf * BUF;
    if (st_on && leader) {
      const long long t0 = clock64();
      mbar_wait(&full[buf], (unsigned)(count >> 1) & 1u);
      cs.mbar += clock64() - t0;
    } else {
      mbar_wait(&full[buf], (unsigned)(count >> 1) & 1u);
    }
    if (y_item) {
#pragma unroll
      for (int k = 0; k < Line::NPHASES; ++k) {
        if (k) group_sync(1 + yr.g, Line::GT);
        Line::phase(k, yr, tb, xg, lp);
      }
      if (INV && leader) publish_pending();
      fence_proxy_async();                      // tile writes -> visible to the TMA store
      __syncthreads();
      if (leader) {
#pragma unroll
        for (int h = 0; h < 512 / Line::BOX_ROWS; ++h)
          tma_store_3d(&tmap, tb + h * Line::BOX_ROWS * Line::ROWB, it.idx * Line::COLS, h * Line::BOX_ROWS, it.plane);
        tma_store_commit();
        if (INV) pending_plane = it.plane;
        if (st_on) cs.yitem += clock64() - t_item;
      }
    } else {
      // one 512-point real line per warp, its input row already in shared memory
      const long long row = (long long)it.plane * p.ny + (long long)it.idx * kChainZLines + warp;
      typename ZP::Regs r;
      r.t = lane; r.l = 0; r.row = row; r.valid = true;
      if (!INV) {
        const cf* line = reinterpret_cast<const cf*>(tb + (size_t)warp * p.nz * sizeof(float));
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = line[lane + e * T];
        line_stage_compute_pre<M, -1>(0, r.v, lane, r.w);
        ZP::write_stage(r, zb, 0);
        stage_twiddles<M>(1, lane, p.twz, r.w);
        __syncwarp();
#pragma unroll
        for (int s = 1; s < ZP::S; ++s) {
          ZP::read_natural(r, zb);
          __syncwarp();
          line_stage_compute_pre<M, -1>(s, r.v, lane, r.w);
          ZP::write_stage(r, zb, s);
          stage_twiddles<M>(s + 1, lane, p.twz, r.w);
          __syncwarp();
        }
        cf* out = p.spec + row * p.P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = lane + e * T;
          const cf zk = zb[zline_idx<M>(kk)];
          const cf zmk = zb[zline_idx<M>(kk == 0 ? 0 : M - kk)];
          out[kk] = ZP::untangle_fwd(zk, zmk, p.twr[kk]);
        }
        if (lane == 0) {
          const cf z0 = zb[zline_idx<M>(0)];
          out[M] = cf{z0.x - z0.y, 0.f};
        }
        __syncthreads();
        if (leader) {
          __threadfence();
          red_release_gpu_add(p.done0 + it.plane, 1u);
          if (st_on) cs.zitem += clock64() - t_item;
        }
      } else {
        if (p.real_in) {
          const cf* u = reinterpret_cast<const cf*>(p.real_in + row * p.nz);
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = u[lane + e * T];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) r.u[e] = cf{0.f, 0.f};
        }
        // the row sits in natural order in the input buffer: X[k] and X[M-k] are both
        // lane-contiguous reads, no staging copy
        const cf* in = reinterpret_cast<const cf*>(tb + (size_t)warp * p.P * sizeof(cf));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = lane + e * T;
          r.v[e] = ZP::untangle_inv(in[kk], in[M - kk], p.twr[kk]);
        }
#pragma unroll
        for (int s = 0; s < ZP::S; ++s) {
          if (s) {
            ZP::read_natural(r, zb);
            __syncwarp();
          }
          line_stage_compute_pre<M, +1>(s, r.v, lane, r.w);
          if (s < ZP::S - 1) {
            ZP::write_stage(r, zb, s);
            stage_twiddles<M>(s + 1, lane, p.twz, r.w);
            __syncwarp();
          }
        }
        cf* out = reinterpret_cast<cf*>(p.real_out + row * p.nz);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[lane + e * T] = cadd(r.v[e], r.u[e]);
        if (leader) publish_pending();
        __syncthreads();
        if (st_on && leader) cs.zitem += clock64() - t_item;
      }
    }
    ++count;
  }
  if (leader) {
    if (INV) publish_pending();
    tma_store_wait_read();            // shared memory must outlive the stores
    if (st_on) {
      unsigned long long* o = p.stats + 8ull * blockIdx.x;
      o[0] = cs.dep; o[1] = cs.mbar; o[2] = cs.zitem; o[3] = cs.yitem; o[4] = clock64() - t_begin;
      o[5] = cs.nwait; o[6] = cs.nearly_fail; o[7] = cs.pub;
    }
  }
}

// ---- host ---------------------------------------------------------------------------
bool chain_supported(int nx, int ny, int nz) {
  return ny == 512 && nz == 512 && nx >= 8 && line_pass_available();
}

static int chain_lag() {
  const char* e = getenv("EVX_FFT_CHAIN_LAG");
  const int v = e ? atoi(e) : 12;
  return v < 1 ? 1 : v;
}
static int chain_ahead() {
  const char* e = getenv("EVX_FFT_CHAIN_AHEAD");
  const int v = e ? atoi(e) : 1;
  return v < 0 ? 0 : (v > 1 ? 1 : v);
}

template <bool INV, bool STATS>
static int chain_launch_t(ChainParams p, const void* tmap, cudaStream_t st) {
  constexpr size_t smem = kChainSmemBytes;
  auto kern = fft_chain_kernel<INV, STATS>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, smem)) return rc;
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kChainThreads, smem) != cudaSuccess || per_sm < 1)
    return EVX_ERR_UNSUPPORTED;
  if (per_sm > 2) per_sm = 2;
  long long grid = (long long)sms * per_sm;
  if (grid > p.sched.total) grid = p.sched.total;
  cudaError_t e = cudaMemsetAsync(p.done0, 0, (size_t)p.sched.nplanes * sizeof(unsigned), st);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all blocks co-resident: the waits cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, *(const CUtensorMap*)tmap, p);
  count_launch();
  return (int)e;
}

int chain_launch(bool inverse, const ChainArgs& a, const void* tmap_y, cudaStream_t st) {
  if (!chain_supported(a.nx, a.ny, a.nz)) return EVX_ERR_UNSUPPORTED;
  ChainParams p;
  const int ytiles = (a.nz / 2 + 1 + 7) / 8;
  const int zitems = a.ny / kChainZLines;
  p.sched = inverse ? make_chain_schedule(a.nx, chain_lag(), ytiles, zitems)
                    : make_chain_schedule(a.nx, chain_lag(), zitems, ytiles);
  p.real_in = a.real_in; p.real_out = a.real_out; p.spec = (cf*)a.spec;
  p.twz = (const cf*)a.twz; p.twr = (const cf*)a.twr; p.twy = (const cf*)a.twy;
  p.ny = a.ny; p.nz = a.nz; p.P = a.P;
  p.done0 = (unsigned*)a.flags;
  p.ahead = chain_ahead();
  const char* es = getenv("EVX_FFT_CHAIN_STATS");
  p.stats = (es && atoi(es) != 0) ? (unsigned long long*)a.stats : nullptr;
  if (p.stats) return inverse ? chain_launch_t<true, true>(p, tmap_y, st) : chain_launch_t<false, true>(p, tmap_y, st);
  return inverse ? chain_launch_t<true, false>(p, tmap_y, st) : chain_launch_t<false, false>(p, tmap_y, st);
}

}  // namespace evx
