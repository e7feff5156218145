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
#include "fft_zline_core.h"
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
  int discard;             // inverse: drop the consumed spectrum rows from L2 without write-back
  int blk_mode;            // x-slab plan: the y tiles are stored to (forward) / loaded from (inverse) the
                           // all-to-all block layout, one tensor map per 256-row box (ChainMaps::blk)
  unsigned long long* stats;   // optional [gridDim.x][16] cycle counters of thread 0 (EVX_FFT_CHAIN_STATS)
};

// tensor maps of a launch: the plain spectrum ([8 x 256 x 1] boxes along y) and, for the x-slab
// plan with ny / W = 256, the two destination (forward) / source (inverse) blocks of the all-to-all
// layout - local block buffers or, forward, a peer's buffer over NVLink
struct ChainMaps {
  CUtensorMap spec;
  CUtensorMap blk[2];
};

constexpr int kChainCompute = 512;                   // compute threads (16 warps)
// + one loader warp + one retirer warp per input buffer (lane 0 of each acts)
constexpr int chain_threads(int nbuf) { return kChainCompute + 32 + 32 * (nbuf <= 2 ? 1 : nbuf); }
constexpr int kChainZLines = kChainCompute / 32;     // z lines per item
// input buffer: a [512 x 8] complex tile (32 KB) or 16 spectrum rows of 264 complex (33 KB)
constexpr int kChainBufBytes = 34 * 1024;
constexpr size_t chain_smem_bytes(int nbuf) {
  return 1024 + (size_t)nbuf * kChainBufBytes + StridedLine<512, 8, PASS_FWD>::X_BYTES + 128;
}
// named barriers: 0 block, 1..4 y groups (128 threads), 5..12 z groups (64 threads)
constexpr int kBarY = 1, kBarZ = 5;

// spin until *flag >= target (bounded: a scheduling bug must trap, not hang the GPU)
__device__ __forceinline__ void wait_count(const unsigned* flag, unsigned target) {
  const long long t0 = clock64();
  for (unsigned spin = 0; ld_acquire_gpu(flag) < target; ++spin)
    if ((spin & 255u) == 255u && clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cycle accounting of the two control threads (only when the launcher passes a stats buffer)
struct ChainStats {
  long long a = 0, b = 0, c = 0, d = 0, n = 0;
};

// Block = 16 compute warps + a loader warp + a retirer warp (warp specialisation):
//   compute threads  wait for the item's input (mbarrier full[b]), run its phases with group
//                    barriers, arrive on done[b] - no flags, no fences, no waits on other blocks
//   loader           walks the item list two items ahead: waits for the buffer (empty[b]) and
//                    for the plane the item depends on, then issues the TMA tile / bulk row copies
//   retirer          waits for done[b]; y tile: TMA store, buffer free once the store has read
//                    it, (inverse) publish the plane counter once the store is complete;
//                    z item: buffer free at once, (forward) publish the plane counter
// so the latencies of the control path (store drain, gpu-scope fences, acquire loads) run next to
// the compute warps instead of inside one of them, and loads and retirements overlap each other.
// NBUF = 2: two blocks per SM, the next item's input in flight while the current one runs;
// NBUF = 4: one block per SM with its input copies issued up to three items ahead.
// TW: roots of unity kept in registers over the whole item loop - 0 none (fetched per item),
// 1 the stage twiddles of the y and z lines, 2 also the untangle roots of the z lines, 3 the
// stage twiddles of the y lines only.
template <bool INV, bool STATS, int NBUF, int TW>
__global__ void __launch_bounds__(chain_threads(NBUF), NBUF <= 2 ? 2 : 1)
    fft_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using Line = StridedLine<512, 8, INV ? PASS_INV : PASS_FWD>;
  using ZG = ZGroupLine<INV>;
  constexpr int BUF = kChainBufBytes;
  static_assert(2 * ZG::XG <= Line::XG, "the two z groups of a y group share its exchange buffer");
  static_assert(Line::TILE_BYTES <= BUF && kChainZLines * ZG::ROWP * sizeof(cf) <= BUF, "input buffer size");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* bufs = sm;                     // NBUF input buffers: a y tile or the rows of a z item
  cf* xall = reinterpret_cast<cf*>(sm + NBUF * BUF);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + NBUF * BUF + Line::X_BYTES);
  unsigned long long* done = full + NBUF;
  unsigned long long* empty = full + 2 * NBUF;

  const int tid = threadIdx.x;
  const ChainSchedule& sc = p.sched;
  const int G = gridDim.x;
  constexpr bool st_on = STATS;
  auto is_y = [](const ChainItem& it) { return INV ? it.stage == 0 : it.stage == 1; };

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&done[b], kChainCompute);
      mbar_init(&empty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  if (tid == kChainCompute) {
    // ==================================== loader ==========================================
    const unsigned zin_bytes = (unsigned)(kChainZLines * (INV ? ZG::ROWP * sizeof(cf) : p.nz * sizeof(float)));
    ChainStats cs;
    const long long t_begin = st_on ? clock64() : 0;
    ChainCursor cur;
    chain_cursor_init(cur, sc, blockIdx.x, G);
    ChainItem it;
    for (int n = 0; chain_cursor_next(cur, sc, it); ++n, chain_cursor_step(cur, sc)) {
      const int buf = n % NBUF;
      long long t0 = st_on ? clock64() : 0;
      if (n >= NBUF) mbar_wait(&empty[buf], (unsigned)(n / NBUF - 1) & 1u);
      long long t1 = st_on ? clock64() : 0;
      if (it.stage == 1) {          // second stage of the pair: the plane's first stage must be done
        wait_count(p.done0 + it.plane, (unsigned)sc.n0);
        fence_proxy_async_global();    // other blocks' stores -> this asynchronous copy
      }
      long long t2 = st_on ? clock64() : 0;
      unsigned char* dst = bufs + buf * BUF;
      if (is_y(it)) {
        mbar_expect_tx(&full[buf], Line::TILE_BYTES);
#pragma unroll
        for (int h = 0; h < 512 / Line::BOX_ROWS; ++h) {
          if (INV && p.blk_mode)
            tma_load_3d(dst + h * Line::BOX_ROWS * Line::ROWB, &maps.blk[h], &full[buf], it.idx * Line::COLS, 0, it.plane);
          else
            tma_load_3d(dst + h * Line::BOX_ROWS * Line::ROWB, &maps.spec, &full[buf], it.idx * Line::COLS,
                        h * Line::BOX_ROWS, it.plane);
        }
      } else {
        const long long row = (long long)it.plane * p.ny + (long long)it.idx * kChainZLines;
        mbar_expect_tx(&full[buf], zin_bytes);
        if (INV) {                  // spectrum rows: the global pitch is the buffer's pitch
          bulk_load_1d(dst, p.spec + row * p.P, zin_bytes, &full[buf]);
          // the u rows are loaded straight into registers: get them into L2
          if (p.real_in) bulk_prefetch_l2(p.real_in + row * p.nz, (unsigned)(kChainZLines * p.nz * sizeof(float)));
        } else {                    // real rows (2 KB) go to the same 264-complex pitch
          const unsigned rb = (unsigned)(p.nz * sizeof(float));
          const char* src = (const char*)(p.real_in + row * p.nz);
          for (int l = 0; l < kChainZLines; ++l)
            bulk_load_1d(dst + (size_t)l * ZG::ROWP * sizeof(cf), src + (size_t)l * rb, rb, &full[buf]);
        }
      }
      if (st_on) { cs.a += t1 - t0; cs.b += t2 - t1; cs.c += clock64() - t2; ++cs.n; }
    }
    if (st_on) {
      unsigned long long* o = p.stats + 16ull * blockIdx.x;
      o[0] = cs.a; o[1] = cs.b; o[2] = cs.c; o[3] = cs.n; o[4] = clock64() - t_begin;
    }
    return;
  }
  if (tid >= kChainCompute + 32 && (tid & 31) == 0) {
    // ============ retirers: one per buffer (NBUF > 2) or one for both (NBUF = 2) ===========
    constexpr int NRET = NBUF <= 2 ? 1 : NBUF;
    const int me = (tid - kChainCompute - 32) >> 5;
    ChainStats cs;
    ChainCursor cur;
    chain_cursor_init(cur, sc, blockIdx.x, G);
    ChainItem it;
    for (int n = 0; chain_cursor_next(cur, sc, it); ++n, chain_cursor_step(cur, sc)) {
      const int buf = n % NBUF;
      if (NRET > 1 && buf != me) continue;
      long long t0 = st_on ? clock64() : 0;
      mbar_wait(&done[buf], (unsigned)(n / NBUF) & 1u);
      long long t1 = st_on ? clock64() : 0;
      if (is_y(it)) {
        const unsigned char* tb = bufs + buf * BUF;
#pragma unroll
        for (int h = 0; h < 512 / Line::BOX_ROWS; ++h) {
          if (!INV && p.blk_mode)
            tma_store_3d(&maps.blk[h], tb + h * Line::BOX_ROWS * Line::ROWB, it.idx * Line::COLS, 0, it.plane);
          else
            tma_store_3d(&maps.spec, tb + h * Line::BOX_ROWS * Line::ROWB, it.idx * Line::COLS, h * Line::BOX_ROWS, it.plane);
        }
        tma_store_commit();
        tma_store_wait_read();      // the tile has left shared memory: the buffer is free
        mbar_arrive(&empty[buf]);
        if (INV) {                  // publish the tile once it is in global memory
          tma_store_wait_all();
          fence_proxy_async_global();
          red_release_gpu_add(p.done0 + it.plane, 1u);
        }
      } else {
        mbar_arrive(&empty[buf]);
        // release: covers the compute threads' stores (ordered before by the done barrier)
        if (!INV) red_release_gpu_add(p.done0 + it.plane, 1u);
      }
      if (st_on) { cs.a += t1 - t0; cs.b += clock64() - t1; }
    }
    tma_store_wait_all();             // shared memory must outlive the stores
    if (st_on && me == 0) {
      unsigned long long* o = p.stats + 16ull * blockIdx.x + 8;
      o[0] = cs.a; o[1] = cs.b;
    }
    return;
  }
  if (tid >= kChainCompute) return;     // idle lanes of the control warps

  // ================================ compute warps ==========================================
  typename Line::Regs yr;
  Line::init(yr, tid);
  cf* xg = xall + yr.g * Line::XG;
  typename ZG::Regs zr;
  ZG::init(zr, tid);
  // the two z groups (64 threads) of a y group (128 threads) split that group's exchange buffer
  cf* zxg = xall + (zr.g >> 1) * Line::XG + (zr.g & 1) * ZG::XG;
  // roots of unity of this thread's butterflies: the same for every item, kept in registers
  // where the block has them to spare (one block per SM)
  constexpr bool TWREG = TW >= 1, TWREG_Z = TW == 1 || TW == 2, ROOTREG = TW == 2;
  cf yw1[3], yw2[3], zw1[3], zw2[3], zroot[8];
  if (TWREG) {
    stage_twiddles<512>(1, yr.t, p.twy, yw1);
    stage_twiddles<512>(2, yr.t, p.twy, yw2);
  }
  if (TWREG_Z) {
    stage_twiddles<ZG::M>(1, zr.t, p.twz, zw1);
    stage_twiddles<ZG::M>(2, zr.t, p.twz, zw2);
  }
  if (ROOTREG) {
#pragma unroll
    for (int e = 0; e < 8; ++e) zroot[e] = p.twr[zr.t + e * ZG::T];
  }
  LineParams lp;
  lp.tw = p.twy;
  ZGroupParams zp;
  zp.tw = p.twz; zp.twr = p.twr; zp.nz = p.nz; zp.P = p.P;
  ChainCursor cur;
  chain_cursor_init(cur, sc, blockIdx.x, G);
  ChainItem it;
  for (int count = 0; chain_cursor_next(cur, sc, it); ++count, chain_cursor_step(cur, sc)) {
    const int buf = count % NBUF;
    unsigned char* tb = bufs + buf * BUF;
    mbar_wait(&full[buf], (unsigned)(count / NBUF) & 1u);
    // the 128 threads of a y group own one exchange buffer across item types: nobody may start
    // writing it while a neighbour still reads the previous item's data
    group_sync(kBarY + yr.g, Line::GT);
    if (is_y(it)) {
#pragma unroll
      for (int k = 0; k < Line::NPHASES; ++k) {
        if (k) group_sync(kBarY + yr.g, Line::GT);
        if (TWREG) Line::phase_tw(k, yr, tb, xg, lp, yw1, yw2);
        else Line::phase(k, yr, tb, xg, lp);
      }
      fence_proxy_async();                      // tile writes -> visible to the TMA store
    } else {
      const long long grow = (long long)it.plane * p.ny + (long long)it.idx * kChainZLines + zr.g * ZG::G + zr.c2;
      cf* rows = reinterpret_cast<cf*>(tb);
      if (INV && p.discard) {
        // The item's 16 spectrum rows (264 cache lines, contiguous) are in shared memory now and
        // nobody reads them from global memory again before the next step overwrites them: drop
        // the lines - dirty since the y tiles were stored in place - so that they are never
        // written back to HBM.
        constexpr int kLines = kChainZLines * ZG::ROWP * (int)sizeof(cf) / 128;
        static_assert(kChainZLines * ZG::ROWP * sizeof(cf) % 128 == 0 && kLines <= kChainCompute, "");
        if (tid < kLines) {
          const char* line = (const char*)(p.spec + ((long long)it.plane * p.ny + (long long)it.idx * kChainZLines) * p.P) + 128 * tid;
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
        }
      }
#pragma unroll
      for (int k = 0; k < ZG::NPHASES; ++k) {
        if (k) group_sync(kBarZ + zr.g, ZG::GT);
        if (TWREG_Z) ZG::template phase_tw<ROOTREG>(k, zr, rows, zxg, zw1, zw2, zroot, p.twr, p.nz, p.P, grow, p.real_in, p.real_out, p.spec);
        else ZG::phase(k, zr, rows, zxg, zp, grow, p.real_in, p.real_out, p.spec);
      }
    }
    mbar_arrive(&done[buf]);
  }
}

// ---- host ---------------------------------------------------------------------------
bool chain_supported(int nx, int ny, int nz) {
  return ny == 512 && nz == 512 && nx >= 8 && line_pass_available();
}

static int chain_lag() {
  const char* e = getenv("EVX_FFT_CHAIN_LAG");
  const int v = e ? atoi(e) : 24;
  return v < 1 ? 1 : v;
}
template <bool INV, bool STATS, int NBUF, int TW>
static int chain_launch_t(ChainParams p, const ChainMaps& maps, cudaStream_t st) {
  constexpr size_t smem = chain_smem_bytes(NBUF);
  auto kern = fft_chain_kernel<INV, STATS, NBUF, TW>;
  static SmemOptIn optin;
  if (int rc = optin.ensure(kern, smem)) return rc;
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, chain_threads(NBUF), smem) != cudaSuccess || per_sm < 1)
    return EVX_ERR_UNSUPPORTED;
  if (per_sm > (NBUF <= 2 ? 2 : 1)) per_sm = NBUF <= 2 ? 2 : 1;
  long long grid = (long long)sms * per_sm;
  if (grid > p.sched.total) grid = p.sched.total;
  cudaError_t e = cudaMemsetAsync(p.done0, 0, (size_t)p.sched.nplanes * sizeof(unsigned), st);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(chain_threads(NBUF));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all blocks co-resident: the waits cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, maps, p);
  count_launch();
  return (int)e;
}

int chain_launch(bool inverse, const ChainArgs& a, const void* tmap_y, cudaStream_t st, const void* tmap_blk0,
                 const void* tmap_blk1) {
  if (!chain_supported(a.nx, a.ny, a.nz)) return EVX_ERR_UNSUPPORTED;
  ChainMaps maps;
  maps.spec = *(const CUtensorMap*)tmap_y;
  maps.blk[0] = *(const CUtensorMap*)(tmap_blk0 ? tmap_blk0 : tmap_y);
  maps.blk[1] = *(const CUtensorMap*)(tmap_blk1 ? tmap_blk1 : tmap_y);
  ChainParams p;
  p.blk_mode = (tmap_blk0 && tmap_blk1) ? 1 : 0;
  const int ytiles = (a.nz / 2 + 1 + 7) / 8;
  const int zitems = a.ny / kChainZLines;
  p.sched = inverse ? make_chain_schedule(a.nx, chain_lag(), ytiles, zitems)
                    : make_chain_schedule(a.nx, chain_lag(), zitems, ytiles);
  p.real_in = a.real_in; p.real_out = a.real_out; p.spec = (cf*)a.spec;
  p.twz = (const cf*)a.twz; p.twr = (const cf*)a.twr; p.twy = (const cf*)a.twy;
  p.ny = a.ny; p.nz = a.nz; p.P = a.P;
  p.done0 = (unsigned*)a.flags;
  const char* ed = getenv("EVX_FFT_CHAIN_DISCARD");
  p.discard = ed ? atoi(ed) : 1;      // inverse: 2.27 -> 1.77 GB of DRAM traffic per launch at 512^3 (ncu)
  const char* es = getenv("EVX_FFT_CHAIN_STATS");
  p.stats = (es && atoi(es) != 0) ? (unsigned long long*)a.stats : nullptr;
  const char* eb = getenv("EVX_FFT_CHAIN_NBUF");
  const int nbuf = eb ? atoi(eb) : 3;     // measured at 512^3: 3 buffers 0.433 / 0.449 ms, 2: 0.425 / 0.476, 4: 0.441 / 0.451
  const char* et = getenv("EVX_FFT_CHAIN_TW");
  // forward: all roots in registers (0.435 -> 0.339 ms at 512^3); the inverse kernel also holds
  // a line's eight u values and spills with more than the y twiddles (0.449 -> 0.421 ms)
  const int tw = et ? atoi(et) : (inverse ? 3 : 2);
#define EVX_CHAIN(NB, TWV)                                                                         \
  {                                                                                                \
    if (p.stats) return inverse ? chain_launch_t<true, true, NB, TWV>(p, maps, st)               \
                                : chain_launch_t<false, true, NB, TWV>(p, maps, st);            \
    return inverse ? chain_launch_t<true, false, NB, TWV>(p, maps, st)                           \
                   : chain_launch_t<false, false, NB, TWV>(p, maps, st);                        \
  }
  if (nbuf == 2) EVX_CHAIN(2, 0)
  if (tw <= 0) EVX_CHAIN(3, 0)
  if (tw == 1) EVX_CHAIN(3, 1)
  if (tw == 3) EVX_CHAIN(3, 3)
  EVX_CHAIN(3, 2)
#undef EVX_CHAIN
}

}  // namespace evx
