// Kernel (b''): second tcgen05 pass of the exactness layer (/root/reference/model/SearchTransfer.py:33-34).
//
// The first pass (relevance_tcs.cu / relevance_tc.cu) keeps the kTopK best bf16 scores per (query, key segment).  A query
// whose list is saturated -- its k-th entry is still at or above the query's threshold T (rescore.cu) -- may have lost keys
// that matter.  Its threshold is known by then, so this pass does not rank anything: it recomputes the bf16 scores of the
// QUEUED queries against ALL keys on the tensor cores and emits every (query, key) pair with score >= T; the pairs are
// rescored exactly and folded into the query's packed maximum.  Cost is proportional to the number of queued queries
// (0.2 - 1.5 % of a 720p frame with the certified window), where the round-1 fallback paid 133 MFLOP of fp32 CUDA-core
// work per queued query.
//
//   flag_pack_kernel          gathers the 3x3x128 bf16 patch of every queued query from the staged query image into a
//                             dense A operand  [tile][tap 9][channel group 16][row 128][8 ch]  = the canonical K-major
//                             no-swizzle UMMA layout (core matrix 8 rows x 16 B; SBO 128 B, LBO 2 KB), so a pipeline
//                             stage (one tap x 32 channels) is one contiguous 8 KB bulk copy.
//   relevance_flagged_kernel  implicit GEMM, M = 128 packed queries, N = 8 x 32 key positions, K = 9 taps x 128 channels
//                             (72 tcgen05.mma per tile pair, fp32 accumulation in TMEM, two accumulators).  The key operand
//                             is the dense kernel's: one halo tile of the staged key image, resident in shared memory while
//                             every packed query tile streams past it; the 9 taps are descriptor start offsets into it.
//                             The A operand streams from L2 (8 KB stages).  Epilogue: thread = query row, score = acc * rk,
//                             warp-aggregated append of the pairs at or above the row's threshold.
//   rescore_emitted_kernel    one warp per emitted pair: exact fp32/fp64 relevance, packed (score, ~key) atomicMax.
//
// Nothing here synchronises with the host: the queue lengths live on the device, the grids are fixed and idle CTAs exit.
// If the queue exceeds the packed-row capacity, or the emission buffer fills, a device flag hands the queued queries to
// the exhaustive fp32 search instead (rescore.cu) -- slow but still exact.
#include <cuda.h>

#include "spei_common.cuh"
#include "tc_ptx.cuh"
#include "exact_score.cuh"

namespace spei {

constexpr int kFThreadsF = 256;
constexpr uint32_t kFHaloU = kTileU + 2;                                  // 10 positions
constexpr uint32_t kFRowBytes = kFHaloU * 16;                             // 160 B: one halo-tile row = SBO of the key operand
constexpr uint32_t kFKLbo = (kFlagNy + 2) * kFRowBytes;                   // 5440 B: one channel-group plane of the halo tile
constexpr uint32_t kFBBytes = kCG * kFKLbo;                               // 87040 B: the resident key tile
constexpr uint32_t kFAStageBytes = kCGS * 128 * 16;                       // 8192 B: one tap x 4 channel groups x 128 rows
constexpr uint32_t kFALbo = 128 * 16;                                     // 2048 B between channel groups of the A operand
constexpr int kFAStages = 6;
constexpr int kFStagesPerPair = 9 * (kCG / kCGS);                         // 36
constexpr uint32_t kFNumBars = 2 * kFAStages + 2 + 4;
constexpr uint32_t kFOffA = (kFBBytes + 1023) / 1024 * 1024;
constexpr uint32_t kFOffBars = kFOffA + kFAStages * kFAStageBytes;
constexpr uint32_t kFOffRk = kFOffBars + kFNumBars * 8 + 16;
constexpr uint32_t kFSmemBytes = kFOffRk + 4 * 256 * 4;
static_assert(kFOffRk % 16 == 0, "key-norm staging must be 16-byte aligned");
constexpr uint32_t kFTmemCols = 512, kFAccCols = 256;

struct FlagParams {
  int n, rf, H, W, Hr, Wr, L, lk1;
  int q_orient, q_Upad, q_Vpad;
  int k_orient, Uk, Vk, k_tu, k_tiles_img, KT;
  int cap_rows;
  const __nv_bfloat16* qbf;
  const int32_t* flag_list;
  int32_t* counters;
  const float* thr;
  __nv_bfloat16* apack;
  float* prow_thr;
  int32_t* prow_q;
  const float* rk;
  int32_t* emit_q;
  int32_t* emit_k;
  const float *q32, *k32, *rq;
  unsigned long long* packed;
  int32_t* stats;
  int* error_flag;
};

// packed 128-row tiles of item `item` and of the items before it (queue lengths are device data)
__device__ __forceinline__ void item_tiles(const int32_t* counters, int item, int& base, int& tiles) {
  base = 0;
  for (int i = 0; i < item; ++i) base += (__ldg(counters + i) + 127) >> 7;
  tiles = (__ldg(counters + item) + 127) >> 7;
}
__device__ __forceinline__ bool flagged_disabled(const FlagParams& p) {
  return *reinterpret_cast<const volatile int32_t*>(p.counters + p.n + kCntExhaust) != 0;
}

// ---------------------------------------------------------------------------------------------
// A operand: one warp per packed row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flag_pack_kernel(const FlagParams p) {
  const int lane = threadIdx.x & 31;
  int total_tiles = 0;
  for (int i = 0; i < p.n; ++i) total_tiles += (__ldg(p.counters + i) + 127) >> 7;
  if (total_tiles == 0) return;
  if (total_tiles * 128 > p.cap_rows) {   // more queued queries than packed rows: the exhaustive search takes them all
    if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[p.n + kCntExhaust] = 1;
    return;
  }
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total_tiles * 128; row += warps) {
    const int tile = row >> 7, r = row & 127;
    int item = 0, base = 0;
    while (true) {
      const int t = (__ldg(p.counters + item) + 127) >> 7;
      if (tile < base + t) break;
      base += t; ++item;
    }
    const int idx = (tile - base) * 128 + r;
    const bool live = idx < __ldg(p.counters + item);
    int ql = 0;
    if (live) ql = __ldg(p.flag_list + (size_t)item * p.L + idx);
    if (lane == 0) {
      p.prow_thr[row] = live ? __ldg(p.thr + (size_t)item * p.L + ql) : INFINITY;
      p.prow_q[row] = live ? item * p.L + ql : -1;
    }
    const int y = ql / p.W, x = ql - y * p.W;
    const int u = p.q_orient == 0 ? x : y, v = p.q_orient == 0 ? y : x;
    const size_t plane = (size_t)p.q_Vpad * p.q_Upad;
    const uint4* src = reinterpret_cast<const uint4*>(p.qbf) + (size_t)item * kCG * plane;
    uint4* dst = reinterpret_cast<uint4*>(p.apack) + (size_t)tile * 9 * kCG * 128 + r;
    for (int e = lane; e < 9 * kCG; e += 32) {
      const int tap = e / kCG, cg = e - tap * kCG;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;                    // image-space offsets (ki, kj) - 1
      const int du = p.q_orient == 0 ? dx : dy, dv = p.q_orient == 0 ? dy : dx;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (live) val = __ldg(src + (size_t)cg * plane + (size_t)(v + 1 + dv) * p.q_Upad + (u + 1 + du));   // zero border = zero padding
      dst[(size_t)e * 128] = val;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bf16 scores of the packed queries against all keys, emission of the pairs at or above threshold
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kFThreadsF, 1)
relevance_flagged_kernel(const __grid_constant__ CUtensorMap tmk, const FlagParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  int total_tiles = 0;
  for (int i = 0; i < p.n; ++i) total_tiles += (__ldg(p.counters + i) + 127) >> 7;
  if (total_tiles == 0 || flagged_disabled(p)) return;   // uniform over the grid: decided before this launch

  const uint32_t sB = smem_u32(smem), sA = sB + kFOffA, bars = sB + kFOffBars;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kFAStages;
  const uint32_t bar_bfull = bars + 16 * kFAStages, bar_bfree = bar_bfull + 8;
  const uint32_t bar_tfull = bar_bfull + 16, bar_tempty = bar_bfull + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kFOffBars + kFNumBars * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, G = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kFAStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_bfull, 1); mbar_init(bar_bfree, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmk) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kFTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work of this CTA: for every item, the (key tile, packed query tile) pairs [pb, pe) of P = KT * tiles in key-major
  // order, so consecutive pairs share the resident key tile.  All three roles walk the same sequence.
  if (warp == 0) {
    // ===================================== producer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, bloads = 0;
      for (int item = 0; item < p.n; ++item) {
        int tbase, tiles;
        item_tiles(p.counters, item, tbase, tiles);
        if (tiles == 0) continue;
        const long long P = (long long)p.KT * tiles, pb = (long long)b * P / G, pe = (long long)(b + 1) * P / G;
        int last_kt = -1;
        for (long long pp = pb; pp < pe; ++pp) {
          const int kt = (int)(pp / tiles), qt = (int)(pp - (long long)kt * tiles);
          if (kt != last_kt) {
            if (bloads > 0) mbar_wait(bar_bfree, (bloads - 1) & 1u, p.error_flag);   // MMAs of the previous key tile are done
            const int f = kt / p.k_tiles_img, kti = kt - f * p.k_tiles_img;
            const int ktv = kti / p.k_tu, ktu = kti - ktv * p.k_tu;
            mbar_arrive_expect_tx(bar_bfull, kFBBytes);
            for (int g4 = 0; g4 < kCG / kCGS; ++g4)
              tma_load_4d(sB + g4 * kCGS * kFKLbo, &tmk, bar_bfull, ktu * kTileU * 8, ktv * kFlagNy, g4 * kCGS, item * p.rf + f);
            ++bloads;
            last_kt = kt;
          }
          const char* asrc = reinterpret_cast<const char*>(p.apack) + (size_t)(tbase + qt) * 9 * kCG * 128 * 16;
          for (int s = 0; s < kFStagesPerPair; ++s) {
            mbar_wait_parked(bar_empty + 8 * stage, phase ^ 1, p.error_flag);
            mbar_arrive_expect_tx(bar_full + 8 * stage, kFAStageBytes);
            bulk_load_1d(sA + stage * kFAStageBytes, asrc + (size_t)s * kFAStageBytes, kFAStageBytes, bar_full + 8 * stage);
            if (++stage == kFAStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tile_i = 0, buses = 0;
      const uint32_t k_dki = p.k_orient == 0 ? kFRowBytes : 16u, k_dkj = p.k_orient == 0 ? 16u : kFRowBytes;
      // kind::f16: D = f32, A = B = bf16, K-major both, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      for (int item = 0; item < p.n; ++item) {
        int tbase, tiles;
        item_tiles(p.counters, item, tbase, tiles);
        if (tiles == 0) continue;
        const long long P = (long long)p.KT * tiles, pb = (long long)b * P / G, pe = (long long)(b + 1) * P / G;
        int last_kt = -1;
        for (long long pp = pb; pp < pe; ++pp, ++tile_i) {
          const int kt = (int)(pp / tiles);
          if (kt != last_kt) {
            mbar_wait(bar_bfull, buses & 1u, p.error_flag);
            ++buses;
            last_kt = kt;
          }
          const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
          mbar_wait_parked(bar_tempty + 8 * acc, (use & 1u) ^ 1u, p.error_flag);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kFAccCols;
          for (uint32_t s = 0; s < (uint32_t)kFStagesPerPair; ++s) {
            mbar_wait_parked(bar_full + 8 * stage, phase, p.error_flag);
            tc_fence_after();
            const uint32_t tap = s >> 2, cq = s & 3u, ki = tap / 3, kj = tap - ki * 3;
            const uint32_t abase = sA + stage * kFAStageBytes;
            const uint32_t kbase = sB + cq * kCGS * kFKLbo + ki * k_dki + kj * k_dkj;
#pragma unroll
            for (uint32_t cgp = 0; cgp < kCGS / 2; ++cgp) {
              const uint64_t adesc = umma_desc_kmajor(abase + cgp * 2 * kFALbo, kFALbo, 128);
              const uint64_t bdesc = umma_desc_kmajor(kbase + cgp * 2 * kFKLbo, kFKLbo, kFRowBytes);
              tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (s | cgp) != 0u);
            }
            tc_commit(bar_empty + 8 * stage);
            if (++stage == kFAStages) { stage = 0; phase ^= 1; }
          }
          tc_commit(bar_tfull + 8 * acc);
          const bool last_of_kt = (pp + 1 == pe) || ((int)((pp + 1) / tiles) != kt);
          if (last_of_kt) tc_commit(bar_bfree);   // the resident key tile may be overwritten once these MMAs have read it
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int ew = warp - 4, m = ew * 32 + lane;
    float* rk_s = reinterpret_cast<float*>(smem + kFOffRk) + ew * 256;
    int32_t* emit_count = p.counters + p.n + kCntEmit;
    uint32_t tile_i = 0;
    for (int item = 0; item < p.n; ++item) {
      int tbase, tiles;
      item_tiles(p.counters, item, tbase, tiles);
      if (tiles == 0) continue;
      const long long P = (long long)p.KT * tiles, pb = (long long)b * P / G, pe = (long long)(b + 1) * P / G;
      int last_kt = -1, f = 0, ku0 = 0, kv0 = 0;
      for (long long pp = pb; pp < pe; ++pp, ++tile_i) {
        const int kt = (int)(pp / tiles), qt = (int)(pp - (long long)kt * tiles);
        if (kt != last_kt) {
          f = kt / p.k_tiles_img;
          const int kti = kt - f * p.k_tiles_img, ktv = kti / p.k_tu, ktu = kti - ktv * p.k_tu;
          ku0 = ktu * kTileU; kv0 = ktv * kFlagNy;
          const float* rkimg = p.rk + ((size_t)item * p.rf + f) * p.lk1;
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = lane * 8 + i, ku = ku0 + (c & 7), kv = kv0 + (c >> 3);
            float r = __int_as_float(0x7fc00000);   // NaN: a position outside the image never compares >=
            if (ku < p.Uk && kv < p.Vk) r = __ldg(rkimg + uv_to_linear(p.k_orient, ku, kv, p.Wr));
            rk_s[c] = r;
          }
          __syncwarp();
          last_kt = kt;
        }
        const int grow = (tbase + qt) * 128 + m;
        const float thr = __ldg(p.prow_thr + grow);
        const int qid = __ldg(p.prow_q + grow);
        const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
        mbar_wait_parked(bar_tfull + 8 * acc, use & 1u, p.error_flag);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * kFAccCols + ((uint32_t)(ew * 32) << 16);
        uint32_t a[16];
        tc_ld16(taddr, a);
#pragma unroll 1
        for (int r2 = 0; r2 < 16; ++r2) {
          tc_wait_ld();
          unsigned msk = 0;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 r = reinterpret_cast<const float4*>(rk_s + r2 * 16)[i4];   // broadcast reads
            msk |= (__uint_as_float(a[4 * i4 + 0]) * r.x >= thr) ? (1u << (4 * i4 + 0)) : 0u;
            msk |= (__uint_as_float(a[4 * i4 + 1]) * r.y >= thr) ? (1u << (4 * i4 + 1)) : 0u;
            msk |= (__uint_as_float(a[4 * i4 + 2]) * r.z >= thr) ? (1u << (4 * i4 + 2)) : 0u;
            msk |= (__uint_as_float(a[4 * i4 + 3]) * r.w >= thr) ? (1u << (4 * i4 + 3)) : 0u;
          }
          if (r2 + 1 < 16) tc_ld16(taddr + (r2 + 1) * 16, a);
          if (__any_sync(0xffffffffu, msk != 0u)) {
            // warp-aggregated append: one atomic per warp and chunk
            const int cnt = __popc(msk);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int up = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0;
            if (lane == 0) base = atomicAdd(emit_count, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            int pos = base + incl - cnt;
            while (msk) {
              const int i = __ffs(msk) - 1;
              msk &= msk - 1;
              if (pos < kFlagMaxEmit) {
                const int ku = ku0 + (i & 7), kv = kv0 + 2 * r2 + (i >> 3);
                p.emit_q[pos] = qid;
                p.emit_k[pos] = f * p.lk1 + uv_to_linear(p.k_orient, ku, kv, p.Wr);
              } else {
                p.counters[p.n + kCntExhaust] = 1;   // emission buffer full: the exhaustive search finishes the job
              }
              ++pos;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kFTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// exact relevance of the emitted pairs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4)
rescore_emitted_kernel(const FlagParams p) {
  __shared__ __align__(16) float qpatch[8][9 * kC3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int raw = *reinterpret_cast<const volatile int32_t*>(p.counters + p.n + kCntEmit);
  const int count = raw < kFlagMaxEmit ? raw : kFlagMaxEmit;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.stats) { p.stats[5] = raw; p.stats[3] = p.counters[p.n + kCntExhaust]; }
  const int warps = gridDim.x * 8;
  int nres = 0;
  for (int e = blockIdx.x * 8 + warp; e < count; e += warps) {
    const int qid = __ldg(p.emit_q + e), jj = __ldg(p.emit_k + e);
    if (qid < 0) continue;
    const int item = qid / p.L, ql = qid - item * p.L;
    __syncwarp();
    load_query_patch_async(&qpatch[warp][0], p.q32 + (size_t)item * p.L * kC3, ql / p.W, ql % p.W, p.H, p.W, lane);
    const float rq = __ldg(p.rq + qid);
    const int f = jj / p.lk1, rem = jj - f * p.lk1, hr = rem / p.Wr, wr = rem - hr * p.Wr;
    const float rk = __ldg(p.rk + ((size_t)item * p.rf + f) * p.lk1 + rem);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const float4* qv = reinterpret_cast<const float4*>(&qpatch[warp][0]);
    const float score = exact_relevance([&](int t) { return qv + t * 32; }, p.k32 + ((size_t)item * p.rf + f) * p.lk1 * kC3, hr, wr, p.Hr, p.Wr, rq, rk, lane);
    if (lane == 0) atomicMax(p.packed + qid, pack_score(score, jj));
    ++nres;
  }
  if (lane == 0 && nres && p.stats) atomicAdd(p.stats + 1, nres);
}

// 4-D map over a staged key image [img][16][Vpad][Upad*8] bf16; box = [1][4 groups][kFlagNy + 2 rows][10 positions x 8 channels]
static int make_map_f(EncodeTiledFn enc, CUtensorMap* tm, void* base, int nimg, const OperandPlan& o) {
  const cuuint64_t dims[4] = {(cuuint64_t)o.Upad * 8, (cuuint64_t)o.Vpad, (cuuint64_t)kCG, (cuuint64_t)nimg};
  const cuuint64_t strides[3] = {(cuuint64_t)o.Upad * 16, (cuuint64_t)o.Vpad * o.Upad * 16, (cuuint64_t)kCG * o.Vpad * o.Upad * 16};
  const cuuint32_t box[4] = {kFHaloU * 8, (cuuint32_t)(kFlagNy + 2), (cuuint32_t)kCGS, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (second pass) failed with CUresult %d", (int)r); return SPEI_ERR_CUDA; }
  return SPEI_OK;
}

int launch_relevance_flagged(const Plan& p, int32_t* stats, char* ws, cudaStream_t st) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  CUtensorMap tmk;
  if ((rc = make_map_f(enc, &tmk, ws + p.off_kbf, p.n * p.rf, p.k))) return rc;
  FlagParams f{};
  f.n = p.n; f.rf = p.rf; f.H = p.H; f.W = p.W; f.Hr = p.Hr; f.Wr = p.Wr; f.L = p.H * p.W; f.lk1 = p.Hr * p.Wr;
  f.q_orient = p.q.orient; f.q_Upad = p.q.Upad; f.q_Vpad = p.q.Vpad;
  f.k_orient = p.k.orient; f.Uk = p.k.U; f.Vk = p.k.V;
  f.k_tu = (p.k.U + kTileU - 1) / kTileU;
  f.k_tiles_img = f.k_tu * ((p.k.V + kFlagNy - 1) / kFlagNy);
  f.KT = p.rf * f.k_tiles_img;
  f.cap_rows = p.flag_rows;
  f.qbf = (const __nv_bfloat16*)(ws + p.off_qbf);
  f.flag_list = (const int32_t*)(ws + p.off_flag);
  f.counters = (int32_t*)(ws + p.off_counters);
  f.thr = (const float*)(ws + p.off_thr);
  f.apack = (__nv_bfloat16*)(ws + p.off_apack);
  f.prow_thr = (float*)(ws + p.off_prow_thr);
  f.prow_q = (int32_t*)(ws + p.off_prow_q);
  f.rk = (const float*)(ws + p.off_rk);
  f.emit_q = (int32_t*)(ws + p.off_emit_q);
  f.emit_k = (int32_t*)(ws + p.off_emit_k);
  f.q32 = (const float*)(ws + p.off_q32); f.k32 = (const float*)(ws + p.off_k32); f.rq = (const float*)(ws + p.off_rq);
  f.packed = (unsigned long long*)(ws + p.off_packed);
  f.stats = stats;
  f.error_flag = (int*)(ws + p.off_errflag);
  int dev = 0, sms = 0;
  SPEI_CUDA(cudaGetDevice(&dev));
  SPEI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  flag_pack_kernel<<<2 * sms, 256, 0, st>>>(f);
  SPEI_CUDA(cudaGetLastError());
  SPEI_CUDA(cudaFuncSetAttribute(relevance_flagged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmemBytes));
  SPEI_CUDA(cudaFuncSetAttribute(relevance_flagged_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  relevance_flagged_kernel<<<sms, kFThreadsF, kFSmemBytes, st>>>(tmk, f);
  SPEI_CUDA(cudaGetLastError());
  rescore_emitted_kernel<<<2 * sms, 256, 0, st>>>(f);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
