// Kernel (b): relevance bmm fused with the row-wise max / argmax
// (/root/reference/model/SearchTransfer.py:33-34) as a tcgen05 / TMEM implicit GEMM.
//
//   R[j, i] = rk[j] * rq[i] * sum_{ki,kj,c} K[c, hr+ki-1, wr+kj-1] * Q[c, y+ki-1, x+kj-1]
//
// The unfolded [9C x L] operands of the reference never exist.  Both operands are staged by kernel
// (a) as channel-group-planar bf16 images ([16 groups][V][U][8 channels], zero border).  In that
// layout the canonical no-swizzle K-major UMMA operand (8 rows x 16 bytes core matrices, row-group
// stride SBO, K-chunk stride LBO) of an 8-wide spatial tile is exactly a window of a halo tile:
// SBO = one halo-tile row (10 pixels x 16 B), LBO = one channel-group plane of the halo tile, and a
// 3x3 patch shift (ki,kj) is just a start-address offset of (ki*10 + kj)*16 bytes.  So one TMA load of
// a (8+2) x (Ty+2) halo tile feeds all 9 shifts: shared-memory fill traffic drops 9x versus loading
// each shift, which is what keeps a single CTA per SM inside the L2 -> SM bandwidth budget.
//
// Per CTA (persistent, one per SM):
//   warp 0   TMA producer: query halo tile (16 groups, 46 KB, resident per query tile) and a ring of
//            key halo-tile stages (4 channel groups = 2 x K16 per stage)
//   warp 1   MMA issuer: per key tile 8 K16-steps x 9 shifts = 72 tcgen05.mma (M=128 queries,
//            N=8*Ny keys, fp32 accumulate in TMEM), two accumulators (2 x 256 TMEM columns) so the
//            epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld (one query row per thread: no cross-lane reduction), scale by the
//            key's reciprocal patch norm, keep the best kTopK (score, key) per query in registers.
// The full relevance matrix (13.3 GB at 720p) is never written; per (query tile, key segment) only the
// kTopK candidates per query leave the SM.  Exact fp32 rescoring + tie-break happen in rescore.cu.
//
// Work split: the P = n*QT*KT (query tile, key tile) pairs are cut into G equal contiguous ranges,
// one per CTA, so all SMs finish together; a query tile whose key range straddles CTAs simply gets
// one candidate list per segment.
#include <cuda.h>

#include "spei_common.cuh"
#include "tc_ptx.cuh"

namespace spei {

constexpr int kThreads = 256;
constexpr uint32_t kHaloU = kTileU + 2;                                     // 10 pixels
constexpr uint32_t kRowBytes = kHaloU * 16;                                 // 160 B = SBO
constexpr uint32_t kQTileBytes = kCG * (kQTileV + 2) * kRowBytes;           // 46080
constexpr uint32_t kQLBO = (kQTileV + 2) * kRowBytes;                       // 2880
constexpr uint32_t kStageBytesMax = kCGS * (kMaxNy + 2) * kRowBytes;        // 21760
constexpr uint32_t kStagesPerTile = kCG / kCGS;                             // 4
constexpr uint32_t kNumBars = 2 * kStages + 6;
constexpr uint32_t kRkSmemOffset = kQTileBytes + kStages * kStageBytesMax + kNumBars * 8 + 16;  // 4 warps x 256 floats
constexpr uint32_t kSmemBytes = kRkSmemOffset + 4 * 256 * 4;
static_assert(kRkSmemOffset % 16 == 0, "key-norm staging must be float4 aligned");
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 256;

struct TcParams {
  int n, rf, QT, KT, G, maxseg;
  long long P;
  int q_tu, q_orient, Uq, Vq, W, L;
  int k_tu, k_tiles_img, k_orient, Ny, Wr, lk1, UkT, VkT;
  uint32_t idesc, stage_bytes, k_lbo;
  float win;         // candidate window (normalised relevance units, slightly wider than the rescoring's eps)
  const float* rq;   // query reciprocal patch norms [n*L]: the epilogue's scores lack this factor
  const float* dq;     // per-query relative bf16 residual norm (stage_norm.cu)
  const int* dkmax;    // per-item maximum over the keys, float bits
  const float* rkpad;
  float* cval;
  int32_t* cidx;
  float* debug_acc;  // optional [128][256] raw accumulator dump of pair 0
  int* error_flag;   // set on a barrier timeout
};

// K-major, no swizzle, SBO = one halo-tile row
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return umma_desc_kmajor(saddr, lbo_bytes, kRowBytes);
}

template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 1)
relevance_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + kQTileBytes;
  const uint32_t bars = sK + kStages * kStageBytesMax;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages;
  const uint32_t bar_qfull = bars + 16 * kStages, bar_qfree = bar_qfull + 8;
  const uint32_t bar_tfull = bar_qfull + 16, bar_tempty = bar_qfull + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kQTileBytes + kStages * kStageBytesMax + kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const long long pb = (long long)b * p.P / p.G, pe = (long long)(b + 1) * p.P / p.G;
  const long long cyc0 = clock64();   // CTA 0 publishes its clock64 span: the SM clock this kernel really ran at (bench.py)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_qfull, 1); mbar_init(bar_qfree, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmq) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmk) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int qloaded = 0;
      PairIdx ix = decode_pair(pb, p.QT, p.KT);
      for (long long pp = pb; pp < pe; ++pp, ix = next_pair(ix, p.QT, p.KT)) {
        if (pp == pb || ix.kt == 0) {
          if (qloaded > 0) mbar_wait(bar_qfree, (uint32_t)((qloaded - 1) & 1), p.error_flag);
          const int qtv = ix.qt / p.q_tu, qtu = ix.qt - qtv * p.q_tu;
          mbar_arrive_expect_tx(bar_qfull, kQTileBytes);
          tma_load_4d(sQ, &tmq, bar_qfull, qtu * kTileU * 8, qtv * kQTileV, 0, ix.item);
          ++qloaded;
        }
        const int f = ix.kt / p.k_tiles_img, kti = ix.kt - f * p.k_tiles_img;
        const int ktv = kti / p.k_tu, ktu = kti - ktv * p.k_tu;
        for (uint32_t s4 = 0; s4 < kStagesPerTile; ++s4) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, p.error_flag);
          mbar_arrive_expect_tx(bar_full + 8 * stage, p.stage_bytes);
          tma_load_4d(sK + stage * kStageBytesMax, &tmk, bar_full + 8 * stage, ktu * kTileU * 8, ktv * p.Ny, (int)(s4 * kCGS),
                      ix.item * p.rf + f);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int qused = 0;
      uint32_t tile_i = 0;
      // byte offset of patch tap (ki,kj) inside a halo tile, per operand orientation
      const uint32_t q_dki = p.q_orient == 0 ? kRowBytes : 16u, q_dkj = p.q_orient == 0 ? 16u : kRowBytes;
      const uint32_t k_dki = p.k_orient == 0 ? kRowBytes : 16u, k_dkj = p.k_orient == 0 ? 16u : kRowBytes;
      PairIdx ix = decode_pair(pb, p.QT, p.KT);
      for (long long pp = pb; pp < pe; ++pp, ++tile_i, ix = next_pair(ix, p.QT, p.KT)) {
        if (pp == pb || ix.kt == 0) {
          mbar_wait(bar_qfull, (uint32_t)(qused & 1), p.error_flag);
          ++qused;
        }
        const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
        mbar_wait(bar_tempty + 8 * acc, (use & 1u) ^ 1u, p.error_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (uint32_t s4 = 0; s4 < kStagesPerTile; ++s4) {
          mbar_wait(bar_full + 8 * stage, phase, p.error_flag);
          tc_fence_after();
          const uint32_t kbase = sK + stage * kStageBytesMax;
#pragma unroll
          for (uint32_t cgp = 0; cgp < kCGS / 2; ++cgp) {
            const uint32_t qa = sQ + (s4 * kCGS + cgp * 2) * kQLBO;
            const uint32_t ka = kbase + (cgp * 2) * p.k_lbo;
#pragma unroll
            for (uint32_t tap = 0; tap < 9; ++tap) {
              const uint32_t ki = tap / 3, kj = tap % 3;
              const uint64_t adesc = umma_desc(qa + ki * q_dki + kj * q_dkj, kQLBO);
              const uint64_t bdesc = umma_desc(ka + ki * k_dki + kj * k_dkj, p.k_lbo);
              tc_mma_bf16(d_tmem, adesc, bdesc, p.idesc, (s4 | cgp | tap) != 0u);
            }
          }
          tc_commit(bar_empty + 8 * stage);  // stage reusable once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);      // accumulator complete -> epilogue
        if (ix.kt == p.KT - 1 && pp + 1 < pe) tc_commit(bar_qfree);  // next query tile may overwrite this one
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int ew = warp - 4;              // TMEM lane quarter this warp may read
    const int m = ew * 32 + lane;         // MMA row = query inside the tile
    float tv[kTopK];
    int ti[kTopK];
    uint32_t tile_i = 0;
    long long qlin = -1;
    float winq = 0.f;  // the window in this query's un-normalised score units
    float* rk_s = reinterpret_cast<float*>(smem + kRkSmemOffset) + ew * kAccCols;  // this warp's copy of the tile's key norms
    const int nchunk = p.Ny / 2;
    // key-norm prefetch: lane l owns float4 #l and #(l+32) of the tile's [Ny][8] reciprocal norms
    auto rk_prefetch = [&](const PairIdx ix, float4 (&pre)[2]) {
      const int f = ix.kt / p.k_tiles_img, kti = ix.kt - f * p.k_tiles_img;
      const int ktv = kti / p.k_tu, ktu = kti - ktv * p.k_tu;
      const float* rkrow = p.rkpad + ((size_t)(ix.item * p.rf + f) * p.VkT + ktv * p.Ny) * p.UkT + ktu * kTileU;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e4 = lane + 32 * j, row = e4 >> 1;
        pre[j] = row < p.Ny ? __ldg(reinterpret_cast<const float4*>(rkrow + (size_t)row * p.UkT) + (e4 & 1))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 pre[2];
    PairIdx ix = decode_pair(pb, p.QT, p.KT);
    if (pb < pe) rk_prefetch(ix, pre);
    for (long long pp = pb; pp < pe; ++pp, ++tile_i, ix = next_pair(ix, p.QT, p.KT)) {
      if (pp == pb || ix.kt == 0) {
#pragma unroll
        for (int s = 0; s < kTopK; ++s) { tv[s] = -INFINITY; ti[s] = -1; }
        const int qtv = ix.qt / p.q_tu, qtu = ix.qt - qtv * p.q_tu;
        const int u = qtu * kTileU + (m & 7), v = qtv * kQTileV + (m >> 3);
        qlin = (u < p.Uq && v < p.Vq) ? (long long)ix.item * p.L + uv_to_linear(p.q_orient, u, v, p.W) : -1;
        winq = 0.f;
        if (qlin >= 0) {
          // fixed window (eps > 0) or the certified one (spei_common.cuh: certified_window)
          const float wn = p.win > 0.f ? p.win
                                       : 1.02f * certified_window(certified_delta(__ldg(p.dq + qlin), __int_as_float(__ldg(p.dkmax + ix.item))));
          winq = wn / __ldg(p.rq + qlin);
        }
#ifdef SPEI_NO_WINDOW  // A/B only: plain top-k insertion
        winq = INFINITY;
#endif
      }
      const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
      const int f = ix.kt / p.k_tiles_img, kti = ix.kt - f * p.k_tiles_img;
      const int ktv = kti / p.k_tu, ktu = kti - ktv * p.k_tu;
      const int ku0 = ktu * kTileU, kv0 = ktv * p.Ny;
      // publish this tile's key norms to the warp (the loads were issued one tile ago), then start the
      // loads for the next tile so their latency hides behind this tile's work
      __syncwarp();
      reinterpret_cast<float4*>(rk_s)[lane] = pre[0];
      reinterpret_cast<float4*>(rk_s)[lane + 32] = pre[1];
      __syncwarp();
      if (pp + 1 < pe) rk_prefetch(next_pair(ix, p.QT, p.KT), pre);
      mbar_wait(bar_tfull + 8 * acc, use & 1u, p.error_flag);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccCols + ((uint32_t)(ew * 32) << 16);

      // TMEM reads are double-buffered in registers: chunk c+1 is in flight while chunk c is processed.
      // The code below exists once (copying 16 registers is cheaper than a second inlined copy: the
      // epilogue shares the SM's instruction cache with the MMA issuer, and a fat epilogue starves it).
      uint32_t a[16];
      tc_ld16(taddr, a);
      for (int r2 = 0; r2 < nchunk; ++r2) {
        tc_wait_ld();
        float v[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 r = reinterpret_cast<const float4*>(rk_s + r2 * 16)[i4];  // broadcast read
          v[4 * i4 + 0] = __uint_as_float(a[4 * i4 + 0]) * r.x;  // NaN for padded keys
          v[4 * i4 + 1] = __uint_as_float(a[4 * i4 + 1]) * r.y;
          v[4 * i4 + 2] = __uint_as_float(a[4 * i4 + 2]) * r.z;
          v[4 * i4 + 3] = __uint_as_float(a[4 * i4 + 3]) * r.w;
        }
        if (kDebug && p.debug_acc && pp == 0) {   // compile-time: the production instantiation carries no debug branch
#pragma unroll
          for (int i = 0; i < 16; ++i) p.debug_acc[(size_t)m * kAccCols + r2 * 16 + i] = __uint_as_float(a[i]);
        }
        if (r2 + 1 < nchunk) tc_ld16(taddr + (r2 + 1) * 16, a);  // v[] holds this chunk; refill a[] asynchronously
        float mx = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(v[8], v[9]), fmaxf(v[10], v[11])), fmaxf(fmaxf(v[12], v[13]), fmaxf(v[14], v[15]))));
        // A score can only matter to the rescoring if it is within `win` of the best score seen so far
        // (the final best is no smaller), so the entry bar is max(k-th best, best - win): once the
        // running best has settled almost nothing passes it, whatever kTopK is.
        const float thr = fmaxf(tv[kTopK - 1], tv[0] - winq);
        if (mx > thr) {
          // compact slow path: bit mask of the qualifying columns, then one sorted insertion per set
          // bit.  Strict '>' keeps earlier keys ahead on ties.
          unsigned msk = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) msk |= (v[i] > thr) ? (1u << i) : 0u;
          while (msk) {
            const int i = __ffs(msk) - 1;
            msk &= msk - 1;
            // 16-way register select as a 4-level tree on the bits of i
            float s8[8], s4[4], s2[2];
#pragma unroll
            for (int j = 0; j < 8; ++j) s8[j] = (i & 1) ? v[2 * j + 1] : v[2 * j];
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = (i & 2) ? s8[2 * j + 1] : s8[2 * j];
#pragma unroll
            for (int j = 0; j < 2; ++j) s2[j] = (i & 4) ? s4[2 * j + 1] : s4[2 * j];
            float x = (i & 8) ? s2[1] : s2[0];
            if (x > fmaxf(tv[kTopK - 1], tv[0] - winq)) {
              const int ku = ku0 + (i & 7), kv = kv0 + 2 * r2 + (i >> 3);
              int xi = f * p.lk1 + uv_to_linear(p.k_orient, ku, kv, p.Wr);
#pragma unroll
              for (int s = 0; s < kTopK; ++s) {
                if (x > tv[s]) {
                  const float tf = tv[s]; tv[s] = x; x = tf;
                  const int tj = ti[s]; ti[s] = xi; xi = tj;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (pp + 1 == pe || ix.kt == p.KT - 1) {
        if (qlin >= 0) {
          const long long p0 = ((long long)ix.item * p.QT + ix.qt) * p.KT;
          const int slot = b - (int)(((p0 + 1) * (long long)p.G - 1) / p.P);
          float4* dv = reinterpret_cast<float4*>(p.cval + ((size_t)qlin * p.maxseg + slot) * kTopK);
          int4* di = reinterpret_cast<int4*>(p.cidx + ((size_t)qlin * p.maxseg + slot) * kTopK);
#pragma unroll
          for (int s4 = 0; s4 < kTopK / 4; ++s4) {
            dv[s4] = make_float4(tv[4 * s4], tv[4 * s4 + 1], tv[4 * s4 + 2], tv[4 * s4 + 3]);
            di[s4] = make_int4(ti[4 * s4], ti[4 * s4 + 1], ti[4 * s4 + 2], ti[4 * s4 + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (b == 0 && threadIdx.x == 0) *reinterpret_cast<long long*>(p.error_flag + 2) = clock64() - cyc0;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SPEI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SPEI_ERR_CUDA; }
    cached = (EncodeTiledFn)fn;
  }
  *out = cached;
  return SPEI_OK;
}

// 4-D map over a staged operand [img][16][Vpad][Upad*8] bf16; box = [1][groups][rows][80]
static int make_map(EncodeTiledFn enc, CUtensorMap* tm, void* base, int nimg, const OperandPlan& o, int box_rows, int box_groups) {
  const cuuint64_t dims[4] = {(cuuint64_t)o.Upad * 8, (cuuint64_t)o.Vpad, (cuuint64_t)kCG, (cuuint64_t)nimg};
  const cuuint64_t strides[3] = {(cuuint64_t)o.Upad * 16, (cuuint64_t)o.Vpad * o.Upad * 16, (cuuint64_t)kCG * o.Vpad * o.Upad * 16};
  const cuuint32_t box[4] = {kHaloU * 8, (cuuint32_t)box_rows, (cuuint32_t)box_groups, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SPEI_ERR_CUDA; }
  return SPEI_OK;
}

static thread_local float* tl_debug_acc = nullptr;  // set by spei_debug_relevance_tile for the next launch

int launch_relevance_tc(const Plan& p, float eps, char* ws, cudaStream_t st) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  CUtensorMap tmq, tmk;
  if ((rc = make_map(enc, &tmq, ws + p.off_qbf, p.n, p.q, kQTileV + 2, kCG))) return rc;
  if ((rc = make_map(enc, &tmk, ws + p.off_kbf, p.n * p.rf, p.k, p.k.tile_v + 2, kCGS))) return rc;

  TcParams t{};
  t.n = p.n; t.rf = p.rf; t.QT = p.QT; t.KT = p.KT; t.G = p.G; t.maxseg = p.maxseg; t.P = p.P;
  t.q_tu = p.q.tu; t.q_orient = p.q.orient; t.Uq = p.q.U; t.Vq = p.q.V; t.W = p.W; t.L = p.H * p.W;
  t.k_tu = p.k.tu; t.k_tiles_img = p.k.tiles(); t.k_orient = p.k.orient; t.Ny = p.k.tile_v; t.Wr = p.Wr; t.lk1 = p.Hr * p.Wr;
  t.UkT = p.k.tu * kTileU; t.VkT = p.k.tv * p.k.tile_v;
  const uint32_t ncols = (uint32_t)(kTileU * p.k.tile_v);
  // kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both,
  // N>>3 at bits 17-22, M>>4 at bits 24-28
  t.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((ncols >> 3) << 17) | ((128u >> 4) << 24);
  t.k_lbo = (uint32_t)(p.k.tile_v + 2) * kRowBytes;
  t.stage_bytes = kCGS * t.k_lbo;
  t.win = eps > 0.f ? eps * 1.02f : 0.f;  // a hair wider than the rescoring window: the two sides round differently
  t.rq = (const float*)(ws + p.off_rq);
  t.dq = (const float*)(ws + p.off_dq);
  t.dkmax = (const int*)(ws + p.off_dkmax);
  t.rkpad = (const float*)(ws + p.off_rkpad);
  t.cval = (float*)(ws + p.off_cval);
  t.cidx = (int32_t*)(ws + p.off_cidx);
  t.debug_acc = tl_debug_acc;
  t.error_flag = (int*)(ws + p.off_errflag);
  tl_debug_acc = nullptr;
  SPEI_CUDA(cudaMemsetAsync(t.error_flag, 0, sizeof(int), st));
  // per device, so set it on every launch (nn.DataParallel replicas call from several devices)
  auto kern = t.debug_acc ? relevance_tc_kernel<true> : relevance_tc_kernel<false>;
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  // ask for the largest shared-memory carve-out: the persistent CTA uses ~137 KB of it and the rest lets
  // kernels of other streams (rescoring / gather / fusion of the previous clip) co-reside on the SM
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  kern<<<p.G, kThreads, kSmemBytes, st>>>(tmq, tmk, t);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

void set_debug_acc(float* ptr) { tl_debug_acc = ptr; }
float* take_debug_acc() { float* r = tl_debug_acc; tl_debug_acc = nullptr; return r; }

}  // namespace spei
