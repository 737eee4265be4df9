// Kernel (b), tap-sharing variant: relevance bmm + row-wise max / argmax
// (/root/reference/model/SearchTransfer.py:33-34) with the 3x3 patch sum split between the tensor
// cores and the epilogue.
//
//   R[j, i] = rk[j] * rq[i] * sum_{tu in -1..1} D[j + tu*e_u, i + tu*e_u]
//   D[j, i] = sum_{tv in -1..1} sum_c K[c, j + tv*e_v] * Q[c, i + tv*e_v]          (e_u, e_v: unit steps in u, v)
//
// D is a dense contraction with K = 3 taps x 128 channels = 384 and runs on tcgen05 exactly like the
// dense kernel (relevance_tc.cu) but with a third of the MMAs.  The missing sum over the u taps adds
// three *neighbouring* accumulator entries: the product for (query i, key j, tap tu) is the tv-partial
// of (query i + tu, key j + tu), which the dense kernel recomputes for every one of them.  Tiles are
// 32 positions wide along u, MMA row m = v*32 + u, MMA column c = v'*32 + u', so the neighbours are
// (m-1, c-1) and (m+1, c+1): one TMEM lane up / down -- a warp shuffle inside the warp that owns the
// tile row -- and one accumulator column left / right, which the thread already holds.  The outer
// positions u = 0 and u = 31 of a tile are halo (consumed, not produced), so tiles advance by 30: 120
// of 128 rows and 30 of 32 columns per key row are productive, and the tensor-core work drops
// 3 * (30/32)^2 = 2.6x against the dense kernel for identical bf16 scores (same products, fp32 sums).
// What bounds the kernel (round-2 measurements): the epilogue, and it is dispatch bound.  Ablation
// ladder at 720p: MMA + TMA only 1.69 ms; + TMEM loads 1.71; + tap-sum adds, maxima, bookkeeping 1.85; + the 60 lane
// shuffles per key row 2.08; + the insertion path 2.27 (kTopK 16, one bound per row) -> 2.15 ms with kTopK 8 and the
// per-group row test below.  More epilogue warps are slower, the query operand from tensor memory is slower (2.71 ms),
// tcgen05.shift / lane-offset tcgen05.ld cannot replace the shuffles (tools/exp/); a scheduler issues one SHFL per 3-4 cycles
// (tools/exp/exp_shfl.cu), so the exchange alone is ~1 600 of a tile's 3 072 MMA cycles.  The chip also runs the kernel at its
// power cap (sw_power_cap, 1.67-1.75 GHz); tensor pipe 62 % active.
//
// Operand layout: the same channel-group-planar bf16 images as the dense kernel ([16][Vpad][Upad][8],
// zero border).  A tile row is exactly 32 positions x 16 B = 512 B, so the canonical no-swizzle
// K-major core matrices (8 rows x 16 B) of MMA rows 8g..8g+7 sit at g*128 B: SBO = 128 B, LBO = one
// channel-group plane of the tile, tap tv = start address + tv*512 B.  One TMA box per operand tile.
//
// Per CTA (persistent, one per SM, 12 warps):
//   warp 0    TMA producer (query tile 48 KB resident per query tile, ring of key stages)
//   warp 1    MMA issuer: per key tile 8 K16-steps x 3 taps = 24 tcgen05.mma (M=128, N=32*Ny)
//   warp 2    TMEM allocator (2 x 256 columns, double-buffered accumulators)
//   warps 4-11 epilogue, two groups of four (warp % 4 = the tile row v it owns): group h handles key rows
//             [h*ceil(Ny/2), ...) of every tile and keeps its own top-k list per query (the rescoring merges lists).
//             Per key row: 2 TMEM loads, 60 shuffles + 32 packed adds (tap sums), maxima over four 8-column groups and ONE
//             test of max_g(max(tap sum of group g) x max(reciprocal key norm of group g)) against the entry bar; only rows
//             that pass multiply out the groups that passed and go through the sorted insertion.  The first tile of a list is swept twice: the
//             first sweep only finds its best score, which seeds the entry bar.  (Alternating whole tiles between
//             the groups was tried and is slower: each accumulator is then held for a full 8-row epilogue.)
#include <cuda.h>
#include <cstdio>

#include "spei_common.cuh"
#include "tc_ptx.cuh"

namespace spei {

#ifndef SPEI_TCS_STAGES
#define SPEI_TCS_STAGES 6
#endif
constexpr int kSStages = SPEI_TCS_STAGES;   // key pipeline depth (4 stages = one key tile)
#ifndef SPEI_TCS_GROUPS
#define SPEI_TCS_GROUPS 2
#endif
constexpr int kSGroups = SPEI_TCS_GROUPS;     // epilogue warp groups (4 warps each); group g drains key rows [g*Ny/G, ...)
constexpr int kSThreads = (4 + 4 * kSGroups) * 32;
// register budget per thread after setmaxnreg: 4 non-epilogue warps shrink to 48, the epilogue warps share the rest
constexpr int kSEpiRegs = ((65536 - 4 * 32 * 48) / (4 * kSGroups * 32)) / 8 * 8 > 232 ? 232 : ((65536 - 4 * 32 * 48) / (4 * kSGroups * 32)) / 8 * 8;
constexpr uint32_t kSRowBytes = kSBoxU * 16;                                  // 512 B
constexpr uint32_t kSSBO = 128;                                               // 8 positions x 16 B
constexpr uint32_t kSQRows = kSQTileV + 2;                                    // 6
constexpr uint32_t kSQLBO = kSQRows * kSRowBytes;                             // 3072
constexpr uint32_t kSQTileBytes = kCG * kSQLBO;                               // 49152
constexpr uint32_t kSStageBytesMax = kCGS * (kSMaxNy + 2) * kSRowBytes;       // 20480
constexpr uint32_t kSStagesPerTile = kCG / kCGS;                              // 4
constexpr uint32_t kSNumBars = 2 * kSStages + 6;
constexpr uint32_t kSRkOffset = kSQTileBytes + kSStages * kSStageBytesMax + kSNumBars * 8 + 16;  // per epilogue warp: 128 floats + row maxima
constexpr uint32_t kSRkWarpFloats = 144;  // [<=4 rows][32] reciprocal key norms + [<=4 rows][4] maxima over 8-column groups
constexpr uint32_t kSSmemBytes = kSRkOffset + 4 * kSGroups * kSRkWarpFloats * 4;
static_assert(kSRkOffset % 16 == 0, "key-norm staging must be float4 aligned");
constexpr uint32_t kSTmemCols = 512;
constexpr uint32_t kSAccCols = 256;

struct TcsParams {
  int n, rf, QT, KT, G, maxseg;   // QT = query work slots per item (tile pairs with kPair), G = workers (CTAs or CTA pairs)
  int QTreal;                     // query tiles per item
  long long P;
  int q_tu, q_orient, Uq, Vq, W, L;
  int k_tu, k_tvn, k_tiles_img, k_orient, Ny, Wr, lk1, UkP, VkT;
  uint32_t idesc, stage_bytes, k_lbo;
  float win;
  const float* rq;
  const float* rkpad;  // [img][VkT][UkP], u border included, NaN outside the image
  const float* dq;     // per-query relative bf16 residual norm (stage_norm.cu)
  const int* dkmax;    // per-item maximum over the keys, float bits
  float* cval;
  int32_t* cidx;
  float* debug_acc;
  int* error_flag;
};

// (reference frame, tile row, tile column) of key tile kt, kept incrementally: the per-tile integer divisions were
// ~15 % of the epilogue warps' instructions (ncu source view, round 1)
struct KeyTile { int f, tv, tu; };
__device__ __forceinline__ KeyTile key_tile_decode(int kt, int tiles_img, int k_tu) {
  KeyTile t;
  t.f = kt / tiles_img;
  const int kti = kt - t.f * tiles_img;
  t.tv = kti / k_tu;
  t.tu = kti - t.tv * k_tu;
  return t;
}
__device__ __forceinline__ KeyTile key_tile_next(KeyTile t, int k_tu, int k_tvn, int rf) {  // mirrors next_pair's kt + 1 (wraps to 0)
  if (++t.tu == k_tu) { t.tu = 0; if (++t.tv == k_tvn) { t.tv = 0; if (++t.f == rf) t.f = 0; } }
  return t;
}

// kPair: the kernel runs as clusters of two CTAs on one TPC (cta_group::2).  Rank 0 issues M = 256 MMAs for both: each CTA
// holds its own query tile (128 accumulator rows in its own tensor memory) and HALF of the key rows of a stage, so a key tile
// is fetched and read from shared memory once per TWO query tiles: per MMA an SM reads 4 KB of A + 4 KB of B instead of
// 4 + 8 KB.  The barriers that gate the MMA issue (key stage full, query tile full, accumulator drained) live in rank 0's
// shared memory: both producers' TMA loads count their bytes there and both CTAs' epilogue warps arrive there; every
// tcgen05.commit is multicast to the barrier of the same name in both CTAs (stage free, query tile free, accumulator full).
template <bool kDebug, bool kPair>
__global__ void __launch_bounds__(kSThreads, 1)
relevance_tcs_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk, const TcsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + kSQTileBytes;
  const uint32_t bars = sK + kSStages * kSStageBytesMax;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kSStages;
  const uint32_t bar_qfull = bars + 16 * kSStages, bar_qfree = bar_qfull + 8;
  const uint32_t bar_tfull = bar_qfull + 16, bar_tempty = bar_qfull + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSQTileBytes + kSStages * kSStageBytesMax + kSNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;     // 0 = the CTA that issues the MMAs
  const int b = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const long long pb = (long long)b * p.P / p.G, pe = (long long)(b + 1) * p.P / p.G;
  const long long cyc0 = clock64();   // CTA 0 publishes its clock64 span: the SM clock this kernel really ran at (bench.py)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kSStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_qfull, 1); mbar_init(bar_qfree, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4 * kSGroups * (kPair ? 2 : 1)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmq) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmk) : "memory");
  }
  if (warp == 2) {
    if (kPair) {   // one warp of EACH CTA of the pair, same shared-memory slot in both
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kSTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kSTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // register re-balancing (one warpgroup = 4 consecutive warps): the producer / issuer / allocator warps need few
  // registers, the epilogue warps hold a top-k list, two TMEM loads and a row of tap sums each
  // (each setmaxnreg sits at the top of the branch it governs, so the compiler's budget for that branch is unambiguous)
  if (warp < 4) {
    if (kSGroups >= 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int qloaded = 0;
      PairIdx ix = decode_pair(pb, p.QT, p.KT);
      KeyTile kc = key_tile_decode(ix.kt, p.k_tiles_img, p.k_tu);
      // kPair: each CTA loads its own query tile and its half of the key rows; the bytes of both are counted on rank 0's barrier
      const uint32_t qfull_dst = kPair ? mapa_u32(bar_qfull, 0) : bar_qfull;
      const int krow0 = kPair ? (int)rank * (p.Ny >> 1) : 0;
      for (long long pp = pb; pp < pe; ++pp, ix = next_pair(ix, p.QT, p.KT), kc = key_tile_next(kc, p.k_tu, p.k_tvn, p.rf)) {
        if (pp == pb || ix.kt == 0) {
          if (qloaded > 0) mbar_wait(bar_qfree, (uint32_t)((qloaded - 1) & 1), p.error_flag);
          // (an odd tile count leaves rank 1 without a tile in the last pair: it re-reads the last one and drops the results)
          const int qt = kPair ? min(2 * ix.qt + (int)rank, p.QTreal - 1) : ix.qt;
          const int qtv = qt / p.q_tu, qtu = qt - qtv * p.q_tu;
          if (rank == 0) mbar_arrive_expect_tx(bar_qfull, kSQTileBytes * (kPair ? 2u : 1u));
          // staged coordinates carry a 1-position border: interior position (u, v) lives at (u+1, v+1); the box
          // starts one position before the tile's first interior position in both directions
          if (kPair) tma_load_4d_pair(sQ, &tmq, qfull_dst, qtu * kSTileU * 8, qtv * kSQTileV, 0, ix.item);
          else tma_load_4d(sQ, &tmq, bar_qfull, qtu * kSTileU * 8, qtv * kSQTileV, 0, ix.item);
          ++qloaded;
        }
        const int f = kc.f, ktv = kc.tv, ktu = kc.tu;
        for (uint32_t s4 = 0; s4 < kSStagesPerTile; ++s4) {
          mbar_wait_parked(bar_empty + 8 * stage, phase ^ 1, p.error_flag);
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, p.stage_bytes * (kPair ? 2u : 1u));
          if (kPair)
            tma_load_4d_pair(sK + stage * kSStageBytesMax, &tmk, mapa_u32(bar_full + 8 * stage, 0), ktu * kSTileU * 8, ktv * p.Ny + krow0,
                             (int)(s4 * kCGS), ix.item * p.rf + f);
          else
            tma_load_4d(sK + stage * kSStageBytesMax, &tmk, bar_full + 8 * stage, ktu * kSTileU * 8, ktv * p.Ny, (int)(s4 * kCGS),
                        ix.item * p.rf + f);
          if (++stage == kSStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0;
      int qused = 0;
      uint32_t tile_i = 0;
      PairIdx ix = decode_pair(pb, p.QT, p.KT);
#ifdef SPEI_TCS_PROF
      long long pt0 = clock64(), p_tempty = 0, p_full = 0;
#define PROF_T(acc_, stmt) { const long long c0_ = clock64(); stmt; acc_ += clock64() - c0_; }
#else
#define PROF_T(acc_, stmt) { stmt; }
#endif
      for (long long pp = pb; pp < pe; ++pp, ++tile_i, ix = next_pair(ix, p.QT, p.KT)) {
        if (pp == pb || ix.kt == 0) {
          mbar_wait(bar_qfull, (uint32_t)(qused & 1), p.error_flag);
          ++qused;
        }
        const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
        PROF_T(p_tempty, mbar_wait_parked(bar_tempty + 8 * acc, (use & 1u) ^ 1u, p.error_flag));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kSAccCols;
        for (uint32_t s4 = 0; s4 < kSStagesPerTile; ++s4) {
          PROF_T(p_full, mbar_wait_parked(bar_full + 8 * stage, phase, p.error_flag));
          tc_fence_after();
          const uint32_t kbase = sK + stage * kSStageBytesMax;
#pragma unroll
          for (uint32_t cgp = 0; cgp < kCGS / 2; ++cgp) {
            const uint32_t qa = sQ + (s4 * kCGS + cgp * 2) * kSQLBO;
            const uint32_t ka = kbase + (cgp * 2) * p.k_lbo;
#pragma unroll
            for (uint32_t tap = 0; tap < 3; ++tap) {
              const uint64_t adesc = umma_desc_kmajor(qa + tap * kSRowBytes, kSQLBO, kSSBO);
              const uint64_t bdesc = umma_desc_kmajor(ka + tap * kSRowBytes, p.k_lbo, kSSBO);
              if (kPair) tc_mma_bf16_pair(d_tmem, adesc, bdesc, p.idesc, (s4 | cgp | tap) != 0u);
              else tc_mma_bf16(d_tmem, adesc, bdesc, p.idesc, (s4 | cgp | tap) != 0u);
            }
          }
          if (kPair) tc_commit_pair(bar_empty + 8 * stage); else tc_commit(bar_empty + 8 * stage);
          if (++stage == kSStages) { stage = 0; phase ^= 1; }
        }
        if (kPair) tc_commit_pair(bar_tfull + 8 * acc); else tc_commit(bar_tfull + 8 * acc);
        if (ix.kt == p.KT - 1 && pp + 1 < pe) { if (kPair) tc_commit_pair(bar_qfree); else tc_commit(bar_qfree); }
      }
#ifdef SPEI_TCS_PROF
      if (b == 3) printf("mma: total %lld wait_tempty %lld wait_full %lld tiles %lld\n", clock64() - pt0, p_tempty, p_full, pe - pb);
#endif
    }
  }
  } else {
    // ======================================= epilogue =======================================
    if (kSGroups >= 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSEpiRegs));
    const int ew = (warp - 4) & 3;        // TMEM lane quarter this warp may read (= warp % 4) = tile row v
    const int half = (warp - 4) >> 2;     // epilogue group: which share of a tile's key rows
    const int m = ew * 32 + lane;         // MMA row: query (qu = lane, qv = ew) of the tile
    const int qu = lane, qv = ew;
    // rows split as evenly as possible, the larger shares first: Ny = 8, 3 groups -> 3 / 3 / 2
    const int r_base = p.Ny / kSGroups, r_rem = p.Ny % kSGroups;
    const int r_lo = half * r_base + (half < r_rem ? half : r_rem), r_hi = r_lo + r_base + (half < r_rem ? 1 : 0);
    float tv[kTopK];
    int ti[kTopK];
    uint32_t tile_i = 0;
    long long qlin = -1;
    float winq = 0.f, floor0 = -INFINITY;
    float* rk_s = reinterpret_cast<float*>(smem + kSRkOffset) + (warp - 4) * kSRkWarpFloats;  // this warp's key norms: [<=4 rows][32]
    // key-norm prefetch: lane l owns float2 #l and #(l+32) of the warp's [<=4 rows][32] reciprocal norms.  32-bit element
    // offsets (the launcher checks the array fits) and lane-invariant parts hoisted: the 64-bit address chain of the
    // round-1 form was ~45 instructions per tile per warp
    const int pf_row = lane >> 4;                                    // rows pf_row and pf_row + 2
    const bool pf_ok0 = r_lo + pf_row < r_hi, pf_ok1 = r_lo + pf_row + 2 < r_hi;
    const uint32_t pf_off0 = (uint32_t)(pf_row * p.UkP + (lane & 15) * 2), pf_off1 = pf_off0 + 2u * (uint32_t)p.UkP;
    auto rk_prefetch = [&](int item, const KeyTile kt, float2 (&pre)[2]) {
      const uint32_t off = (uint32_t)(((item * p.rf + kt.f) * p.VkT + kt.tv * p.Ny + r_lo) * p.UkP + kt.tu * kSTileU);
      const float* base = p.rkpad + off;
      pre[0] = pf_ok0 ? __ldg(reinterpret_cast<const float2*>(base + pf_off0)) : make_float2(0.f, 0.f);
      pre[1] = pf_ok1 ? __ldg(reinterpret_cast<const float2*>(base + pf_off1)) : make_float2(0.f, 0.f);
    };
    // tap sums of one key row from its 32 accumulator columns: (m-1, col-1) + (m, col) + (m+1, col+1).  Columns 0 and 31
    // are halo: s[0] / s[31] are never looked at and simply alias the raw entries.  Packed f32x2 adds for columns 2..29;
    // columns 1 and 30 take scalar adds (same rounding order) so that no zero has to be moved into a register pair.
    auto tap_sums = [&](const uint32_t (&xr)[32], float (&s)[32]) {
      float x[32], up[32], dn[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(xr[i]);
#pragma unroll
      for (int i = 1; i < 31; ++i) {
#ifdef SPEI_TCS_NOSHFL   // timing experiment only (wrong sums): what the lane exchange costs
        up[i] = x[i - 1];
        dn[i] = x[i + 1];
#else
        up[i] = __shfl_up_sync(0xffffffffu, x[i - 1], 1);
        dn[i] = __shfl_down_sync(0xffffffffu, x[i + 1], 1);
#endif
      }
      s[0] = x[0];
      s[31] = x[31];
      s[1] = __fadd_rn(__fadd_rn(up[1], dn[1]), x[1]);
      s[30] = __fadd_rn(__fadd_rn(up[30], dn[30]), x[30]);
#pragma unroll
      for (int i = 2; i < 30; i += 2) {
        float t0, t1;
        fadd2(t0, t1, up[i], up[i + 1], dn[i], dn[i + 1]);
        fadd2(s[i], s[i + 1], t0, t1, x[i], x[i + 1]);
      }
    };
    float2 pre[2];
    PairIdx ix = decode_pair(pb, p.QT, p.KT);
    KeyTile kc = key_tile_decode(ix.kt, p.k_tiles_img, p.k_tu);
    if (pb < pe) rk_prefetch(ix.item, kc, pre);
#ifdef SPEI_TCS_PROF
    long long et0 = clock64(), e_tfull = 0, e_rows = 0, e_rc0 = 0;
#endif
    for (long long pp = pb; pp < pe; ++pp, ++tile_i) {
      if (pp == pb || ix.kt == 0) {
#pragma unroll
        for (int s = 0; s < kTopK; ++s) { tv[s] = -INFINITY; ti[s] = -1; }
        const int qt = kPair ? 2 * ix.qt + (int)rank : ix.qt;   // (kPair, odd tile count: the last pair's rank 1 has no tile)
        const int qtv = qt / p.q_tu, qtu = qt - qtv * p.q_tu;
        const int u = qtu * kSTileU + qu - 1, v = qtv * kSQTileV + qv;
        qlin = (qt < p.QTreal && qu >= 1 && qu <= kSTileU && u < p.Uq && v < p.Vq) ? (long long)ix.item * p.L + uv_to_linear(p.q_orient, u, v, p.W)
                                                                                  : -1;
        winq = 0.f;
        if (qlin >= 0) {
          // fixed window (eps > 0) or the certified one (spei_common.cuh: certified_window)
          const float wn = p.win > 0.f ? p.win
                                       : 1.02f * certified_window(certified_delta(__ldg(p.dq + qlin), __int_as_float(__ldg(p.dkmax + ix.item))));
          winq = wn / __ldg(p.rq + qlin);
        }
        floor0 = -INFINITY;
      }
      const bool fresh = pp == pb || ix.kt == 0;   // first tile of a (query tile, key segment)
      const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
      const int f = kc.f;
      const int ku0 = kc.tu * kSTileU - 1, kv0 = kc.tv * p.Ny;
      // largest reciprocal norm of every key row (NaN = keys outside the image are ignored by fmaxf): the row test
      // below bounds all 30 scores of a row by (largest tap sum) x (largest reciprocal norm)
      // (four lanes hold the 8 columns of a group: rows lane >> 4 (+2), group (lane & 15) >> 2)
      float rmax0 = fmaxf(pre[0].x, pre[0].y), rmax1 = fmaxf(pre[1].x, pre[1].y);
#pragma unroll
      for (int o = 2; o >= 1; o >>= 1) {
        rmax0 = fmaxf(rmax0, __shfl_xor_sync(0xffffffffu, rmax0, o));
        rmax1 = fmaxf(rmax1, __shfl_xor_sync(0xffffffffu, rmax1, o));
      }
      __syncwarp();
      reinterpret_cast<float2*>(rk_s)[lane] = pre[0];
      reinterpret_cast<float2*>(rk_s)[lane + 32] = pre[1];
      if ((lane & 3) == 0) { rk_s[128 + (lane >> 2)] = rmax0; rk_s[136 + (lane >> 2)] = rmax1; }
      __syncwarp();
      const PairIdx nx = next_pair(ix, p.QT, p.KT);
      const KeyTile kn = key_tile_next(kc, p.k_tu, p.k_tvn, p.rf);
      if (pp + 1 < pe) rk_prefetch(nx.item, kn, pre);
      PROF_T(e_tfull, mbar_wait_parked(bar_tfull + 8 * acc, use & 1u, p.error_flag));
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kSAccCols + ((uint32_t)(ew * 32) << 16);
#ifdef SPEI_TCS_PROF
      e_rc0 = clock64();
#endif

      // One key row (32 accumulator columns = one TMEM load) per iteration.  Row addresses (TMEM row, group maxima and
      // key norms of the row in shared memory) are running values behind an opaque copy: re-derived from the special
      // registers in every iteration they were ~15 of the ~150 instructions of a row.
      uint32_t xr[32];
      const uint32_t trow0 = opaque_u32(taddr + (uint32_t)r_lo * 32u);
      if (r_lo < r_hi) tc_ld32(trow0, xr);
      if (fresh && r_lo < r_hi) {
        // First tile of a list: one extra sweep over this group's rows finds their best score, and the entry bar
        // starts at (that - window) instead of -inf.  Without it the first rows push every key through the sorted
        // insertion (the lists restart for every query tile: that transient was ~15 % of the epilogue instructions).
        float best0 = -INFINITY;
#pragma unroll 1
        for (int r = r_lo; r < r_hi; ++r) {
          tc_wait_ld();
          float s[32];
          tap_sums(xr, s);
          const int rn = r + 1 < r_hi ? r + 1 : r_lo;   // the last refill re-reads the first row for the main sweep
          tc_ld32(taddr + rn * 32, xr);
          const float* rkr = rk_s + (r - r_lo) * 32;
          float vmax = -INFINITY;
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 t4 = reinterpret_cast<const float4*>(rkr)[i4];
            float v0, v1, v2, v3;
            fmul2(v0, v1, s[4 * i4], s[4 * i4 + 1], t4.x, t4.y);
            fmul2(v2, v3, s[4 * i4 + 2], s[4 * i4 + 3], t4.z, t4.w);
            if (i4 == 0) v0 = -INFINITY;     // halo column 0
            if (i4 == 7) v3 = -INFINITY;     // halo column 31
            vmax = fmaxf(vmax, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));   // fmaxf drops NaN (keys outside the image)
          }
          best0 = fmaxf(best0, vmax);
        }
        floor0 = best0 - winq;
      }
#ifdef SPEI_TCS_NOEPI     // timing experiment only: the MMA / TMA pipeline without any epilogue work
      tc_wait_ld();
      if (false)
#endif
      uint32_t trow = trow0;                                              // TMEM address of row r
      uint32_t srow = opaque_u32(smem_u32(rk_s));                         // shared-memory address of row r's key norms
      uint32_t sgm = srow + 128u * 4u;                                    // ... and of its four group maxima
#pragma unroll 1
      for (int r = r_lo; r < r_hi; ++r, trow += 32u, srow += 32u * 4u, sgm += 4u * 4u) {
        tc_wait_ld();
#ifdef SPEI_TCS_LDONLY   // timing experiment only: TMEM loads without any arithmetic
        if (r + 1 < r_hi) tc_ld32(trow + 32u, xr);
        if (xr[0] == 0x7fc12345u && xr[19] == 0x7fc12345u) tv[0] = 1.f;
        continue;
#endif
        if (kDebug && p.debug_acc && pp == 0 && rank == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) p.debug_acc[(size_t)m * kSAccCols + r * 32 + i] = __uint_as_float(xr[i]);
        }
        float s[32];
        tap_sums(xr, s);
#ifndef SPEI_TCS_NOLD      // (timing experiment: arithmetic without the TMEM reloads)
        if (r + 1 < r_hi)
#else
        if (false)
#endif
        {  // xr[] is consumed: refill while the scores are examined
          tc_ld32(trow + 32u, xr);
        }
        // Row test on the un-normalised tap sums: score[i] = s[i] * rk[i] <= max(s) * max(rk) when max(s) > 0 and
        // <= 0 otherwise (rk > 0), so a row whose bound does not beat the entry bar holds no candidate and its key
        // norms are never read (one broadcast LDS + 16 FMNMX instead of 8 LDS.128 + 16 FMUL2 + 16 FMNMX per row).
        // ... per group of 8 columns: half as many rows take the slow path as with one bound per row (round-2 A/B)
        float gb[4], bound;
        {
          const float g0 = fmax3(fmax3(s[1], s[2], s[3]), fmax3(s[4], s[5], s[6]), s[7]);
          const float g1 = fmax3(fmax3(s[8], s[9], s[10]), fmax3(s[11], s[12], s[13]), fmaxf(s[14], s[15]));
          const float g2 = fmax3(fmax3(s[16], s[17], s[18]), fmax3(s[19], s[20], s[21]), fmaxf(s[22], s[23]));
          const float g3 = fmax3(fmax3(s[24], s[25], s[26]), fmax3(s[27], s[28], s[29]), s[30]);
          const float4 gm = lds_f4(sgm);   // broadcast read
          // a non-positive maximum bounds its group's scores by 0; a NaN product (no key of the group inside the image)
          // fails every comparison and is dropped by fmaxf
          gb[0] = fmaxf(g0, 0.f) * gm.x; gb[1] = fmaxf(g1, 0.f) * gm.y; gb[2] = fmaxf(g2, 0.f) * gm.z; gb[3] = fmaxf(g3, 0.f) * gm.w;
          bound = fmaxf(fmaxf(gb[0], gb[1]), fmaxf(gb[2], gb[3]));
        }
        const float thr = fmax3(tv[kTopK - 1], tv[0] - winq, floor0);
#ifdef SPEI_TCS_NOSLOW    // timing experiment only (wrong results): tap sums + row maximum, never the insertion path
        tv[0] = fmaxf(tv[0], bound);
        if (false)
#else
        if (qlin >= 0 && bound > thr)
#endif
        {
          // compact slow path: only the 8-column groups whose bound passed are multiplied out (products s * rk, NaN for
          // keys outside the image) and compared -> bit mask of the qualifying columns (halo columns 0 / 31 excluded), then
          // one sorted insertion per set bit.  Strict '>' keeps earlier keys ahead on ties.
          const float* rkr = rk_s + (r - r_lo) * 32;
          unsigned msk = 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (gb[g] > thr) {
              const float4 ta = lds_f4(srow + 32u * g), tb = lds_f4(srow + 32u * g + 16u);
              const float rk8[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int i = 8 * g + j;
                if (i >= 1 && i <= 30) msk |= (s[i] * rk8[j] > thr) ? (1u << i) : 0u;
              }
            }
          }
          while (msk) {
            const int i = __ffs(msk) - 1;
            msk &= msk - 1;
            // 32-way register select as a 5-level tree on the bits of i
            float s16[16], s8[8], s4[4], s2[2];
#pragma unroll
            for (int j = 0; j < 16; ++j) s16[j] = (i & 1) ? s[2 * j + 1] : s[2 * j];
#pragma unroll
            for (int j = 0; j < 8; ++j) s8[j] = (i & 2) ? s16[2 * j + 1] : s16[2 * j];
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = (i & 4) ? s8[2 * j + 1] : s8[2 * j];
#pragma unroll
            for (int j = 0; j < 2; ++j) s2[j] = (i & 8) ? s4[2 * j + 1] : s4[2 * j];
            float x = ((i & 16) ? s2[1] : s2[0]) * rkr[i];   // the same fp32 product the mask compared
            if (x > fmax3(tv[kTopK - 1], tv[0] - winq, floor0)) {
              int xi = f * p.lk1 + uv_to_linear(p.k_orient, ku0 + i, kv0 + r, p.Wr);
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
#ifdef SPEI_TCS_PROF
      e_rows += clock64() - e_rc0;
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(mapa_u32(bar_tempty + 8 * acc, 0));   // the issuing CTA's barrier counts both CTAs' warps
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      if (pp + 1 == pe || ix.kt == p.KT - 1) {
        if (qlin >= 0) {
          const long long p0 = ((long long)ix.item * p.QT + ix.qt) * p.KT;
          const int slot = (b - (int)(((p0 + 1) * (long long)p.G - 1) / p.P)) * kSGroups + half;
          float4* dv = reinterpret_cast<float4*>(p.cval + ((size_t)qlin * p.maxseg * kSGroups + slot) * kTopK);
          int4* di = reinterpret_cast<int4*>(p.cidx + ((size_t)qlin * p.maxseg * kSGroups + slot) * kTopK);
#pragma unroll
          for (int s4 = 0; s4 < kTopK / 4; ++s4) {
            dv[s4] = make_float4(tv[4 * s4], tv[4 * s4 + 1], tv[4 * s4 + 2], tv[4 * s4 + 3]);
            di[s4] = make_int4(ti[4 * s4], ti[4 * s4 + 1], ti[4 * s4 + 2], ti[4 * s4 + 3]);
          }
        }
      }
      ix = nx;
      kc = kn;
    }
#ifdef SPEI_TCS_PROF
    if (b == 3 && lane == 0) printf("epi warp %d: total %lld wait_tfull %lld rows %lld\n", warp, clock64() - et0, e_tfull, e_rows);
#endif
  }

  tc_fence_before();
  if (kPair) cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while its peer's MMAs / commits can still touch it
  else __syncthreads();
  if (b == 0 && rank == 0 && threadIdx.x == 0) *reinterpret_cast<long long*>(p.error_flag + 2) = clock64() - cyc0;
  if (warp == 2) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kSTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kSTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 4-D map over a staged operand [img][16][Vpad][Upad*8] bf16; box = [1][groups][rows][16 positions x 8 channels]
static int make_map_s(EncodeTiledFn enc, CUtensorMap* tm, void* base, int nimg, const OperandPlan& o, int box_rows, int box_groups) {
  const cuuint64_t dims[4] = {(cuuint64_t)o.Upad * 8, (cuuint64_t)o.Vpad, (cuuint64_t)kCG, (cuuint64_t)nimg};
  const cuuint64_t strides[3] = {(cuuint64_t)o.Upad * 16, (cuuint64_t)o.Vpad * o.Upad * 16, (cuuint64_t)kCG * o.Vpad * o.Upad * 16};
  const cuuint32_t box[4] = {kSBoxU * 8, (cuuint32_t)box_rows, (cuuint32_t)box_groups, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SPEI_ERR_CUDA; }
  return SPEI_OK;
}

int tcs_epilogue_groups() { return kSGroups; }
// -DSPEI_TCS_PAIR=1 builds the kernel for CTA pairs.  Measured at 720p (round 2, parity suite green in both builds): 2.25 ms
// against 2.21 ms on the same box -- a third fewer operand reads from shared memory and 40 % less key traffic from L2 do not
// buy anything because neither bounds the kernel (the epilogue does), and the pair's two epilogues now have to finish
// before either accumulator is free (max of two tile times instead of one).  Default: single CTAs.
#ifndef SPEI_TCS_PAIR
#define SPEI_TCS_PAIR 0
#endif
bool tcs_cta_pairs() { return SPEI_TCS_PAIR != 0; }

template <bool kDebug, bool kPair>
static int launch_tcs_t(const Plan& p, const CUtensorMap& tmq, const CUtensorMap& tmk, const TcsParams& t, cudaStream_t st) {
  auto kern = relevance_tcs_kernel<kDebug, kPair>;
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmemBytes));
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(p.G * (kPair ? 2 : 1)));
  cfg.blockDim = dim3(kSThreads);
  cfg.dynamicSmemBytes = kSSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPEI_CUDA(cudaLaunchKernelEx(&cfg, kern, tmq, tmk, t));
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_relevance_tcs(const Plan& p, float eps, char* ws, cudaStream_t st) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const bool pair = p.pair != 0;
  if (pair && (p.k.tile_v & 1)) { set_error("relevance_tcs: CTA pairs need an even key tile height"); return SPEI_ERR_ARG; }
  const int rows_cta = pair ? p.k.tile_v / 2 : p.k.tile_v;   // key rows of a tile one CTA stages
  CUtensorMap tmq, tmk;
  if ((rc = make_map_s(enc, &tmq, ws + p.off_qbf, p.n, p.q, kSQTileV + 2, kCG))) return rc;
  if ((rc = make_map_s(enc, &tmk, ws + p.off_kbf, p.n * p.rf, p.k, rows_cta + 2, kCGS))) return rc;

  TcsParams t{};
  t.n = p.n; t.rf = p.rf; t.QT = p.QTs; t.QTreal = p.QT; t.KT = p.KT; t.G = p.G; t.maxseg = p.maxseg; t.P = p.P;
  t.q_tu = p.q.tu; t.q_orient = p.q.orient; t.Uq = p.q.U; t.Vq = p.q.V; t.W = p.W; t.L = p.H * p.W;
  t.k_tu = p.k.tu; t.k_tvn = p.k.tv; t.k_tiles_img = p.k.tiles(); t.k_orient = p.k.orient; t.Ny = p.k.tile_v; t.Wr = p.Wr; t.lk1 = p.Hr * p.Wr;
  t.UkP = p.k.Upad; t.VkT = p.k.tv * p.k.tile_v;
  if ((long long)p.n * p.rf * t.VkT * t.UkP >= (1ll << 31)) { set_error("relevance_tcs: key-norm array exceeds 32-bit offsets"); return SPEI_ERR_ARG; }
  const uint32_t ncols = (uint32_t)(kSBoxU * p.k.tile_v);
  // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at bits 17-22, M>>4 at bits 24-28 (M = 256 across a CTA pair)
  t.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((ncols >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24);
  t.k_lbo = (uint32_t)(rows_cta + 2) * kSRowBytes;
  t.stage_bytes = kCGS * t.k_lbo;
  t.win = eps > 0.f ? eps * 1.02f : 0.f;   // <= 0: certified per-query window
  t.rq = (const float*)(ws + p.off_rq);
  t.dq = (const float*)(ws + p.off_dq);
  t.dkmax = (const int*)(ws + p.off_dkmax);
  t.rkpad = (const float*)(ws + p.off_rkpad);
  t.cval = (float*)(ws + p.off_cval);
  t.cidx = (int32_t*)(ws + p.off_cidx);
  t.debug_acc = take_debug_acc();
  t.error_flag = (int*)(ws + p.off_errflag);
  SPEI_CUDA(cudaMemsetAsync(t.error_flag, 0, sizeof(int), st));
#if SPEI_TCS_PAIR
  if (pair) return t.debug_acc ? launch_tcs_t<true, true>(p, tmq, tmk, t, st) : launch_tcs_t<false, true>(p, tmq, tmk, t, st);
#endif
  return t.debug_acc ? launch_tcs_t<true, false>(p, tmq, tmk, t, st) : launch_tcs_t<false, false>(p, tmq, tmk, t, st);
}

}  // namespace spei
