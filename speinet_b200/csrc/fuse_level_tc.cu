// Kernel (d) on the 5th-generation tensor cores: one fusion line of SPEINet._decode
// (/root/reference/model/speinet.py:93-94, 96-97, 108-109)
//
//     out = dec + (W . cat(dec, T) + b) * bicubic_up(S, scale)
//
// as ONE pass over HBM (read dec, read T, write out = 3 x tensor size, the algorithmic minimum).
//
// Two tcgen05 kernels live in this file (DESIGN.md section 4(d) has the measurements that led from one to the next):
//   fuse_level_ts_kernel   DEFAULT.  Activations = A operand from tensor memory, resident weight tiles, TMA-fed residual
//                          and TMA-store epilogue.                                    (second half of the file)
//   fuse_level_tc_kernel   LDG-fed, no alignment requirement: the path for planes whose byte stride is not a multiple of
//                          16 (TMA needs that) or whose base pointers are not 16-byte aligned, described next.
//
// The 1x1 convolution is a skinny GEMM  D[pixel, o] = sum_k X[k, pixel] * W[o, k]  (K = 2C) that must keep
// fp32 accuracy (1e-4 bar), so it runs as 3xTF32 on tcgen05: x = hi + lo with hi = x truncated to TF32 and
// lo = x - hi (exact), D ~= Xlo.Whi + Xhi.Wlo + Xhi.Whi with fp32 accumulation in TMEM (dropped term
// ~2^-22).  M = 128 pixels (one TMEM lane per pixel), N = C output channels, kind::tf32, K = 8 per MMA.
//
// The split needs a register pass over the activations anyway, so the producer warps do the layout
// change at the same time: coalesced loads along pixels (NCHW planes), split, 16-byte stores straight into
// the canonical no-swizzle K-major UMMA layout (core matrix = 8 rows x 16 B = 8 pixels x 4 channels,
// rows contiguous: SBO = 128 B, LBO = 128 rows x 16 B).  No TMA, no alignment requirement on the planes.
//
// CTA = 160 threads: warps 0-3 produce (thread = pixel; it also owns one weight row) and later run the
// epilogue (thread = TMEM lane = pixel: every channel store is a coalesced 128-byte row segment), warp 4
// issues the MMAs.  Three 16-channel stages, loads prefetched one 32-channel set ahead in registers, two to
// three CTAs per SM.
#include <cstdio>
#include <cuda_bf16.h>

#include "spei_common.cuh"
#include "tc_ptx.cuh"

#ifdef SPEI_FUSE_PROF   // in-kernel clock64 accounting of who waits for whom (experiments only)
#define FPROF_T(acc_, stmt) { const long long c0_ = clock64(); stmt; acc_ += clock64() - c0_; }
#else
#define FPROF_T(acc_, stmt) { stmt; }
#endif

namespace spei {

constexpr int kFM = 128;       // pixels per CTA = MMA M
constexpr int kFKC = 16;       // channels per pipeline stage = 2 MMA K-steps of 8
constexpr int kFStages = 3;
constexpr int kFThreads = 160;

__device__ __forceinline__ float ft_cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float ft_cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// F.interpolate(S, scale_factor, mode='bicubic'): align_corners=False, A=-0.75, border-clamped taps
// (torch/include/ATen/native/UpSample.h:289-300,400-423)
__device__ __forceinline__ float ft_bicubic(const float* __restrict__ S, int h, int w, int oy, int ox, float rscale) {
  const float ry = rscale * (oy + 0.5f) - 0.5f, rx = rscale * (ox + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  const int iy = (int)fy, ix = (int)fx;
  const float ty = ry - fy, tx = rx - fx;
  const float cx[4] = {ft_cubic2(tx + 1.f), ft_cubic1(tx), ft_cubic1(1.f - tx), ft_cubic2(2.f - tx)};
  const float cy[4] = {ft_cubic2(ty + 1.f), ft_cubic1(ty), ft_cubic1(1.f - ty), ft_cubic2(2.f - ty)};
  float rows[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int yy = min(max(iy - 1 + i, 0), h - 1);
    const float* r = S + (size_t)yy * w;
    const float v0 = __ldg(r + min(max(ix - 1, 0), w - 1)), v1 = __ldg(r + min(max(ix, 0), w - 1));
    const float v2 = __ldg(r + min(max(ix + 1, 0), w - 1)), v3 = __ldg(r + min(max(ix + 2, 0), w - 1));
    rows[i] = v0 * cx[0] + v1 * cx[1] + v2 * cx[2] + v3 * cx[3];
  }
  return rows[0] * cy[0] + rows[1] * cy[1] + rows[2] * cy[2] + rows[3] * cy[3];
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// hi = x truncated to TF32, lo = x - hi (exact; the tensor core ignores lo's 13 low mantissa bits: 2^-20 |x|)
__device__ __forceinline__ void ft_split4(const float (&x)[4], uint4& hi, uint4& lo) {
  hi.x = __float_as_uint(x[0]) & 0xffffe000u; hi.y = __float_as_uint(x[1]) & 0xffffe000u;
  hi.z = __float_as_uint(x[2]) & 0xffffe000u; hi.w = __float_as_uint(x[3]) & 0xffffe000u;
  lo.x = __float_as_uint(x[0] - __uint_as_float(hi.x)); lo.y = __float_as_uint(x[1] - __uint_as_float(hi.y));
  lo.z = __float_as_uint(x[2] - __uint_as_float(hi.z)); lo.w = __float_as_uint(x[3] - __uint_as_float(hi.w));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


template <int C>
struct FuseTcSmem {
  static constexpr uint32_t kALBO = kFM * 16;                 // 2048: stride between 4-channel core-matrix columns of X
  static constexpr uint32_t kBLBO = C * 16;                   // ... of W
  static constexpr uint32_t kABytes = (kFKC / 4) * kALBO;     // 8192 per hi / lo tile
  static constexpr uint32_t kBBytes = (kFKC / 4) * kBLBO;
  static constexpr uint32_t kStage = 2 * kABytes + 2 * kBBytes;
  static constexpr uint32_t kBars = kFStages * kStage;        // full[3], empty[3], accfull, tmem slot
  static constexpr uint32_t kTotal = kBars + 8 * (2 * kFStages + 1) + 16;
};

// grid: persistent, up to 2 (C = 128) or 3 CTAs per SM; each CTA walks 128-pixel tiles with stride gridDim.x
template <int C>
__global__ void __launch_bounds__(kFThreads, C == 128 ? 2 : 3)
fuse_level_tc_kernel(const float* __restrict__ dec, const float* __restrict__ tt, const float* __restrict__ S,
                     const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out, int n_items,
                     int h, int w, int scale) {
  using L = FuseTcSmem<C>;
  int* error_flag = nullptr;   // bounded waits trap on a starved barrier (tc_ptx.cuh): no sticky global state
  constexpr int K = 2 * C, NCH = K / kFKC;
  constexpr uint32_t kCols = C < 32 ? 32 : C;  // TMEM columns (power of two >= 32)
  extern __shared__ __align__(1024) uint8_t fsmem[];
  const uint32_t s0 = smem_u32(fsmem);
  const uint32_t bar_full = s0 + L::kBars, bar_empty = bar_full + 8 * kFStages, bar_acc = bar_empty + 8 * kFStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fsmem + L::kBars + 8 * (2 * kFStages + 1));

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int hs = h * scale, wsz = w * scale;
  const size_t plane = (size_t)hs * wsz;

  if (t == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(bar_full + 8 * s, kFM); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long tpi = (long long)((plane + kFM - 1) / kFM), total = tpi * n_items;

  if (warp < 4) {
    // ============================ producer: thread = pixel (and weight row) ============================
    const int px = t;
    const bool wrow = px < C;
    const float* w_row = weight + (size_t)(wrow ? px : 0) * K;

    struct TileCtx { bool pin; size_t base; float sw; };  // base = offset of (item, channel 0, this pixel)
    auto make_ctx = [&](long long tile) {
      TileCtx c;
      const int n = (int)(tile / tpi);
      const size_t p = (size_t)(tile - (long long)n * tpi) * kFM + px;
      c.pin = p < plane;
      c.base = (size_t)n * C * plane + (c.pin ? p : 0);
      c.sw = 0.f;  // soft-attention weight of this pixel (bicubic upsampled S), kept in a register for the epilogue
      if (c.pin) {
        const int oy = (int)(p / wsz), ox = (int)(p % wsz);
        const float* S_n = S + (size_t)n * h * w;
        c.sw = scale == 1 ? __ldg(S_n + (size_t)oy * w + ox) : ft_bicubic(S_n, h, w, oy, ox, 1.0f / (float)scale);
      }
      return c;
    };

    // Loads run one 32-channel set (= two pipeline stages) ahead of the stores in a two-set register ring, so
    // every thread keeps 32 coalesced 4-byte loads in flight (16 KB per CTA) while it splits and stores the
    // previous set; the first set of the NEXT tile is issued before the epilogue of the current one.  A set
    // never straddles the dec / T halves of the concatenation (C is a multiple of 32).
    constexpr int kSet = 2 * kFKC, NSC = K / kSet;
    auto load_set = [&](const TileCtx& c, int sc, float (&x)[kSet]) {
      const int ch0 = sc * kSet;
      const float* src = (ch0 < C ? dec + (size_t)ch0 * plane : tt + (size_t)(ch0 - C) * plane) + c.base;
#pragma unroll
      for (int i = 0; i < kSet; ++i) x[i] = c.pin ? __ldg(src + (size_t)i * plane) : 0.f;
    };
    // The weight row (L2 resident, 64 bytes per stage) is prefetched one stage ahead (wrapping into the next tile).
    float4 wnext[kFKC / 4];
    auto load_w = [&](int kc) {
      if (kc == NCH) kc = 0;
#pragma unroll
      for (int j = 0; j < kFKC / 4; ++j)
        wnext[j] = wrow ? __ldg(reinterpret_cast<const float4*>(w_row + kc * kFKC) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_set = [&](int g0, int sc, const float (&x)[kSet]) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int kc = 2 * sc + hf, g = g0 + kc, s = g % kFStages;
        float4 wv[kFKC / 4];
#pragma unroll
        for (int j = 0; j < kFKC / 4; ++j) wv[j] = wnext[j];
        load_w(kc + 1);
        if (g >= kFStages) mbar_wait(bar_empty + 8 * s, (uint32_t)((g / kFStages - 1) & 1), error_flag);
        const uint32_t a_hi = s0 + s * L::kStage, a_lo = a_hi + L::kABytes, b_hi = a_lo + L::kABytes, b_lo = b_hi + L::kBBytes;
#pragma unroll
        for (int j = 0; j < kFKC / 4; ++j) {
          const int e = hf * kFKC + 4 * j;
          const float x4[4] = {x[e], x[e + 1], x[e + 2], x[e + 3]};
          uint4 hi, lo;
          ft_split4(x4, hi, lo);
          st_shared_v4(a_hi + j * L::kALBO + px * 16, hi);
          st_shared_v4(a_lo + j * L::kALBO + px * 16, lo);
          if (wrow) {
            const float4 w4v = wv[j];
            const float w4[4] = {w4v.x, w4v.y, w4v.z, w4v.w};
            ft_split4(w4, hi, lo);
            st_shared_v4(b_hi + j * L::kBLBO + px * 16, hi);
            st_shared_v4(b_lo + j * L::kBLBO + px * 16, lo);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA (async proxy)
        mbar_arrive(bar_full + 8 * s);
      }
    };

    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float xa[kSet], xb[kSet];
    long long tile = blockIdx.x;
    TileCtx cur = make_ctx(tile < total ? tile : 0);
    if (tile < total) { load_set(cur, 0, xa); load_w(0); }
    int it = 0;
#pragma unroll 1
    for (; tile < total; tile += gridDim.x, ++it) {
      const int g0 = it * NCH;
#pragma unroll 1
      for (int sc = 0; sc < NSC; sc += 2) {
        load_set(cur, sc + 1, xb);
        store_set(g0, sc, xa);
        if (sc + 2 < NSC) load_set(cur, sc + 2, xa);
        store_set(g0, sc + 1, xb);
      }
      // next tile: context (bicubic taps) and first set of loads go out before this tile's epilogue
      const bool more = tile + gridDim.x < total;
      TileCtx nxt = cur;
      if (more) { nxt = make_ctx(tile + gridDim.x); load_set(nxt, 0, xa); }

      // ---------------- epilogue: thread = TMEM lane = pixel ----------------
      mbar_wait(bar_acc, (uint32_t)(it & 1), error_flag);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 16) {
        uint32_t a[16];
        tc_ld16(taddr + c0, a);
        float dv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dv[i] = cur.pin ? __ldg(dec + cur.base + (size_t)(c0 + i) * plane) : 0.f;  // residual (L2 hit)
        tc_wait_ld();
        if (cur.pin) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            out[cur.base + (size_t)(c0 + i) * plane] = dv[i] + (__uint_as_float(a[i]) + __ldg(bias + c0 + i)) * cur.sw;
        }
      }
      tc_fence_before();  // accumulator reads are complete before this thread's next barrier arrival lets the MMA overwrite it
      cur = nxt;
    }
  } else if (lane == 0) {
    // ======================================== MMA issuer ========================================
    // kind::tf32 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = tf32 (bits 7-9, 10-12 = 2), K-major both,
    // N >> 3 at bits 17-22, M >> 4 at bits 24-28
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(kFM >> 4) << 24);
    int g = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int kc = 0; kc < NCH; ++kc, ++g) {
        const int s = g % kFStages;
        mbar_wait(bar_full + 8 * s, (uint32_t)((g / kFStages) & 1), error_flag);
        tc_fence_after();
        const uint32_t a_hi = s0 + s * L::kStage, a_lo = a_hi + L::kABytes, b_hi = a_lo + L::kABytes, b_lo = b_hi + L::kBBytes;
#pragma unroll
        for (uint32_t kk = 0; kk < kFKC / 8; ++kk) {
          const uint64_t dah = umma_desc_kmajor(a_hi + kk * 2 * L::kALBO, L::kALBO, 128);
          const uint64_t dal = umma_desc_kmajor(a_lo + kk * 2 * L::kALBO, L::kALBO, 128);
          const uint64_t dbh = umma_desc_kmajor(b_hi + kk * 2 * L::kBLBO, L::kBLBO, 128);
          const uint64_t dbl = umma_desc_kmajor(b_lo + kk * 2 * L::kBLBO, L::kBLBO, 128);
          tc_mma_tf32(tmem_base, dal, dbh, idesc, (kc | kk) != 0);  // small terms first
          tc_mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          tc_mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        tc_commit(bar_empty + 8 * s);
      }
      tc_commit(bar_acc);
    }
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-fed pipeline constants (planes whose byte stride is a multiple of 16, i.e. plane % 4 == 0 -- every shape the
// model produces): one elected thread streams raw [16 channel][128 pixel] fp32 boxes of dec / T through a TMA ring.
// ---------------------------------------------------------------------------------------------------------
#ifdef SPEI_FUSE_SPIN
#define FUSE_WAIT mbar_wait_spin
#else
#define FUSE_WAIT mbar_wait
#endif
constexpr uint32_t kFRawBytes = kFKC * kFM * 4;           // 8192: one raw activation box
constexpr int kFOutCh = 16;                               // channels per epilogue staging pass

// ---------------------------------------------------------------------------------------------------------
// v3 ("TS"): the activations are the A operand FROM TENSOR MEMORY.
//
// Its predecessor (round 1, removed) kept both operands in shared memory; clock64 accounting showed the converter threads
// busy 87 % of the time, the MMA issuer waiting for operands 60 %, the TMA producer waiting for a free slot 86 %: every
// activation float crossed shared memory as 4 B TMA write + 4 B LDS + 8 B hi/lo STS + 12 B UMMA operand reads (28 B per
// 4 B of HBM traffic), and the weights again per 128-pixel tile: the shared-memory pipe, not HBM, bounded that design.  Here
//   * the converter thread (= pixel = TMEM lane) reads its 16 channels of the raw TMA box, splits them and writes hi / lo
//     straight into a ring of A tiles in tensor memory (tcgen05.st, 32 columns per stage); the MMA reads A from there
//     (tcgen05.mma [d], [a], b-desc): per activation float shared memory now carries 4 B TMA write + 4 B LDS;
//   * for C <= 64 the hi / lo weight tiles of ALL stages are converted once per CTA and stay resident in shared memory
//     (64 KB / 16 KB); C = 128 (256 KB of hi + lo) keeps the per-stage conversion of the 16 weight columns, with two
//     converter threads per pixel.
// TMEM columns: [accumulator 0 | accumulator 1 | A ring: kU x (16 hi + 16 lo)].
// ---------------------------------------------------------------------------------------------------------
template <int C>
struct FuseTsCfg {
  static constexpr bool kResidentW = C <= 64;
  static constexpr int kConvWarps = C == 128 ? 8 : 4;
  static constexpr int kSplit = kConvWarps / 4;
  static constexpr int kThreads = (kConvWarps + 6) * 32;
  static constexpr int kCtasPerSm = C == 128 ? 1 : 2;
  static constexpr int kRaw = C == 128 ? 5 : (C == 64 ? 3 : 4);   // raw TMA ring depth
  static constexpr int kE = C == 64 ? 3 : 4;                       // residual / output box ring depth (epilogue)
  static constexpr int kU = 4;                                     // A ring depth in TMEM (= B ring depth for C = 128)
  static constexpr int NCH = 2 * C / kFKC;
  static constexpr uint32_t kBLBO = C * 16;                        // stride between 4-channel core-matrix columns of W
  static constexpr uint32_t kBBytes = (kFKC / 4) * kBLBO;          // one hi (or lo) weight tile of a stage
  static constexpr uint32_t kWBox = kResidentW ? 0u : (uint32_t)(C * kFKC * 4);
  static constexpr uint32_t kRawStage = kFRawBytes + kWBox;
  static constexpr uint32_t kOffB = kRaw * kRawStage;
  static constexpr uint32_t kOffOut = kOffB + (uint32_t)(kResidentW ? NCH : kU) * 2 * kBBytes;
  static constexpr uint32_t kOutBox = kFOutCh * kFM * 4;           // [16 channel][128 pixel] fp32, the TMA box of dec / out
  static constexpr uint32_t kOffBars = kOffOut + kE * kOutBox;
  static constexpr uint32_t kNumBars = 2 * kRaw + 2 * kU + 4 + kE;
  static constexpr uint32_t kTotal = kOffBars + 8 * kNumBars + 16;
  static constexpr uint32_t kACol0 = 2 * C;
  static constexpr uint32_t kColsNeeded = 2 * C + kU * 32;
  static constexpr uint32_t kCols = kColsNeeded <= 128 ? 128 : (kColsNeeded <= 256 ? 256 : 512);
};

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// kBf16: dec / T / out are bf16 tensors (native bf16 I/O, north_star's 1e-2 mode).  A bf16 value is exactly representable in
// TF32, so the activation split has no low part: the converter writes only the hi tile, the MMA issues 2 instead of 3
// products per step (X.Wlo + X.Whi), the TMA boxes are half the bytes, and the epilogue updates / stores a bf16 box.
// The arithmetic (fp32 weights split 3xTF32-style, fp32 accumulation, fp32 residual add, one rounding to bf16 at the
// store) is exactly what the fp32 kernel computes on the up-cast inputs.
template <int C, bool kBf16>
__global__ void __launch_bounds__(FuseTsCfg<C>::kThreads, FuseTsCfg<C>::kCtasPerSm)
fuse_level_ts_kernel(const __grid_constant__ CUtensorMap tm_dec, const __grid_constant__ CUtensorMap tm_t,
                     const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ S,
                     const float* __restrict__ weight, const float* __restrict__ bias, int n_items, int h, int w, int scale) {
  using F = FuseTsCfg<C>;
  constexpr uint32_t kRawTx = (kBf16 ? kFRawBytes / 2 : kFRawBytes) + F::kWBox;   // bytes one producer stage brings in
  constexpr uint32_t kBoxTx = kBf16 ? F::kOutBox / 2 : F::kOutBox;               // bytes of a residual / output box
  constexpr int kRaw = F::kRaw, kU = F::kU, kE = F::kE, CW = F::kConvWarps, kConvThreads = CW * 32, NCH = F::NCH;
  constexpr int kCh = kFKC / F::kSplit;          // channels of a stage per converter thread (16 or 8)
  constexpr int K = 2 * C;
  int* error_flag = nullptr;   // bounded waits trap on a starved barrier (tc_ptx.cuh): no sticky global state
  extern __shared__ __align__(1024) uint8_t fsmem[];
  const uint32_t s0 = smem_u32(fsmem);
  const uint32_t bar_rfull = s0 + F::kOffBars, bar_rempty = bar_rfull + 8 * kRaw, bar_afull = bar_rempty + 8 * kRaw,
                 bar_aempty = bar_afull + 8 * kU, bar_dfull = bar_aempty + 8 * kU, bar_dempty = bar_dfull + 16,
                 bar_efull = bar_dempty + 16;   // residual box of an epilogue pass has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fsmem + F::kOffBars + 8 * F::kNumBars);
  float* ostage = reinterpret_cast<float*>(fsmem + F::kOffOut);   // [kE][16 ch][128 px]
  const float* raw = reinterpret_cast<const float*>(fsmem);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int hs = h * scale, wsz = w * scale;
  const size_t plane = (size_t)hs * wsz;
  const long long tpi = (long long)((plane + kFM - 1) / kFM), total = tpi * n_items;

  if (t == 0) {
    for (int s = 0; s < kRaw; ++s) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, kConvThreads); }
    for (int s = 0; s < kU; ++s) { mbar_init(bar_afull + 8 * s, kConvThreads); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_dempty + 8 * s, kFM); }
    for (int s = 0; s < kE; ++s) mbar_init(bar_efull + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_dec) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_t) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_out) : "memory");
    if (!F::kResidentW) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (warp == CW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(F::kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (F::kResidentW) {
    // hi / lo weight tiles of every stage, once per CTA: element (o, k) -> stage k/16, chunk (k%16)/4, row o.
    // Consecutive threads take consecutive rows o (conflict-free 16-byte stores; the strided reads are L2 hits).
    for (int idx = t; idx < C * (K / 4); idx += F::kThreads) {
      const int o = idx % C, k4 = idx / C;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(weight + (size_t)o * K) + k4);
      const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
      uint4 hi, lo;
      ft_split4(w4, hi, lo);
      const uint32_t b_hi = s0 + F::kOffB + (uint32_t)(k4 >> 2) * 2 * F::kBBytes + (uint32_t)(k4 & 3) * F::kBLBO + (uint32_t)o * 16;
      st_shared_v4(b_hi, hi);
      st_shared_v4(b_hi + F::kBBytes, lo);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < CW) {
    // ====== convert: thread = (pixel = TMEM lane, channel part) ======
    const int px = t & (kFM - 1), part = t / kFM;
    const bool wrow = !F::kResidentW && px < C;
    const uint32_t a_lane = tmem_base + F::kACol0 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * kCh);
    long long tile = blockIdx.x;
    int g = 0;
    [[maybe_unused]] long long pf_t0 = clock64(), pf_rfull = 0, pf_aempty = 0;
#pragma unroll 1
    for (; tile < total; tile += gridDim.x) {
#pragma unroll 1
      for (int kc = 0; kc < NCH; ++kc, ++g) {
        const int r = g % kRaw, u = g % kU;
        FPROF_T(pf_rfull, FUSE_WAIT(bar_rfull + 8 * r, (uint32_t)((g / kRaw) & 1), error_flag));
        const float* rbox = raw + (size_t)r * (F::kRawStage / 4);
        float x[kCh];
        if (kBf16) {
          const uint16_t* rb = reinterpret_cast<const uint16_t*>(rbox);
#pragma unroll
          for (int c = 0; c < kCh; ++c) x[c] = __uint_as_float((uint32_t)rb[(part * kCh + c) * kFM + px] << 16);
        } else {
#pragma unroll
          for (int c = 0; c < kCh; ++c) x[c] = rbox[(part * kCh + c) * kFM + px];
        }
        constexpr int kWChunks = kCh / 4;          // 16-byte chunks of weight row px this thread converts (C = 128 only)
        const int jrot = (px >> 1) & (kWChunks - 1);
        float4 wv[kWChunks];
        if (!F::kResidentW) {
#pragma unroll
          for (int j = 0; j < kWChunks; ++j)
            wv[j] = wrow ? reinterpret_cast<const float4*>(rbox + kFRawBytes / 4 + px * kFKC)[part * kWChunks + ((j + jrot) & (kWChunks - 1))]
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // hi = x truncated to TF32, lo = x - hi (exact).  Computing them here consumes every value read from the box, so
        // the loads have completed before the slot is handed back to the TMA producer (no proxy fence needed for that
        // write-after-read; the fence cost ~1/4 of the converter's stage time in the shared-memory-operand kernel).
        uint32_t xhi[kCh], xlo[kCh];
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
          xhi[i] = __float_as_uint(x[i]) & 0xffffe000u;
          xlo[i] = __float_as_uint(x[i] - __uint_as_float(xhi[i]));
        }
        uint4 whi[kWChunks], wlo[kWChunks];
        if (!F::kResidentW) {
#pragma unroll
          for (int j = 0; j < kWChunks; ++j) {
            const float w4[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
            ft_split4(w4, whi[j], wlo[j]);
          }
        }
        mbar_arrive(bar_rempty + 8 * r);
        if (g >= kU) FPROF_T(pf_aempty, FUSE_WAIT(bar_aempty + 8 * u, (uint32_t)((g / kU - 1) & 1), error_flag));
        tc_fence_after();
        // columns [u*32 + c] (hi) and [u*32 + 16 + c] (lo) of this thread's TMEM lane
#pragma unroll
        for (int c8 = 0; c8 < kCh / 8; ++c8) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { hi[i] = xhi[8 * c8 + i]; lo[i] = xlo[8 * c8 + i]; }
          tc_st8(a_lane + (uint32_t)(u * 32 + 8 * c8), hi);
          if (!kBf16) tc_st8(a_lane + (uint32_t)(u * 32 + 16 + 8 * c8), lo);   // bf16 activations: lo == 0, never read
        }
        if (!F::kResidentW && wrow) {
          const uint32_t b_hi = s0 + F::kOffB + (uint32_t)u * 2 * F::kBBytes, b_lo = b_hi + F::kBBytes;
#pragma unroll
          for (int j = 0; j < kWChunks; ++j) {
            const uint32_t jj = (uint32_t)(part * kWChunks + ((j + jrot) & (kWChunks - 1)));
            st_shared_v4(b_hi + jj * F::kBLBO + px * 16, whi[j]);
            st_shared_v4(b_lo + jj * F::kBLBO + px * 16, wlo[j]);
          }
        }
        tc_wait_st();
        if (!F::kResidentW) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight tile stores -> MMA (async proxy)
        tc_fence_before();
        mbar_arrive(bar_afull + 8 * u);
      }
    }
#ifdef SPEI_FUSE_PROF
    if (blockIdx.x == 3 && (t == 0 || t == 100)) printf("TS C=%d conv t%d: total %lld wait_rfull %lld wait_aempty %lld stages %d\n", C, t, clock64() - pf_t0, pf_rfull, pf_aempty, g);
#endif
  } else if (warp == CW) {
    // ======================================== MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(kFM >> 4) << 24);
      int g = 0, it = 0;
      [[maybe_unused]] long long pf_t0 = clock64(), pf_dempty = 0, pf_afull = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1;
        if (it >= 2) FPROF_T(pf_dempty, FUSE_WAIT(bar_dempty + 8 * ab, (uint32_t)((it / 2 - 1) & 1), error_flag));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * C;
        for (int kc = 0; kc < NCH; ++kc, ++g) {
          const int u = g % kU;
          FPROF_T(pf_afull, FUSE_WAIT(bar_afull + 8 * u, (uint32_t)((g / kU) & 1), error_flag));
          tc_fence_after();
          const uint32_t b_hi = s0 + F::kOffB + (uint32_t)(F::kResidentW ? kc : u) * 2 * F::kBBytes, b_lo = b_hi + F::kBBytes;
          const uint32_t a_hi = tmem_base + F::kACol0 + (uint32_t)(u * 32), a_lo = a_hi + 16;
#pragma unroll
          for (uint32_t kk = 0; kk < kFKC / 8; ++kk) {
            const uint64_t dbh = umma_desc_kmajor(b_hi + kk * 2 * F::kBLBO, F::kBLBO, 128);
            const uint64_t dbl = umma_desc_kmajor(b_lo + kk * 2 * F::kBLBO, F::kBLBO, 128);
            if (!kBf16) tc_mma_tf32_ts(d_tmem, a_lo + kk * 8, dbh, idesc, (kc | kk) != 0);  // small terms first
            tc_mma_tf32_ts(d_tmem, a_hi + kk * 8, dbl, idesc, kBf16 ? (uint32_t)((kc | kk) != 0) : 1u);
            tc_mma_tf32_ts(d_tmem, a_hi + kk * 8, dbh, idesc, 1u);
          }
          tc_commit(bar_aempty + 8 * u);
        }
        tc_commit(bar_dfull + 8 * ab);
      }
#ifdef SPEI_FUSE_PROF
      if (blockIdx.x == 3) printf("TS C=%d mma: total %lld wait_dempty %lld wait_afull %lld tiles %d\n", C, clock64() - pf_t0, pf_dempty, pf_afull, it);
#endif
    }
  } else if (warp == CW + 1) {
    // ======================================== TMA producer ========================================
    if (lane == 0) {
      int g = 0;
      [[maybe_unused]] long long pf_t0 = clock64(), pf_rempty = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = (int)(tile / tpi);
        const int p0 = (int)((tile - (long long)n * tpi) * kFM);
        for (int kc = 0; kc < NCH; ++kc, ++g) {
          const int r = g % kRaw;
          if (g >= kRaw) FPROF_T(pf_rempty, FUSE_WAIT(bar_rempty + 8 * r, (uint32_t)((g / kRaw - 1) & 1), error_flag));
          mbar_arrive_expect_tx(bar_rfull + 8 * r, kRawTx);
          const int ch0 = kc * kFKC;
          tma_load_2d(s0 + r * F::kRawStage, ch0 < C ? &tm_dec : &tm_t, bar_rfull + 8 * r, p0, n * C + (ch0 < C ? ch0 : ch0 - C));
          if (!F::kResidentW) tma_load_2d(s0 + r * F::kRawStage + kFRawBytes, &tm_w, bar_rfull + 8 * r, ch0, 0);  // weight columns ch0..ch0+15
        }
      }
#ifdef SPEI_FUSE_PROF
      if (blockIdx.x == 3) printf("TS C=%d tma: total %lld wait_rempty %lld\n", C, clock64() - pf_t0, pf_rempty);
#endif
    }
  } else {
    // ================================ epilogue (thread = TMEM lane = pixel) ================================
    // One pass = 16 output channels of a tile = one [16 channel][128 pixel] box.  The residual box of `dec` arrives by
    // TMA into a ring slot (prefetched kE - 1 passes ahead: the shared-memory-operand kernel read it with LDGs one
    // pass ahead and was bound by that latency), every thread adds (acc + bias) * soft weight to its pixel in place,
    // and one thread stores the box with a TMA store.  Partial last tiles are clipped by the tensor maps.
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int px = q * 32 + lane;
    const bool elected = warp == CW + 2 && lane == 0;
    auto soft_weight = [&](long long tile) {  // bicubic upsampled S at this thread's pixel of `tile`
      const int n = (int)(tile / tpi);
      const size_t p = (size_t)(tile - (long long)n * tpi) * kFM + px;
      if (p >= plane) return 0.f;
      const int oy = (int)(p / wsz), ox = (int)(p % wsz);
      const float* S_n = S + (size_t)n * h * w;
      return scale == 1 ? __ldg(S_n + (size_t)oy * w + ox) : ft_bicubic(S_n, h, w, oy, ox, 1.0f / (float)scale);
    };
    constexpr int kPasses = C / kFOutCh;
    // residual box of global pass number `pn` (tile = blockIdx.x + (pn / kPasses) * gridDim.x, channels (pn % kPasses) * 16)
    auto issue_residual = [&](long long pn) {
      const long long tl = (long long)blockIdx.x + (pn / kPasses) * (long long)gridDim.x;
      if (tl >= total) return;
      const int n = (int)(tl / tpi);
      const int p0 = (int)((tl - (long long)n * tpi) * kFM);
      const int e = (int)(pn % kE);
      mbar_arrive_expect_tx(bar_efull + 8 * e, kBoxTx);
      tma_load_2d(s0 + F::kOffOut + e * F::kOutBox, &tm_dec, bar_efull + 8 * e, p0, n * C + (int)(pn % kPasses) * kFOutCh);
    };
    if (elected) {
      for (int i = 0; i < kE - 1; ++i) issue_residual(i);
    }
    long long tile = blockIdx.x;
    float sw = tile < total ? soft_weight(tile) : 0.f;
    int it = 0;
    long long pass = 0;
    [[maybe_unused]] long long pf_t0 = clock64(), pf_dfull = 0, pf_efull = 0;
#pragma unroll 1
    for (; tile < total; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      const float sw_next = tile + gridDim.x < total ? soft_weight(tile + gridDim.x) : 0.f;  // taps in flight during this tile
      const int n = (int)(tile / tpi);
      const int p0 = (int)((tile - (long long)n * tpi) * kFM);
      FPROF_T(pf_dfull, FUSE_WAIT(bar_dfull + 8 * ab, (uint32_t)((it / 2) & 1), error_flag));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * C + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += kFOutCh, ++pass) {
        const int e = (int)(pass % kE);
        float* ost = ostage + e * (kFOutCh * kFM);
        uint32_t a[16];
        tc_ld16(taddr + c0, a);
        FPROF_T(pf_efull, FUSE_WAIT(bar_efull + 8 * e, (uint32_t)((pass / kE) & 1), error_flag));
        tc_wait_ld();
        if (kBf16) {
          uint16_t* ob = reinterpret_cast<uint16_t*>(ost);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = __uint_as_float((uint32_t)ob[i * kFM + px] << 16) + (__uint_as_float(a[i]) + __ldg(bias + c0 + i)) * sw;
            ob[i * kFM + px] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) ost[i * kFM + px] = ost[i * kFM + px] + (__uint_as_float(a[i]) + __ldg(bias + c0 + i)) * sw;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the TMA store (async proxy)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (elected) {
          tma_store_2d(&tm_out, s0 + F::kOffOut + e * F::kOutBox, p0, n * C + c0);
          bulk_commit_group();
          // the slot of the PREVIOUS pass is free once its store has read shared memory: refill it kE - 1 passes ahead
          bulk_wait_group_read<1>();
          issue_residual(pass >= 1 ? pass - 1 + kE : kE - 1);   // (pass 0: the one slot the prologue left empty)
        }
      }
      tc_fence_before();
      mbar_arrive(bar_dempty + 8 * ab);   // this accumulator may be overwritten by the tile after next
      sw = sw_next;
    }
    if (elected) bulk_wait_group<0>();     // all output boxes are in global memory before the CTA exits
#ifdef SPEI_FUSE_PROF
    if (blockIdx.x == 3 && px == 0) printf("TS C=%d epi: total %lld wait_dfull %lld wait_efull %lld tiles %d\n", C, clock64() - pf_t0, pf_dfull, pf_efull, it);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(F::kCols) : "memory");
  }
}

// 2-D map over an NCHW fp32 (or bf16) tensor seen as [n*C rows][plane]; box = 16 channels x 128 pixels
static int make_plane_map(EncodeTiledFn enc, CUtensorMap* tm, const void* base, size_t rows, size_t plane, bool bf16 = false) {
  const cuuint64_t dims[2] = {(cuuint64_t)plane, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)plane * (bf16 ? 2 : 4)};
  const cuuint32_t box[2] = {kFM, kFKC};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fuse_level) failed with CUresult %d", (int)r); return SPEI_ERR_CUDA; }
  return SPEI_OK;
}

template <int C, bool kBf16>
static int launch_fuse_tma_t(int n, int h, int w, int scale, const void* dec, const void* t, const float* S, const float* weight,
                             const float* bias, void* out, cudaStream_t st) {
  const size_t plane = (size_t)h * scale * w * scale;
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  CUtensorMap tmd, tmt, tmw;
  if ((rc = make_plane_map(enc, &tmd, dec, (size_t)n * C, plane, kBf16))) return rc;
  if ((rc = make_plane_map(enc, &tmt, t, (size_t)n * C, plane, kBf16))) return rc;
  {  // Conv2d 1x1 weight [C rows][2C] fp32; box = all C rows x 16 input channels
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * C), (cuuint64_t)C};
    const cuuint64_t strides[1] = {(cuuint64_t)(2 * C) * 4};
    const cuuint32_t box[2] = {kFKC, (cuuint32_t)C};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(weight), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fuse_level weight) failed with CUresult %d", (int)r); return SPEI_ERR_CUDA; }
  }
  const int smem = (int)FuseTsCfg<C>::kTotal;
  auto kern = fuse_level_ts_kernel<C, kBf16>;
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  SPEI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  int dev = 0, sms = 0;
  SPEI_CUDA(cudaGetDevice(&dev));
  SPEI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (long long)((plane + kFM - 1) / kFM) * n, slots = (long long)sms * FuseTsCfg<C>::kCtasPerSm;
  CUtensorMap tmo;
  if ((rc = make_plane_map(enc, &tmo, out, (size_t)n * C, plane, kBf16))) return rc;
  kern<<<(unsigned)(tiles < slots ? tiles : slots), FuseTsCfg<C>::kThreads, smem, st>>>(tmd, tmt, tmw, tmo, S, weight, bias, n, h, w, scale);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

template <int C>
static int launch_fuse_tc_t(int n, int h, int w, int scale, const float* dec, const float* t, const float* S, const float* weight,
                            const float* bias, float* out, cudaStream_t st) {
  const size_t plane = (size_t)h * scale * w * scale;
  const int smem = (int)FuseTcSmem<C>::kTotal;
  SPEI_CUDA(cudaFuncSetAttribute(fuse_level_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  SPEI_CUDA(cudaFuncSetAttribute(fuse_level_tc_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  int dev = 0, sms = 0;
  SPEI_CUDA(cudaGetDevice(&dev));
  SPEI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (long long)((plane + kFM - 1) / kFM) * n, slots = (long long)sms * (C == 128 ? 2 : 3);
  fuse_level_tc_kernel<C><<<(unsigned)(tiles < slots ? tiles : slots), kFThreads, smem, st>>>(dec, t, S, weight, bias, out, n, h, w, scale);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_fuse_level(int n, int c, int h, int w, int scale, const float* dec, const float* t, const float* S,
                      const float* weight, const float* bias, float* out, cudaStream_t st) {
  if (n > 65535) { set_error("fuse_level: n too large"); return SPEI_ERR_ARG; }
  const size_t plane = (size_t)h * scale * w * scale;
  // TMA needs 16-byte plane strides and 16-byte aligned bases; anything else takes the LDG-fed kernel
  const bool aligned = (((uintptr_t)dec | (uintptr_t)t | (uintptr_t)out | (uintptr_t)weight) & 15) == 0;
  if (aligned && plane % 4 == 0 && (long long)n * c < (1ll << 31) && plane < (1ull << 31)) {
    if (c == 128) return launch_fuse_tma_t<128, false>(n, h, w, scale, dec, t, S, weight, bias, out, st);
    if (c == 64) return launch_fuse_tma_t<64, false>(n, h, w, scale, dec, t, S, weight, bias, out, st);
    if (c == 32) return launch_fuse_tma_t<32, false>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  }
  if (c == 128) return launch_fuse_tc_t<128>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  if (c == 64) return launch_fuse_tc_t<64>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  if (c == 32) return launch_fuse_tc_t<32>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  set_error("fuse_level: unsupported channel count %d", c);
  return SPEI_ERR_ARG;
}

// bf16 dec / t / out (TMA-fed kernel only: planes must be a multiple of 8 pixels and the bases 16-byte aligned)
int launch_fuse_level_bf16(int n, int c, int h, int w, int scale, const void* dec, const void* t, const float* S,
                           const float* weight, const float* bias, void* out, cudaStream_t st) {
  if (n > 65535) { set_error("fuse_level: n too large"); return SPEI_ERR_ARG; }
  const size_t plane = (size_t)h * scale * w * scale;
  const bool aligned = (((uintptr_t)dec | (uintptr_t)t | (uintptr_t)out | (uintptr_t)weight) & 15) == 0;
  if (!aligned || plane % 8 != 0 || (long long)n * c >= (1ll << 31) || plane >= (1ull << 31)) {
    set_error("fuse_level (bf16): planes must be a multiple of 8 pixels (got %zu) and every base 16-byte aligned; "
              "up-cast to fp32 for other shapes", plane);
    return SPEI_ERR_ARG;
  }
  if (c == 128) return launch_fuse_tma_t<128, true>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  if (c == 64) return launch_fuse_tma_t<64, true>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  if (c == 32) return launch_fuse_tma_t<32, true>(n, h, w, scale, dec, t, S, weight, bias, out, st);
  set_error("fuse_level: unsupported channel count %d", c);
  return SPEI_ERR_ARG;
}

}  // namespace spei
