// Kernel (d): one fusion line of SPEINet._decode
// (/root/reference/model/speinet.py:93-94, 96-97, 108-109):
//
//     out = dec + (W . cat(dec, T) + b) * bicubic_up(S, scale)
//
// fused into one pass over HBM (read dec, read T, write out -- 3 x tensor size, the algorithmic
// minimum): the concatenation is never materialised (two row ranges of one K loop), the x2 / x4
// bicubic upsampling of the one-channel map S (align_corners=False, A=-0.75, border-clamped taps,
// torch/include/ATen/native/UpSample.h:289-300,400-423) is evaluated inline, and bias, multiply and
// residual add are applied before the single store.
//
// The 1x1 convolution is a skinny [C x 2C] x [2C x pixels] GEMM that must stay fp32-accurate (1e-4
// bar) yet not become the bottleneck of an HBM-bound kernel, so it runs on the tensor cores with the
// 3xTF32 split: x = hi + lo (both TF32), W.X ~= Wlo.Xhi + Whi.Xlo + Whi.Xhi with fp32 accumulation
// (error ~2^-21 per product, i.e. fp32 class).  Operands are streamed through a double-buffered
// cp.async pipeline; shared-memory pitches (W: K-chunk+4, X: 128+8 floats) make every mma.sync
// fragment load bank-conflict free.
#include <cstdlib>

#include "spei_common.cuh"

namespace spei {

constexpr int kFN = 128;            // pixels per block
constexpr int kFK = 32;             // input channels per pipeline step
constexpr int kXP = kFN + 8;        // X pitch (floats): 8*k + n -> 32 distinct banks
constexpr int kWP = kFK + 4;        // W pitch (floats): 4*m + k -> 32 distinct banks

__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ float bicubic_tap(const float* __restrict__ S, int h, int w, int oy, int ox, float rscale) {
  const float ry = rscale * (oy + 0.5f) - 0.5f, rx = rscale * (ox + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  const int iy = (int)fy, ix = (int)fx;
  const float ty = ry - fy, tx = rx - fx;
  const float cx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
  const float cy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
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

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x = hi + lo with hi = x truncated to TF32 (top 19 bits) and lo = x - hi, which is exact in fp32.  The tensor
// core ignores the 13 low mantissa bits of a .tf32 operand, so lo needs no conversion either: its own
// truncation error is 2^-10 |lo| <= 2^-20 |x|.  One LOP3 + one FADD instead of two cvt.rna (the conversion
// pipe, not the MMA, was what bounded this kernel).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// grid: (ceil(plane/128), 1, n)   block: 256 = (C/32) warps along channels x the rest along pixels
template <int C, bool kVec>
__global__ void __launch_bounds__(256)
fuse_level_kernel(const float* __restrict__ dec, const float* __restrict__ tt, const float* __restrict__ S,
                  const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out, int h, int w,
                  int scale) {
  constexpr int K = 2 * C, WM = C / 32, WN = 8 / WM, WPX = kFN / WN, NTW = WPX / 8;
  extern __shared__ __align__(16) float fsm[];
  float (*Xs)[kFK][kXP] = reinterpret_cast<float (*)[kFK][kXP]>(fsm);                       // [2][32][136]
  float (*Ws)[C][kWP] = reinterpret_cast<float (*)[C][kWP]>(fsm + 2 * kFK * kXP);           // [2][C][36]
  float* sw = fsm + 2 * kFK * kXP + 2 * C * kWP;                                            // [128]

  const int hs = h * scale, wsz = w * scale;
  const size_t plane = (size_t)hs * wsz;
  const int n = blockIdx.z;
  const size_t p0 = (size_t)blockIdx.x * kFN;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const float* dec_n = dec + (size_t)n * C * plane;
  const float* t_n = tt + (size_t)n * C * plane;

  auto load_chunk = [&](int buf, int k0) {
    if (kVec) {
#pragma unroll
      for (int it = 0; it < (kFK * kFN / 4) / 256; ++it) {
        const int e = t + 256 * it, row = e >> 5, c4 = e & 31, ch = k0 + row;
        const float* src = (ch < C ? dec_n + (size_t)ch * plane : t_n + (size_t)(ch - C) * plane) + p0 + c4 * 4;
        cp_async16(&Xs[buf][row][c4 * 4], src, p0 + c4 * 4 < plane);
      }
    } else {
#pragma unroll 4
      for (int it = 0; it < (kFK * kFN) / 256; ++it) {
        const int e = t + 256 * it, row = e >> 7, c = e & 127, ch = k0 + row;
        const float* src = (ch < C ? dec_n + (size_t)ch * plane : t_n + (size_t)(ch - C) * plane) + p0 + c;
        cp_async4(&Xs[buf][row][c], src, p0 + c < plane);
      }
    }
#pragma unroll
    for (int it = 0; it < (C * kFK / 4 + 255) / 256; ++it) {
      const int e = t + 256 * it, row = e >> 3, c4 = e & 7;
      if (row < C) cp_async16(&Ws[buf][row][c4 * 4], weight + (size_t)row * K + k0 + c4 * 4, true);
    }
    cp_async_commit();
  };

  load_chunk(0, 0);
  // soft-attention weight of this block's pixels (bicubic upsampled S), once per block
  if (t < kFN) {
    const size_t p = p0 + t;
    float v = 0.f;
    if (p < plane) {
      const int oy = (int)(p / wsz), ox = (int)(p % wsz);
      const float* S_n = S + (size_t)n * h * w;
      v = scale == 1 ? __ldg(S_n + (size_t)oy * w + ox) : bicubic_tap(S_n, h, w, oy, ox, 1.0f / (float)scale);
    }
    sw[t] = v;
  }

  float acc[2][NTW][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

  constexpr int NCH = K / kFK;
  const int g = lane >> 2, q4 = lane & 3;
  for (int kc = 0; kc < NCH; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < NCH) { load_chunk(buf ^ 1, (kc + 1) * kFK); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll
    for (int k8 = 0; k8 < kFK / 8; ++k8) {
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* wr = &Ws[buf][wm * 32 + mt * 16 + g][k8 * 8 + q4];
        split_tf32(wr[0], ahi[mt][0], alo[mt][0]);
        split_tf32(wr[8 * kWP], ahi[mt][1], alo[mt][1]);
        split_tf32(wr[4], ahi[mt][2], alo[mt][2]);
        split_tf32(wr[8 * kWP + 4], ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const float* xr = &Xs[buf][k8 * 8 + q4][wn * WPX + nt * 8 + g];
        uint32_t bhi[2], blo[2];
        split_tf32(xr[0], bhi[0], blo[0]);
        split_tf32(xr[4 * kXP], bhi[1], blo[1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], alo[mt], bhi);   // small terms first
          mma_tf32(acc[mt][nt], ahi[mt], blo);
          mma_tf32(acc[mt][nt], ahi[mt], bhi);
        }
      }
    }
    __syncthreads();
  }

  // epilogue: bias, multiply by the soft-attention weight, residual add, one store
  const bool pair_ok = (plane & 1) == 0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int o = wm * 32 + mt * 16 + g + hh * 8;
      const float b = __ldg(bias + o);
      const size_t row = ((size_t)n * C + o) * plane;
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int pl = wn * WPX + nt * 8 + 2 * q4;
        const size_t p = p0 + pl;
        const float r0 = (acc[mt][nt][hh * 2 + 0] + b) * sw[pl], r1 = (acc[mt][nt][hh * 2 + 1] + b) * sw[pl + 1];
        if (pair_ok && p + 1 < plane) {
          const float2 d = __ldg(reinterpret_cast<const float2*>(dec + row + p));
          *reinterpret_cast<float2*>(out + row + p) = make_float2(d.x + r0, d.y + r1);
        } else {
          if (p < plane) out[row + p] = __ldg(dec + row + p) + r0;
          if (p + 1 < plane) out[row + p + 1] = __ldg(dec + row + p + 1) + r1;
        }
      }
    }
}

template <int C, bool kVec>
static int launch_fuse_t(int n, int h, int w, int scale, const float* dec, const float* t, const float* S, const float* weight,
                         const float* bias, float* out, cudaStream_t st) {
  const size_t plane = (size_t)h * scale * w * scale;
  const int smem = (2 * kFK * kXP + 2 * C * kWP + kFN) * (int)sizeof(float);
  SPEI_CUDA(cudaFuncSetAttribute(fuse_level_kernel<C, kVec>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)((plane + kFN - 1) / kFN), 1, n);
  fuse_level_kernel<C, kVec><<<grid, 256, smem, st>>>(dec, t, S, weight, bias, out, h, w, scale);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_fuse_level_tc(int n, int c, int h, int w, int scale, const float* dec, const float* t, const float* S,
                         const float* weight, const float* bias, float* out, cudaStream_t st);

int launch_fuse_level(int n, int c, int h, int w, int scale, const float* dec, const float* t, const float* S,
                      const float* weight, const float* bias, float* out, cudaStream_t st) {
  if (n > 65535) { set_error("fuse_level: n too large"); return SPEI_ERR_ARG; }
  static const bool legacy = getenv("SPEI_FUSE_MMA_SYNC") != nullptr;  // A/B switch: the mma.sync version below
  if (!legacy) return launch_fuse_level_tc(n, c, h, w, scale, dec, t, S, weight, bias, out, st);
  const size_t plane = (size_t)h * scale * w * scale;
  const bool vec = (plane % 4) == 0;
#define FUSE(C_) (vec ? launch_fuse_t<C_, true>(n, h, w, scale, dec, t, S, weight, bias, out, st) \
                      : launch_fuse_t<C_, false>(n, h, w, scale, dec, t, S, weight, bias, out, st))
  if (c == 128) return FUSE(128);
  if (c == 64) return FUSE(64);
  if (c == 32) return FUSE(32);
#undef FUSE
  set_error("fuse_level: unsupported channel count %d", c);
  return SPEI_ERR_ARG;
}

}  // namespace spei
