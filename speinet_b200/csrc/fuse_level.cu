// Kernel (d): one fusion line of SPEINet._decode
// (/root/reference/model/speinet.py:93-94, 96-97, 108-109):
//
//     out = dec + (W . cat(dec, T) + b) * bicubic_up(S, scale)
//
// fused into one pass: the concatenation is never materialised (two row ranges of one K loop), the
// 1x1 convolution is a [C x 2C] x [2C x pixels] fp32 register-tiled GEMM, the x2 / x4 bicubic
// upsampling of the one-channel map S (align_corners=False, A=-0.75, border-clamped taps,
// torch/include/ATen/native/UpSample.h:289-300,400-423) is evaluated inline in the epilogue, and the
// multiply + residual add are applied before the single store.
#include "spei_common.cuh"

namespace spei {

constexpr int kFM = 32;   // output channels per block
constexpr int kFN = 128;  // pixels per block
constexpr int kFK = 32;   // input channels per step

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

// grid: (ceil(plane/128), C/32, n)   block: 256 = 32 pixel-quads x 8 channel-quads
__global__ void __launch_bounds__(256)
fuse_level_kernel(const float* __restrict__ dec, const float* __restrict__ tt, const float* __restrict__ S,
                  const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out, int C, int h,
                  int w, int scale) {
  __shared__ __align__(16) float Ws[kFK][kFM];
  __shared__ __align__(16) float Xs[kFK][kFN];
  const int hs = h * scale, wsz = w * scale;
  const size_t plane = (size_t)hs * wsz;
  const int n = blockIdx.z, o0 = blockIdx.y * kFM;
  const size_t p0 = (size_t)blockIdx.x * kFN;
  const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
  const float* dec_n = dec + (size_t)n * C * plane;
  const float* t_n = tt + (size_t)n * C * plane;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int K = 2 * C;
  for (int k0 = 0; k0 < K; k0 += kFK) {
    __syncthreads();
    // weight chunk, transposed: Ws[k][o] = W[o0+o][k0+k]
#pragma unroll
    for (int it = 0; it < (kFK * kFM) / 256; ++it) {
      // lanes walk the output channel so the shared-memory store is conflict-free; the strided global
      // read hits a <=128 KB matrix that stays in L1/L2
      const int e = t + 256 * it, oo = e & (kFM - 1), kk = e / kFM;
      Ws[kk][oo] = __ldg(weight + (size_t)(o0 + oo) * K + k0 + kk);
    }
    // input chunk: rows k0..k0+31 of cat(dec, T)
#pragma unroll
    for (int it = 0; it < (kFK * kFN) / 256; ++it) {
      const int e = t + 256 * it, pp = e & (kFN - 1), kk = e / kFN;
      const int ch = k0 + kk;
      const float* src = ch < C ? dec_n + (size_t)ch * plane : t_n + (size_t)(ch - C) * plane;
      Xs[kk][pp] = (p0 + pp < plane) ? __ldg(src + p0 + pp) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kFK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Ws[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // epilogue: soft-attention weight (bicubic upsampled S), bias, multiply, residual add
  const float* S_n = S + (size_t)n * h * w;
  const float rscale = 1.0f / (float)scale;
  float sw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t p = p0 + tx * 4 + j;
    sw[j] = 0.f;
    if (p < plane) {
      const int oy = (int)(p / wsz), ox = (int)(p % wsz);
      sw[j] = scale == 1 ? __ldg(S_n + (size_t)oy * w + ox) : bicubic_tap(S_n, h, w, oy, ox, rscale);
    }
  }
  const bool vec_ok = (plane & 3) == 0 && (p0 + tx * 4 + 3 < plane);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + ty * 4 + i;
    const float b = __ldg(bias + o);
    const size_t idx0 = ((size_t)n * C + o) * plane + p0 + tx * 4;
    if (vec_ok) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dec + idx0));
      float4 r;
      r.x = d.x + (acc[i][0] + b) * sw[0]; r.y = d.y + (acc[i][1] + b) * sw[1];
      r.z = d.z + (acc[i][2] + b) * sw[2]; r.w = d.w + (acc[i][3] + b) * sw[3];
      *reinterpret_cast<float4*>(out + idx0) = r;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + tx * 4 + j < plane) out[idx0 + j] = __ldg(dec + idx0 + j) + (acc[i][j] + b) * sw[j];
    }
  }
}

int launch_fuse_level(int n, int c, int h, int w, int scale, const float* dec, const float* t, const float* S,
                      const float* weight, const float* bias, float* out, cudaStream_t st) {
  const size_t plane = (size_t)h * scale * w * scale;
  if (n > 65535) { set_error("fuse_level: n too large"); return SPEI_ERR_ARG; }
  dim3 grid((unsigned)((plane + kFN - 1) / kFN), c / kFM, n);
  fuse_level_kernel<<<grid, 256, 0, st>>>(dec, t, S, weight, bias, out, c, h, w, scale);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
