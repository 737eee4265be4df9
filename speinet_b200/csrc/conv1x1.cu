// Channel mix of the `relu(conv1x1(bicubic_x2(x)))` chains (/root/reference/model/SearchTransfer.py:70-76 SelfTransfer,
// model/speinet.py:99-100, 111-112 _decode) at LOW resolution:  y[n, o, p] = sum_c W[o, c] * x[n, c, p]   (no bias).
// A 1x1 convolution mixes channels per pixel and the bicubic resize mixes pixels per channel, so they commute; the resize +
// bias + activation half is upsample_act.cu.  fp32 FMA on the CUDA cores (the 1e-4 bar rules out a single TF32 pass and the
// whole GEMM is 0.5 GFMA at 720p: ~25 us, a third of what the fused resize that follows moves through HBM).
//
// One block = 128 consecutive pixels of one item x all COUT outputs; thread = 4 pixels x COUT/8 outputs in registers.
// The transposed weight [cin][COUT] stays in shared memory for the block's lifetime; activations stream through a
// [32 channel][128 pixel] shared tile (coalesced 16-byte loads along pixels, broadcast weight reads).
#include "spei_common.cuh"

namespace spei {

constexpr int kMixPx = 128, kMixCk = 32;

// grid: (ceil(P / 128), n)   block: 256   dynamic smem: (cin * COUT + 32 * 128) floats
template <int COUT>
__global__ void __launch_bounds__(256)
conv1x1_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int cin, int P) {
  constexpr int OPT = COUT / 8;   // outputs per thread
  extern __shared__ __align__(16) float smem_mix[];
  float* ws = smem_mix;                       // [cin][COUT]
  float* xs = smem_mix + (size_t)cin * COUT;  // [32][128]
  const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
  const int n = blockIdx.y, p0 = blockIdx.x * kMixPx;
  const float* xn = x + (size_t)n * cin * P;
  for (int e = t; e < cin * COUT; e += 256) {
    const int o = e / cin, c = e - o * cin;   // coalesced read of w[o][c], transposed store
    ws[c * COUT + o] = __ldg(w + e);
  }
  float acc[4][OPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < OPT; ++j) acc[i][j] = 0.f;
  const bool vec = (P & 3) == 0 && (((uintptr_t)x) & 15) == 0;
  for (int c0 = 0; c0 < cin; c0 += kMixCk) {
    __syncthreads();   // previous chunk consumed (and, first time, the weights are in place)
    for (int e = t; e < kMixCk * (kMixPx / 4); e += 256) {
      const int c = e / (kMixPx / 4), q4 = e - c * (kMixPx / 4), p = p0 + q4 * 4;
      const float* src = xn + (size_t)(c0 + c) * P + p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c < cin) {
        if (vec && p + 3 < P) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (p < P) v.x = __ldg(src);
          if (p + 1 < P) v.y = __ldg(src + 1);
          if (p + 2 < P) v.z = __ldg(src + 2);
          if (p + 3 < P) v.w = __ldg(src + 3);
        }
      }
      reinterpret_cast<float4*>(xs + c * kMixPx)[q4] = v;
    }
    __syncthreads();
    const int cmax = min(kMixCk, cin - c0);
#pragma unroll 8
    for (int c = 0; c < cmax; ++c) {
      const float4 xv = reinterpret_cast<const float4*>(xs + c * kMixPx)[tx];
      const float* wr = ws + (size_t)(c0 + c) * COUT + ty * OPT;
#pragma unroll
      for (int j = 0; j < OPT; ++j) {
        const float wv = wr[j];
        acc[0][j] = fmaf(xv.x, wv, acc[0][j]); acc[1][j] = fmaf(xv.y, wv, acc[1][j]);
        acc[2][j] = fmaf(xv.z, wv, acc[2][j]); acc[3][j] = fmaf(xv.w, wv, acc[3][j]);
      }
    }
  }
  const int p = p0 + tx * 4;
  float* yn = y + (size_t)n * COUT * P;
  const bool vst = (P & 3) == 0 && (((uintptr_t)y) & 15) == 0;
#pragma unroll
  for (int j = 0; j < OPT; ++j) {
    float* dst = yn + (size_t)(ty * OPT + j) * P + p;
    if (vst && p + 3 < P) *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    else {
      if (p < P) dst[0] = acc[0][j];
      if (p + 1 < P) dst[1] = acc[1][j];
      if (p + 2 < P) dst[2] = acc[2][j];
      if (p + 3 < P) dst[3] = acc[3][j];
    }
  }
}

template <int COUT>
static int launch_mix(int n, int cin, long long P, const float* x, const float* w, float* y, cudaStream_t st) {
  const int smem = (cin * COUT + kMixCk * kMixPx) * (int)sizeof(float);
  SPEI_CUDA(cudaFuncSetAttribute(conv1x1_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  conv1x1_kernel<COUT><<<dim3((unsigned)((P + kMixPx - 1) / kMixPx), n), 256, smem, st>>>(x, w, y, cin, (int)P);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_conv1x1(int n, int cin, int cout, long long P, const float* x, const float* w, float* y, cudaStream_t st) {
  if (n > 65535 || P >= (1ll << 31) || cin > 512) { set_error("conv1x1: problem too large (n=%d cin=%d P=%lld)", n, cin, P); return SPEI_ERR_ARG; }
  if (cout == 64) return launch_mix<64>(n, cin, P, x, w, y, st);
  if (cout == 32) return launch_mix<32>(n, cin, P, x, w, y, st);
  if (cout == 128) return launch_mix<128>(n, cin, P, x, w, y, st);
  if (cout == 16) return launch_mix<16>(n, cin, P, x, w, y, st);
  if (cout == 8) return launch_mix<8>(n, cin, P, x, w, y, st);
  set_error("conv1x1: output channels must be 8, 16, 32, 64 or 128 (got %d)", cout);
  return SPEI_ERR_ARG;
}

}  // namespace spei
