// the `relu(conv1x1(bicubic_x2(x)))` chains next to the path
//   /root/reference/model/SearchTransfer.py:70-76   SelfTransfer: T_lv2 = relu(search1(up2(x))), T_lv1 = relu(search2(up2(T_lv2)))
//   /root/reference/model/speinet.py:99-100, 111-112   search_1 = relu(search1(up2(f_lv3))), search_13 = relu(search13(up2(f_v3)))
// A 1x1 convolution mixes channels per pixel and the bicubic resize mixes pixels per channel, so they commute:
//   relu(conv1x1(up2(x)) + b) == relu(up2(W . x) + b)            (F.interpolate's border-clamped taps are linear too)
// The host runs the channel mix at LOW resolution (a plain library GEMM on a quarter of the pixels) and this kernel does
// the rest in one pass: out = act(bicubic_x2(y) + bias).  The reference materialises up2(x) (4x the input) and the
// pre-activation; here HBM sees one read of y and one write of the result.
//
// F.interpolate(scale_factor=2, mode='bicubic'): align_corners=False, A=-0.75, source index (dst+0.5)/2-0.5, taps clamped
// to the border (torch/include/ATen/native/UpSample.h:289-300,400-423).  Output (2Y+a, 2X+b) reads rows Y-2+a .. Y+1+a
// with t = 0.75 (a = 0) or 0.25 (a = 1): a thread owns one low-resolution pixel and writes its 2x2 outputs from the
// 5x5 window around it, horizontal interpolation first, then vertical, exactly as upsample_bicubic2d does per pixel.
#include "spei_common.cuh"

namespace spei {

__device__ __forceinline__ float ua_cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float ua_cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// grid: (ceil(w/32), ceil(h/8), n*c)  block (32, 8)
template <bool kRelu>
__global__ void __launch_bounds__(256)
upsample2_bias_act_kernel(const float* __restrict__ y, const float* __restrict__ bias, float* __restrict__ out, int C, int h, int w) {
  const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y;
  if (X >= w || Y >= h) return;
  const int plane_id = blockIdx.z;
  const float* src = y + (size_t)plane_id * h * w;
  const float b = bias ? __ldg(bias + plane_id % C) : 0.f;
  // coefficient sets of get_cubic_upsample_coefficients for t = 0.75 and t = 0.25
  const float c75[4] = {ua_cubic2(1.75f), ua_cubic1(0.75f), ua_cubic1(0.25f), ua_cubic2(1.25f)};
  const float c25[4] = {ua_cubic2(1.25f), ua_cubic1(0.25f), ua_cubic1(0.75f), ua_cubic2(1.75f)};
  float hrow[5][2];   // horizontal interpolation of window row r for the two output columns
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int yy = min(max(Y - 2 + r, 0), h - 1);
    const float* p = src + (size_t)yy * w;
    float v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = __ldg(p + min(max(X - 2 + i, 0), w - 1));
    hrow[r][0] = v[0] * c75[0] + v[1] * c75[1] + v[2] * c75[2] + v[3] * c75[3];   // output column 2X:   cols X-2 .. X+1, t = 0.75
    hrow[r][1] = v[1] * c25[0] + v[2] * c25[1] + v[3] * c25[2] + v[4] * c25[3];   // output column 2X+1: cols X-1 .. X+2, t = 0.25
  }
  float o[2][2];
#pragma unroll
  for (int bcol = 0; bcol < 2; ++bcol) {
    o[0][bcol] = hrow[0][bcol] * c75[0] + hrow[1][bcol] * c75[1] + hrow[2][bcol] * c75[2] + hrow[3][bcol] * c75[3] + b;  // row 2Y
    o[1][bcol] = hrow[1][bcol] * c25[0] + hrow[2][bcol] * c25[1] + hrow[3][bcol] * c25[2] + hrow[4][bcol] * c25[3] + b;  // row 2Y+1
  }
  float* dst = out + (size_t)plane_id * (4 * (size_t)h * w) + (size_t)(2 * Y) * (2 * w) + 2 * X;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    float2 v = make_float2(o[a][0], o[a][1]);
    if (kRelu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
    *reinterpret_cast<float2*>(dst + (size_t)a * (2 * w)) = v;
  }
}

int launch_upsample2_bias_act(int n, int c, int h, int w, const float* y, const float* bias, int relu, float* out, cudaStream_t st) {
  const long long planes = (long long)n * c;
  if (planes > 65535) { set_error("upsample2_bias_act: n*c too large"); return SPEI_ERR_ARG; }
  if ((h + 7) / 8 > 65535) { set_error("upsample2_bias_act: h too large"); return SPEI_ERR_ARG; }
  const dim3 grid((w + 31) / 32, (h + 7) / 8, (unsigned)planes), block(32, 8);
  if (relu) upsample2_bias_act_kernel<true><<<grid, block, 0, st>>>(y, bias, out, c, h, w);
  else upsample2_bias_act_kernel<false><<<grid, block, 0, st>>>(y, bias, out, c, h, w);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
