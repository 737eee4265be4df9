// Kernel (a): the unfold + normalize pre-pass of SearchTransfer.forward
// (/root/reference/model/SearchTransfer.py:26-31) WITHOUT materialising the unfolded tensors.
//
// Per operand (query set / key set) it writes, from the NCHW fp32 input:
//   * bf  [img][16][Vpad][Upad][8] bf16  -- channel-group-planar image with a 1-pixel zero border,
//         (u,v) orientation chosen by the plan.  A 3x3 patch shift is a +-16-byte / +-row offset in
//         this layout, which is what lets the tcgen05 kernel read all 9 shifts of a tile from one
//         halo tile in shared memory.
//   * x32 [img][H][W][128] fp32 (NHWC)   -- read by the exact fp32 rescoring with 512-byte rows.
//   * r   [img][H*W] fp32                -- 1 / max(||3x3x128 patch||_2, 1e-12)  (F.normalize, :30-31)
//   * rkpad (keys) [img][tv*Ny][tu*8] (dense) or [img][tv*Ny][Upad] (tap-sharing, u border included)
//         -- r in tile-padded (u,v) order, NaN for padded positions so
//         that a padded key can never win a comparison in the relevance epilogue.
#include "spei_common.cuh"

namespace spei {

constexpr int kPx = 32;  // pixels (consecutive x) per block

__global__ void __launch_bounds__(256)
stage_transpose_kernel(const float* __restrict__ x, int H, int W, int orient, int Upad, int Vpad,
                       __nv_bfloat16* __restrict__ bf, float* __restrict__ x32, float* __restrict__ ss) {
  __shared__ float tile[kC3][kPx + 1];
  __shared__ float part[8][kPx];
  const int img = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * kPx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  const float* src = x + (size_t)img * kC3 * plane + (size_t)y * W + x0;
  const bool in = (x0 + lane) < W;
#pragma unroll 4
  for (int c = warp; c < kC3; c += 8) tile[c][lane] = in ? __ldg(src + (size_t)c * plane + lane) : 0.f;
  __syncthreads();

  // NHWC fp32 copy: 128 consecutive floats per pixel
  float* dst32 = x32 + ((size_t)img * plane + (size_t)y * W + x0) * kC3;
  for (int px = warp; px < kPx; px += 8) {
    if (x0 + px < W) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dst32[(size_t)px * kC3 + lane + 32 * j] = tile[lane + 32 * j][px];
    }
  }
  // bf16 channel-group-planar copy, 16 bytes per (channel group, pixel)
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = threadIdx.x + 256 * it, px = item & 31, cg = item >> 5;
    if (x0 + px < W) {
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __float2bfloat16_rn(tile[cg * 8 + i][px]);
      const int xx = x0 + px;
      const int u = orient == 0 ? xx : y, vv = orient == 0 ? y : xx;
      const size_t o = ((((size_t)img * kCG + cg) * Vpad + (vv + 1)) * Upad + (u + 1)) * 8;
      *reinterpret_cast<uint4*>(bf + o) = *reinterpret_cast<const uint4*>(v);
    }
  }
  // per-pixel sum of squares over the 128 channels (fixed order: 8 partials of 16 channels)
  {
    const int px = lane, pt = warp;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float a = tile[pt * 16 + i][px]; s = fmaf(a, a, s); }
    part[pt][px] = s;
  }
  __syncthreads();
  if (warp == 0 && in) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    ss[(size_t)img * plane + (size_t)y * W + x0 + lane] = s;
  }
}

__device__ __forceinline__ float patch_rnorm(const float* __restrict__ ss, int H, int W, int y, int x) {
  float s = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += __ldg(ss + (size_t)yy * W + xx);
    }
  return 1.0f / fmaxf(sqrtf(s), 1e-12f);  // F.normalize: v / max(||v||, eps)
}

__global__ void __launch_bounds__(256)
patch_norm_kernel(const float* __restrict__ ss, int nimg, int H, int W, float* __restrict__ r) {
  const size_t plane = (size_t)H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane * nimg) return;
  const int img = (int)(i / plane), rem = (int)(i % plane);
  r[i] = patch_rnorm(ss + (size_t)img * plane, H, W, rem / W, rem % W);
}

__global__ void __launch_bounds__(256)
key_norm_padded_kernel(const float* __restrict__ ss, int nimg, int H, int W, int orient, int UT, int VT, int border,
                       float* __restrict__ rkpad) {
  const size_t per = (size_t)UT * VT;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per * nimg) return;
  const int img = (int)(i / per), rem = (int)(i % per);
  const int v = rem / UT, u = rem % UT - border;
  const int x = orient == 0 ? u : v, y = orient == 0 ? v : u;
  float out = __int_as_float(0x7fc00000);  // NaN: padded keys never compare greater
  if (u >= 0 && x < W && y < H) out = patch_rnorm(ss + (size_t)img * H * W, H, W, y, x);
  rkpad[i] = out;
}

// Zeroes only what stage_transpose_kernel does not write: the 1-position border and the tile padding of the
// [Vpad][Upad] planes (3-5 % of the buffer; a full cudaMemset of both operands cost ~10 us per call at 720p).
__global__ void __launch_bounds__(256)
zero_padding_kernel(__nv_bfloat16* __restrict__ bf, int nimg, int U, int V, int Upad, int Vpad) {
  const int per = Upad * Vpad;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int v = i / Upad, u = i - v * Upad;
  if (u >= 1 && u <= U && v >= 1 && v <= V) return;   // interior: written by the transpose kernel
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int pl = blockIdx.y; pl < nimg * kCG; pl += gridDim.y)
    *reinterpret_cast<uint4*>(bf + ((size_t)pl * per + i) * 8) = z;
}

static int stage_operand(const float* x, int nimg, int H, int W, const OperandPlan& o, __nv_bfloat16* bf, float* x32,
                         float* ss, float* r, float* rkpad, cudaStream_t st) {
  {
    const int per = o.Upad * o.Vpad;
    const int planes = nimg * kCG;
    zero_padding_kernel<<<dim3((per + 255) / 256, planes < 64 ? planes : 64), 256, 0, st>>>(bf, nimg, o.U, o.V, o.Upad, o.Vpad);
    SPEI_CUDA(cudaGetLastError());
  }
  dim3 grid((W + kPx - 1) / kPx, H, nimg);
  stage_transpose_kernel<<<grid, 256, 0, st>>>(x, H, W, o.orient, o.Upad, o.Vpad, bf, x32, ss);
  SPEI_CUDA(cudaGetLastError());
  const size_t tot = (size_t)nimg * H * W;
  patch_norm_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ss, nimg, H, W, r);
  SPEI_CUDA(cudaGetLastError());
  if (rkpad) {
    // dense tiling: [tv*Ny][tu*8]; tap-sharing tiling: [tv*Ny][Upad] with the staged image's 1-position u border
    const bool shared = o.tile_u == kSTileU;
    const int UT = shared ? o.Upad : o.tu * kTileU, VT = o.tv * o.tile_v;
    const size_t totp = (size_t)nimg * UT * VT;
    key_norm_padded_kernel<<<(unsigned)((totp + 255) / 256), 256, 0, st>>>(ss, nimg, H, W, o.orient, UT, VT, shared ? 1 : 0, rkpad);
    SPEI_CUDA(cudaGetLastError());
  }
  return SPEI_OK;
}

int launch_stage_norm(const Plan& p, const float* q, const float* k, char* ws, cudaStream_t st) {
  int rc = stage_operand(q, p.n, p.H, p.W, p.q, (__nv_bfloat16*)(ws + p.off_qbf), (float*)(ws + p.off_q32),
                         (float*)(ws + p.off_qss), (float*)(ws + p.off_rq), nullptr, st);
  if (rc) return rc;
  return stage_operand(k, p.n * p.rf, p.Hr, p.Wr, p.k, (__nv_bfloat16*)(ws + p.off_kbf), (float*)(ws + p.off_k32),
                       (float*)(ws + p.off_kss), (float*)(ws + p.off_rk), (float*)(ws + p.off_rkpad), st);
}

}  // namespace spei
