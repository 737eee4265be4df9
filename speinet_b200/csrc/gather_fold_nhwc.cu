// Kernel (c), channels-last variant: the same closed form as gather_fold.cu, but the gathered
// reference is read from an NHWC copy so that every (query cell, neighbour) read is one 512-byte
// contiguous run (all sectors useful) instead of 32 scattered 4/8/16-byte pieces.  With a random match
// field the NCHW gather moves ~8x / 4x / 2x more L2 sectors than it uses at lv3 / lv2 / lv1; this
// variant moves exactly the 9 x output bytes the algorithm needs.  The output stays NCHW: results are
// transposed through shared memory and written as full rows.
//
//   stage_ref_nhwc_kernel   NCHW -> NHWC copy of one pyramid level (lv3 reuses the staged k32 copy
//                           when ref_lv3 aliases refsr_lv3, as in speinet.py:135)
//   gather_fold_nhwc_kernel one block = 32 consecutive query cells of one cell row; one warp iteration =
//                           one (cell, pixel row): lanes cover (pixel-in-cell, channel quad) = S * C/4 = 32.
#include "spei_common.cuh"

namespace spei {

// grid: (ceil(Ws/32), Hs, nimg)  block 256.  TIn = float or __nv_bfloat16 (native bf16 I/O; the copy is fp32 either way)
template <typename TIn>
__global__ void __launch_bounds__(256)
stage_ref_nhwc_kernel(const TIn* __restrict__ x, int C, int Hs, int Ws, float* __restrict__ out) {
  extern __shared__ float tile[];  // [C][33]
  const int img = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)Hs * Ws;
  const TIn* src = x + (size_t)img * C * plane + (size_t)y * Ws + x0;
  const bool in = (x0 + lane) < Ws;
  for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = in ? (float)src[(size_t)c * plane + lane] : 0.f;
  __syncthreads();
  float* dst = out + ((size_t)img * plane + (size_t)y * Ws + x0) * C;
  const int c4n = C >> 2;
  for (int e = threadIdx.x; e < 32 * c4n; e += 256) {
    const int px = e / c4n, c4 = e - px * c4n;
    if (x0 + px < Ws) {
      const float4 v = make_float4(tile[(c4 * 4 + 0) * 33 + px], tile[(c4 * 4 + 1) * 33 + px], tile[(c4 * 4 + 2) * 33 + px],
                                   tile[(c4 * 4 + 3) * 33 + px]);
      *reinterpret_cast<float4*>(dst + (size_t)px * C + c4 * 4) = v;
    }
  }
}

int launch_stage_ref_nhwc(const void* ref, int in_bf16, int nimg, int C, int Hs, int Ws, float* dst, cudaStream_t st) {
  if (nimg > 65535 || Hs > 65535) { set_error("stage_ref_nhwc: grid too large"); return SPEI_ERR_ARG; }
  const dim3 grid((Ws + 31) / 32, Hs, nimg);
  if (in_bf16) stage_ref_nhwc_kernel<<<grid, 256, C * 33 * sizeof(float), st>>>((const __nv_bfloat16*)ref, C, Hs, Ws, dst);
  else stage_ref_nhwc_kernel<<<grid, 256, C * 33 * sizeof(float), st>>>((const float*)ref, C, Hs, Ws, dst);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

__device__ const float4 g_zero16n = {0.f, 0.f, 0.f, 0.f};  // source of every non-contributing neighbour

__device__ __forceinline__ void store_out(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <bool kTrueDiv> __device__ __forceinline__ float ninth_n(float a) {
  return kTrueDiv ? __fdiv_rn(a, 9.0f) : __fmul_rn(a, 1.0f / 9.0f);
}

// grid: (ceil(W/32), H, n)  block 256; dynamic smem C * S * (32*S + 1) floats
template <int S, int C, bool kCpuOrder, bool kTrueDiv, typename TOut>
__global__ void __launch_bounds__(256)
gather_fold_nhwc_kernel(const int32_t* __restrict__ arg, const float* __restrict__ ref, TOut* __restrict__ out, int rf, int H,
                        int W, int Hr, int Wr) {
  static_assert(S * (C / 4) == 32, "one warp covers one pixel row of one cell");
  constexpr int kPitch = 32 * S + 1;
  extern __shared__ float tile[];  // [C][S][kPitch]
  const int X0 = blockIdx.x * 32, Y = blockIdx.y, n = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rx = lane / (C / 4), c4 = lane % (C / 4);
  const int lk1 = Hr * Wr, jmax = rf * lk1 - 1;
  const size_t ref_pix = (size_t)(S * Hr) * (S * Wr);  // pixels per reference frame
  const int ref_pitch = S * Wr;
  const int32_t* a = arg + (size_t)n * H * W;
  const float* rbase = ref + (size_t)n * rf * ref_pix * C;

  for (int ci = warp; ci < 32; ci += 8) {   // 4 cells per warp
    const int X = X0 + ci;
    if (X >= W) break;
    // decode the <= 9 neighbours of this cell: lane t decodes neighbour t, then broadcast
    long long mine = -1;
    if (lane < 9) {
      const int tt = kCpuOrder ? 8 - lane : lane;
      const int dy = tt / 3 - 1, dx = tt % 3 - 1;
      const int qy = Y + dy, qx = X + dx;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        int j = __ldg(a + qy * W + qx);
        j = min(max(j, 0), jmax);
        const int f = j / lk1, rem = j - f * lk1;
        const int hr = rem / Wr, wr = rem - hr * Wr;
        const int cy = Y + hr - qy, cx = X + wr - qx;  // source cell
        if (cy >= 0 && cy < Hr && cx >= 0 && cx < Wr)
          mine = ((long long)f * ref_pix + (long long)(cy * S) * ref_pitch + (long long)cx * S) * C;
      }
    }
    // branch-free: a neighbour without a contribution reads a zero constant with row stride 0
    const float4* base[9];
    unsigned step[9];  // float4 stride between the S pixel rows of the run
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const long long o = __shfl_sync(0xffffffffu, mine, t);
      base[t] = o >= 0 ? reinterpret_cast<const float4*>(rbase + o + (long long)rx * C) + c4 : &g_zero16n;
      step[t] = o >= 0 ? (unsigned)(ref_pitch * (C / 4)) : 0u;
    }
#pragma unroll
    for (int ry = 0; ry < S; ++ry) {
      float4 v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = __ldg(base[t] + (size_t)ry * step[t]);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        acc.x = __fadd_rn(acc.x, v[t].x); acc.y = __fadd_rn(acc.y, v[t].y);
        acc.z = __fadd_rn(acc.z, v[t].z); acc.w = __fadd_rn(acc.w, v[t].w);
      }
      const int xt = ci * S + rx;
      float* trow = tile + ((size_t)(c4 * 4) * S + ry) * kPitch + xt;
      trow[0] = ninth_n<kTrueDiv>(acc.x);
      trow[(size_t)S * kPitch] = ninth_n<kTrueDiv>(acc.y);
      trow[(size_t)2 * S * kPitch] = ninth_n<kTrueDiv>(acc.z);
      trow[(size_t)3 * S * kPitch] = ninth_n<kTrueDiv>(acc.w);
    }
  }
  __syncthreads();
  // write out: rows of 32*S consecutive pixels per (channel, pixel row)
  const int wpx = min(32, W - X0) * S;  // valid pixels in this tile row
  const size_t out_plane = (size_t)(S * H) * (S * W);
  TOut* obase = out + (size_t)n * C * out_plane + (size_t)(Y * S) * (S * W) + (size_t)X0 * S;
  for (int r = warp; r < C * S; r += 8) {
    const int c = r / S, ry = r - c * S;
    const float* trow = tile + (size_t)r * kPitch;
    TOut* orow = obase + (size_t)c * out_plane + (size_t)ry * (S * W);
    for (int xx = lane; xx < wpx; xx += 32) store_out(orow + xx, trow[xx]);   // fp32: streaming store; bf16: one rounding of the fp32 sum
  }
}

template <int S, int C, typename TOut>
static int launch_gf_nhwc_t(int n, int rf, int h, int w, int hr, int wr, int fold_mode, const int32_t* arg32, const float* ref,
                            TOut* out, cudaStream_t st) {
  const int smem = C * S * (32 * S + 1) * (int)sizeof(float);
  const bool cpu_order = (fold_mode & SPEI_FOLD_ORDER_CPU) != 0, true_div = (fold_mode & SPEI_FOLD_TRUE_DIV) != 0;
  dim3 grid((w + 31) / 32, h, n);
#define GFN(O_, D_)                                                                                                   \
  do {                                                                                                                \
    SPEI_CUDA(cudaFuncSetAttribute(gather_fold_nhwc_kernel<S, C, O_, D_, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    gather_fold_nhwc_kernel<S, C, O_, D_, TOut><<<grid, 256, smem, st>>>(arg32, ref, out, rf, h, w, hr, wr);         \
  } while (0)
  if (cpu_order) { if (true_div) GFN(true, true); else GFN(true, false); }
  else { if (true_div) GFN(false, true); else GFN(false, false); }
#undef GFN
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_gather_fold_nhwc(int n, int rf, int c, int h, int w, int hr, int wr, int scale, int fold_mode, const int32_t* arg32,
                            const float* ref_nhwc, void* out, int out_bf16, cudaStream_t st) {
  if (n > 65535 || h > 65535) { set_error("gather_fold: grid too large"); return SPEI_ERR_ARG; }
  if (out_bf16) {
    if (scale == 1 && c == 128) return launch_gf_nhwc_t<1, 128>(n, rf, h, w, hr, wr, fold_mode, arg32, ref_nhwc, (__nv_bfloat16*)out, st);
    if (scale == 2 && c == 64) return launch_gf_nhwc_t<2, 64>(n, rf, h, w, hr, wr, fold_mode, arg32, ref_nhwc, (__nv_bfloat16*)out, st);
  } else {
    if (scale == 1 && c == 128) return launch_gf_nhwc_t<1, 128>(n, rf, h, w, hr, wr, fold_mode, arg32, ref_nhwc, (float*)out, st);
    if (scale == 2 && c == 64) return launch_gf_nhwc_t<2, 64>(n, rf, h, w, hr, wr, fold_mode, arg32, ref_nhwc, (float*)out, st);
  }
  set_error("gather_fold (channels-last): unsupported scale/channels %d/%d", scale, c);
  return SPEI_ERR_ARG;
}

}  // namespace spei
