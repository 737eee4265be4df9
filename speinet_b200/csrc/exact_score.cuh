// Exact relevance of one (query, key) pair, shared by the rescoring kernels (rescore.cu, relevance_flagged.cu):
// R = <q patch, k patch> / (max(||q patch||, 1e-12) * max(||k patch||, 1e-12))   (/root/reference/model/SearchTransfer.py:30-33)
// from the fp32 channels-last copies of the operands, one warp per pair.
#pragma once

#include "spei_common.cuh"

namespace spei {

__device__ __forceinline__ unsigned flip_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unflip_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// (score, key) packed so that an unsigned 64-bit max picks the highest score and, on equal scores, the LOWEST key index:
// torch.max's first-index tie-break (SearchTransfer.py:34)
__device__ __forceinline__ unsigned long long pack_score(float s, int j) {
  return ((unsigned long long)flip_f32(s) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j);
}
__device__ __forceinline__ float packed_score(unsigned long long b) { return unflip_f32((unsigned)(b >> 32)); }
__device__ __forceinline__ int packed_key(unsigned long long b) { return (int)(0xffffffffu - (unsigned)(b & 0xffffffffull)); }

// The warp's query patch is read through `qtap(t)`: a pointer to the 32 float4 (128 channels) of tap t, zero for taps outside
// the image (lane l reads channels 4l .. 4l+3); kimg is the NHWC fp32 image of the key's reference frame.  Each lane owns 4
// channels of every tap: their 4 products are summed in fp32 (one rounding of ~6e-8 relative per product, ~2e-9 absolute on
// a normalised score -- four orders of magnitude inside the 1e-5 near-tie rule) and the 9 tap partials, then the 32 lanes,
// are accumulated in fp64 in a fixed order.  Returns the same value in every lane.
template <typename QTap>
__device__ __forceinline__ float exact_relevance(QTap qtap, const float* __restrict__ kimg, int hr, int wr, int Hr, int Wr, float rq,
                                                 float rk, int lane) {
  // all nine key rows are requested before the first one is used (branch-free: a tap outside the image re-reads the centre
  // row and is zeroed); the kernel is latency bound, so loads in flight are what matters
  const float4* kc = reinterpret_cast<const float4*>(kimg) + ((size_t)hr * Wr + wr) * (kC3 / 4) + lane;
  float4 kv[9];
  if ((unsigned)(hr - 1) < (unsigned)(Hr - 2) && (unsigned)(wr - 1) < (unsigned)(Wr - 2)) {
    // interior key (the whole warp works on the same key: uniform branch): nine plain loads, no bounds logic
    const int rowq = Wr * (kC3 / 4);
    const float4* k0 = kc - rowq - (kC3 / 4);
#pragma unroll
    for (int t = 0; t < 9; ++t) kv[t] = __ldg(k0 + (t / 3) * rowq + (t % 3) * (kC3 / 4));
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const bool in = (unsigned)(hr + dy) < (unsigned)Hr && (unsigned)(wr + dx) < (unsigned)Wr;
      kv[t] = __ldg(kc + (in ? (dy * Wr + dx) * (kC3 / 4) : 0));
      if (!in) kv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // three independent chains (taps t % 3), fixed order
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 qq = qtap(t)[lane];
    float part = qq.x * kv[t].x;
    part = fmaf(qq.y, kv[t].y, part);
    part = fmaf(qq.z, kv[t].z, part);
    part = fmaf(qq.w, kv[t].w, part);
    if (t % 3 == 0) acc0 += (double)part;
    else if (t % 3 == 1) acc1 += (double)part;
    else acc2 += (double)part;
  }
  double acc = (acc0 + acc1) + acc2;
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return (float)(acc * (double)rq * (double)rk);
}

// cp.async of the 9 x 128 fp32 query patch of pixel (y, x) into `dst` (one warp; zero-filled taps outside the image)
__device__ __forceinline__ void load_query_patch_async(float* dst, const float* __restrict__ qimg, int y, int x, int H, int W, int lane) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const float* src = qimg + ((size_t)(in ? yy : y) * W + (in ? xx : x)) * kC3 + lane * 4;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + t * kC3 + lane * 4);
    const int sz = in ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

}  // namespace spei
