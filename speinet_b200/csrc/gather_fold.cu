// Kernel (c): unfold(ref) -> bis gather -> fold -> /9 of one pyramid level
// (/root/reference/model/SearchTransfer.py:36-46) in closed form:
//
//   T_s[n,c,y,x] = (1/9) * sum over the <=9 query cells q=(Y+dy, X+dx), Y=y/s, X=x/s, inside the HxW grid, of
//                  ref_s[n, f(q), c, y + (hr(q)-(Y+dy))*s, x + (wr(q)-(X+dx))*s]     (0 outside the image)
//
// with (f, hr, wr) decoded from arg[n,q].  The 1.86 GB of unfolded + gathered patches the reference
// materialises per 720p frame never exist.  One thread owns the s consecutive output pixels of one
// query cell row (a float / float2 / float4), so every neighbour contribution is one aligned vector
// load; a warp covers 32 consecutive cells = 32*s consecutive pixels, i.e. fully coalesced stores.
// The <=9 terms are added in the order of torch's col2im so the result is bit-identical to
// F.fold: ascending patch origin on CUDA (ATen/native/cuda/im2col.cuh:139-154), ascending (ki,kj)
// on CPU (ATen/native/im2col.h:131-146); then x*(1/9f) resp. x/9 (hard part 4).
#include "spei_common.cuh"

namespace spei {

template <int S> struct Vec;
template <> struct Vec<1> { using T = float; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

__device__ __forceinline__ void vadd(float& a, const float& b) { a = __fadd_rn(a, b); }
__device__ __forceinline__ void vadd(float2& a, const float2& b) { a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); }
__device__ __forceinline__ void vadd(float4& a, const float4& b) {
  a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
}
template <bool kTrueDiv> __device__ __forceinline__ float ninth(float a) {
  return kTrueDiv ? __fdiv_rn(a, 9.0f) : __fmul_rn(a, 1.0f / 9.0f);
}
template <bool D> __device__ __forceinline__ float fin(float a) { return ninth<D>(a); }
template <bool D> __device__ __forceinline__ float2 fin(float2 a) { return make_float2(ninth<D>(a.x), ninth<D>(a.y)); }
template <bool D> __device__ __forceinline__ float4 fin(float4 a) {
  return make_float4(ninth<D>(a.x), ninth<D>(a.y), ninth<D>(a.z), ninth<D>(a.w));
}
__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float2 vzero(float2) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ const float4 g_zero16 = {0.f, 0.f, 0.f, 0.f};  // source of every non-contributing neighbour

// Finest level (S = 4, 118 MB of reference per 720p item), gathered straight from the planar NCHW input (its 16-byte runs
// are already whole cell rows).  With a scattered match field every reference pixel is covered by ~9 different gathered
// patches at unrelated times, and the whole level does not stay in L2: a cell-major order re-read it ~9x from DRAM (ncu,
// round 1: 1.19 GB per launch).  Here the slowest grid dimension is an 8-channel slab (29.5 MB at 720p, L2 resident while
// every query cell is processed for it), so DRAM sees each slab once; one block handles the S sub-rows of a cell row so the
// index decode is shared by S x 8 x 32 outputs.
// TIO = float, or __nv_bfloat16 for native bf16 I/O: a cell row is then an 8-byte run; the <= 9 terms are summed in fp32
// in the same order and rounded to bf16 once at the store (= what the fp32 kernel followed by a cast produces).
__device__ __forceinline__ float4 load_run4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load_run4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void store_run4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void store_run4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  __stcs(reinterpret_cast<uint2*>(p), u);
}

// grid: (ceil(W/32), H, n * C/8)   block: (32 cells, 8 channels)
template <int S, bool kCpuOrder, bool kTrueDiv, typename TIO>
__global__ void __launch_bounds__(256)
gather_fold_slab_kernel(const int32_t* __restrict__ arg, const TIO* __restrict__ ref, TIO* __restrict__ out, int rf, int C,
                        int H, int W, int Hr, int Wr) {
  static_assert(S == 4, "one thread = the 4 pixels of a cell row");
  using V = float4;
  __shared__ long long s_off[9][32];  // per (neighbour, cell): float offset of the source run of sub-row 0, -1 = none
  const int X0 = blockIdx.x * 32, X = X0 + threadIdx.x;
  const int Y = blockIdx.y;
  const int slabs = C / 8, n = blockIdx.z / slabs, c = (blockIdx.z - n * slabs) * 8 + threadIdx.y;
  const int lk1 = Hr * Wr, jmax = rf * lk1 - 1;
  const size_t ref_plane = (size_t)(S * Hr) * (S * Wr);
  const int ref_pitch = S * Wr;
  const int32_t* a = arg + (size_t)n * H * W;
  for (int e = threadIdx.y * 32 + threadIdx.x; e < 9 * 32; e += 256) {
    const int t = e >> 5, cell = e & 31;
    const int tt = kCpuOrder ? 8 - t : t;
    const int dy = tt / 3 - 1, dx = tt % 3 - 1;
    const int Xc = X0 + cell, qy = Y + dy, qx = Xc + dx;
    long long o = -1;
    if (Xc < W && qy >= 0 && qy < H && qx >= 0 && qx < W) {
      int j = __ldg(a + qy * W + qx);
      j = min(max(j, 0), jmax);
      const int f = j / lk1, rem = j - f * lk1;
      const int hr = rem / Wr, wr = rem - hr * Wr;
      const int cy = Y + hr - qy, cx = Xc + wr - qx;  // source cell
      if (cy >= 0 && cy < Hr && cx >= 0 && cx < Wr)
        o = (long long)f * C * (long long)ref_plane + (long long)(cy * S) * ref_pitch + (long long)cx * S;
    }
    s_off[t][cell] = o;
  }
  __syncthreads();
  if (X >= W) return;
  const TIO* rbase = ref + (size_t)n * rf * C * ref_plane + (size_t)c * ref_plane;
  const TIO* base[9];
  unsigned step[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const long long o = s_off[t][threadIdx.x];
    base[t] = o >= 0 ? rbase + o : reinterpret_cast<const TIO*>(&g_zero16);   // 16 zero bytes: zero in either type
    step[t] = o >= 0 ? (unsigned)ref_pitch : 0u;   // sub-row stride (0 for the zero constant)
  }
  const size_t out_plane = (size_t)(S * H) * (S * W);
  TIO* obase = out + ((size_t)n * C + c) * out_plane + (size_t)(Y * S) * (S * W) + (size_t)X * S;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    V v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = load_run4(base[t] + (size_t)r * step[t]);
    V acc = vzero(V{});
#pragma unroll
    for (int t = 0; t < 9; ++t) vadd(acc, v[t]);
    store_run4(obase + (size_t)r * (S * W), fin<kTrueDiv>(acc));
  }
}

int launch_gather_fold(int n, int rf, int c, int h, int w, int hr, int wr, int scale, int fold_mode, const int32_t* arg32,
                       const void* ref, void* out, int io_bf16, cudaStream_t st) {
  if (c % 8) { set_error("gather_fold: channels must be a multiple of 8"); return SPEI_ERR_ARG; }
  if (scale != 4) { set_error("gather_fold (planar source): only the finest level (scale 4) takes this kernel"); return SPEI_ERR_ARG; }
  if ((long long)n * (c / 8) > 65535 || h > 65535) { set_error("gather_fold: grid too large (n * c / 8 = %lld)", (long long)n * (c / 8)); return SPEI_ERR_ARG; }
  const bool cpu_order = (fold_mode & SPEI_FOLD_ORDER_CPU) != 0, true_div = (fold_mode & SPEI_FOLD_TRUE_DIV) != 0;
  const dim3 sgrid((w + 31) / 32, h, n * (c / 8)), block(32, 8);
#define GFL(O_, D_)                                                                                                                  \
  do {                                                                                                                               \
    if (io_bf16) gather_fold_slab_kernel<4, O_, D_, __nv_bfloat16><<<sgrid, block, 0, st>>>(arg32, (const __nv_bfloat16*)ref,        \
                                                                                           (__nv_bfloat16*)out, rf, c, h, w, hr, wr); \
    else gather_fold_slab_kernel<4, O_, D_, float><<<sgrid, block, 0, st>>>(arg32, (const float*)ref, (float*)out, rf, c, h, w, hr, wr); \
  } while (0)
  if (cpu_order) { if (true_div) GFL(true, true); else GFL(true, false); }
  else { if (true_div) GFL(false, true); else GFL(false, false); }
#undef GFL
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
