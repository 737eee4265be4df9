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

// Finest level (S = 4, 118 MB of reference per 720p item).  With a scattered match field every reference pixel is covered by
// ~9 different gathered patches at unrelated times, and the whole level does not stay in L2: a cell-major order re-read it
// ~9x from DRAM (ncu, round 1: 1.19 GB per launch).  Here the slowest grid dimension is an 8-channel slab (29.5 MB at 720p,
// L2 resident while every query cell is processed for it), so DRAM sees each slab once; one block handles a cell row of 32
// cells so the index decode is shared by 4 x 8 x 32 output runs.
//
// Two source layouts, chosen on the device from the match field itself (no host round trip):
//  * planar (the caller's NCHW tensor; its 16-byte runs are already whole cell rows): thread = (cell, channel), 4 pixel rows.
//    On a coherent field neighbouring lanes share sectors and L1 lines: 70 us on the identity field, 125 us at +-2 cells of
//    jitter — but on a scattered field half of every 32-byte sector is the neighbouring cell's row, never used: the kernel then
//    runs at the L2 -> L1 bandwidth of the chip (2.1 GB in 208 us at 720p).
//  * cell-major copy (stage_ref_cell_kernel: the 4 x 4 pixels of a cell are 64 contiguous bytes, 32 in bf16 = one sector):
//    thread = (cell, pixel row) x 4 channels, four lanes read one cell, whole sectors, 8 instead of 32 distinct lines per warp
//    request: ~105 us on a scattered field after a ~50 us re-tiling pass (236 MB).
// stage_ref_cell_kernel samples 1024 queries first: a query is "near" when its match lies within +-2 cells of where its left
// neighbour's match would put it; the copy (and the cell-major gather) run only when fewer than half are near.
// The closed form and the order of the <= 9 additions are the same in both: bit-identical output.
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

// Every block samples the same 1024 queries of the match field (4 L2 hits per thread) and reaches the same verdict: cell-major
// when fewer than half of the samples are "near"; block 0 publishes it in mode[0] for the gather.  Blocks of a coherent field
// return at once; the others re-tile: persistent grid, block 256 = 64 cells x 4 pixel rows, work item = (plane, cell row, 64-cell
// chunk).  force: -1 = decide, 0 / 1 = planar / cell-major (A/B builds).
template <typename TIO>
__global__ void __launch_bounds__(256, 8)
stage_ref_cell_kernel(const TIO* __restrict__ x, int planes, int Hr, int Wr, TIO* __restrict__ out, const int32_t* __restrict__ arg, int n,
                      int rf, int H, int W, int* __restrict__ mode, int force) {
  __shared__ int s_near[8], s_all[8], s_mode;
  if (force < 0) {
    const int lk1 = Hr * Wr, jmax = rf * lk1 - 1;
    const unsigned total = (unsigned)n * (unsigned)H * (unsigned)W;   // (< 2^31: the launcher checks)
    constexpr int kPer = 4;
    const unsigned stride = total > 256u * kPer ? total / (256u * kPer) : 1u;
    int j1[kPer], j0[kPer];
    bool ok[kPer];
#pragma unroll
    for (int s = 0; s < kPer; ++s) {   // all loads first
      const unsigned q = (threadIdx.x + 256u * s) * stride;
      ok[s] = q < total && (q % (unsigned)W) != 0u;
      j1[s] = ok[s] ? __ldg(arg + q) : 0;
      j0[s] = ok[s] ? __ldg(arg + q - 1) : 0;
    }
    int near = 0, all = 0;
#pragma unroll
    for (int s = 0; s < kPer; ++s) {
      const int a1 = min(max(j1[s], 0), jmax), a0 = min(max(j0[s], 0), jmax);
      const int f1 = a1 / lk1, r1 = a1 - f1 * lk1, f0 = a0 / lk1, r0 = a0 - f0 * lk1;
      const int y1 = r1 / Wr, y0 = r0 / Wr;
      const int dy = y1 - y0, dx = (r1 - y1 * Wr) - (r0 - y0 * Wr) - 1;
      near += (ok[s] && f1 == f0 && dy >= -2 && dy <= 2 && dx >= -2 && dx <= 2) ? 1 : 0;
      all += ok[s] ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { near += __shfl_xor_sync(0xffffffffu, near, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    if ((threadIdx.x & 31) == 0) { s_near[threadIdx.x >> 5] = near; s_all[threadIdx.x >> 5] = all; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int a = 0, b = 0;
      for (int i = 0; i < 8; ++i) { a += s_near[i]; b += s_all[i]; }
      s_mode = (2 * a < b) ? 1 : 0;
      if (blockIdx.x == 0) { mode[0] = s_mode; mode[1] = a; mode[2] = b; }
    }
  } else if (threadIdx.x == 0) {
    s_mode = force;
    if (blockIdx.x == 0) mode[0] = force;
  }
  __syncthreads();
  if (s_mode == 0) return;   // coherent match field: the gather reads the planar input
  const int r = threadIdx.x & 3, xc = threadIdx.x >> 2;
  const int chunks = (Wr + 63) / 64;
  const long long items = (long long)planes * Hr * chunks;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int chunk = (int)(it % chunks);
    const long long py = it / chunks;
    const int Y = (int)(py % Hr);
    const size_t plane = (size_t)(py / Hr) * (size_t)(16 * Hr) * Wr;
    const int X = chunk * 64 + xc;
    if (X >= Wr) continue;
    const TIO* src = x + plane + (size_t)(4 * Y + r) * (4 * Wr) + (size_t)X * 4;
    TIO* dst = out + plane + ((size_t)Y * Wr + X) * 16 + r * 4;
    if (sizeof(TIO) == 4) *reinterpret_cast<float4*>(dst) = __ldcs(reinterpret_cast<const float4*>(src));
    else *reinterpret_cast<uint2*>(dst) = __ldcs(reinterpret_cast<const uint2*>(src));
  }
}

// grid: (ceil(W/32), H, n * C/8)   block 256
//   planar source:     thread = (cell = tid & 31, channel = tid >> 5), iterates the 4 pixel rows
//   cell-major source: thread = (cell = (warp & 3) * 8 + lane / 4, pixel row = lane & 3), iterates 4 channels ((warp >> 2) * 4 ..)
template <bool kCpuOrder, bool kTrueDiv, typename TIO>
__global__ void __launch_bounds__(256, 4)
gather_fold_lv1_kernel(const int32_t* __restrict__ arg, const TIO* __restrict__ ref, const TIO* __restrict__ refc, TIO* __restrict__ out,
                       int rf, int C, int H, int W, int Hr, int Wr, const int* __restrict__ mode) {
  constexpr int S = 4;
  __shared__ int s_src[9][32];   // per (neighbour, cell): element offset of the source run inside the item's reference, -1 = none
  const bool cells = mode != nullptr && __ldg(mode) != 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int X0 = blockIdx.x * 32, Y = blockIdx.y;
  const int slabs = C / 8, n = blockIdx.z / slabs, slab0 = (blockIdx.z - n * slabs) * 8;
  const int lk1 = Hr * Wr, jmax = rf * lk1 - 1;
  const int ref_pitch = S * Wr;
  const int32_t* a = arg + (size_t)n * H * W;
  for (int e = tid; e < 9 * 32; e += 256) {
    const int t = e >> 5, cell = e & 31;
    const int tt = kCpuOrder ? 8 - t : t;
    const int dy = tt / 3 - 1, dx = tt % 3 - 1;
    const int Xc = X0 + cell, qy = Y + dy, qx = Xc + dx;
    int o = -1;
    if (Xc < W && qy >= 0 && qy < H && qx >= 0 && qx < W) {
      int j = __ldg(a + qy * W + qx);
      j = min(max(j, 0), jmax);
      const int f = j / lk1, rem = j - f * lk1;
      const int hr = rem / Wr, wr = rem - hr * Wr;
      const int cy = Y + hr - qy, cx = Xc + wr - qx;  // source cell
      if (cy >= 0 && cy < Hr && cx >= 0 && cx < Wr)
        o = cells ? (f * C * lk1 + cy * Wr + cx) * 16 : f * C * lk1 * 16 + (cy * S) * ref_pitch + cx * S;
    }
    s_src[t][cell] = o;
  }
  __syncthreads();
  const int cell = cells ? (warp & 3) * 8 + (lane >> 2) : lane;
  const int X = X0 + cell;
  if (X >= W) return;
  const int c = slab0 + (cells ? (warp >> 2) * 4 : warp);        // (first) channel of this thread
  const int r = cells ? (lane & 3) : 0;                          // (first) pixel row of this thread
  const size_t plane = (size_t)lk1 * 16;
  const TIO* rbase = (cells ? refc + r * 4 : ref) + ((size_t)n * rf * C + c) * plane;
  const unsigned in_step = cells ? (unsigned)plane : (unsigned)ref_pitch;   // next channel / next pixel row of the source
  const TIO* base[9];
  unsigned step[9];
  bool same = true;   // all nine neighbours point at the same source cell (a locally rigid match field: the usual case on video)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int o = s_src[t][cell];
    base[t] = o >= 0 ? rbase + o : reinterpret_cast<const TIO*>(&g_zero16);   // 16 zero bytes: zero in either type
    step[t] = o >= 0 ? in_step : 0u;                                          // (stride 0 for the zero constant)
    same = same && o >= 0 && o == s_src[0][cell];
  }
  const size_t out_plane = (size_t)(S * H) * (S * W);
  TIO* obase = out + ((size_t)n * C + c) * out_plane + (size_t)(Y * S + r) * (S * W) + (size_t)X * S;
  const size_t out_step = cells ? out_plane : (size_t)(S * W);
  if (same) {   // one request instead of nine; the value is added nine times in the same order: identical sum
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = load_run4(base[0] + (size_t)i * step[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 9; ++t) vadd(acc, v[i]);
      store_run4(obase + (size_t)i * out_step, fin<kTrueDiv>(acc));
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = load_run4(base[t] + (size_t)i * step[t]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) vadd(acc, v[t]);
    store_run4(obase + (size_t)i * out_step, fin<kTrueDiv>(acc));
  }
}

// -DSPEI_GATHER_LV1=1 / =2 pin the planar / cell-major path (A/B builds); 0 = decided from the match field
#ifndef SPEI_GATHER_LV1
#define SPEI_GATHER_LV1 0
#endif
int launch_gather_fold(int n, int rf, int c, int h, int w, int hr, int wr, int scale, int fold_mode, const int32_t* arg32,
                       const void* ref, void* ref_cells, int* mode, void* out, int io_bf16, cudaStream_t st) {
  if (c % 8) { set_error("gather_fold: channels must be a multiple of 8"); return SPEI_ERR_ARG; }
  if (scale != 4) { set_error("gather_fold (planar source): only the finest level (scale 4) takes this kernel"); return SPEI_ERR_ARG; }
  if ((long long)n * (c / 8) > 65535 || h > 65535) { set_error("gather_fold: grid too large (n * c / 8 = %lld)", (long long)n * (c / 8)); return SPEI_ERR_ARG; }
  if ((long long)rf * c * hr * wr >= (1ll << 31) / 16 || (long long)n * h * w >= (1ll << 31)) { set_error("gather_fold: level exceeds 32-bit offsets"); return SPEI_ERR_ARG; }
  const bool cpu_order = (fold_mode & SPEI_FOLD_ORDER_CPU) != 0, true_div = (fold_mode & SPEI_FOLD_TRUE_DIV) != 0;
  const int force = (fold_mode & SPEI_FOLD_LV1_PLANAR) ? 0 : ((fold_mode & SPEI_FOLD_LV1_CELLS) ? 1 : (SPEI_GATHER_LV1 == 0 ? -1 : SPEI_GATHER_LV1 - 1));
  int sms = 0, dev = 0;
  SPEI_CUDA(cudaGetDevice(&dev));
  SPEI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long items = (long long)n * rf * c * hr * ((wr + 63) / 64);
  const int rgrid = (int)(items < (long long)sms * 8 ? items : (long long)sms * 8);
  // a level that fits L2 several times over (<= 32 MB: 256 x 256 clips, BSD crops) never pays for sector waste: planar, one launch
  const bool small = (long long)n * rf * c * hr * wr * 16 * (io_bf16 ? 2 : 4) <= (32ll << 20);
  if (force == 0 || (small && force < 0)) mode = nullptr;
  else if (io_bf16)
    stage_ref_cell_kernel<<<rgrid, 256, 0, st>>>((const __nv_bfloat16*)ref, n * rf * c, hr, wr, (__nv_bfloat16*)ref_cells, arg32, n, rf, h, w, mode, force);
  else
    stage_ref_cell_kernel<<<rgrid, 256, 0, st>>>((const float*)ref, n * rf * c, hr, wr, (float*)ref_cells, arg32, n, rf, h, w, mode, force);
  const dim3 sgrid((w + 31) / 32, h, n * (c / 8));
#define GFL(O_, D_)                                                                                                                \
  do {                                                                                                                             \
    if (io_bf16)                                                                                                                   \
      gather_fold_lv1_kernel<O_, D_, __nv_bfloat16><<<sgrid, 256, 0, st>>>(arg32, (const __nv_bfloat16*)ref,                       \
                                                                          (const __nv_bfloat16*)ref_cells, (__nv_bfloat16*)out,   \
                                                                          rf, c, h, w, hr, wr, mode);                             \
    else                                                                                                                           \
      gather_fold_lv1_kernel<O_, D_, float><<<sgrid, 256, 0, st>>>(arg32, (const float*)ref, (const float*)ref_cells, (float*)out, \
                                                                  rf, c, h, w, hr, wr, mode);                                     \
  } while (0)
  if (cpu_order) { if (true_div) GFL(true, true); else GFL(true, false); }
  else { if (true_div) GFL(false, true); else GFL(false, false); }
#undef GFL
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
