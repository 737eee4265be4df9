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
//   * d   [img][H*W] fp32 (queries) / dkmax [item] (keys: maximum) -- ||patch - bf16(patch)||_2 / ||patch||_2, the measured
//         relative rounding residual of the patch.  By Cauchy-Schwarz the bf16 score of (query i, key j) is within
//         d_i + (1 + d_i) d_j of the exact normalised relevance: the candidate window of the tcgen05 pass and the
//         rescoring threshold are derived from these numbers instead of from an assumed worst case (certified_delta()).
//   * rkpad (keys) [img][tv*Ny][tu*8] (dense) or [img][tv*Ny][Upad] (tap-sharing, u border included)
//         -- r in tile-padded (u,v) order, NaN for padded positions so
//         that a padded key can never win a comparison in the relevance epilogue.
#include "spei_common.cuh"

namespace spei {

constexpr int kPx = 32;  // pixels (consecutive x) per block

// One operand (query set or key set) of the staging pass
struct StageOp {
  const void* x;           // [nimg][128][H][W] fp32, or bf16 when in_bf16 (native bf16 I/O: the staged bf16 operand is then exact)
  int in_bf16;
  __nv_bfloat16* bf;       // [nimg][16][Vpad][Upad][8]
  float* x32;              // [nimg][H][W][128]
  float* ss;               // [nimg][H][W] per-pixel sum of squares
  float* rs;               // [nimg][H][W] per-pixel sum of squared bf16 rounding residuals
  float* r;                // [nimg][H*W] reciprocal patch norms
  float* d;                // queries: [nimg][H*W] relative residual norm of the patch (else nullptr)
  int* dmax;               // keys: [nimg / frames] maximum relative residual norm over the item's keys, as float bits (else nullptr)
  int frames;              // images per item (1 for queries, rf for keys)
  float* rkpad;            // keys only (else nullptr): tile-padded reciprocal norms, NaN outside the image
  int nimg, H, W, orient, U, V, Upad, Vpad;
  int UT, VT, border;      // rkpad geometry
};

// grid: (ceil(maxW/32), maxH, nimg_q + nimg_k): both operands in ONE launch (the staging pass was 7 launches of
// 2-17 us each; the gaps between them were a quarter of its time)
__global__ void __launch_bounds__(256)
stage_transpose_kernel(const StageOp oq, const StageOp ok) {
  __shared__ float tile[kC3][kPx + 1];
  __shared__ float part[8][kPx];
  __shared__ float partr[8][kPx];
  const bool is_k = (int)blockIdx.z >= oq.nimg;
  const StageOp& o = is_k ? ok : oq;
  const int img = is_k ? blockIdx.z - oq.nimg : blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * kPx;
  const int H = o.H, W = o.W;
  if (y >= H || x0 >= W) return;
  const float* __restrict__ x = reinterpret_cast<const float*>(o.x);
  const __nv_bfloat16* __restrict__ xb = reinterpret_cast<const __nv_bfloat16*>(o.x);
  __nv_bfloat16* __restrict__ bf = o.bf;
  float* __restrict__ x32 = o.x32;
  float* __restrict__ ss = o.ss;
  const int orient = o.orient, Upad = o.Upad, Vpad = o.Vpad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  const size_t soff = (size_t)img * kC3 * plane + (size_t)y * W + x0;
  const bool in = (x0 + lane) < W;
  if (o.in_bf16) {
#pragma unroll 4
    for (int c = warp; c < kC3; c += 8) tile[c][lane] = in ? __bfloat162float(xb[soff + (size_t)c * plane + lane]) : 0.f;
  } else {
#pragma unroll 4
    for (int c = warp; c < kC3; c += 8) tile[c][lane] = in ? __ldg(x + soff + (size_t)c * plane + lane) : 0.f;
  }
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
      const size_t off = ((((size_t)img * kCG + cg) * Vpad + (vv + 1)) * Upad + (u + 1)) * 8;
      *reinterpret_cast<uint4*>(bf + off) = *reinterpret_cast<const uint4*>(v);
    }
  }
  // per-pixel sum of squares over the 128 channels (fixed order: 8 partials of 16 channels)
  {
    const int px = lane, pt = warp;
    float s = 0.f, sr = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = tile[pt * 16 + i][px];
      s = fmaf(a, a, s);
      const float res = a - __bfloat162float(__float2bfloat16_rn(a));   // exact: the operand the tensor cores see differs by this
      sr = fmaf(res, res, sr);
    }
    part[pt][px] = s;
    partr[pt][px] = sr;
  }
  __syncthreads();
  if (warp == 0 && in) {
    float s = 0.f, sr = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += part[i][lane]; sr += partr[i][lane]; }
    ss[(size_t)img * plane + (size_t)y * W + x0 + lane] = s;
    o.rs[(size_t)img * plane + (size_t)y * W + x0 + lane] = sr;
  }
}

__device__ __forceinline__ float patch_sum(const float* __restrict__ ss, int H, int W, int y, int x) {
  float s = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += __ldg(ss + (size_t)yy * W + xx);
    }
  return s;
}
__device__ __forceinline__ float patch_rnorm(const float* __restrict__ ss, int H, int W, int y, int x) {
  return 1.0f / fmaxf(sqrtf(patch_sum(ss, H, W, y, x)), 1e-12f);  // F.normalize: v / max(||v||, eps)
}

// rq, rk and the tile-padded key norms in ONE launch: thread i covers [q positions | k positions | padded k positions]
__global__ void __launch_bounds__(256)
patch_norms_kernel(const StageOp oq, const StageOp ok) {
  const size_t nq = (size_t)oq.nimg * oq.H * oq.W, nk = (size_t)ok.nimg * ok.H * ok.W;
  const size_t per = (size_t)ok.UT * ok.VT, np = ok.rkpad ? per * ok.nimg : 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq + nk) {
    const StageOp& o = i < nq ? oq : ok;
    if (i >= nq) i -= nq;
    const size_t plane = (size_t)o.H * o.W;
    const int img = (int)(i / plane), rem = (int)(i % plane);
    const float s = patch_sum(o.ss + (size_t)img * plane, o.H, o.W, rem / o.W, rem % o.W);
    // F.normalize: v / max(||v||, 1e-12).  An all-zero QUERY patch is marked with +inf (the rescoring short-circuits it to
    // S = 0, arg = 0 -- every relevance is 0 and torch.max returns the first index; nothing else multiplies by it)
    o.r[i] = (s == 0.f && o.dmax == nullptr) ? INFINITY : 1.0f / fmaxf(sqrtf(s), 1e-12f);
    // relative rounding residual of the patch (the 1 % factor of certified_delta() covers the fp32 sums, sqrt and division)
    const float sr = patch_sum(o.rs + (size_t)img * plane, o.H, o.W, rem / o.W, rem % o.W);
    const float dl = s > 0.f ? fminf(sqrtf(sr / s), 1.f) : 0.f;
    if (o.d) o.d[i] = dl;
    if (o.dmax) {
      // per-item maximum: one atomic per run of lanes that share the item (non-negative floats order like their bit patterns)
      const int item = img / o.frames;
      const unsigned peers = __match_any_sync(__activemask(), item);
      const int mx = __reduce_max_sync(peers, __float_as_int(dl));
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicMax(o.dmax + item, mx);
    }
    return;
  }
  i -= nq + nk;
  if (i >= np) return;
  const int img = (int)(i / per), rem = (int)(i % per);
  const int v = rem / ok.UT, u = rem % ok.UT - ok.border;
  const int x = ok.orient == 0 ? u : v, y = ok.orient == 0 ? v : u;
  float out = __int_as_float(0x7fc00000);  // NaN: padded keys never compare greater
  if (u >= 0 && x < ok.W && y < ok.H) out = patch_rnorm(ok.ss + (size_t)img * ok.H * ok.W, ok.H, ok.W, y, x);
  ok.rkpad[i] = out;
}

// Zeroes only what stage_transpose_kernel does not write: the 1-position border and the tile padding of the
// [Vpad][Upad] planes of both operands (1-5 % of the buffers; a full cudaMemset cost ~10 us per call at 720p).
// Threads enumerate the padding positions only: first the full rows v = 0 and v > V, then the columns u = 0 and u > U
// of the interior rows.   grid: (ceil(max padding positions / 256), 1, 2 operands)
__global__ void __launch_bounds__(256)
zero_padding_kernel(const StageOp oq, const StageOp ok) {
  const StageOp& o = blockIdx.z ? ok : oq;
  if (blockIdx.z == 1 && blockIdx.x == 0)   // per-item key maxima start at 0 (patch_norms_kernel runs two launches later)
    for (int i = threadIdx.x; i < ok.nimg / ok.frames; i += blockDim.x) ok.dmax[i] = 0;
  const int full_rows = o.Vpad - o.V, side_cols = o.Upad - o.U;
  const int n_full = full_rows * o.Upad, n_side = o.V * side_cols;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_full + n_side) return;
  int u, v;
  if (i < n_full) {
    const int row = i / o.Upad;
    u = i - row * o.Upad;
    v = row == 0 ? 0 : o.V + row;
  } else {
    const int j = i - n_full, row = j / side_cols, c = j - row * side_cols;
    v = 1 + row;
    u = c == 0 ? 0 : o.U + c;
  }
  const size_t per = (size_t)o.Upad * o.Vpad, pos = (size_t)v * o.Upad + u;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int pl = 0; pl < o.nimg * kCG; ++pl) *reinterpret_cast<uint4*>(o.bf + (pl * per + pos) * 8) = z;
}

static StageOp make_op(const void* x, int in_bf16, int nimg, int H, int W, const OperandPlan& o, __nv_bfloat16* bf, float* x32, float* ss, float* rs,
                       float* r, float* rkpad, float* d, int* dmax, int frames) {
  StageOp s{};
  s.x = x; s.in_bf16 = in_bf16; s.bf = bf; s.x32 = x32; s.ss = ss; s.rs = rs; s.r = r; s.rkpad = rkpad; s.d = d; s.dmax = dmax; s.frames = frames;
  s.nimg = nimg; s.H = H; s.W = W; s.orient = o.orient; s.U = o.U; s.V = o.V; s.Upad = o.Upad; s.Vpad = o.Vpad;
  // dense tiling: [tv*Ny][tu*8]; tap-sharing tiling: [tv*Ny][Upad] with the staged image's 1-position u border
  const bool shared = o.tile_u == kSTileU;
  s.UT = shared ? o.Upad : o.tu * kTileU; s.VT = o.tv * o.tile_v; s.border = shared ? 1 : 0;
  return s;
}

int launch_stage_norm(const Plan& p, const void* q, const void* k, char* ws, cudaStream_t st) {
  const StageOp oq = make_op(q, p.io_bf16, p.n, p.H, p.W, p.q, (__nv_bfloat16*)(ws + p.off_qbf), (float*)(ws + p.off_q32),
                             (float*)(ws + p.off_qss), (float*)(ws + p.off_qrs), (float*)(ws + p.off_rq), nullptr,
                             (float*)(ws + p.off_dq), nullptr, 1);
  const StageOp ok = make_op(k, p.io_bf16, p.n * p.rf, p.Hr, p.Wr, p.k, (__nv_bfloat16*)(ws + p.off_kbf), (float*)(ws + p.off_k32),
                             (float*)(ws + p.off_kss), (float*)(ws + p.off_krs), (float*)(ws + p.off_rk), (float*)(ws + p.off_rkpad),
                             nullptr, (int*)(ws + p.off_dkmax), p.rf);
  if ((long long)oq.nimg + ok.nimg > 65535) { set_error("stage_norm: too many images"); return SPEI_ERR_ARG; }
  const int padq = (oq.Vpad - oq.V) * oq.Upad + oq.V * (oq.Upad - oq.U), padk = (ok.Vpad - ok.V) * ok.Upad + ok.V * (ok.Upad - ok.U);
  zero_padding_kernel<<<dim3(((padq > padk ? padq : padk) + 255) / 256, 1, 2), 256, 0, st>>>(oq, ok);
  SPEI_CUDA(cudaGetLastError());
  const int maxW = oq.W > ok.W ? oq.W : ok.W, maxH = oq.H > ok.H ? oq.H : ok.H;
  stage_transpose_kernel<<<dim3((maxW + kPx - 1) / kPx, maxH, oq.nimg + ok.nimg), 256, 0, st>>>(oq, ok);
  SPEI_CUDA(cudaGetLastError());
  const size_t tot = (size_t)oq.nimg * oq.H * oq.W + (size_t)ok.nimg * ok.H * ok.W + (size_t)ok.nimg * ok.UT * ok.VT;
  patch_norms_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(oq, ok);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
