// the Richardson-Lucy "edge prior" of the reference
// (/root/reference/model/rcl.py:22-51, r_l_per_channel; call sites speinet.py:81 with 1 iteration and
// speinet.py:129,141 with 5 iterations) as ONE kernel.
//
// Per channel and iteration the reference launches conv2d(5x5) -> div -> two masked assignments -> conv2d(3x3
// Laplacian) -> axpy -> mul on full-size planes (about 8 launches x 3 channels x 5 iterations per frame, every
// intermediate through HBM).  Here a block keeps a (64 + 2R) x (32 + 2R) window of the frame in shared memory,
// R = iterations * (ks / 2), and runs ALL iterations on it (temporal blocking: the region that is still exact
// shrinks by ks/2 per iteration and ends as the 64 x 32 output tile), so HBM sees one read of the frame and one
// write of the result.  Positions outside the image stay 0 in every iteration = the zero padding F.conv2d applies
// to each iteration's input.
//
//   blurred = sum_{ky,kx} k[ky][kx] * d[y+ky-p][x+kx-p]          (cross-correlation, ascending (ky,kx), fp32 FMA)
//   cf      = x / blurred ;  NaN -> 0 ;  cf < 0 -> 0              (rcl.py:38-40; +inf is kept, as in the reference)
//   reg     = d + lambda * (4 d - up - left - right - down)        (rcl.py:43)
//   d       = cf * reg                                             (rcl.py:46)
#include "spei_common.cuh"

namespace spei {

constexpr int kRlTX = 64, kRlTY = 32, kRlThreads = 256;

constexpr int kRlVec = 4;  // consecutive ROWS per thread and iteration: a register tile of (4 + KS - 1) x KS loads feeds
                           // 4 x KS x KS taps, and consecutive lanes read consecutive x (conflict-free scalar LDS)

__host__ __device__ inline int rl_rows(int sh) { return sh + kRlVec; }   // row slack: the last strip may read past the region

template <int KS>
__global__ void __launch_bounds__(kRlThreads)
rl_deconv_kernel(const float* __restrict__ img, const float* __restrict__ kern, float* __restrict__ out, int H, int W,
                 int iters, float lambda) {
  constexpr int P = KS / 2;
  constexpr int NV = kRlVec + KS - 1;   // window rows a thread reads
  extern __shared__ float rl_smem[];
  const int R = iters * P;
  const int SW = kRlTX + 2 * R, SH = kRlTY + 2 * R;
  const int SN = SW * rl_rows(SH);
  float* sx = rl_smem;            // the observed frame x (never changes)
  float* da = sx + SN;            // current estimate d
  float* db = da + SN;            // next estimate
  const size_t plane = (size_t)H * W;
  const float* src = img + (size_t)blockIdx.z * plane;
  const int x0 = blockIdx.x * kRlTX - R, y0 = blockIdx.y * kRlTY - R;

  float wk[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wk[i] = __ldg(kern + i);

  for (int i = threadIdx.x; i < SN; i += kRlThreads) {
    const int sy = i / SW, sxx = i - sy * SW;
    const int gy = y0 + sy, gx = x0 + sxx;
    const float v = (sy < SH && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(src + (size_t)gy * W + gx) : 0.f;
    sx[i] = v;
    da[i] = v;    // deblurred_channel = channel_tensor.clone() (rcl.py:29)
    db[i] = 0.f;
  }
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    const int m = P * (it + 1);                 // margin of the region that is exact after this iteration
    const int RW = SW - 2 * m, RH = SH - 2 * m;
    const int strips = (RH + kRlVec - 1) / kRlVec;
    for (int i = threadIdx.x; i < strips * RW; i += kRlThreads) {
      const int rs = i / RW, rx = i - rs * RW;
      const int sy0 = rs * kRlVec + m, sxx = rx + m;
      const int gy0 = y0 + sy0, gx = x0 + sxx;
      const float* c = da + sy0 * SW + sxx;
      float blurred[kRlVec], ctr[kRlVec + 2], lft[kRlVec], rgt[kRlVec];
#pragma unroll
      for (int j = 0; j < kRlVec; ++j) blurred[j] = 0.f;
#pragma unroll
      for (int e = 0; e < NV; ++e) {            // window row e - P relative to the strip's first row
        float v[KS];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) v[kx] = c[(e - P) * SW + (kx - P)];
#pragma unroll
        for (int j = 0; j < kRlVec; ++j) {
          const int ky = e - j;                 // this window row is tap row ky of output row j
          if (ky >= 0 && ky < KS) {
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) blurred[j] = fmaf(v[kx], wk[ky * KS + kx], blurred[j]);
          }
        }
        if (e >= P - 1 && e <= P + kRlVec) ctr[e - (P - 1)] = v[P];       // rows -1 .. kRlVec at the centre column
        if (e >= P && e < P + kRlVec) { lft[e - P] = v[P - 1]; rgt[e - P] = v[P + 1]; }
      }
#pragma unroll
      for (int j = 0; j < kRlVec; ++j) {
        const int sy = sy0 + j, gy = gy0 + j;
        if (sy < SH - m) {
          float dn = 0.f;                          // outside the image: conv2d's zero padding of the next iteration
          if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const float d0 = ctr[j + 1];
            const float lap = 4.f * d0 - ctr[j] - lft[j] - rgt[j] - ctr[j + 2];
            float cf = __fdiv_rn(sx[sy * SW + sxx], blurred[j]);
            if (cf != cf) cf = 0.f;                // rcl.py:39
            if (cf < 0.f) cf = 0.f;                // rcl.py:40
            dn = cf * fmaf(lambda, lap, d0);
          }
          db[sy * SW + sxx] = dn;
        }
      }
    }
    __syncthreads();
    float* t = da; da = db; db = t;
  }

  float* dst = out + (size_t)blockIdx.z * plane;
  for (int i = threadIdx.x; i < kRlTX * kRlTY; i += kRlThreads) {
    const int ty = i / kRlTX, tx = i - ty * kRlTX;
    const int gy = blockIdx.y * kRlTY + ty, gx = blockIdx.x * kRlTX + tx;
    if (gy < H && gx < W) dst[(size_t)gy * W + gx] = da[(ty + R) * SW + (tx + R)];
  }
}

template <int KS>
static int launch_rl_t(int planes, int h, int w, int iters, float lambda, const float* img, const float* kern, float* out,
                       cudaStream_t st) {
  const int R = iters * (KS / 2);
  const size_t smem = (size_t)3 * (kRlTX + 2 * R) * rl_rows(kRlTY + 2 * R) * sizeof(float);
  if (smem > 200 * 1024) { set_error("rl_deconv: %d iterations of a %dx%d kernel need %zu bytes of shared memory", iters, KS, KS, smem); return SPEI_ERR_ARG; }
  SPEI_CUDA(cudaFuncSetAttribute(rl_deconv_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((w + kRlTX - 1) / kRlTX, (h + kRlTY - 1) / kRlTY, planes);
  rl_deconv_kernel<KS><<<grid, kRlThreads, smem, st>>>(img, kern, out, h, w, iters, lambda);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_rl_deconv(int n, int c, int h, int w, int ks, int iters, float lambda, const float* img, const float* kern, float* out,
                     cudaStream_t st) {
  const long long planes = (long long)n * c;
  if (planes > 65535) { set_error("rl_deconv: n*c too large"); return SPEI_ERR_ARG; }
  if (iters < 1) { set_error("rl_deconv: num_iterations must be >= 1"); return SPEI_ERR_ARG; }
  if (ks == 3) return launch_rl_t<3>((int)planes, h, w, iters, lambda, img, kern, out, st);
  if (ks == 5) return launch_rl_t<5>((int)planes, h, w, iters, lambda, img, kern, out, st);
  if (ks == 7) return launch_rl_t<7>((int)planes, h, w, iters, lambda, img, kern, out, st);
  set_error("rl_deconv: blur kernel size %d not supported (3, 5 or 7; the reference uses 5, rcl.py:18)", ks);
  return SPEI_ERR_ARG;
}

}  // namespace spei
