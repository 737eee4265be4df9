// Kernel (b'): exactness layer behind the bf16 tcgen05 relevance pass
// (/root/reference/model/SearchTransfer.py:33-34: R = bmm(K^, Q^); R_star, arg = max(R, dim=1)).
//
//  rescore_kernel       one warp per query: merges the per-segment top-k candidate lists of the bf16
//                       pass, keeps those within eps of the best bf16 score, recomputes their relevance
//                       exactly (fp32 operands, 4-term fp32 partial dots accumulated in fp64) and picks the maximum with
//                       torch.max's first-index tie-break.  A query whose candidate list is saturated
//                       (its k-th candidate is still inside the eps window, so a better key might have
//                       been dropped) is queued for the exhaustive search below.
//  exact_search_kernel  exhaustive fp32 search on CUDA cores for a list of queries (or all of them:
//                       SPEI_SEARCH_EXACT, the on-GPU checker).  64 queries x 64 keys register-tiled
//                       implicit GEMM over the 9 taps x 128 channels, packed (score, ~index) atomicMax.
//  unpack_kernel        writes S / arg for exhaustively searched queries.
#include "spei_common.cuh"

namespace spei {

__host__ __device__ inline long long cta_of_pair_d(long long p, long long P, int G) {
  return ((p + 1) * (long long)G - 1) / P;
}

__device__ __forceinline__ unsigned flip_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unflip_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long pack_score(float s, int j) {
  return ((unsigned long long)flip_f32(s) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j);
}

struct RescoreParams {
  int n, rf, H, W, Hr, Wr;
  int q_orient, q_tu, q_tile_u, q_tile_v, nlist;  // query tile grid (to find a query's tile -> its segment count)
  int QT, KT, G, maxseg;
  long long P;
  float eps;
  const float *q32, *k32, *rq, *rk, *qss;
  const float* cval;
  const int32_t* cidx;
  float* S;
  int32_t* arg32;
  int64_t* arg64;
  int32_t* flag_list;   // [n][L]
  int32_t* flag_count;  // [n]
  int32_t* stats;
};

__global__ void __launch_bounds__(256, 4)
rescore_kernel(const RescoreParams p) {
  // The kernel is latency bound (every phase is a dependent L2 round trip), so it is written for
  // occupancy and early issue: the query's 9x128 patch goes to shared memory with cp.async (no register
  // staging, <= 64 registers -> 32 warps per SM) and every load that does not depend on another is
  // issued before the first branch.
  __shared__ __align__(16) float qpatch[8][9 * kC3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wq = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int L = p.H * p.W;
  if (wq >= (long long)p.n * L) return;
  const int n = (int)(wq / L), ql = (int)(wq % L);
  const int y = ql / p.W, x = ql % p.W;
  const int lk1 = p.Hr * p.Wr;

  // query patch -> shared memory (zero-filled taps outside the image)
  {
    const float* qimg = p.q32 + (size_t)n * L * kC3;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const float* src = qimg + ((size_t)(in ? yy : y) * p.W + (in ? xx : x)) * kC3 + lane * 4;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&qpatch[warp][t * kC3 + lane * 4]);
      const int sz = in ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // which query tile is this, and into how many key segments was it split?
  const int u = p.q_orient == 0 ? x : y, v = p.q_orient == 0 ? y : x;
  const int qt = (v / p.q_tile_v) * p.q_tu + (u / p.q_tile_u);
  const long long p0 = ((long long)n * p.QT + qt) * p.KT;
  const int nseg = (int)(cta_of_pair_d(p0 + p.KT - 1, p.P, p.G) - cta_of_pair_d(p0, p.P, p.G)) + 1;
  const int nlists = nseg * p.nlist;  // lists of a query are contiguous: [segment][list][kTopK]
  const int ncand = nlists * kTopK;
  const float* cv = p.cval + (size_t)wq * p.maxseg * p.nlist * kTopK;
  const int32_t* ci = p.cidx + (size_t)wq * p.maxseg * p.nlist * kTopK;

  // independent loads first: patch energy (zero test), query norm, first 32 candidates
  float s = 0.f;
  {
    const float* ss = p.qss + (size_t)n * L;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) s += __ldg(ss + yy * p.W + xx);
    }
  }
  const float rq = __ldg(p.rq + wq);
  int j0 = -1;
  float v0 = 0.f;
  if (lane < ncand) { j0 = __ldg(ci + lane); v0 = __ldg(cv + lane); }

  // zero query patch: every relevance is 0 -> first index, S = 0 (reference semantics)
  if (s == 0.f) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) {
      p.S[wq] = 0.f; p.arg32[wq] = 0;
      if (p.arg64) p.arg64[wq] = 0;
    }
    return;
  }

  // pass 1: best bf16 score
  float bn = j0 >= 0 ? v0 * rq : -INFINITY;
  for (int e = lane + 32; e < ncand; e += 32) {
    const int j = __ldg(ci + e);
    if (j >= 0) bn = fmaxf(bn, __ldg(cv + e) * rq);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) bn = fmaxf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
  const float thr = bn - p.eps;

  // saturation: the last (smallest) entry of some segment is still inside the window
  bool sat = false;
  for (int sg = lane; sg < nlists; sg += 32) {
    const int e = sg * kTopK + (kTopK - 1);
    if (__ldg(ci + e) >= 0 && __ldg(cv + e) * rq >= thr) sat = true;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  if (__any_sync(0xffffffffu, sat)) {
    if (lane == 0) {
      const int pos = atomicAdd(p.flag_count + n, 1);
      p.flag_list[(size_t)n * L + pos] = ql;
      if (p.stats) atomicAdd(p.stats + 0, 1);
    }
    return;  // S / arg are written by unpack_kernel after the exhaustive search
  }
  const float4* qv = reinterpret_cast<const float4*>(&qpatch[warp][0]) + lane;  // tap t at qv[t * 32]

  // pass 2: exact relevance of every kept candidate
  unsigned long long best = 0ull;
  int nres = 0;
  float max_err = 0.f;  // largest |bf16 score - exact score| seen: evidence for the eps window
  for (int e0 = 0; e0 < ncand; e0 += 32) {
    const int e = e0 + lane;
    int j = -1;
    float vb = 0.f;
    if (e < ncand) {
      j = e0 == 0 ? j0 : __ldg(ci + e);
      vb = (e0 == 0 ? v0 : __ldg(cv + e)) * rq;
      if (j >= 0 && !(vb >= thr)) j = -1;
    }
    unsigned m = __ballot_sync(0xffffffffu, j >= 0);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int jj = __shfl_sync(0xffffffffu, j, src);
      const float vbb = __shfl_sync(0xffffffffu, vb, src);
      const int f = jj / lk1, rem = jj - f * lk1, hr = rem / p.Wr, wr = rem - hr * p.Wr;
      const float* kimg = p.k32 + ((size_t)n * p.rf + f) * lk1 * kC3;
      // Each lane owns 4 channels of every tap: their 4 products are summed in fp32 (one rounding of ~6e-8 relative per
      // product, ~2e-9 absolute on a normalised score -- four orders of magnitude inside the 1e-5 near-tie rule) and the 9
      // tap partials, then the 32 lanes, are accumulated in fp64.  The all-fp64 version (8 F2F + 4 DFMA per lane and tap)
      // kept the conversion pipe 40 % busy (ncu, round 1) and was the kernel's top pipe.
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // three independent chains (taps t % 3), fixed order
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = hr + t / 3 - 1, xx = wr + t % 3 - 1;
        if (yy >= 0 && yy < p.Hr && xx >= 0 && xx < p.Wr) {
          const float4 kv = __ldg(reinterpret_cast<const float4*>(kimg + ((size_t)yy * p.Wr + xx) * kC3) + lane);
          const float4 qq = qv[t * 32];
          float part = qq.x * kv.x;
          part = fmaf(qq.y, kv.y, part);
          part = fmaf(qq.z, kv.z, part);
          part = fmaf(qq.w, kv.w, part);
          if (t % 3 == 0) acc0 += (double)part;
          else if (t % 3 == 1) acc1 += (double)part;
          else acc2 += (double)part;
        }
      }
      double acc = (acc0 + acc1) + acc2;
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float rk = __ldg(p.rk + ((size_t)n * p.rf + f) * lk1 + rem);
      const float score = (float)(acc * (double)rq * (double)rk);
      const unsigned long long key = pack_score(score, jj);
      best = key > best ? key : best;
      max_err = fmaxf(max_err, fabsf(vbb - score));
      ++nres;
    }
  }
  if (lane == 0) {
    const float s = unflip_f32((unsigned)(best >> 32));
    const int j = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
    p.S[wq] = s; p.arg32[wq] = j;
    if (p.arg64) p.arg64[wq] = j;
    if (p.stats) {
      atomicAdd(p.stats + 1, nres);
      atomicMax(p.stats + 2, (int)(max_err * 1e9f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exhaustive fp32 search
// ---------------------------------------------------------------------------------------------
struct ExactParams {
  int n, rf, H, W, Hr, Wr;
  int key_splits;
  int small_max;              // list mode: items with <= small_max queued queries belong to exact_small_kernel
  const float *q32, *k32, *rq, *rk;
  const int32_t* list;        // [n][L] query ids, or NULL = all queries
  const int32_t* list_count;  // [n], or NULL
  unsigned long long* packed; // [n][L]
};

constexpr int kEQ = 64, kEK = 64, kEC = 32;

// grid: (query blocks, key splits, n)   block: 256
__global__ void __launch_bounds__(256)
exact_search_kernel(const ExactParams p) {
  __shared__ __align__(16) float As[kEC][kEQ];
  __shared__ __align__(16) float Bs[kEC][kEK];
  const int n = blockIdx.z;
  const int L = p.H * p.W, lk1 = p.Hr * p.Wr, Lk = p.rf * lk1;
  const int nq = p.list ? __ldg(p.list_count + n) : L;
  if (p.list && nq <= p.small_max) return;  // handled by exact_small_kernel
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lq = t & 63, lpart = t >> 6;  // loader role: one of 64 rows, 8 of the 32 channels
  // grid-stride over blocks of 64 queries: the queued-query count is only known on the device
  for (int qb = blockIdx.x * kEQ; qb < nq; qb += gridDim.x * kEQ) {

  // loader: this thread's query
  int my_q = -1, my_qy = 0, my_qx = 0;
  if (qb + lq < nq) {
    my_q = p.list ? __ldg(p.list + (size_t)n * L + qb + lq) : qb + lq;
    my_qy = my_q / p.W; my_qx = my_q % p.W;
  }
  // compute role: 4 queries x 4 keys
  int cq[4];
  float crq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = qb + ty * 4 + i;
    cq[i] = r < nq ? (p.list ? __ldg(p.list + (size_t)n * L + r) : r) : -1;
    crq[i] = cq[i] >= 0 ? __ldg(p.rq + (size_t)n * L + cq[i]) : 0.f;
  }
  unsigned long long best[4] = {0ull, 0ull, 0ull, 0ull};

  const int per = ((Lk + p.key_splits - 1) / p.key_splits + kEK - 1) / kEK * kEK;
  const int k_lo = blockIdx.y * per, k_hi = min(Lk, k_lo + per);
  const float* qimg = p.q32 + (size_t)n * L * kC3;

  for (int kb = k_lo; kb < k_hi; kb += kEK) {
    const int my_j = kb + lq;
    int kf = 0, khr = 0, kwr = 0;
    const bool kin = my_j < k_hi;
    if (kin) { kf = my_j / lk1; const int rem = my_j - kf * lk1; khr = rem / p.Wr; kwr = rem - khr * p.Wr; }
    const float* kimg = p.k32 + ((size_t)n * p.rf + kf) * lk1 * kC3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const int qy = my_qy + dy, qx = my_qx + dx;
      const bool qok = my_q >= 0 && qy >= 0 && qy < p.H && qx >= 0 && qx < p.W;
      const int ky = khr + dy, kx = kwr + dx;
      const bool kok = kin && ky >= 0 && ky < p.Hr && kx >= 0 && kx < p.Wr;
      const float* qsrc = qimg + ((size_t)qy * p.W + qx) * kC3 + lpart * 8;
      const float* ksrc = kimg + ((size_t)ky * p.Wr + kx) * kC3 + lpart * 8;
      for (int c0 = 0; c0 < kC3; c0 += kEC) {
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
        if (qok) { a0 = __ldg(reinterpret_cast<const float4*>(qsrc + c0)); a1 = __ldg(reinterpret_cast<const float4*>(qsrc + c0) + 1); }
        if (kok) { b0 = __ldg(reinterpret_cast<const float4*>(ksrc + c0)); b1 = __ldg(reinterpret_cast<const float4*>(ksrc + c0) + 1); }
        __syncthreads();
        const int cb = lpart * 8;
        As[cb + 0][lq] = a0.x; As[cb + 1][lq] = a0.y; As[cb + 2][lq] = a0.z; As[cb + 3][lq] = a0.w;
        As[cb + 4][lq] = a1.x; As[cb + 5][lq] = a1.y; As[cb + 6][lq] = a1.z; As[cb + 7][lq] = a1.w;
        Bs[cb + 0][lq] = b0.x; Bs[cb + 1][lq] = b0.y; Bs[cb + 2][lq] = b0.z; Bs[cb + 3][lq] = b0.w;
        Bs[cb + 4][lq] = b1.x; Bs[cb + 5][lq] = b1.y; Bs[cb + 6][lq] = b1.z; Bs[cb + 7][lq] = b1.w;
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kEC; ++c) {
          const float4 a = *reinterpret_cast<const float4*>(&As[c][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
      }
    }
    // epilogue: normalise, running packed max (first index wins ties)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kj = kb + tx * 4 + j;
      if (kj < k_hi) {
        const int f = kj / lk1, rem = kj - f * lk1;
        const float rk = __ldg(p.rk + ((size_t)n * p.rf + f) * lk1 + rem);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned long long key = pack_score(acc[i][j] * crq[i] * rk, kj);
          best[i] = key > best[i] ? key : best[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned long long b = best[i];
#pragma unroll
    for (int o = 8; o; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
      b = other > b ? other : b;
    }
    if (tx == 0 && cq[i] >= 0) atomicMax(p.packed + (size_t)n * L + cq[i], b);
  }
  }  // query blocks
}

// Few queued queries (the common case: a handful per frame): the 64-query tile of the kernel above
// would be almost empty, so a warp brute-forces (query, key) pairs instead -- the query's 9x128 patch
// lives in registers, each key costs nine coalesced 512-byte reads and an fp64 dot, exactly the
// arithmetic of rescore_kernel.  grid: (key blocks, 1, n); items with more than kSmallMax queued
// queries are left to exact_search_kernel (which skips the others).
constexpr int kSmallMax = 64;

// grid: (key blocks, query lanes, n)
__global__ void __launch_bounds__(256, 4)
exact_small_kernel(const ExactParams p) {
  __shared__ __align__(16) float qpatch[9 * kC3];
  const int n = blockIdx.z;
  const int cnt = __ldg(p.list_count + n);
  if (cnt == 0 || cnt > kSmallMax || (int)blockIdx.y >= cnt) return;
  const int L = p.H * p.W, lk1 = p.Hr * p.Wr, Lk = p.rf * lk1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (Lk + gridDim.x - 1) / gridDim.x;
  const int k_lo = blockIdx.x * per, k_hi = min(Lk, k_lo + per);
  const float* qimg = p.q32 + (size_t)n * L * kC3;
  const float4* qv = reinterpret_cast<const float4*>(qpatch) + lane;  // tap t at qv[t * 32]
  for (int qi = blockIdx.y; qi < cnt; qi += gridDim.y) {
    const int q = __ldg(p.list + (size_t)n * L + qi);
    const int y = q / p.W, x = q % p.W;
    const float rq = __ldg(p.rq + (size_t)n * L + q);
    __syncthreads();  // previous query's patch no longer in use
    for (int e = threadIdx.x; e < 9 * 32; e += 256) {
      const int t = e >> 5, l4 = e & 31;
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      reinterpret_cast<float4*>(qpatch)[e] =
          in ? __ldg(reinterpret_cast<const float4*>(qimg + ((size_t)yy * p.W + xx) * kC3) + l4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    unsigned long long best = 0ull;
    for (int j = k_lo + warp; j < k_hi; j += 8) {
      const int f = j / lk1, rem = j - f * lk1, hr = rem / p.Wr, wr = rem - hr * p.Wr;
      const float* kimg = p.k32 + ((size_t)n * p.rf + f) * lk1 * kC3;
      // fp32 accumulation like exact_search_kernel (the exhaustive paths agree with each other and are
      // within ~1e-7 of the fp64 rescoring, far inside the 1e-5 near-tie rule)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = hr + t / 3 - 1, xx = wr + t % 3 - 1;
        if (yy >= 0 && yy < p.Hr && xx >= 0 && xx < p.Wr) {
          const float4 kv = __ldg(reinterpret_cast<const float4*>(kimg + ((size_t)yy * p.Wr + xx) * kC3) + lane);
          const float4 qq = qv[t * 32];
          a0 = fmaf(qq.x, kv.x, a0); a1 = fmaf(qq.y, kv.y, a1); a2 = fmaf(qq.z, kv.z, a2); a3 = fmaf(qq.w, kv.w, a3);
        }
      }
      float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float rk = __ldg(p.rk + ((size_t)n * p.rf + f) * lk1 + rem);
      const unsigned long long key = pack_score(acc * rq * rk, j);
      best = key > best ? key : best;
    }
    if (lane == 0 && best) atomicMax(p.packed + (size_t)n * L + q, best);
  }
}

__global__ void __launch_bounds__(256)
clear_packed_kernel(unsigned long long* packed, int32_t* flag_count, size_t total, int n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) packed[i] = 0ull;
  if (i < (size_t)n) flag_count[i] = 0;
}

// grid: (blocks, 1, n)
__global__ void __launch_bounds__(256)
unpack_kernel(const unsigned long long* __restrict__ packed, const int32_t* __restrict__ list,
              const int32_t* __restrict__ list_count, int L, float* S, int32_t* arg32, int64_t* arg64) {
  const int n = blockIdx.z;
  const int cnt = list ? __ldg(list_count + n) : L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    const int q = list ? __ldg(list + (size_t)n * L + i) : i;
    const unsigned long long b = packed[(size_t)n * L + q];
    const float s = unflip_f32((unsigned)(b >> 32));
    const int j = (int)(0xffffffffu - (unsigned)(b & 0xffffffffull));
    S[(size_t)n * L + q] = s;
    arg32[(size_t)n * L + q] = j;
    if (arg64) arg64[(size_t)n * L + q] = j;
  }
}

static int clear_lists(const Plan& p, char* ws, cudaStream_t st) {
  const size_t tot = (size_t)p.n * p.H * p.W;
  clear_packed_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((unsigned long long*)(ws + p.off_packed),
                                                                      (int32_t*)(ws + p.off_counters), tot, p.n);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

static ExactParams exact_params(const Plan& p, char* ws) {
  ExactParams e{};
  e.n = p.n; e.rf = p.rf; e.H = p.H; e.W = p.W; e.Hr = p.Hr; e.Wr = p.Wr;
  e.q32 = (const float*)(ws + p.off_q32); e.k32 = (const float*)(ws + p.off_k32);
  e.rq = (const float*)(ws + p.off_rq); e.rk = (const float*)(ws + p.off_rk);
  e.packed = (unsigned long long*)(ws + p.off_packed);
  return e;
}

int launch_exact_all(const Plan& p, float* S, int32_t* arg32, int64_t* arg64, char* ws, cudaStream_t st) {
  if (p.n > 65535) { set_error("exact search: n too large"); return SPEI_ERR_ARG; }
  int rc = clear_lists(p, ws, st);
  if (rc) return rc;
  ExactParams e = exact_params(p, ws);
  const int L = p.H * p.W, qblocks = (L + kEQ - 1) / kEQ;
  // enough CTAs to fill the machine a few times over
  int splits = (148 * 4 + qblocks * p.n - 1) / (qblocks * p.n);
  const int max_splits = (p.rf * p.Hr * p.Wr + kEK - 1) / kEK;
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  e.key_splits = splits;
  exact_search_kernel<<<dim3(qblocks, splits, p.n), 256, 0, st>>>(e);
  SPEI_CUDA(cudaGetLastError());
  unpack_kernel<<<dim3((L + 255) / 256, 1, p.n), 256, 0, st>>>(e.packed, nullptr, nullptr, L, S, arg32, arg64);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

int launch_rescore(const Plan& p, float eps, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats, char* ws,
                   cudaStream_t st) {
  if (p.n > 65535) { set_error("rescore: n too large"); return SPEI_ERR_ARG; }
  int rc = clear_lists(p, ws, st);
  if (rc) return rc;
  RescoreParams r{};
  r.n = p.n; r.rf = p.rf; r.H = p.H; r.W = p.W; r.Hr = p.Hr; r.Wr = p.Wr;
  r.q_orient = p.q.orient; r.q_tu = p.q.tu; r.q_tile_u = p.q.tile_u; r.q_tile_v = p.q.tile_v; r.nlist = p.nlist; r.QT = p.QT; r.KT = p.KT; r.G = p.G; r.maxseg = p.maxseg; r.P = p.P;
  r.eps = eps;
  r.q32 = (const float*)(ws + p.off_q32); r.k32 = (const float*)(ws + p.off_k32);
  r.rq = (const float*)(ws + p.off_rq); r.rk = (const float*)(ws + p.off_rk); r.qss = (const float*)(ws + p.off_qss);
  r.cval = (const float*)(ws + p.off_cval); r.cidx = (const int32_t*)(ws + p.off_cidx);
  r.S = S; r.arg32 = arg32; r.arg64 = arg64;
  r.flag_list = (int32_t*)(ws + p.off_flag); r.flag_count = (int32_t*)(ws + p.off_counters);
  r.stats = stats;
  const long long nq = (long long)p.n * p.H * p.W;
  rescore_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(r);
  SPEI_CUDA(cudaGetLastError());

  // exhaustive search for the queued queries.  The count lives on the device (no host sync): launch a
  // grid that covers the worst case per item in chunks; blocks beyond the count exit immediately.
  ExactParams e = exact_params(p, ws);
  e.list = r.flag_list; e.list_count = r.flag_count;
  const int L = p.H * p.W, qblocks = (L + kEQ - 1) / kEQ;
  const int max_splits = (p.rf * p.Hr * p.Wr + kEK - 1) / kEK;
  // few queries are expected here: many key splits (short serial loops), few query-block columns
  e.key_splits = max_splits < 450 ? max_splits : 450;
  e.small_max = kSmallMax;
  const int key_blocks = (p.rf * p.Hr * p.Wr + 63) / 64;
  exact_small_kernel<<<dim3(key_blocks < 296 ? key_blocks : 296, 8, p.n), 256, 0, st>>>(e);
  SPEI_CUDA(cudaGetLastError());
  exact_search_kernel<<<dim3(qblocks < 8 ? qblocks : 8, e.key_splits, p.n), 256, 0, st>>>(e);
  SPEI_CUDA(cudaGetLastError());
  unpack_kernel<<<dim3(148, 1, p.n), 256, 0, st>>>(e.packed, e.list, e.list_count, L, S, arg32, arg64);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
