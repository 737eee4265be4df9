// Kernel (b'): exactness layer behind the bf16 tcgen05 relevance pass
// (/root/reference/model/SearchTransfer.py:33-34: R = bmm(K^, Q^); R_star, arg = max(R, dim=1)).
//
//  rescore_kernel       one warp per query.  Merges the per-segment top-k candidate lists of the bf16 pass, recomputes the
//                       relevance of the best bf16 candidate exactly (fp32 operands, fp64 accumulation) -> E, and from E
//                       the threshold T below which no key can be the argmax:
//                         certified mode (eps <= 0, default):  T = E - Delta_i, Delta_i = the bound on |bf16 score - exact
//                           score| over ALL keys derived from the measured rounding residuals (stage_norm.cu,
//                           certified_delta()).  The true argmax j* has exact score >= E, hence bf16 score >= E - Delta_i.
//                         fixed window (eps > 0, uncertified): T = best bf16 score - eps.
//                       Every other candidate with bf16 score >= T is rescored exactly and the maximum is taken with
//                       torch.max's first-index tie-break.  A query one of whose lists is SATURATED (its k-th entry is
//                       still >= T, so keys >= T may have been dropped) is queued, with its threshold, for the second
//                       tcgen05 pass (relevance_flagged.cu), which enumerates every key >= T of that query.
//  exact_search_kernel  exhaustive fp32 search on CUDA cores: SPEI_SEARCH_EXACT (the on-GPU checker) and the last resort
//                       when the second pass runs out of capacity.  64 queries x 64 keys register-tiled implicit GEMM over
//                       the 9 taps x 128 channels, packed (score, ~index) atomicMax.
//  unpack_kernel        writes S / arg of the queued queries from the packed maxima.
#include "spei_common.cuh"
#include "exact_score.cuh"

namespace spei {

// persistent CTA that owns tile pair p (the split of relevance_tc*.cu); 32-bit arithmetic whenever (P + 1) * G fits
__host__ __device__ inline long long cta_of_pair_d(long long p, long long P, int G) {
  if ((P + 1) * (long long)G < (1ll << 32)) return (long long)(((unsigned)(p + 1) * (unsigned)G - 1u) / (unsigned)P);
  return ((p + 1) * (long long)G - 1) / P;
}

// the same with the division by P as one multiply-high (FastDiv, spei_common.cuh) when (P + 1) * G stays below 2^31
__device__ __forceinline__ long long cta_of_pair_f(long long p, long long P, int G, const FastDiv divP) {
  if ((P + 1) * (long long)G < (1ll << 31)) return (long long)fast_div((int)(p + 1) * G - 1, divP);
  return cta_of_pair_d(p, P, G);
}

struct RescoreParams {
  int n, rf, H, W, Hr, Wr;
  int q_orient, q_tu, q_tile_u, q_tile_v, nlist;  // query tile grid (to find a query's tile -> its segment count)
  int QT, pair, KT, G, maxseg;
  long long P;
  FastDiv div_lk1, div_wr;    // key index -> (frame, row, column)
  FastDiv div_tu, div_tv, div_P;   // query position -> tile; work step -> persistent CTA
  float eps;                  // > 0: fixed window; <= 0: certified
  const float *q32, *k32, *rq, *rk, *qss, *dq;
  const int* dkmax;
  const float* cval;
  const int32_t* cidx;
  float* S;
  int32_t* arg32;
  int64_t* arg64;
  int32_t* flag_list;         // [n][L] queries queued for the second pass
  int32_t* counters;          // [n] queued per item, then the kCnt* words
  float* thr;                 // [n*L] threshold of a queued query in accumulator units (T / rq)
  unsigned long long* packed; // [n*L] running (score, ~key) maxima of the queued queries
  int32_t* stats;
};

// Query block of one CTA: kRB x kRB queries, one warp each.  Their 3x3 patches overlap, so the block stages the
// (kRB + 2)^2 pixel rows (512 B each) it needs ONCE in shared memory -- 1.15 KB per query instead of the 4.6 KB a private
// patch costs; the kernel is bound by L2 -> SM traffic (every exact score also reads the 4.6 KB key patch).
constexpr int kRB = 4, kRBH = kRB + 2;

// grid: (ceil(W / 4), ceil(H / 4), n)   block: 512 (16 warps)
__global__ void __launch_bounds__(kRB * kRB * 32, 2)
rescore_kernel(const RescoreParams p) {
  __shared__ __align__(16) float qrows[kRBH * kRBH][kC3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = p.H * p.W;
  const int n = blockIdx.z, y0 = blockIdx.y * kRB, x0 = blockIdx.x * kRB;
  {  // halo rows of the block, zero-filled outside the image (= the zero padding of F.unfold)
    const float* qimg = p.q32 + (size_t)n * L * kC3;
    for (int e = threadIdx.x; e < kRBH * kRBH * 32; e += kRB * kRB * 32) {
      const int row = e >> 5, l4 = e & 31;
      const int yy = y0 + row / kRBH - 1, xx = x0 + row % kRBH - 1;
      const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const float* src = qimg + ((size_t)(in ? yy : 0) * p.W + (in ? xx : 0)) * kC3 + l4 * 4;
      const unsigned d = (unsigned)__cvta_generic_to_shared(&qrows[row][l4 * 4]);
      const int sz = in ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int wy = warp / kRB, wx = warp % kRB;
  const int y = y0 + wy, x = x0 + wx;
  const bool live = y < p.H && x < p.W;
  const int ql = live ? y * p.W + x : 0;
  const long long wq = (long long)n * L + ql;
  const int lk1 = p.Hr * p.Wr;

  // which query tile is this, and into how many key segments was it split?
  const int u = p.q_orient == 0 ? x : y, v = p.q_orient == 0 ? y : x;
  const int qt = fast_div(v, p.div_tv) * p.q_tu + fast_div(u, p.div_tu);
  const long long p0 = ((long long)n * p.QT + (qt >> p.pair)) * p.KT;   // (QT = work slots per item: tile pairs with cta_group::2)
  const int nseg = (int)(cta_of_pair_f(p0 + p.KT - 1, p.P, p.G, p.div_P) - cta_of_pair_f(p0, p.P, p.G, p.div_P)) + 1;
  const int nlists = nseg * p.nlist;  // lists of a query are contiguous: [segment][list][kTopK]
  const int ncand = nlists * kTopK;
  const float* cv = p.cval + (size_t)wq * p.maxseg * p.nlist * kTopK;
  const int32_t* ci = p.cidx + (size_t)wq * p.maxseg * p.nlist * kTopK;

  // independent loads first: query norm (+inf marks an all-zero patch, stage_norm.cu), residual bound, first 32 candidates
  const float rq = __ldg(p.rq + wq);
  const float delta = p.eps > 0.f ? 0.f : certified_delta(__ldg(p.dq + wq), __int_as_float(__ldg(p.dkmax + n)));
  int j0 = -1;
  float v0 = 0.f;
  if (lane < ncand) { j0 = __ldg(ci + lane); v0 = __ldg(cv + lane); }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();   // the block's query rows are in shared memory (no block-wide barrier after this point)
  if (!live) return;

  // zero query patch: every relevance is 0 -> first index, S = 0 (reference semantics)
  if (rq == INFINITY) {
    if (lane == 0) {
      p.S[wq] = 0.f; p.arg32[wq] = 0;
      if (p.arg64) p.arg64[wq] = 0;
    }
    return;
  }

  // pass 1: best bf16 score and its key
  float bn = j0 >= 0 ? v0 * rq : -INFINITY;
  int bj = j0;
  for (int e = lane + 32; e < ncand; e += 32) {
    const int j = __ldg(ci + e);
    const float vb = __ldg(cv + e) * rq;
    if (j >= 0 && vb > bn) { bn = vb; bj = j; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, bn, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > bn || (ob == bn && (unsigned)oj < (unsigned)bj)) { bn = ob; bj = oj; }   // (unsigned): -1 loses
  }
  auto qtap = [&](int t) {   // tap (t / 3, t % 3) of this warp's query = block row (wy + t / 3, wx + t % 3)
    return reinterpret_cast<const float4*>(&qrows[(wy + t / 3) * kRBH + wx + t % 3][0]);
  };
  auto exact = [&](int jj) {
    const int f = p.rf == 1 ? 0 : fast_div(jj, p.div_lk1), rem = jj - f * lk1, hr = fast_div(rem, p.div_wr), wr = rem - hr * p.Wr;
    const float rk = __ldg(p.rk + ((size_t)n * p.rf + f) * lk1 + rem);
    return exact_relevance(qtap, p.k32 + ((size_t)n * p.rf + f) * lk1 * kC3, hr, wr, p.Hr, p.Wr, rq, rk, lane);
  };
  // every query has at least one candidate: its lists hold the best keys of every segment (bj >= 0)
  const float E = exact(bj);
  const float thr = p.eps > 0.f ? bn - p.eps : E - delta;
  float max_err = fabsf(bn - E);
  int nres = 1, nviol = (p.eps <= 0.f && max_err > delta) ? 1 : 0;

  // saturation: the last (smallest) entry of some list is still at or above the threshold -- or, certified mode, the
  // threshold lies below what the tensor-core pass kept (best bf16 score off by more than kWindowMargin)
  bool sat = p.eps <= 0.f && thr < bn - certified_window(delta);
  for (int sg = lane; sg < nlists; sg += 32) {
    const int e = sg * kTopK + (kTopK - 1);
    if (__ldg(ci + e) >= 0 && __ldg(cv + e) * rq >= thr) sat = true;
  }
  if (__any_sync(0xffffffffu, sat)) {
    if (lane == 0) {
      const int pos = atomicAdd(p.counters + n, 1);
      p.flag_list[(size_t)n * L + pos] = ql;
      p.thr[wq] = thr / rq;
      p.packed[wq] = pack_score(E, bj);
      if (p.stats) { atomicAdd(p.stats + 0, 1); atomicAdd(p.stats + 1, 1); atomicMax(p.stats + 2, (int)(max_err * 1e9f)); }
    }
    return;  // S / arg are written by unpack_kernel after the second pass
  }

  // pass 2: exact relevance of every other candidate at or above the threshold
  unsigned long long best = pack_score(E, bj);
  for (int e0 = 0; e0 < ncand; e0 += 32) {
    const int e = e0 + lane;
    int j = -1;
    float vb = 0.f;
    if (e < ncand) {
      j = e0 == 0 ? j0 : __ldg(ci + e);
      vb = (e0 == 0 ? v0 : __ldg(cv + e)) * rq;
      if (j == bj || !(vb >= thr)) j = -1;
    }
    unsigned m = __ballot_sync(0xffffffffu, j >= 0);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int jj = __shfl_sync(0xffffffffu, j, src);
      const float vbb = __shfl_sync(0xffffffffu, vb, src);
      const float score = exact(jj);
      const unsigned long long key = pack_score(score, jj);
      best = key > best ? key : best;
      const float err = fabsf(vbb - score);
      max_err = fmaxf(max_err, err);
      if (p.eps <= 0.f && err > delta) ++nviol;
      ++nres;
    }
  }
  if (lane == 0) {
    const int j = packed_key(best);
    p.S[wq] = packed_score(best); p.arg32[wq] = j;
    if (p.arg64) p.arg64[wq] = j;
    if (p.stats) {
      atomicAdd(p.stats + 1, nres);
      atomicMax(p.stats + 2, (int)(max_err * 1e9f));
      if (nviol) atomicAdd(p.stats + 6, nviol);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exhaustive fp32 search
// ---------------------------------------------------------------------------------------------
struct ExactParams {
  int n, rf, H, W, Hr, Wr;
  int key_splits;
  const float *q32, *k32, *rq, *rk;
  const int32_t* list;        // [n][L] query ids, or NULL = all queries
  const int32_t* list_count;  // [n], or NULL
  const int32_t* enable;      // list mode: run only if *enable != 0 (the second pass ran out of capacity); NULL = always
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
  if (p.enable && *reinterpret_cast<const volatile int32_t*>(p.enable) == 0) return;
  const int nq = p.list ? __ldg(p.list_count + n) : L;
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
    if (crq[i] == INFINITY) crq[i] = 0.f;   // all-zero query patch (marked by stage_norm.cu): every relevance is 0 -> first index
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

// clears the packed maxima, the per-item queue counters + kCnt* words and the caller's stats block
__global__ void __launch_bounds__(256)
clear_packed_kernel(unsigned long long* packed, int32_t* counters, size_t total, int n, int32_t* stats) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) packed[i] = 0ull;
  if (i < (size_t)(n + kCntWords)) counters[i] = 0;
  if (stats && i < (size_t)kStatsWords) stats[i] = 0;
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
    const int j = packed_key(b);
    S[(size_t)n * L + q] = packed_score(b);
    arg32[(size_t)n * L + q] = j;
    if (arg64) arg64[(size_t)n * L + q] = j;
  }
}

static int clear_lists(const Plan& p, char* ws, int32_t* stats, cudaStream_t st) {
  const size_t tot = (size_t)p.n * p.H * p.W;
  const size_t cover = tot > (size_t)(p.n + kCntWords) ? tot : (size_t)(p.n + kCntWords);
  clear_packed_kernel<<<(unsigned)((cover + 255) / 256), 256, 0, st>>>((unsigned long long*)(ws + p.off_packed),
                                                                        (int32_t*)(ws + p.off_counters), tot, p.n, stats);
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

int launch_exact_all(const Plan& p, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats, char* ws, cudaStream_t st) {
  if (p.n > 65535) { set_error("exact search: n too large"); return SPEI_ERR_ARG; }
  int rc = clear_lists(p, ws, stats, st);
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
  int rc = clear_lists(p, ws, stats, st);
  if (rc) return rc;
  RescoreParams r{};
  r.n = p.n; r.rf = p.rf; r.H = p.H; r.W = p.W; r.Hr = p.Hr; r.Wr = p.Wr;
  r.q_orient = p.q.orient; r.q_tu = p.q.tu; r.q_tile_u = p.q.tile_u; r.q_tile_v = p.q.tile_v; r.nlist = p.nlist; r.div_lk1 = make_fastdiv(p.Hr * p.Wr); r.div_wr = make_fastdiv(p.Wr);
  r.div_tu = make_fastdiv(p.q.tile_u); r.div_tv = make_fastdiv(p.q.tile_v); r.div_P = make_fastdiv((int)(p.P < (1ll << 31) ? p.P : 1)); r.QT = p.QTs; r.pair = p.pair; r.KT = p.KT; r.G = p.G; r.maxseg = p.maxseg; r.P = p.P;
  r.eps = eps;
  r.q32 = (const float*)(ws + p.off_q32); r.k32 = (const float*)(ws + p.off_k32);
  r.rq = (const float*)(ws + p.off_rq); r.rk = (const float*)(ws + p.off_rk); r.qss = (const float*)(ws + p.off_qss);
  r.dq = (const float*)(ws + p.off_dq); r.dkmax = (const int*)(ws + p.off_dkmax);
  r.cval = (const float*)(ws + p.off_cval); r.cidx = (const int32_t*)(ws + p.off_cidx);
  r.S = S; r.arg32 = arg32; r.arg64 = arg64;
  r.flag_list = (int32_t*)(ws + p.off_flag); r.counters = (int32_t*)(ws + p.off_counters);
  r.thr = (float*)(ws + p.off_thr); r.packed = (unsigned long long*)(ws + p.off_packed);
  r.stats = stats;
  if ((p.H + kRB - 1) / kRB > 65535) { set_error("rescore: grid too large"); return SPEI_ERR_ARG; }
  rescore_kernel<<<dim3((p.W + kRB - 1) / kRB, (p.H + kRB - 1) / kRB, p.n), kRB * kRB * 32, 0, st>>>(r);
  SPEI_CUDA(cudaGetLastError());

  // queued queries (saturated candidate lists): second tcgen05 pass enumerating every key at or above each query's
  // threshold, exact rescoring of what it emits.  The counts live on the device (no host sync): fixed grids, blocks
  // beyond the count exit immediately.
  if ((rc = launch_relevance_flagged(p, stats, ws, st))) return rc;
  // last resort (the pass ran out of packed rows or emission slots): exhaustive fp32 search of the queued queries
  ExactParams e = exact_params(p, ws);
  e.list = r.flag_list; e.list_count = r.counters;
  e.enable = r.counters + p.n + kCntExhaust;
  const int L = p.H * p.W, qblocks = (L + kEQ - 1) / kEQ;
  const int max_splits = (p.rf * p.Hr * p.Wr + kEK - 1) / kEK;
  // (a small grid: this launch is idle in every ordinary call and only has to exit quickly; when it does run, the
  // grid-stride loops cover any queue length)
  e.key_splits = max_splits < 74 ? max_splits : 74;
  exact_search_kernel<<<dim3(qblocks < 8 ? qblocks : 8, e.key_splits, p.n), 256, 0, st>>>(e);
  SPEI_CUDA(cudaGetLastError());
  unpack_kernel<<<dim3(148, 1, p.n), 256, 0, st>>>(e.packed, e.list, e.list_count, L, S, arg32, arg64);
  SPEI_CUDA(cudaGetLastError());
  return SPEI_OK;
}

}  // namespace spei
