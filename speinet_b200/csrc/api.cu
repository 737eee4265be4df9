// C-ABI entry points of libspeinet_b200 (see include/speinet_b200.h): argument checks, the tiling
// plan, the workspace carve-up and the stage sequence of SearchTransfer.forward
// (/root/reference/model/SearchTransfer.py:24-51).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "spei_common.cuh"

namespace spei {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return SPEI_ERR_CUDA;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Pick the orientation (and, for keys, the tile height Ny) that wastes the fewest padded positions.
// dense:       tiles are 8 (u) x tile_v, MMA N = 8*Ny
// tap-sharing: tiles are 30 interior (32 with halo) x tile_v, MMA N = 32*Ny; cost counts MMA columns
static OperandPlan plan_operand(int H, int W, bool is_key, bool shared, int force_orient = -1, double* cost_out = nullptr,
                                bool even_ny = false) {
  OperandPlan best{};
  double best_cost = 1e300;
  const int tile_u = shared ? kSTileU : kTileU, cols_u = shared ? kSBoxU : kTileU;
  const int qv = shared ? kSQTileV : kQTileV, max_ny = shared ? kSMaxNy : kMaxNy;
  for (int orient = 0; orient < 2; ++orient) {
    if (force_orient >= 0 && orient != force_orient) continue;
    const int U = orient == 0 ? W : H, V = orient == 0 ? H : W;
    const int ny_lo = is_key ? (shared ? 1 : 2) : qv, ny_hi = is_key ? max_ny : qv;
    for (int ny = ny_hi; ny >= ny_lo; ny -= (shared ? 1 : 2)) {
      if (is_key && even_ny && (ny & 1)) continue;   // CTA pairs split the key rows of a tile between two CTAs
      OperandPlan o{};
      o.orient = orient; o.U = U; o.V = V; o.tile_u = tile_u; o.tile_v = ny;
      o.tu = ceil_div(U, tile_u); o.tv = ceil_div(V, ny);
      o.Upad = o.tu * tile_u + 2; o.Vpad = o.tv * ny + 2;
      double cost = (double)o.tu * cols_u * o.tv * ny;
      if (is_key) {
        // narrower MMA N re-reads the 4 KB query operand more often per flop: shared-memory
        // bytes per cycle = 8192/N + 64; penalise N below ~192
        const double n_cols = (double)cols_u * ny, bpc = 8192.0 / n_cols + 64.0;
        if (bpc > 104.0) cost *= bpc / 104.0;
      }
      if (cost < best_cost - 1e-9) { best_cost = cost; best = o; }
    }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}

static long long cta_of_pair(long long p, long long P, int G) { return ((p + 1) * (long long)G - 1) / P; }

int make_plan(const SpeiShape& s, int num_sms, Plan* out) {
  Plan p{};
  const bool shared = s.search == SPEI_SEARCH_TCS;
  p.mode = shared ? SPEI_SEARCH_TCS : SPEI_SEARCH_TC;
  p.nlist = shared ? tcs_epilogue_groups() : 1;
  // tap-sharing kernel on CTA pairs (cta_group::2): two query tiles per step share every key tile
  const bool pair = shared && tcs_cta_pairs() && num_sms >= 2;
  p.pair = pair ? 1 : 0;
  p.n = s.n; p.rf = s.rf; p.H = s.h; p.W = s.w; p.Hr = s.hr; p.Wr = s.wr;
  p.io_bf16 = s.io_dtype == SPEI_IO_BF16;
  if (!shared) {
    // the MMA applies all nine taps with per-operand address offsets: each operand picks its own orientation
    p.q = plan_operand(s.h, s.w, false, false);
    p.k = plan_operand(s.hr, s.wr, true, false);
  } else {
    // the u taps are summed in the epilogue for both operands at once: u must be the same image axis for both
    double best = 1e300;
    for (int orient = 0; orient < 2; ++orient) {
      double cq = 0, ck = 0;
      const OperandPlan oq = plan_operand(s.h, s.w, false, true, orient, &cq);
      const OperandPlan ok = plan_operand(s.hr, s.wr, true, true, orient, &ck, pair);
      if (cq * ck < best - 1e-9) { best = cq * ck; p.q = oq; p.k = ok; }
    }
  }
  p.QT = p.q.tiles();
  p.KT = p.rf * p.k.tiles();
  p.QTs = pair ? (p.QT + 1) / 2 : p.QT;
  p.P = (long long)p.n * p.QTs * p.KT;
  const int workers = pair ? num_sms / 2 : num_sms;   // persistent CTAs, or CTA pairs
  p.G = (int)((p.P < (long long)workers) ? p.P : (long long)workers);
  p.maxseg = 1;
  for (long long t = 0; t < (long long)p.n * p.QTs; ++t) {
    const long long p0 = t * p.KT;
    const int nseg = (int)(cta_of_pair(p0 + p.KT - 1, p.P, p.G) - cta_of_pair(p0, p.P, p.G)) + 1;
    if (nseg > p.maxseg) p.maxseg = nseg;
  }
  const size_t L = (size_t)s.h * s.w, Lk1 = (size_t)s.hr * s.wr, nq = (size_t)s.n, nk = (size_t)s.n * s.rf;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  p.off_qbf = take(nq * kCG * p.q.Vpad * p.q.Upad * 16);
  p.off_kbf = take(nk * kCG * p.k.Vpad * p.k.Upad * 16);
  p.off_q32 = take(nq * L * kC3 * 4);
  p.off_k32 = take(nk * Lk1 * kC3 * 4);
  p.off_qss = take(nq * L * 4);
  p.off_kss = take(nk * Lk1 * 4);
  p.off_qrs = take(nq * L * 4);
  p.off_krs = take(nk * Lk1 * 4);
  p.off_dq = take(nq * L * 4);
  p.off_dkmax = take((nq + 16) * 4);
  p.off_rq = take(nq * L * 4);
  p.off_rk = take(nk * Lk1 * 4);
  // dense: [tv*Ny][tu*8] tile-padded; tap-sharing: [tv*Ny][tu*14 + 2] with a 1-position border
  p.off_rkpad = take(nk * (size_t)(p.k.tv * p.k.tile_v) * (shared ? p.k.Upad : p.k.tu * kTileU) * 4);
  p.off_cval = take(nq * L * p.maxseg * p.nlist * kTopK * 4);
  p.off_cidx = take(nq * L * p.maxseg * p.nlist * kTopK * 4);
  p.off_flag = take(nq * L * 4);
  p.off_packed = take(nq * L * 8);
  p.off_arg32 = take(nq * L * 4);
  p.off_counters = take((nq + kCntWords + 8) * 4);  // per-item count of queued queries, then the kCnt* words
  p.off_thr = take(nq * L * 4);
  {
    const size_t rows = (nq * L + 127) / 128 * 128 + 128 * nq;  // every item's last tile may be partial
    p.flag_rows = (int)(rows < (size_t)kFlagMaxRows ? rows : (size_t)kFlagMaxRows);
  }
  p.off_apack = take((size_t)p.flag_rows * 9 * kC3 * 2);
  p.off_prow_thr = take((size_t)p.flag_rows * 4);
  p.off_prow_q = take((size_t)p.flag_rows * 4);
  p.off_emit_q = take((size_t)kFlagMaxEmit * 4);
  p.off_emit_k = take((size_t)kFlagMaxEmit * 4);
  p.off_errflag = take(64);
  p.off_ref3n = take(nk * Lk1 * kC3 * 4);        // channels-last copies of ref_lv3 (when it is not the searched tensor)
  p.off_ref2n = take(nk * Lk1 * 4 * (kC3 / 2) * 4);  // ... and of ref_lv2 ([2hr][2wr][64])
  p.off_ref1c = take(nk * Lk1 * 16 * (size_t)s.c1 * 4);   // cell-major copy of ref_lv1 ([c1][hr][wr][4 x 4])
  p.off_gmode = take(64);                                 // lv1 gather: path chosen from the match field (gather_fold.cu)
  p.total = off;
  *out = p;
  return SPEI_OK;
}

static int check_device(int* num_sms) {
  int dev = 0;
  SPEI_CUDA(cudaGetDevice(&dev));
  int major = 0, sms = 0;
  SPEI_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SPEI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; libspeinet_b200 is sm_100a only (no fallback)", dev, major);
    return SPEI_ERR_ARCH;
  }
  *num_sms = sms;
  return SPEI_OK;
}

static int check_shape(const SpeiShape* s) {
  if (!s) { set_error("shape is NULL"); return SPEI_ERR_ARG; }
  if (s->n < 1 || s->h < 1 || s->w < 1 || s->hr < 1 || s->wr < 1 || s->rf < 1) {
    set_error("empty or negative dimension: n=%d h=%d w=%d hr=%d wr=%d rf=%d", s->n, s->h, s->w, s->hr, s->wr, s->rf);
    return SPEI_ERR_ARG;
  }
  if (s->c3 != 128 || s->c2 != 64 || s->c1 != 32) {
    set_error("unsupported channels c3=%d c2=%d c1=%d (this build: 128/64/32, n_feat=32)", s->c3, s->c2, s->c1);
    return SPEI_ERR_ARG;
  }
  if ((long long)s->rf * s->hr * s->wr >= (1ll << 31) || (long long)s->n * s->h * s->w >= (1ll << 31)) {
    set_error("index space exceeds int32"); return SPEI_ERR_ARG;
  }
  if (s->fold_mode < 0 || s->fold_mode > 15 || (s->fold_mode & 12) == 12) { set_error("bad fold_mode %d", s->fold_mode); return SPEI_ERR_ARG; }
  if (s->io_dtype != SPEI_IO_F32 && s->io_dtype != SPEI_IO_BF16) { set_error("bad io_dtype %d", s->io_dtype); return SPEI_ERR_ARG; }
  if (s->search != SPEI_SEARCH_TC && s->search != SPEI_SEARCH_EXACT && s->search != SPEI_SEARCH_TCS) {
    set_error("bad search %d", s->search); return SPEI_ERR_ARG;
  }
  return SPEI_OK;
}

static int check_ptr(const void* p, const char* name, size_t align) {
  if (!p) { set_error("%s is NULL", name); return SPEI_ERR_ARG; }
  if (((uintptr_t)p) % align) { set_error("%s is not %zu-byte aligned", name, align); return SPEI_ERR_ARG; }
  return SPEI_OK;
}

static int prepare(const SpeiShape* shape, void* ws, size_t ws_bytes, Plan* plan) {
  int rc = check_shape(shape);
  if (rc) return rc;
  int sms = 0;
  rc = check_device(&sms);
  if (rc) return rc;
  make_plan(*shape, sms, plan);
  rc = check_ptr(ws, "workspace", 256);
  if (rc) return rc;
  if (ws_bytes < plan->total) {
    set_error("workspace too small: %zu < %zu bytes", ws_bytes, plan->total);
    return SPEI_ERR_WORKSPACE;
  }
  return SPEI_OK;
}

}  // namespace spei

using namespace spei;

extern "C" {

int spei_version(void) { return SPEI_VERSION; }

const char* spei_last_error(void) { return g_err; }

int spei_workspace_bytes(const SpeiShape* shape, size_t* bytes) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!bytes) { set_error("bytes is NULL"); return SPEI_ERR_ARG; }
  int sms = 0;
  rc = check_device(&sms);
  if (rc) return rc;
  Plan p;
  make_plan(*shape, sms, &p);
  *bytes = p.total;
  return SPEI_OK;
}

int spei_stage_norm(const SpeiShape* shape, const void* q, const void* k, void* workspace, size_t workspace_bytes,
                    void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if ((rc = check_ptr(q, "q", 16)) || (rc = check_ptr(k, "k", 16))) return rc;
  return launch_stage_norm(p, q, k, (char*)workspace, (cudaStream_t)stream);
}

int spei_relevance_candidates(const SpeiShape* shape, void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  return p.mode == SPEI_SEARCH_TCS ? launch_relevance_tcs(p, shape->eps, (char*)workspace, (cudaStream_t)stream)
                                   : launch_relevance_tc(p, shape->eps, (char*)workspace, (cudaStream_t)stream);
}

int spei_rescore(const SpeiShape* shape, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats, void* workspace,
                 size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if ((rc = check_ptr(S, "S", 4)) || (rc = check_ptr(arg32, "arg32", 4))) return rc;
  return launch_rescore(p, shape->eps, S, arg32, arg64, stats, (char*)workspace, (cudaStream_t)stream);
}

int spei_relevance_argmax(const SpeiShape* shape, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats,
                          void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if ((rc = check_ptr(S, "S", 4)) || (rc = check_ptr(arg32, "arg32", 4))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  if (shape->search == SPEI_SEARCH_EXACT) return launch_exact_all(p, S, arg32, arg64, stats, ws, st);
  const float eps = shape->eps;
  rc = p.mode == SPEI_SEARCH_TCS ? launch_relevance_tcs(p, eps, ws, st) : launch_relevance_tc(p, eps, ws, st);
  if (rc) return rc;
  return launch_rescore(p, eps, S, arg32, arg64, stats, ws, st);
}

int spei_debug_relevance_tile(const SpeiShape* shape, float* acc_out, void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if ((rc = check_ptr(acc_out, "acc_out", 16))) return rc;
  set_debug_acc(acc_out);
  const float eps = shape->eps;
  return p.mode == SPEI_SEARCH_TCS ? launch_relevance_tcs(p, eps, (char*)workspace, (cudaStream_t)stream)
                                   : launch_relevance_tc(p, eps, (char*)workspace, (cudaStream_t)stream);
}

int spei_debug_error_flag(const SpeiShape* shape, void* workspace, size_t workspace_bytes, void* stream, int32_t* host_out) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if (!host_out) { set_error("host_out is NULL"); return SPEI_ERR_ARG; }
  SPEI_CUDA(cudaMemcpyAsync(host_out, (char*)workspace + p.off_errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SPEI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return SPEI_OK;
}

int spei_debug_search_cycles(const SpeiShape* shape, void* workspace, size_t workspace_bytes, void* stream, int64_t* host_out2) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if (!host_out2) { set_error("host_out2 is NULL"); return SPEI_ERR_ARG; }
  host_out2[1] = 0;
  SPEI_CUDA(cudaMemcpyAsync(host_out2, (char*)workspace + p.off_errflag + 8, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SPEI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return SPEI_OK;
}

int spei_plan_info(const SpeiShape* shape, int32_t* out16) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!out16) { set_error("out16 is NULL"); return SPEI_ERR_ARG; }
  int sms = 0;
  if ((rc = check_device(&sms))) return rc;
  Plan p;
  make_plan(*shape, sms, &p);
  const int32_t v[16] = {p.q.orient, p.q.tu, p.q.tv, p.q.Upad, p.q.Vpad, p.k.orient, p.k.tu, p.k.tv, p.k.tile_v, p.k.Upad, p.k.Vpad,
                         p.QT, p.KT, p.G * (p.pair + 1), p.maxseg, sms};   // [13] = persistent CTAs (two per work range with CTA pairs)
  memcpy(out16, v, sizeof(v));
  return SPEI_OK;
}

// One pyramid level of the transfer.  lv1 (32 channels, 16-byte runs already) gathers straight from the
// planar input; lv3 / lv2 gather from a channels-last copy (512-byte runs), which for lv3 is the fp32 copy
// spei_stage_norm already made whenever ref_lv3 is the searched tensor itself (speinet.py:135).
static int gather_level(const Plan& p, const SpeiShape* shape, int level, const int32_t* arg32, const void* ref, void* out,
                        const void* staged_k, char* ws, cudaStream_t st) {
  const int scale = level == 3 ? 1 : (level == 2 ? 2 : 4);
  const int c = level == 3 ? shape->c3 : (level == 2 ? shape->c2 : shape->c1);
  if (level == 1)
    return launch_gather_fold(shape->n, shape->rf, c, shape->h, shape->w, shape->hr, shape->wr, scale, shape->fold_mode, arg32, ref,
                              ws + p.off_ref1c, (int*)(ws + p.off_gmode), out, p.io_bf16, st);
  const float* src;
  if (level == 3 && staged_k != nullptr && ref == staged_k) {
    src = (const float*)(ws + p.off_k32);
  } else {
    float* dst = (float*)(ws + (level == 3 ? p.off_ref3n : p.off_ref2n));
    int rc = launch_stage_ref_nhwc(ref, p.io_bf16, shape->n * shape->rf, c, scale * shape->hr, scale * shape->wr, dst, st);
    if (rc) return rc;
    src = dst;
  }
  return launch_gather_fold_nhwc(shape->n, shape->rf, c, shape->h, shape->w, shape->hr, shape->wr, scale, shape->fold_mode, arg32, src,
                                 out, p.io_bf16, st);
}

int spei_gather_fold(const SpeiShape* shape, int level, const int32_t* arg32, const void* ref, void* out, const void* staged_k,
                     void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if (level < 1 || level > 3) { set_error("level must be 1, 2 or 3 (got %d)", level); return SPEI_ERR_ARG; }
  if ((rc = check_ptr(arg32, "arg32", 4)) || (rc = check_ptr(ref, "ref", 16)) || (rc = check_ptr(out, "out", 16))) return rc;
  return gather_level(p, shape, level, arg32, ref, out, staged_k, (char*)workspace, (cudaStream_t)stream);
}

int spei_fuse_level(int32_t n, int32_t c, int32_t h, int32_t w, int32_t scale, const float* dec, const float* t,
                    const float* S, const float* weight, const float* bias, float* out, void* stream) {
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (n < 1 || h < 1 || w < 1 || (scale != 1 && scale != 2 && scale != 4)) {
    set_error("bad fuse_level dims n=%d h=%d w=%d scale=%d", n, h, w, scale); return SPEI_ERR_ARG;
  }
  if (c != 128 && c != 64 && c != 32) { set_error("fuse_level: c must be 128, 64 or 32 (got %d)", c); return SPEI_ERR_ARG; }
  // dec / t / out: any float alignment (bases that are not 16-byte aligned take the LDG-fed kernel instead of TMA)
  if ((rc = check_ptr(dec, "dec", 4)) || (rc = check_ptr(t, "t", 4)) || (rc = check_ptr(S, "S", 4)) ||
      (rc = check_ptr(weight, "weight", 16)) || (rc = check_ptr(bias, "bias", 4)) || (rc = check_ptr(out, "out", 4)))
    return rc;
  if (out == dec || out == t) { set_error("fuse_level: out must not alias an input"); return SPEI_ERR_ARG; }
  return launch_fuse_level(n, c, h, w, scale, dec, t, S, weight, bias, out, (cudaStream_t)stream);
}

int spei_fuse_level_bf16(int32_t n, int32_t c, int32_t h, int32_t w, int32_t scale, const void* dec, const void* t,
                         const float* S, const float* weight, const float* bias, void* out, void* stream) {
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (n < 1 || h < 1 || w < 1 || (scale != 1 && scale != 2 && scale != 4)) {
    set_error("bad fuse_level dims n=%d h=%d w=%d scale=%d", n, h, w, scale); return SPEI_ERR_ARG;
  }
  if (c != 128 && c != 64 && c != 32) { set_error("fuse_level: c must be 128, 64 or 32 (got %d)", c); return SPEI_ERR_ARG; }
  if ((rc = check_ptr(dec, "dec", 16)) || (rc = check_ptr(t, "t", 16)) || (rc = check_ptr(S, "S", 4)) ||
      (rc = check_ptr(weight, "weight", 16)) || (rc = check_ptr(bias, "bias", 4)) || (rc = check_ptr(out, "out", 16)))
    return rc;
  if (out == dec || out == t) { set_error("fuse_level: out must not alias an input"); return SPEI_ERR_ARG; }
  return launch_fuse_level_bf16(n, c, h, w, scale, dec, t, S, weight, bias, out, (cudaStream_t)stream);
}

int spei_rl_deconv(int32_t n, int32_t c, int32_t h, int32_t w, int32_t ks, int32_t num_iterations, float regularization_strength,
                   const float* image, const float* blur_kernel, float* out, void* stream) {
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (n < 1 || c < 1 || h < 1 || w < 1) { set_error("bad rl_deconv dims n=%d c=%d h=%d w=%d", n, c, h, w); return SPEI_ERR_ARG; }
  if ((rc = check_ptr(image, "image", 4)) || (rc = check_ptr(blur_kernel, "blur_kernel", 4)) || (rc = check_ptr(out, "out", 4))) return rc;
  if (out == image) { set_error("rl_deconv: out must not alias image"); return SPEI_ERR_ARG; }
  return launch_rl_deconv(n, c, h, w, ks, num_iterations, regularization_strength, image, blur_kernel, out, (cudaStream_t)stream);
}

int spei_conv1x1(int32_t n, int32_t cin, int32_t cout, int64_t pixels, const float* x, const float* weight, float* y, void* stream) {
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (n < 1 || cin < 1 || cout < 1 || pixels < 1) { set_error("bad conv1x1 dims n=%d cin=%d cout=%d pixels=%lld", n, cin, cout, (long long)pixels); return SPEI_ERR_ARG; }
  if ((rc = check_ptr(x, "x", 4)) || (rc = check_ptr(weight, "weight", 4)) || (rc = check_ptr(y, "y", 4))) return rc;
  if ((const void*)y == (const void*)x) { set_error("conv1x1: y must not alias x"); return SPEI_ERR_ARG; }
  return launch_conv1x1(n, cin, cout, pixels, x, weight, y, (cudaStream_t)stream);
}

int spei_upsample2_bias_act(int32_t n, int32_t c, int32_t h, int32_t w, const float* y, const float* bias, int32_t relu, float* out,
                            void* stream) {
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (n < 1 || c < 1 || h < 1 || w < 1) { set_error("bad upsample2_bias_act dims n=%d c=%d h=%d w=%d", n, c, h, w); return SPEI_ERR_ARG; }
  if ((rc = check_ptr(y, "y", 4)) || (rc = check_ptr(out, "out", 8))) return rc;
  if ((const void*)out == (const void*)y) { set_error("upsample2_bias_act: out must not alias y"); return SPEI_ERR_ARG; }
  return launch_upsample2_bias_act(n, c, h, w, y, bias, relu, out, (cudaStream_t)stream);
}

int spei_search_transfer(const SpeiShape* shape, const void* q, const void* k, const void* ref1, const void* ref2,
                         const void* ref3, float* S, void* T3, void* T2, void* T1, int64_t* arg, int32_t* stats,
                         void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = prepare(shape, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if ((rc = check_ptr(q, "q", 16)) || (rc = check_ptr(k, "k", 16)) || (rc = check_ptr(S, "S", 4))) return rc;
  if ((ref1 == nullptr) != (T1 == nullptr) || (ref2 == nullptr) != (T2 == nullptr) || (ref3 == nullptr) != (T3 == nullptr)) {
    set_error("each pyramid level needs both its ref and its T pointer (or neither)");
    return SPEI_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int32_t* arg32 = (int32_t*)(ws + p.off_arg32);
  // (a) SearchTransfer.py:26-31
  if ((rc = launch_stage_norm(p, q, k, ws, st))) return rc;
  // (b) SearchTransfer.py:33-34
  if ((rc = spei_relevance_argmax(shape, S, arg32, arg, stats, workspace, workspace_bytes, stream))) return rc;
  // (c) SearchTransfer.py:36-46
  struct Lvl { const void* ref; void* out; int level; };
  const Lvl lv[3] = {{ref3, T3, 3}, {ref2, T2, 2}, {ref1, T1, 1}};
  for (const Lvl& l : lv) {
    if (!l.ref) continue;
    if ((rc = check_ptr(l.ref, "ref", 16)) || (rc = check_ptr(l.out, "T", 16))) return rc;
    if ((rc = gather_level(p, shape, l.level, arg32, l.ref, l.out, k, ws, st))) return rc;
  }
  return SPEI_OK;
}

}  // extern "C"
