// Shared host/device definitions for libspeinet_b200: tiling plan, workspace layout, error plumbing.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/speinet_b200.h"

#include "fastdiv.h"

namespace spei {

// ---- fixed geometry of the search -------------------------------------------------------------
constexpr int kC3 = 128;          // lv3 channels (n_feat*4, speinet.py:53)
constexpr int kCG = kC3 / 8;      // 16 channel groups of 8 bf16 = 16 bytes
constexpr int kTileU = 8;         // tile extent along the fast (contiguous) image axis, both operands
constexpr int kQTileV = 16;       // query tile: 8 x 16 positions = 128 MMA rows
constexpr int kMaxNy = 32;        // key tile: 8 x Ny positions = up to 256 MMA columns
#ifndef SPEI_TOPK
#define SPEI_TOPK 8
#endif
// bf16-pass candidates kept per query, key segment and epilogue group (multiple of 4).  8 since round 2: a list that cannot
// hold a query's in-window keys is no longer a 133-MFLOP exhaustive search but a slot in the second tcgen05 pass, so the
// shorter sorted insertion wins (720p, certified window: randn -1 %, image-like features -7 %, x16 low-noise -10 % of
// search + exactness time against 16; profiles/r02_window_cost.json)
constexpr int kTopK = SPEI_TOPK;
// tap-sharing search (relevance_tcs.cu): tiles are 32 wide along u (30 interior + a 1-position halo each side,
// the u taps are summed in the epilogue), queries 4 rows, keys up to 8 rows
constexpr int kSTileU = 30;       // interior tile extent along u
constexpr int kSBoxU = 32;        // shared-memory tile extent along u = one 512-byte row = one warp of TMEM lanes
constexpr int kSQTileV = 4;       // query tile: 32 x 4 positions = 128 MMA rows (120 of them interior)
constexpr int kSMaxNy = 8;        // key tile: 32 x Ny positions = up to 256 MMA columns
constexpr int kCGS = 4;           // channel groups per key pipeline stage (32 channels = 2 x K16)
constexpr int kStages = 6;        // key pipeline depth (1.5 key tiles in flight)
// second tcgen05 pass over the queries whose candidate lists saturated (relevance_flagged.cu)
constexpr int kFlagNy = 32;           // key tile: 8 x 32 positions = 256 MMA columns
constexpr int kFlagMaxRows = 16384;   // packed query rows per call (128 tiles of 128); beyond that the exhaustive search takes over
constexpr int kFlagMaxEmit = 1 << 20; // (query, key) pairs the pass may emit per call
constexpr int kStatsWords = 8;        // int32 counters of spei_search_transfer / spei_rescore (include/speinet_b200.h)
// slack added to the certified bf16-score error bound: fp32 accumulation of 1152 products in tensor memory (n * 2^-23 at
// worst) plus the epilogue's tap-sum adds, relative to a normalised score of magnitude <= 1
constexpr float kAccSlack = 1.6e-4f;

// One operand (query set or key set) staged in (u, v) coordinates: u is the fast axis in memory.
// orient 0: u = x, v = y.   orient 1: u = y, v = x (image transposed so the tile grid wastes less).
struct OperandPlan {
  int orient;
  int U, V;        // valid extent
  int tile_u;      // tile stride along u (8 dense; 14 tap-sharing)
  int tile_v;      // tile height along v (16 / 8 for queries, Ny for keys)
  int tu, tv;      // tiles along u / v
  int Upad, Vpad;  // staged bf16 plane = [Vpad][Upad] pixels, 1-pixel zero border included
  __host__ __device__ int tiles() const { return tu * tv; }
};

struct Plan {
  int mode;        // SPEI_SEARCH_TC (dense 9-tap MMA) or SPEI_SEARCH_TCS (tap-sharing); decides the operand tiling
  int io_bf16;     // SpeiShape.io_dtype == SPEI_IO_BF16: q / k / ref* / T* are bf16 tensors
  int nlist;       // candidate lists per (query, key segment): 1 dense, 2 tap-sharing (two epilogue warp groups)
  int n, rf;
  int H, W, Hr, Wr;
  OperandPlan q, k;
  int QT;          // query tiles per item
  int pair;        // 1: the tap-sharing kernel runs on CTA pairs (cta_group::2), one work step = two query tiles x one key tile
  int QTs;         // query work slots per item: QT, or ceil(QT / 2) with CTA pairs
  int KT;          // key tiles per item (all reference frames)
  long long P;     // total work steps = n*QTs*KT
  int G;           // persistent CTAs (CTA pairs when pair = 1)
  int maxseg;      // max key segments a query tile is split into
  // workspace offsets (bytes)
  size_t off_qbf, off_kbf, off_q32, off_k32, off_rq, off_rk, off_rkpad, off_qss, off_kss;
  size_t off_qrs, off_krs;     // per-pixel energy of the bf16 rounding residual  sum_c (x - bf16(x))^2
  size_t off_dq, off_dkmax;    // per-query relative residual norm of its patch; per-item maximum over the keys (float bits)
  size_t off_cval, off_cidx, off_flag, off_packed, off_counters, off_arg32, off_errflag, off_ref3n, off_ref2n, off_ref1c, off_gmode;
  // second pass (relevance_flagged.cu): packed A operand, per-row threshold / query id, emitted pairs
  size_t off_thr, off_apack, off_prow_thr, off_prow_q, off_emit_q, off_emit_k;
  int flag_rows;               // capacity of the packed A operand in query rows (multiple of 128)
  size_t total;
};

int make_plan(const SpeiShape& s, int num_sms, Plan* out);

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define SPEI_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return ::spei::cuda_fail(e__, #call); \
  } while (0)

// ---- stage launchers (defined in the .cu files) -----------------------------------------------
int launch_stage_norm(const Plan& p, const void* q, const void* k, char* ws, cudaStream_t st);
int launch_relevance_tc(const Plan& p, float eps, char* ws, cudaStream_t st);
int launch_relevance_tcs(const Plan& p, float eps, char* ws, cudaStream_t st);
int tcs_epilogue_groups();   // candidate lists per (query, key segment) the tap-sharing kernel writes
bool tcs_cta_pairs();        // the tap-sharing kernel was built for CTA pairs (cta_group::2)
void set_debug_acc(float* ptr);
int launch_rescore(const Plan& p, float eps, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats, char* ws,
                   cudaStream_t st);
int launch_relevance_flagged(const Plan& p, int32_t* stats, char* ws, cudaStream_t st);
// candidate window of the bf16 pass: eps > 0 = fixed (uncertified) window in normalised relevance units; eps <= 0 = the
// certified per-query window 2 * Delta_i (Delta_i bounds |bf16 score - exact score| for every key, from the measured rounding
// residuals of the query patch and of the worst key patch, Cauchy-Schwarz; DESIGN.md section 4(b'))
__host__ __device__ inline float certified_delta(float dq, float dkmax) {
  return 1.01f * (dq + (1.f + dq) * dkmax) + kAccSlack;
}
// Window of the tcgen05 pass in certified mode: Delta + kWindowMargin below a list's best bf16 score.  The rescoring needs
// every key with bf16 score >= E - Delta (E = exact relevance of the best bf16 candidate b); the lists hold all keys above
// b - (Delta + margin), which covers that set whenever E >= b - margin.  A query whose best candidate's bf16 score is off by
// more than the margin (rigorously possible up to Delta, measured <= 6e-4) is queued for the second pass like a saturated one,
// so the result stays certified while the common case pays for a 1.6x narrower window than the worst-case 2 * Delta.
#ifndef SPEI_WINDOW_MARGIN
#define SPEI_WINDOW_MARGIN 1.0e-3f
#endif
constexpr float kWindowMargin = SPEI_WINDOW_MARGIN;
__host__ __device__ inline float certified_window(float delta) { return delta + kWindowMargin; }
int launch_exact_all(const Plan& p, float* S, int32_t* arg32, int64_t* arg64, int32_t* stats, char* ws, cudaStream_t st);
int launch_gather_fold(int n, int rf, int c, int h, int w, int hr, int wr, int scale, int fold_mode, const int32_t* arg32,
                       const void* ref, void* ref_cells, int* mode, void* out, int io_bf16, cudaStream_t st);
int launch_stage_ref_nhwc(const void* ref, int in_bf16, int nimg, int C, int Hs, int Ws, float* dst, cudaStream_t st);
int launch_gather_fold_nhwc(int n, int rf, int c, int h, int w, int hr, int wr, int scale, int fold_mode,
                            const int32_t* arg32, const float* ref_nhwc, void* out, int out_bf16, cudaStream_t st);
int launch_fuse_level_bf16(int n, int c, int h, int w, int scale, const void* dec, const void* t, const float* S,
                           const float* weight, const float* bias, void* out, cudaStream_t st);
int launch_fuse_level(int n, int c, int h, int w, int scale, const float* dec, const float* t, const float* S,
                      const float* weight, const float* bias, float* out, cudaStream_t st);
// counters at workspace + off_counters (int32): [0, n) queries queued per item for the second pass; then
constexpr int kCntEmit = 0;       // + n: pairs emitted by the second pass
constexpr int kCntExhaust = 1;    // + n: != 0 -> capacity exceeded, the queued queries take the exhaustive fp32 search
constexpr int kCntWords = 8;

int launch_rl_deconv(int n, int c, int h, int w, int ks, int iters, float lambda, const float* img, const float* kern, float* out,
                     cudaStream_t st);

int launch_conv1x1(int n, int cin, int cout, long long P, const float* x, const float* w, float* y, cudaStream_t st);
int launch_upsample2_bias_act(int n, int c, int h, int w, const float* y, const float* bias, int relu, float* out, cudaStream_t st);

// position <-> index helpers shared by kernels
__host__ __device__ inline int uv_to_linear(int orient, int u, int v, int W) {
  return orient == 0 ? v * W + u : u * W + v;  // y*W + x
}

}  // namespace spei
