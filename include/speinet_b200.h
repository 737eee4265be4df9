/*
 * speinet_b200.h -- C-ABI of the B200-native SearchTransfer hot path.
 *
 * Drop-in boundary for yangt1013/SPEINet (reference paths relative to /root/reference):
 *   model/SearchTransfer.py:24-51   SearchTransfer.forward   -> spei_search_transfer
 *   model/SearchTransfer.py:59-72   SelfTransfer search half -> spei_search_transfer with NULL pyramids
 *   model/speinet.py:93-94,96-97,108-109  _decode fusion     -> spei_fuse_level
 *   model/rcl.py:22-51              r_l_per_channel          -> spei_rl_deconv
 *   relu(conv1x1(bicubic_x2(x))) chains (SearchTransfer.py:70-76, speinet.py:99-100,111-112) -> spei_conv1x1 + spei_upsample2_bias_act
 *
 * Conventions
 *   - extern "C", plain pointers + sizes + a CUDA stream handle (void*, a cudaStream_t).  No torch types.
 *   - every pointer is DEVICE memory owned by the caller; tensors are contiguous NCHW, fp32 unless SpeiShape.io_dtype /
 *     the _bf16 entry points say bf16.
 *   - the library never allocates persistent device memory, never frees caller memory and never
 *     synchronises the device; all work is enqueued on `stream`.
 *   - return value 0 = ok, negative = error (SPEI_ERR_*); spei_last_error() returns a thread-local message.
 *   - sm_100a only.  There is no CPU path and no fallback: on any other device the calls fail.
 */
#ifndef SPEINET_B200_H_
#define SPEINET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPEI_VERSION 200

#define SPEI_OK 0
#define SPEI_ERR_ARG (-1)        /* NULL pointer / bad shape / misaligned pointer            */
#define SPEI_ERR_ARCH (-2)       /* current device is not compute capability 10.x            */
#define SPEI_ERR_CUDA (-3)       /* a CUDA runtime / driver call failed                      */
#define SPEI_ERR_WORKSPACE (-4)  /* workspace too small                                      */

/* summation order + divisor of the overlap-add (SearchTransfer.py:44-46) */
#define SPEI_FOLD_ORDER_CPU 1 /* bit 0: add in ascending (ki,kj) (torch CPU col2im) instead of ascending patch origin (CUDA col2im) */
#define SPEI_FOLD_TRUE_DIV 2  /* bit 1: true division by 9 instead of x * (1.0f/9.0f)          */
#define SPEI_FOLD_CUDA 0      /* what torch does on a CUDA device: CUDA col2im order, x*(1/9f) */
#define SPEI_FOLD_CPU 3       /* what torch does on the CPU: CPU col2im order, x/9             */
/* Finest level (lv1) of spei_gather_fold: source layout.  Default (neither bit): chosen on the device from the match field
 * (planar input when coherent or when the level is small, a cell-major copy when scattered); results are identical. */
#define SPEI_FOLD_LV1_PLANAR 4 /* bit 2: always gather from the caller's planar tensor           */
#define SPEI_FOLD_LV1_CELLS 8  /* bit 3: always re-tile ref_lv1 cell-major first                 */

/* element type of the feature tensors that cross the boundary */
#define SPEI_IO_F32 0
#define SPEI_IO_BF16 1 /* q, k, ref1/2/3 and T3/T2/T1 are bf16 (north_star's 1e-2 mode): every kernel reads / writes bf16 natively,
                          arithmetic stays fp32 (search candidates bf16 -- now EXACT operands --, rescoring / fold sums fp32).
                          S, arg and stats keep their types. */

/* relevance engine */
#define SPEI_SEARCH_TC 0    /* tcgen05 bf16 candidate pass, dense 9-tap implicit GEMM, + exact fp32 rescoring */
#define SPEI_SEARCH_EXACT 1 /* fp32 CUDA-core exhaustive search (checker / debugging)        */
#define SPEI_SEARCH_TCS 2   /* tcgen05 bf16 candidate pass with tap sharing: the MMA contracts channels x the 3 taps
                               along one image axis (K = 384), the epilogue adds the 3 taps along the other axis from
                               neighbouring accumulator entries -- 2.3x fewer tensor-core flops for the same scores;
                               same exact fp32 rescoring behind it.  What the Python modules use by default. */

/* Problem description.  The reference instantiates c3=128, c2=64, c1=32 (n_feat=32,
 * speinet.py:53); lv2 / lv1 tensors are 2x / 4x the lv3 grids (SearchTransfer.py:36-38,44-46). */
typedef struct SpeiShape {
  int32_t n;         /* batch items (clips)                                                    */
  int32_t h, w;      /* query grid  (lrsr_lv3 is [n, c3, h, w])                                 */
  int32_t hr, wr;    /* reference grid per sharp frame (refsr_lv3 is [n, rf, c3, hr, wr])       */
  int32_t rf;        /* sharp reference frames per item; 1 in the reference      */
  int32_t c3, c2, c1;/* channels of lv3 / lv2 / lv1 (must be 128 / 64 / 32 in this version)    */
  int32_t fold_mode; /* SPEI_FOLD_* (bit mask)                                                 */
  int32_t search;    /* SPEI_SEARCH_*                                                          */
  float eps;         /* candidate window of the bf16 pass in normalised relevance units.
                        <= 0 (default): CERTIFIED window.  The staging pass measures the bf16 rounding residual of every query
                        and key patch; by Cauchy-Schwarz |bf16 score - exact score| <= Delta_i for every key of query i
                        (Delta_i = 1.01 (d_i + (1 + d_i) max_j d_j) + 1.6e-4, typically 3e-3 - 4e-3).  The exact rescoring
                        keeps every key whose bf16 score is >= E - Delta_i (E = exact relevance of the best bf16 candidate):
                        the true argmax cannot be outside that set.  The tensor-core pass keeps every key within
                        Delta_i + 1e-3 of the best bf16 score; a query whose set is not covered by what was kept (its
                        candidate list overflowed, or E is more than 1e-3 below the best bf16 score) takes a second
                        tensor-core pass that enumerates the set completely.
                        > 0: fixed window eps (round-1 behaviour with 2e-3; cheaper, exact only while no bf16 score is
                        further than eps/2 from its exact value -- NOT certified). */
  int32_t io_dtype;  /* SPEI_IO_F32 (0) or SPEI_IO_BF16                                          */
} SpeiShape;

/* Counters written by spei_search_transfer / spei_relevance_argmax / spei_rescore into caller memory (device,
 * SPEI_STATS_WORDS x int32) when `stats` is non-NULL:
 *  [0] queries whose candidate list saturated (handled by the second tensor-core pass)
 *  [1] (query, key) pairs rescored exactly in fp32/fp64
 *  [2] max |bf16 candidate score - exact score| x 1e9 over the candidates rescored in the first round
 *  [3] != 0: the second pass ran out of capacity and the queued queries took the exhaustive fp32 search (slow, still exact)
 *  [4] reserved
 *  [5] (query, key) pairs emitted by the second pass
 *  [6] certified mode: rescored candidates whose bf16 score was further than Delta_i from the exact one.  MUST be 0;
 *      anything else means the bound does not hold on this device and the argmax is not certified.
 *  [7] reserved */
#define SPEI_STATS_WORDS 8

int spei_version(void);
const char *spei_last_error(void);

/* Bytes of scratch the pipeline needs for `shape` (depends on the SM count of the current device). */
int spei_workspace_bytes(const SpeiShape *shape, size_t *bytes);

/*
 * The whole of SearchTransfer.forward (SearchTransfer.py:24-51):
 *   q      [n, c3, h, w]              lrsr_lv3
 *   k      [n, rf, c3, hr, wr]        refsr_lv3 (one or more sharp frames)
 *   ref1   [n, rf, c1, 4hr, 4wr]      ref_lv1   (may be NULL together with T1)
 *   ref2   [n, rf, c2, 2hr, 2wr]      ref_lv2   (may be NULL together with T2)
 *   ref3   [n, rf, c3, hr, wr]        ref_lv3   (may be NULL together with T3; may alias k)
 *   S      [n, 1, h, w]   fp32        R_lv3_star viewed as a map           (:49)
 *   T3/T2/T1 [n, c3, h, w] / [n, c2, 2h, 2w] / [n, c1, 4h, 4w]             (:44-46)
 *   arg    [n, h*w] int64             R_lv3_star_arg, key index j = f*hr*wr + y*wr + x   (:34)  (may be NULL)
 *   stats  SPEI_STATS_WORDS x int32 device counters (may be NULL)
 */
int spei_search_transfer(const SpeiShape *shape, const void *q, const void *k, const void *ref1,
                         const void *ref2, const void *ref3, float *S, void *T3, void *T2, void *T1,
                         int64_t *arg, int32_t *stats, void *workspace, size_t workspace_bytes, void *stream);

/* ---- individual stages (the pipeline above is exactly these, in this order) ---- */

/* (a) unfold + normalize pre-pass (SearchTransfer.py:26-31) without materialising the unfolded
 * tensors: stages q and k into the layouts the search kernels read and computes the per-patch
 * reciprocal L2 norms.  Fills the staging part of `workspace`. */
int spei_stage_norm(const SpeiShape *shape, const void *q, const void *k, void *workspace,
                    size_t workspace_bytes, void *stream);

/* (b) relevance bmm + max/argmax (SearchTransfer.py:33-34) on the staged operands.
 * Writes S [n,1,h,w] and arg32 [n, h*w] int32 (and arg64 if non-NULL). */
int spei_relevance_argmax(const SpeiShape *shape, float *S, int32_t *arg32, int64_t *arg64, int32_t *stats,
                          void *workspace, size_t workspace_bytes, void *stream);

/* (b) split in its two halves, for callers that time or schedule them separately:
 *   spei_relevance_candidates : the tcgen05 bf16 pass; leaves per-query candidate lists in `workspace`
 *   spei_rescore              : exact fp32 rescoring + tie-break + exhaustive fallback -> S, arg32 (, arg64) */
int spei_relevance_candidates(const SpeiShape *shape, void *workspace, size_t workspace_bytes, void *stream);
int spei_rescore(const SpeiShape *shape, float *S, int32_t *arg32, int64_t *arg64, int32_t *stats, void *workspace,
                 size_t workspace_bytes, void *stream);

/* (c) unfold(ref) + bis gather + fold + /9 of one pyramid level (SearchTransfer.py:36-46),
 * level = 3, 2 or 1; ref is [n, rf, c, s*hr, s*wr], out is [n, c, s*h, s*w], s = 1, 2, 4.
 * `arg32` is [n, h*w] int32 key indices (any values in [0, rf*hr*wr)).
 * lv3 / lv2 gather from a channels-last copy kept in `workspace`; `staged_k` is the refsr_lv3 pointer
 * last given to spei_stage_norm with this workspace (or NULL): when `ref == staged_k` at level 3 the copy
 * made there is reused (ref_lv3 and refsr_lv3 are the same tensor at speinet.py:135). */
int spei_gather_fold(const SpeiShape *shape, int level, const int32_t *arg32, const void *ref, void *out,
                     const void *staged_k, void *workspace, size_t workspace_bytes, void *stream);

/* (d) one fusion line of SPEINet._decode (speinet.py:93-94 / 96-97 / 108-109):
 *   out = dec + (W . cat(dec, t) + b) * bicubic_up(S, scale),  scale in {1,2,4}
 *   dec, t, out [n, c, scale*h, scale*w]; weight [c, 2c] (Conv2d 1x1 weight, speinet.py:55-57);
 *   bias [c]; S [n, 1, h, w].  out may alias neither input. */
int spei_fuse_level(int32_t n, int32_t c, int32_t h, int32_t w, int32_t scale, const float *dec, const float *t,
                    const float *S, const float *weight, const float *bias, float *out, void *stream);

/* The same fusion line with bf16 dec / t / out (S, weight, bias stay fp32; fp32 accumulation, one rounding at the store).
 * TMA-fed kernel only: scale*h * scale*w must be a multiple of 8 and dec / t / out / weight 16-byte aligned. */
int spei_fuse_level_bf16(int32_t n, int32_t c, int32_t h, int32_t w, int32_t scale, const void *dec, const void *t,
                         const float *S, const float *weight, const float *bias, void *out, void *stream);

/* ---- next to the path ---- */

/* The Richardson-Lucy edge prior, model/rcl.py:22-51 r_l_per_channel(image_tensor, blur_kernel, num_iterations,
 * regularization_strength) (call sites speinet.py:81 with 1 iteration, :129 / :141 with 5), all iterations of all
 * channels in one kernel:
 *   image       [n, c, h, w] fp32      the frame (channels are processed independently, rcl.py:27-28)
 *   blur_kernel [ks, ks] fp32 (device) the [1,1,ks,ks] weight of rcl.py:33 (create_blur_kernel: 5x5 box, rcl.py:18-20);
 *                                      ks in {3, 5, 7}
 *   out         [n, c, h, w] fp32      must not alias image
 * NaN -> 0 and negative -> 0 on the correction factor exactly as rcl.py:39-40. */
int spei_rl_deconv(int32_t n, int32_t c, int32_t h, int32_t w, int32_t ks, int32_t num_iterations,
                   float regularization_strength, const float *image, const float *blur_kernel, float *out, void *stream);

/* y = W . x, the channel mix of a 1x1 convolution WITHOUT its bias (fp32 FMA): the first half of the
 * `relu(conv1x1(F.interpolate(x, scale_factor=2, mode='bicubic')))` chains, applied at low resolution (see below).
 *   x [n, cin, pixels] fp32, weight [cout, cin] fp32 (Conv2d 1x1 weight), y [n, cout, pixels] fp32; cout in {8,16,32,64,128} */
int spei_conv1x1(int32_t n, int32_t cin, int32_t cout, int64_t pixels, const float *x, const float *weight, float *y,
                 void *stream);

/* out = act(bicubic_x2(y) + bias): the second half of the `relu(conv1x1(F.interpolate(x, scale_factor=2, mode='bicubic')))`
 * chains of SelfTransfer (model/SearchTransfer.py:70-76) and _decode (model/speinet.py:99-100, 111-112).  A 1x1 convolution
 * and a per-channel resize commute, so the caller applies the channel mix W . x at LOW resolution (spei_conv1x1) and passes
 * the result as y:
 *   y    [n, c, h, w] fp32      W . x (no bias)
 *   bias [c] fp32 or NULL       the convolution's bias, added after the resize
 *   relu 0 / 1                  activation
 *   out  [n, c, 2h, 2w] fp32 */
int spei_upsample2_bias_act(int32_t n, int32_t c, int32_t h, int32_t w, const float *y, const float *bias, int32_t relu,
                            float *out, void *stream);

/* ---- diagnostics (used by tests and tools; not part of the reference-facing path) ---- */

/* Runs the tcgen05 relevance kernel on the staged operands and additionally copies the raw fp32
 * accumulator of (item 0, query tile 0, key tile 0) -- un-normalised bf16 dot products, row m =
 * query m of the tile, column c = key c of the tile -- into acc_out [128][256] (device). */
int spei_debug_relevance_tile(const SpeiShape *shape, float *acc_out, void *workspace, size_t workspace_bytes,
                              void *stream);

/* Copies the pipeline watchdog word of `workspace` to *host_out and synchronises `stream`.
 * 0 = no barrier of the tcgen05 kernels timed out.  (A starved barrier traps: the launch fails with a CUDA error and
 * every later call of the process reports it; the word is readable only in -DSPEI_WATCHDOG_DRAIN debug builds.) */
int spei_debug_error_flag(const SpeiShape *shape, void *workspace, size_t workspace_bytes, void *stream,
                          int32_t *host_out);

/* host_out2[0] = clock64 span (SM cycles) of CTA 0 of the last tcgen05 search launch on `workspace`; synchronises
 * `stream`.  Divided by the event-timed duration of that launch it gives the SM clock the kernel really ran at. */
int spei_debug_search_cycles(const SpeiShape *shape, void *workspace, size_t workspace_bytes, void *stream,
                             int64_t *host_out2);

/* Tiling plan for `shape` on the current device, 16 x int32 (host):
 * q.orient q.tu q.tv q.Upad q.Vpad  k.orient k.tu k.tv k.Ny k.Upad k.Vpad  QT KT G maxseg num_sms */
int spei_plan_info(const SpeiShape *shape, int32_t *out16);

#ifdef __cplusplus
}
#endif
#endif /* SPEINET_B200_H_ */
