/*
 * vnlb_b200.h -- C ABI of libvnlb_b200.so: the B200 (sm_100a) implementation of
 * the gauenk/vnlb `vnlb.denoise` hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; every pointer is DEVICE memory unless it is a
 *     parameter struct (host memory, read during the call) ;
 *   - the caller allocates every buffer, including the workspace reported by
 *     the matching *_workspace_bytes() (16-byte aligned device memory, private
 *     to one stream at a time); the library never calls cudaMalloc / cudaFree;
 *   - stream-ordered and asynchronous: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*, NULL = default stream); no entry point
 *     synchronises the device or a stream, so every call can be captured in a
 *     CUDA graph;
 *   - return value 0 on success, VNLB_ERR_* (< 0) otherwise;
 *     vnlb_last_error() returns a thread-local description of the last failure;
 *   - re-entrant per stream; one host thread per device.
 *   - images are float32 [T,C,H,W] (C <= 4), the index codec everywhere is
 *         ind = t*C*H*W + y*W + x
 *     = ravelled offset of channel 0 of the patch's top-left-front corner
 *     (reference: lib/vnlb/search_mask/mask.py:69-71, lib/vnlb/agg/comp_agg.py:119-121);
 *   - a row of `inds` ([B,K] int64) is VALID iff none of its K entries is -1
 *     (reference: lib/vnlb/proc_nl.py:160-177 get_valid_patches /
 *     fill_valid_patches).  Kernels skip invalid rows on the device, which
 *     replaces the reference's host-side row compaction.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository root).
 */
#ifndef VNLB_B200_H
#define VNLB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNLB_OK 0
#define VNLB_ERR_BAD_ARG (-1)      /* null pointer, bad shape, bad parameter     */
#define VNLB_ERR_UNSUPPORTED (-2)  /* valid request this build cannot serve      */
#define VNLB_ERR_CUDA (-3)         /* CUDA runtime error (see vnlb_last_error)   */
#define VNLB_ERR_WORKSPACE (-4)    /* workspace missing or too small             */

#define VNLB_WINDOW_SHIFT 0 /* search window shifted to stay inside the frame (C++ VNLB) */
#define VNLB_WINDOW_CLIP 1  /* window clipped, out-of-frame candidates dropped           */

/* Parameters of the similarity search: the fields vpss.exec_sim_search_burst
 * reads from the reference's `args` (lib/vnlb/params.py:109-197 shortcuts
 * ps, pt, w_s, nWt_f, nWt_b, npatches). */
typedef struct {
    int32_t ps;          /* spatial patch size  (sizePatch)                        */
    int32_t pt;          /* temporal patch size (sizePatchTime)                    */
    int32_t w_s;         /* spatial search window, odd (sizeSearchWindow)          */
    int32_t nWt_f;       /* frames searched forward  (sizeSearchTimeFwd)           */
    int32_t nWt_b;       /* frames searched backward (sizeSearchTimeBwd)           */
    int32_t k;           /* neighbours kept (nSimilarPatches)                      */
    int32_t dist_chnls;  /* channels entering the distance: 1 in step 1, C in step 2 */
    int32_t window_mode; /* VNLB_WINDOW_SHIFT | VNLB_WINDOW_CLIP                   */
} VnlbSearchParams;

/* Parameters of the Bayes estimate: the fields bayes_est.denoise reads from
 * `args` (lib/vnlb/deno/bayes_est.py:17-62). */
typedef struct {
    int32_t step;            /* 0 = first step, 1 = second step (args.step)          */
    int32_t k;               /* patches per group (n)                                */
    int32_t ps;              /* spatial patch size                                   */
    int32_t pt;              /* temporal patch size                                  */
    int32_t c;               /* channels; each is an independent p x p problem       */
    int32_t rank;            /* eigenpairs kept (args.rank)                          */
    float sigma2;            /* args.sigma2                                          */
    float sigmab2;           /* args.sigmab2 (sigmaBasic^2)                          */
    float thresh;            /* args.thresh (variThres)                              */
    int32_t cov_from_basic;  /* args.cpatches == "basic"                             */
    int32_t eig_method;      /* VNLB_EIG_*                                           */
} VnlbBayesParams;

#define VNLB_EIG_TRIDIAG 0 /* Householder tridiagonalisation + bisection + inverse iteration */
#define VNLB_EIG_JACOBI 1  /* cyclic Jacobi in shared memory (slow, cross-check)             */

const char *vnlb_last_error(void);
int vnlb_version(void);
/* Number of CUDA kernels the library has launched in this process (monotonic; bench.py reports the
 * difference over the timed region as gpu_launches). */
unsigned long long vnlb_kernel_launches(void);

/* rgb2yuv_cpp, lib/vnlb/utils/color.py:52-77 (out of place; src may equal dst). */
int vnlb_rgb2yuv(const float *rgb, float *yuv, int T, int C, int H, int W, void *stream);
/* apply_yuv2rgb, lib/vnlb/utils/color.py:31-50 (src may equal dst). */
int vnlb_yuv2rgb(const float *yuv, float *rgb, int T, int C, int H, int W, void *stream);

/* init_mask -> comp_params -> fill_mask, lib/vnlb/search_mask/mask.py:190-213,
 * 252-288,315-358 (whole-frame case).  mask: int8 [T,H,W], fully written.
 * Row bands [y_begin, y_end) restrict the SET pixels to rows of one partition
 * (multi-GPU sharding; pass 0, H for the reference behaviour). */
int vnlb_init_mask(int8_t *mask, int T, int H, int W, int ps, int pt, int proc_step,
                   int y_begin, int y_end, void *stream);
/* The same lattice for a TILE of a larger frame (multi-GPU: one row band + halo per GPU): `mask` is [T,H,W] and holds
 * rows [y_offset, y_offset + H) of a frame of H_total rows; the lattice phase, the always-set first / last row and
 * the valid range follow the global row, y_begin / y_end are local rows.  The union of the tiles' masks over
 * disjoint [y_begin, y_end) bands equals the whole-frame mask of vnlb_init_mask. */
int vnlb_init_mask_tile(int8_t *mask, int T, int H, int W, int ps, int pt, int proc_step,
                        int y_begin, int y_end, int y_offset, int H_total, void *stream);

/* vpss.exec_sim_search_burst, call site lib/vnlb/search/search.py:86-89.
 * img [T,C,H,W]; qinds int64 [Q,3] = (t,y,x) of each query's patch corner;
 * fflow/bflow [T,2,H,W] (ch0 = dx, ch1 = dy) or both NULL for zero flow;
 * vals float32 [Q,k], inds int64 [Q,k] written in place, ascending by
 * (distance, candidate enumeration order t->y->x); slots that cannot be filled
 * (fewer than k candidates) are set to +inf / -1. */
/* Search kernel selection (A/B measurements and tests; every path returns identical bits): 0 = automatic -- for
 * 7x7x2 patches, a 27x27 window and at most 13 frames the "quad" kernel (4x9 candidates per thread, distances in
 * registers over all channel / patch-frame phases, two-histogram selection), else the 1-column tiled kernel or the
 * generic kernel; 1 = never the quad kernel; 2 = same as 0.  Returns the previous setting.
 * (Environment: VNLB_SEARCH_PATH at start-up.) */
int vnlb_set_search_path(int path);
size_t vnlb_search_workspace_bytes(int Q, const VnlbSearchParams *p);
int vnlb_search_topk(const float *img, int T, int C, int H, int W, const int64_t *qinds, int Q,
                     const float *fflow, const float *bflow, const VnlbSearchParams *p,
                     float *vals, int64_t *inds, void *ws, size_t ws_bytes, void *stream);

/* vpss.fill_patches, call site lib/vnlb/search/search.py:91-98.
 * patches float32 [B,K,pt,C,ps,ps]; entries with ind == -1 are left untouched. */
int vnlb_fill_patches(float *patches, const float *img, const int64_t *inds, int B, int K,
                      int T, int C, int H, int W, int ps, int pt, void *stream);

/* update_mask_inds + agg_boost, lib/vnlb/search_mask/mask.py:37-86,104-187:
 * clears mask at every index of every VALID row and, if boost, at its 4
 * spatial neighbours. */
int vnlb_mask_update(int8_t *mask, const int64_t *inds, int B, int K, int T, int C, int H, int W,
                     int boost, void *stream);

/* Device-side replacement of mask2inds (lib/vnlb/search_mask/mask.py:18-31) for
 * the throughput schedule.  vnlb_count_mask adds the number of set pixels to
 * counters[0].  vnlb_select_queries appends every set pixel whose hash-based
 * uniform draw (seed, round) falls below `prob` to qinds (int64 [cap,3] =
 * (t,y,x)), clears it from the mask, and adds the number drawn to counters[1];
 * draws beyond `cap` are dropped (they stay set in the mask).  counters:
 * uint32 [2], zeroed by the caller. */
int vnlb_count_mask(const int8_t *mask, int T, int H, int W, uint32_t *counters, void *stream);
/* vnlb_pad_queries turns the rows of qinds beyond min(counters[1], cap) into invalid queries (-1,-1,-1):
 * vnlb_search_topk answers them with all -1 rows, which every later kernel skips, so a whole round can be
 * enqueued with `cap` rows before the host knows how many were drawn. */
int vnlb_pad_queries(int64_t *qinds, const uint32_t *counters, int cap, void *stream);
int vnlb_select_queries(int8_t *mask, int T, int H, int W, double prob, uint32_t seed,
                        uint32_t round, int64_t *qinds, int cap, uint32_t *counters, void *stream);

/* Greedy conflict resolution inside a round of the throughput schedule (the reference resolves it by processing
 * sub-batches of 128 sequentially, lib/vnlb/search/search.py:38-64): after vnlb_search_topk and BEFORE
 * vnlb_mask_update, a row of `inds` whose reference pixel (its row of `qinds`) lies in the clear-set (found patches +
 * boost neighbours) of a valid row of the same round with a HIGHER PRIORITY (a hash of the reference pixel and the
 * round: deterministic, unlike the row order) is dropped -- its K indices become -1, so every later
 * kernel skips it -- and its pixel is set in the mask again (a later round draws it if nothing ends up covering it).
 * owner: uint32 [T,H,W] scratch, filled with 0xFFFFFFFF once per step by the caller; round: 0, 1, ... within the step
 * (stamps carry the round, the map is never cleared in between); dropped: uint32 counter, incremented. */
int vnlb_round_dedup(const int64_t *qinds, int64_t *inds, int B, int K, uint32_t *owner, uint32_t round,
                     int8_t *mask, int T, int C, int H, int W, int boost, uint32_t *dropped, void *stream);

/* Device-controlled round of the throughput schedule (what a CUDA graph of a round replays): one call = zero the
 * counters, advance the round index state[0] (uint32 in device memory, zeroed by the caller at the start of a step),
 * count the masked pixels (counters[0]), draw -- the probability is computed on the device from that live count with the
 * rule of the host loop: target = clamp(remaining * frac, qmin, (rows - 256) / 1.25), expected draw 0.97 target -- and
 * pad qinds ([rows,3]) with invalid queries beyond counters[1].  No scalar of the call depends on the round, so the
 * launches can be captured once and replayed.  vnlb_round_dedup_dev is vnlb_round_dedup with the round read from
 * state[0] - 1. */
int vnlb_round_draw(int8_t *mask, int T, int H, int W, double frac, int qmin, int rows, uint32_t seed,
                    uint32_t *state, int64_t *qinds, uint32_t *counters, void *stream);
int vnlb_round_dedup_dev(const int64_t *qinds, int64_t *inds, int B, int K, uint32_t *owner, const uint32_t *state,
                         int8_t *mask, int T, int C, int H, int W, int boost, uint32_t *dropped, void *stream);

/* exec_flat_areas, lib/vnlb/utils/flat_areas.py:16-34.  flat: uint8 [B]
 * (0/1), written for every row (invalid rows get 0). thresh = gamma*sigma2. */
int vnlb_flat_areas(const float *pnoisy, const int64_t *inds, uint8_t *flat, int B, int K,
                    int C, int ps, int pt, float thresh, void *stream);

/* bayes_est.denoise, lib/vnlb/deno/bayes_est.py:17-62: in place on
 * pnoisy [B,K,pt,C,ps,ps]; pbasic (step 2) is read only (the reference
 * re-centres it back to its input values, :52).  Rows that are not valid in
 * `inds` ([B,K], may be NULL = all valid) are skipped.  rank_var float32 [B]
 * (may be NULL).  flat uint8 [B] (may be NULL in step 1).
 * Workspace: the production shapes (7x7x2 patches, k = 100 in step 1 / k = 60 in step 2) run as a chain of kernels
 * that hand (d, e, tau, mean, packed reflectors, trailing matrices) to each other through `ws`: 37 KB (step 1) /
 * 12 KB (step 2) per (group, channel).  vnlb_bayes_workspace_bytes(B, p) = the size for min(B, 16384) groups (0 for
 * shapes served by the single-kernel path); a smaller workspace is legal -- the call is then processed in
 * stream-ordered chunks of as many groups as fit -- but one that cannot hold a single group is VNLB_ERR_WORKSPACE. */
size_t vnlb_bayes_workspace_bytes(int B, const VnlbBayesParams *p);
int vnlb_bayes_filter(float *pnoisy, const float *pbasic, const uint8_t *flat, const int64_t *inds,
                      int B, const VnlbBayesParams *p, float *rank_var, void *ws, size_t ws_bytes,
                      void *stream);

/* Parity hook for compute_cov_mat / denoise_eigvals / bayes_filter_coeff (lib/vnlb/deno/bayes_est.py:112-144):
 * vnlb_bayes_filter that also exports, per (group, channel) problem b*C + ch,
 *   mat  float32 [B*C, q, q] : the matrix that is eigen-decomposed -- the covariance Y^T Y / n (q = p = pt*ps*ps), or, when
 *                              k + 8 <= p (step 2: k = 60), the Gram matrix Y Y^T / n (q = k) whose non-zero spectrum is the
 *                              covariance's; vnlb_bayes_matrix_dim() returns q and tells which (0 = shape not served);
 *   lam  float32 [B*C, 40]   : the eigenvalues above the Wiener threshold thresh*sigma2 + sigmab2, descending, zero padded;
 *   coef float32 [B*C, 40]   : their filter coefficients;   m int32 [B*C] : their number (<= rank).
 * Any of the four may be NULL.  VNLB_EIG_TRIDIAG only; same workspace contract as vnlb_bayes_filter. */
int vnlb_bayes_matrix_dim(const VnlbBayesParams *p, int *is_gram);
int vnlb_bayes_debug(float *pnoisy, const float *pbasic, const uint8_t *flat, const int64_t *inds, int B,
                     const VnlbBayesParams *p, float *mat, float *lam, float *coef, int32_t *m, void *ws,
                     size_t ws_bytes, void *stream);

/* Implementation switch of vnlb_bayes_filter / vnlb_bayes_aggregate_fused for the production patch shapes: on (default)
 * = covariance / Gram matrix + Householder tridiagonalisation with the matrix in registers, in phase kernels, followed
 * by the eigen/filter kernel; off = everything in one shared-memory kernel (needs no workspace).  Same algorithm;
 * results agree to rounding.  Returns the previous setting.  (Environment: VNLB_BAYES_SPLIT=0 at start-up.)
 * vnlb_bayes_workspace_bytes() follows the current setting. */
int vnlb_set_bayes_split(int on);
/* Wiener filter of the Bayes kernels on the tensor cores (3xTF32 mma.sync m16n8k8 for the two products of the filter)
 * instead of FFMA2.  Same results to FP32 rounding; measured slower on B200, hence off by default.
 * Returns the previous setting.  (Environment: VNLB_FILTER_MMA=1 at start-up.) */
int vnlb_set_filter_mma(int on);

/* Fusion of vpss.fill_patches (search.py:91-98) + exec_flat_areas
 * (flat_areas.py:16-34) + bayes_est.denoise (bayes_est.py:17-62) + agg_patches
 * (comp_agg.py:47-138) for one round of groups: patches are gathered from the
 * (YUV) images through `inds` [B,K], filtered, and added into deno [T,C,H,W] /
 * weights [T,H,W]; the patch stacks never reach HBM.  flat_thresh =
 * gamma*sigma2 (used in step 2 only).  Same numerics as vnlb_bayes_filter with
 * VNLB_EIG_TRIDIAG, same workspace contract (vnlb_bayes_workspace_bytes).
 * vnlb_bayes_fused_supported() tells whether the patch shape fits the kernel
 * (p = pt*ps*ps <= 128, rank <= 40); T*C*H*W must be < 2^31 (32-bit image offsets). */
int vnlb_bayes_fused_supported(const VnlbBayesParams *p);
int vnlb_bayes_aggregate_fused(const float *img_noisy, const float *img_basic, const int64_t *inds,
                               int B, int T, int C, int H, int W, const VnlbBayesParams *p,
                               float flat_thresh, float *deno, float *weights, void *ws, size_t ws_bytes,
                               void *stream);

/* agg_patches -> exec_agg_simple_numba, lib/vnlb/agg/comp_agg.py:47-60,106-138:
 * deno[t+dt,ch,y+dy,x+dx] += patch, weights[t+dt,y+dy,x+dx] += 1 for every
 * patch of every VALID row (uniform weight). */
int vnlb_aggregate(const float *patches, const int64_t *inds, int B, int K, float *deno,
                   float *weights, int T, int C, int H, int W, int ps, int pt, void *stream);

/* lib/vnlb/proc_nl.py:118-125: deno /= weights where weights != 0, else
 * deno = fill (basic in step 2, noisy in step 1). */
int vnlb_normalize(float *deno, const float *weights, const float *fill, int T, int C, int H,
                   int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VNLB_B200_H */
