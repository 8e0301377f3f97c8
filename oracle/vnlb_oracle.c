/*
 * oracle/vnlb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, gcc) of the two entry points of the un-vendored
 * third-party package `vpss` that the reference calls for its similarity
 * search, plus the reference's (CPU, numba) aggregation loop:
 *
 *   vpss.exec_sim_search_burst  call site: lib/vnlb/search/search.py:86-89
 *   vpss.fill_patches           call site: lib/vnlb/search/search.py:91-98
 *   exec_agg_simple_numba       lib/vnlb/agg/comp_agg.py:106-138
 *
 * PARITY UNPINNED for the search: `vpss` (no version pin in lib/setup.py:29)
 * is neither under /root/reference nor installable offline, and the reference
 * ships no golden vectors for it (tests/test_gpu_sim_search.py needs vpss,
 * svnlb, a GPU and downloaded data).  This file therefore restates the
 * PUBLISHED algorithm the reference's test asserts equality with (the C++ VNLB
 * patch search of Arias & Morel, JMIV 2018 / IPOL `estimateSimilarPatches`,
 * README.md:65-69, tests/test_gpu_sim_search.py:302-303,423) under the
 * call-site contract of search.py (shapes, sentinels, index codec):
 *
 *   - index codec  ind = t*C*H*W + y*W + x  of the patch's top-left-front
 *     corner (lib/vnlb/search_mask/mask.py:69-71, agg/comp_agg.py:119-121);
 *   - temporal range [t0-nWt_b, t0+nWt_f] shifted to stay inside
 *     [0, T-pt] ("shift" mode) -- the window keeps its size at sequence ends;
 *   - spatial window w_s x w_s centred on the flow trajectory of the query's
 *     top-left corner, shifted to stay inside [0, W-ps] x [0, H-ps];
 *   - trajectory: forward  c[t+1] = clamp(round(c[t] + fflow[t][c[t]]))
 *                 backward c[t-1] = clamp(round(c[t] + bflow[t][c[t]]))
 *     (flow channel 0 = x, channel 1 = y: lib/vnlb/testing/file_io.py:36-65);
 *   - squared L2 distance over `dist_chnls` channels (1 in step 1 -- luminance
 *     only -- all C in step 2), accumulated in FP32 in the fixed order
 *     channel -> frame -> row -> column with one fused multiply-add per
 *     element:  dist = fmaf(d, d, dist),  d = query - candidate;
 *   - top-k ascending by (distance, candidate enumeration order t -> y -> x);
 *     rows with fewer than k candidates keep the caller's sentinels.
 *
 * The accumulation order and the tie-break are the canonical ones the CUDA
 * kernel reproduces bit for bit.  "clip" window mode (no shifting, candidates
 * outside the image dropped) is the documented alternative.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int ps;          /* spatial patch size            (args.ps)     */
    int pt;          /* temporal patch size           (args.pt)     */
    int w_s;         /* spatial search window (odd)   (args.w_s)    */
    int nWt_f;       /* frames searched forward       (args.nWt_f)  */
    int nWt_b;       /* frames searched backward      (args.nWt_b)  */
    int k;           /* neighbours kept               (args.npatches) */
    int dist_chnls;  /* channels entering the distance */
    int window_mode; /* 0 = shift (C++ VNLB), 1 = clip */
} OracleSearchParams;

typedef struct { float d; int32_t order; int64_t ind; } Cand;

static int cand_cmp(const void *a, const void *b) {
    const Cand *x = (const Cand *)a, *y = (const Cand *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->order > y->order) - (x->order < y->order);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* FP32 squared distance in the canonical order; fmaf() is correctly rounded
 * whether or not the host has an FMA unit. */
#if defined(__x86_64__)
__attribute__((target_clones("fma", "default")))
#endif
static float patch_dist(const float *img, int C, int H, int W, int ps, int pt, int dc,
                        int t0, int y0, int x0, int qt, int qy, int qx) {
    float dist = 0.f;
    const size_t HW = (size_t)H * W, CHW = (size_t)C * HW;
    for (int c = 0; c < dc; ++c)
        for (int ht = 0; ht < pt; ++ht) {
            const float *a = img + (size_t)(t0 + ht) * CHW + c * HW;
            const float *b = img + (size_t)(qt + ht) * CHW + c * HW;
            for (int hy = 0; hy < ps; ++hy) {
                const float *ar = a + (size_t)(y0 + hy) * W + x0;
                const float *br = b + (size_t)(qy + hy) * W + qx;
                for (int hx = 0; hx < ps; ++hx) {
                    float d = ar[hx] - br[hx];
                    dist = fmaf(d, d, dist);
                }
            }
        }
    return dist;
}

/* temporal search range, restating the published C++ VNLB logic */
static void temporal_range(int t0, int T, const OracleSearchParams *p, int *r0, int *r1) {
    if (p->window_mode == 0) {
        int shift = imin(0, t0 - p->nWt_b) + imax(0, t0 + p->nWt_f - T + p->pt);
        *r0 = imax(0, t0 - p->nWt_b - shift);
        *r1 = imin(T - p->pt, t0 + p->nWt_f - shift);
    } else {
        *r0 = imax(0, t0 - p->nWt_b);
        *r1 = imin(T - p->pt, t0 + p->nWt_f);
    }
}

static void spatial_range(int c, int L, const OracleSearchParams *p, int *a0, int *a1) {
    int half = (p->w_s - 1) / 2;
    if (p->window_mode == 0) {
        int shift = imin(0, c - half) + imax(0, c + half - L + p->ps);
        *a0 = imax(0, c - half - shift);
        *a1 = imin(L - p->ps, c + half - shift);
    } else {
        *a0 = imax(0, c - half);
        *a1 = imin(L - p->ps, c + half);
    }
}

/* trajectory of the search centre through the flows (NULL flow = zero flow) */
static void trajectory(int t0, int y0, int x0, int r0, int r1, int T, int H, int W,
                       const float *fflow, const float *bflow, int *cx, int *cy) {
    const size_t HW = (size_t)H * W;
    cx[t0] = x0; cy[t0] = y0;
    for (int qt = t0 + 1; qt <= r1; ++qt) {
        int px = cx[qt - 1], py = cy[qt - 1];
        if (fflow) {
            float dx = fflow[((size_t)(qt - 1) * 2 + 0) * HW + (size_t)py * W + px];
            float dy = fflow[((size_t)(qt - 1) * 2 + 1) * HW + (size_t)py * W + px];
            px = clampi((int)roundf((float)px + dx), 0, W - 1);
            py = clampi((int)roundf((float)py + dy), 0, H - 1);
        }
        cx[qt] = px; cy[qt] = py;
    }
    for (int qt = t0 - 1; qt >= r0; --qt) {
        int px = cx[qt + 1], py = cy[qt + 1];
        if (bflow) {
            float dx = bflow[((size_t)(qt + 1) * 2 + 0) * HW + (size_t)py * W + px];
            float dy = bflow[((size_t)(qt + 1) * 2 + 1) * HW + (size_t)py * W + px];
            px = clampi((int)roundf((float)px + dx), 0, W - 1);
            py = clampi((int)roundf((float)py + dy), 0, H - 1);
        }
        cx[qt] = px; cy[qt] = py;
    }
    (void)T;
}

/* vpss.exec_sim_search_burst restated.  qinds: [Q,3] int64 (t,y,x).
 * vals [Q,k] / inds [Q,k] are written in place; slots that cannot be filled
 * keep the caller's sentinels (search.py:84-85 pre-fills -1 / +inf).
 * Returns 0, or -1 on bad arguments. */
int oracle_search_topk(const float *img, int T, int C, int H, int W,
                       const int64_t *qinds, int Q,
                       const float *fflow, const float *bflow,
                       const OracleSearchParams *p, float *vals, int64_t *inds) {
    if (!img || !qinds || !p || !vals || !inds) return -1;
    if (p->ps < 1 || p->pt < 1 || p->w_s < 1 || p->k < 1) return -1;
    if (p->dist_chnls < 1 || p->dist_chnls > C) return -1;
    if (H < p->ps || W < p->ps || T < p->pt) return -1;
    const int nfr_max = p->nWt_f + p->nWt_b + 1;
    const size_t ncand_max = (size_t)nfr_max * p->w_s * p->w_s;
    const int64_t CHW = (int64_t)C * H * W;
    int rc = 0;
#pragma omp parallel
    {
        Cand *cands = (Cand *)malloc(ncand_max * sizeof(Cand));
        int *cx = (int *)malloc((size_t)T * sizeof(int));
        int *cy = (int *)malloc((size_t)T * sizeof(int));
#pragma omp for schedule(dynamic, 4)
        for (int q = 0; q < Q; ++q) {
            int t0 = (int)qinds[3 * q], y0 = (int)qinds[3 * q + 1], x0 = (int)qinds[3 * q + 2];
            if (t0 < 0 || t0 > T - p->pt || y0 < 0 || y0 > H - p->ps || x0 < 0 || x0 > W - p->ps) {
                rc = -1;
                continue;
            }
            int r0, r1;
            temporal_range(t0, T, p, &r0, &r1);
            trajectory(t0, y0, x0, r0, r1, T, H, W, fflow, bflow, cx, cy);
            int n = 0;
            for (int qt = r0; qt <= r1; ++qt) {
                int ax0, ax1, ay0, ay1;
                spatial_range(cx[qt], W, p, &ax0, &ax1);
                spatial_range(cy[qt], H, p, &ay0, &ay1);
                for (int qy = ay0; qy <= ay1; ++qy)
                    for (int qx = ax0; qx <= ax1; ++qx) {
                        cands[n].d = patch_dist(img, C, H, W, p->ps, p->pt, p->dist_chnls,
                                                t0, y0, x0, qt, qy, qx);
                        cands[n].order = n;
                        cands[n].ind = (int64_t)qt * CHW + (int64_t)qy * W + qx;
                        ++n;
                    }
            }
            qsort(cands, (size_t)n, sizeof(Cand), cand_cmp);
            int m = n < p->k ? n : p->k;
            for (int i = 0; i < m; ++i) {
                vals[(size_t)q * p->k + i] = cands[i].d;
                inds[(size_t)q * p->k + i] = cands[i].ind;
            }
        }
        free(cands); free(cx); free(cy);
    }
    return rc;
}

/* Distances of EVERY candidate of one query in enumeration order (test helper:
 * lets the tests classify exact-distance ties).  out must hold
 * (nWt_f+nWt_b+1)*w_s*w_s entries; returns the candidate count. */
int oracle_search_all(const float *img, int T, int C, int H, int W,
                      int t0, int y0, int x0, const float *fflow, const float *bflow,
                      const OracleSearchParams *p, float *out_d, int64_t *out_ind) {
    int *cx = (int *)malloc((size_t)T * sizeof(int));
    int *cy = (int *)malloc((size_t)T * sizeof(int));
    int r0, r1, n = 0;
    const int64_t CHW = (int64_t)C * H * W;
    temporal_range(t0, T, p, &r0, &r1);
    trajectory(t0, y0, x0, r0, r1, T, H, W, fflow, bflow, cx, cy);
    for (int qt = r0; qt <= r1; ++qt) {
        int ax0, ax1, ay0, ay1;
        spatial_range(cx[qt], W, p, &ax0, &ax1);
        spatial_range(cy[qt], H, p, &ay0, &ay1);
        for (int qy = ay0; qy <= ay1; ++qy)
            for (int qx = ax0; qx <= ax1; ++qx) {
                out_d[n] = patch_dist(img, C, H, W, p->ps, p->pt, p->dist_chnls, t0, y0, x0, qt, qy, qx);
                out_ind[n] = (int64_t)qt * CHW + (int64_t)qy * W + qx;
                ++n;
            }
    }
    free(cx); free(cy);
    return n;
}

/* vpss.fill_patches restated (search.py:91-98): patches[b,n,dt,ch,dy,dx] =
 * img[t+dt, ch, y+dy, x+dx] for (t,y,x) = decode(inds[b,n]); entries with
 * ind == -1 are left untouched. */
int oracle_fill_patches(float *patches, const float *img, const int64_t *inds,
                        int B, int K, int T, int C, int H, int W, int ps, int pt) {
    const int64_t HW = (int64_t)H * W, CHW = (int64_t)C * HW;
    const size_t pdim = (size_t)pt * C * ps * ps;
#pragma omp parallel for schedule(static)
    for (int64_t bn = 0; bn < (int64_t)B * K; ++bn) {
        int64_t ind = inds[bn];
        if (ind < 0) continue;
        int t = (int)(ind / CHW), y = (int)((ind % HW) / W), x = (int)(ind % W);
        float *dst = patches + (size_t)bn * pdim;
        for (int dt = 0; dt < pt; ++dt)
            for (int ch = 0; ch < C; ++ch)
                for (int dy = 0; dy < ps; ++dy)
                    for (int dx = 0; dx < ps; ++dx) {
                        int tt = t + dt, yy = y + dy, xx = x + dx;
                        float v = 0.f;
                        if (tt < T && yy < H && xx < W)
                            v = img[(size_t)tt * CHW + (size_t)ch * HW + (size_t)yy * W + xx];
                        *dst++ = v;
                    }
    }
    return 0;
}

/* exec_agg_simple_numba restated (lib/vnlb/agg/comp_agg.py:106-138):
 * sequential scatter-add in row order, uniform weight 1 per patch. */
int oracle_aggregate(float *deno, float *weights, const float *patches, const int64_t *inds,
                     int B, int K, int T, int C, int H, int W, int ps, int pt) {
    const int64_t HW = (int64_t)H * W, CHW = (int64_t)C * HW;
    const size_t pdim = (size_t)pt * C * ps * ps;
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < K; ++n) {
            int64_t ind = inds[(size_t)b * K + n];
            if (ind == -1) continue;
            int t0 = (int)(ind / CHW), h0 = (int)((ind % HW) / W), w0 = (int)(ind % W);
            const float *pp = patches + ((size_t)b * K + n) * pdim;
            for (int dt = 0; dt < pt; ++dt)
                for (int pi = 0; pi < ps; ++pi)
                    for (int pj = 0; pj < ps; ++pj) {
                        int t1 = t0 + dt, h1 = h0 + pi, w1 = w0 + pj;
                        if (t1 < 0 || t1 >= T || h1 < 0 || h1 >= H || w1 < 0 || w1 >= W) continue;
                        for (int ci = 0; ci < C; ++ci)
                            deno[(size_t)t1 * CHW + (size_t)ci * HW + (size_t)h1 * W + w1] +=
                                pp[(((size_t)dt * C + ci) * ps + pi) * ps + pj];
                        weights[(size_t)t1 * HW + (size_t)h1 * W + w1] += 1.f;
                    }
        }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
