// bayes_tridiag.cu -- production path of the per-group Bayes estimate
// (VNLB_EIG_TRIDIAG).  Replaces bayes_est.denoise (lib/vnlb/deno/bayes_est.py:17-62)
// including torch.linalg.eigh (:122), the dominant cost of the reference.
//
// One CTA of 128 threads per (group, channel) problem.  Everything between the
// first read of the patches and the write of the filtered patches stays in ONE
// shared-memory region R of LD*LD floats (40 KB for p = 98) that is re-used
// phase by phase (+ ~4 KB of vectors), so 5 CTAs fit on an SM:
//
//   0. centre + covariance   C = Y^T Y / n      4x4 register tiles; patches staged 32 at a time in R
//   1. Householder tridiagonalisation of C in R  -> (d, e); reflector k is written PACKED
//      over the already-dead rows 0..k of R, so the tail of R is free afterwards
//   2. eigenvalues above the Wiener threshold only: one Sturm count gives their number
//      m (<= rank); parallel multisection (interleaved Sturm chains) isolates them
//   3. eigenvectors of the tridiagonal matrix by twisted factorisation, one thread per
//      eigenpair, Z in the free tail of R (accuracy study: tools/proto_tridiag.py)
//   4. back-transformation by the packed reflectors
//   5. Wiener filter  Xhat = X V diag(w) V^T + mean : V re-laid as Vt[j][r] at the head of
//      R, noisy patches staged behind it in chunks
//
// Only the m eigenpairs that can receive a non-zero filter coefficient are computed:
// eigenvalue > thresh*sigma^2 + sigma_b^2 (bayes_est.py:129-144).
//
// Two front ends share the code:
//   * stack mode  (vnlb_bayes_filter): reads/writes the reference's patch stacks
//     [B,K,pt,C,ps,ps] -- the drop-in operator;
//   * fused mode  (vnlb_bayes_aggregate_fused): gathers the patches straight from the
//     noisy/basic images through `inds`, decides the flat flag, filters all channels and
//     scatters the result into the aggregation accumulators with float atomics -- the
//     patch stacks, vpss.fill_patches, exec_flat_areas and agg_patches never touch HBM.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace vnlb {

constexpr int TT = 128;  // threads per CTA
constexpr int MR = 40;   // eigenpairs kept at most (args.rank <= MR)
constexpr int ILP = 1;   // Sturm chains per thread in a multisection round
constexpr int CH = 32;   // patches staged per covariance chunk

struct TriLayout {
    int p, LD, XS, n;
    int gram;                      // 1: eigen-decompose the n x n Gram matrix Y Y^T/n instead of the p x p covariance
    int q, LDq, ZSq;               // dimension of the eigenproblem and its pitches
    int oR, rsize, oZ, oVt, oX, xrows;  // inside R: packed reflectors at 0, Z at oZ; later (Ut at 0,) Vt at oVt, X at oX
    int oD, oE, oE2, oTau, oV, oW, oMean, oLam, oCoef, oLo, oHi, oNlo, oNhi, oRed, oT, oCnt, oPb, total;
};

static TriLayout tri_layout(int n, int p, bool split = false) {
    TriLayout L;
    L.p = p; L.n = n;
    L.LD = (p + 3) & ~3;
    L.XS = L.LD + 1;
    // Gram trick: with fewer patches than patch dimensions the non-zero spectrum of
    // C = Y^T Y/n equals that of G = Y Y^T/n and v = Y^T u / sqrt(n lambda); step 2 (n = 60 < p = 98).
    L.gram = (n + 8 <= p && n >= 3 && n <= 60) ? 1 : 0;   // n <= 60: one 4x4 Gram tile per thread
    L.q = L.gram ? n : p;
    L.LDq = (L.q + 3) & ~3;
    L.ZSq = L.q | 1;
    const int q = L.q;
    const int nref = ((q - 1) * q / 2 + 3) & ~3;            // packed reflectors
    int rsize;
    if (!L.gram) {
        L.oVt = 0;                                          // Vt[p][MR] at the head of R
        L.oX = (p * MR + 3) & ~3;                           // X[xrows][XS] behind it
        L.oZ = nref > L.oX ? nref : L.oX;                   // Z must not overlap the reflectors nor Vt
        rsize = split ? 0 : L.LD * L.LD;                    // split path: the matrix never enters this kernel
        if (!split && rsize < CH * L.LD) rsize = CH * L.LD;
        if (rsize < nref) rsize = nref;
        if (rsize < L.oZ + MR * L.ZSq) rsize = L.oZ + MR * L.ZSq;
        const int want_rows = n < 32 ? n : 32;
        if (rsize < L.oX + want_rows * L.XS) rsize = L.oX + want_rows * L.XS;
    } else {
        // A (LDq^2), the packed reflectors and the staging chunk live below oZ; Z, then Ut (transposed in
        // place through registers) at oZ; Vt at the head once the reflectors are dead; X over Ut at the end.
        L.oVt = 0;
        L.oZ = (p * MR + 3) & ~3;
        if (L.oZ < L.LDq * L.LDq) L.oZ = L.LDq * L.LDq;
        if (L.oZ < 32 * L.LDq) L.oZ = 32 * L.LDq;
        if (L.oZ < nref) L.oZ = nref;
        L.oX = L.oZ;
        int zsz = MR * L.ZSq;
        if (zsz < n * MR) zsz = n * MR;
        const int want_rows = n < 24 ? n : 24;
        if (zsz < want_rows * L.XS) zsz = want_rows * L.XS;
        rsize = L.oZ + zsz;
    }
    rsize = (rsize + 3) & ~3;
    L.rsize = rsize;
    L.xrows = (rsize - L.oX) / L.XS;
    if (L.xrows > n) L.xrows = n;
    if (L.xrows > TT) L.xrows = TT;                         // one or two threads per staged patch
    const int vlen = L.LD > L.LDq ? L.LD : L.LDq;
    int o = 0;
    L.oR = o; o += rsize;
    L.oD = o; o += vlen;
    L.oE = o; o += vlen;
    L.oE2 = o; o += vlen;
    L.oTau = o; o += vlen;
    L.oV = o; o += vlen;
    L.oW = o; o += vlen;
    L.oMean = o; o += vlen;
    L.oLam = o; o += MR;
    L.oCoef = o; o += MR;
    L.oLo = o; o += MR;
    L.oHi = o; o += MR;
    L.oNlo = o; o += MR;
    L.oNhi = o; o += MR;
    L.oRed = o; o += 16;
    L.oT = o; o += 10 * ((vlen + 3) / 4) + 2;               // compact-WY factors of the back-transformation: 10 floats per block of 4 reflectors
    o = (o + 3) & ~3;
    L.oCnt = o; o += TT;                                     // Sturm counts of the 128 section points of round 0
    L.oPb = o; o += 2 * ((n + 3) & ~3);                     // fused mode: per-patch image / weight offsets
    L.total = o;
    return L;
}

struct BayesArgs {
    // stack mode
    float *pnoisy;
    const float *pbasic;
    const unsigned char *flat;
    // fused mode
    const float *img_noisy;
    const float *img_basic;
    float *deno;
    float *weights;
    int T, H, W;
    float flat_thresh;
    // common
    const long long *inds;
    float *rank_var;
    // parity hook (vnlb_bayes_debug, stack mode): the matrix that is eigen-decomposed, the eigenvalues above the Wiener
    // threshold (descending), their filter coefficients and their number, per (group, channel) problem; all may be null
    float *dbg_mat;                // [B*C, q, q]   covariance (q = p) or Gram matrix (q = n), natural index order
    float *dbg_lam;                // [B*C, MR]
    float *dbg_coef;               // [B*C, MR]
    int *dbg_m;                    // [B*C]
    float *ws;                     // split path: per-problem workspace (see cov_tridiag_kernel)
    int ws_stride;                 // floats per problem
    int ws_pitch;                  // floats per (d | e | tau) vector in the workspace
    int filter_mma;                // Wiener filter on the tensor cores (3xTF32 mma.sync, filter_chunk_mma)
    VnlbBayesParams P;
    TriLayout L;
};

// named barrier over the first `nthr` threads' warps (a multiple of 32); barrier 0 is __syncthreads
__device__ __forceinline__ void bar_sync_n(int id, int nthr) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory"); }
// Warp w of every CTA runs on sub-partition w % 4 of its SM, and in several phases the low warps carry most of
// the work (rows retire from the top in the tridiagonalisation, one thread per eigenpair in the twisted
// factorisation).  Rotating the warp numbering by a per-CTA hash spreads that load over the 4 schedulers.
__device__ __forceinline__ int warp_rotation() { return (int)((blockIdx.x * 0x9E3779B1u) >> 30); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
// block-wide reductions broadcast to every thread; `red` has 8 slots, `phase` alternates 0/1.  The partial sums are
// stored and added in LOGICAL (rotated) warp order, so the result does not depend on warp_rotation(), i.e. on which
// block a problem runs in.
__device__ __forceinline__ int logical_warp() { return ((threadIdx.x >> 5) + warp_rotation()) & ((blockDim.x >> 5) - 1); }
__device__ __forceinline__ float block_sum(float v, float *red, int &phase) {
    v = warp_sum(v);
    float *r = red + 4 * phase;
    if ((threadIdx.x & 31) == 0) r[logical_warp()] = v;
    __syncthreads();
    phase ^= 1;
    return (r[0] + r[1]) + (r[2] + r[3]);
}
__device__ __forceinline__ float block_max(float v, float *red, int &phase) {
    v = warp_max(v);
    float *r = red + 4 * phase;
    if ((threadIdx.x & 31) == 0) r[logical_warp()] = v;
    __syncthreads();
    phase ^= 1;
    return fmaxf(fmaxf(r[0], r[1]), fmaxf(r[2], r[3]));
}

// Wiener filter of `rows` centred patches staged in X (pitch XS): xhat = sum_r coef_r <x, v_r> v_r + mean
// (bayes_est.py:146-151,51).  MB = number of 8-eigenpair blocks; 4 / 2 / 1 threads share a patch
// (rows <= 32 / <= 64 / more), each projecting and reconstructing its slice of the patch.
template <int MB>
__device__ __forceinline__ void filter_chunk(float *X, const float *Vt, const float *coef, const float *mean, int rows,
                                             int p, int XS, int m, int tid) {
    const int lsp = (4 * rows <= TT) ? 2 : ((2 * rows <= TT) ? 1 : 0);   // log2(threads per patch): 4, 2 or 1
    const int row = tid >> lsp, part = tid & ((1 << lsp) - 1);
    const bool on = row < rows;
    float *xr = X + min(row, rows - 1) * XS;
    const int ph = (p + (1 << lsp) - 1) >> lsp;
    const int j0 = min(p, part * ph), j1 = min(p, j0 + ph);
    if (MB == 0) {
        if (on) for (int j = j0; j < j1; ++j) xr[j] = mean[j];
        return;
    }
    float zc[MB > 0 ? 8 * MB : 1];
#pragma unroll
    for (int r = 0; r < 8 * MB; ++r) zc[r] = 0.f;
    for (int j = j0; j < j1; ++j) {
        const float xv = xr[j];
        const float2 xd = make_float2(xv, xv);
        const float4 *vt = reinterpret_cast<const float4 *>(Vt + j * MR);
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const float4 f = vt[2 * b], h = vt[2 * b + 1];
            const float2 z0 = __ffma2_rn(xd, make_float2(f.x, f.y), make_float2(zc[8 * b + 0], zc[8 * b + 1]));
            const float2 z1 = __ffma2_rn(xd, make_float2(f.z, f.w), make_float2(zc[8 * b + 2], zc[8 * b + 3]));
            const float2 z2 = __ffma2_rn(xd, make_float2(h.x, h.y), make_float2(zc[8 * b + 4], zc[8 * b + 5]));
            const float2 z3 = __ffma2_rn(xd, make_float2(h.z, h.w), make_float2(zc[8 * b + 6], zc[8 * b + 7]));
            zc[8 * b + 0] = z0.x; zc[8 * b + 1] = z0.y; zc[8 * b + 2] = z1.x; zc[8 * b + 3] = z1.y;
            zc[8 * b + 4] = z2.x; zc[8 * b + 5] = z2.y; zc[8 * b + 6] = z3.x; zc[8 * b + 7] = z3.y;
        }
    }
#pragma unroll
    for (int r = 0; r < 8 * MB; ++r) {
        if (lsp >= 1) zc[r] += __shfl_xor_sync(0xffffffffu, zc[r], 1);
        if (lsp >= 2) zc[r] += __shfl_xor_sync(0xffffffffu, zc[r], 2);
        zc[r] *= (r < m) ? coef[r] : 0.f;
    }
    for (int j = j0; j < j1; ++j) {
        const float4 *vt = reinterpret_cast<const float4 *>(Vt + j * MR);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const float4 f = vt[2 * b], h = vt[2 * b + 1];
            o0 = fmaf(zc[8 * b + 0], f.x, o0); o1 = fmaf(zc[8 * b + 1], f.y, o1);
            o0 = fmaf(zc[8 * b + 2], f.z, o0); o1 = fmaf(zc[8 * b + 3], f.w, o1);
            o0 = fmaf(zc[8 * b + 4], h.x, o0); o1 = fmaf(zc[8 * b + 5], h.y, o1);
            o0 = fmaf(zc[8 * b + 6], h.z, o0); o1 = fmaf(zc[8 * b + 7], h.w, o1);
        }
        if (on) xr[j] = (o0 + o1) + mean[j];
    }
}

// ---------------------------------------------------------------------------------------------
// Wiener filter on the tensor cores (VNLB_FILTER_MMA): the two products of the filter, Z = X V (rows x p x m) and
// Xhat = (Z diag(coef)) V^T (rows x m x p), as mma.sync.m16n8k8 TF32 tiles with 3xTF32 operand splitting (a = hi + lo,
// hi = cvt.rna.tf32(a), lo = cvt.rna.tf32(a - hi); d += hi hi + lo hi + hi lo: FP32-level accuracy).  filter_chunk is
// bound by the shared-memory data pipe -- every thread streams the whole of Vt twice (4 MB LDS.128 per patch element for
// 4 MB FFMA2 + 8 MB FFMA; ncu: 48 % of the kernel's shared-memory wavefronts); as mma fragments one 4-byte load per lane
// feeds 128 / 256 multiply-adds.  A warp owns a tile of 16 patches (rows of X): Z stays in registers as accumulator
// fragments, is scaled by the coefficients, and serves as the A operand of the second product with the K index
// permuted (k-slot t <-> eigenpair 8 b + 2 t, slot t + 4 <-> 8 b + 2 t + 1), which is exactly the accumulator layout;
// B then is one 8-byte load of Vt[j][8 b + 2 t .. + 1].  Same X / Vt / coef / mean layout as filter_chunk.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MB>
__device__ __forceinline__ void filter_chunk_mma(float *X, const float *Vt, const float *coef, const float *mean, int rows,
                                                 int p, int XS, int m, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int rt = warp; rt * 16 < rows; rt += TT / 32) {
        const int r0 = rt * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < rows, ok1 = r1 < rows;
        float *x0 = X + min(r0, rows - 1) * XS, *x1 = X + min(r1, rows - 1) * XS;   // rows beyond the chunk: read row rows-1, never stored
        float z[MB][4], zx[MB][4];               // hi x hi products / cross terms: two independent mma chains per tile
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            z[b][0] = 0.f; z[b][1] = 0.f; z[b][2] = 0.f; z[b][3] = 0.f;
            zx[b][0] = 0.f; zx[b][1] = 0.f; zx[b][2] = 0.f; zx[b][3] = 0.f;
        }
        // ---- Z = X V
#pragma unroll 2
        for (int k0 = 0; k0 < p; k0 += 8) {
            const int ka = k0 + t, kb = k0 + t + 4;
            const bool va = ka < p, vb = kb < p;
            uint32_t ah[4], al[4];
            tf32_split(va ? x0[ka] : 0.f, ah[0], al[0]);
            tf32_split(va ? x1[ka] : 0.f, ah[1], al[1]);
            tf32_split(vb ? x0[kb] : 0.f, ah[2], al[2]);
            tf32_split(vb ? x1[kb] : 0.f, ah[3], al[3]);
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                uint32_t bh[2], bl[2];
                tf32_split(va ? Vt[ka * MR + 8 * b + g] : 0.f, bh[0], bl[0]);
                tf32_split(vb ? Vt[kb * MR + 8 * b + g] : 0.f, bh[1], bl[1]);
                mma_tf32(zx[b], al, bh);
                mma_tf32(z[b], ah, bh);
                mma_tf32(zx[b], ah, bl);
            }
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) { z[b][0] += zx[b][0]; z[b][1] += zx[b][1]; z[b][2] += zx[b][2]; z[b][3] += zx[b][3]; }
        // ---- scale by the Wiener coefficients; accumulator layout = A layout of the second product (permuted K)
        uint32_t zh[MB][4], zl[MB][4];
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const int ra = 8 * b + 2 * t;
            const float c0 = ra < m ? coef[ra] : 0.f, c1 = ra + 1 < m ? coef[ra + 1] : 0.f;
            tf32_split(z[b][0] * c0, zh[b][0], zl[b][0]);   // (row g,     k-slot t)
            tf32_split(z[b][2] * c0, zh[b][1], zl[b][1]);   // (row g + 8, k-slot t)
            tf32_split(z[b][1] * c1, zh[b][2], zl[b][2]);   // (row g,     k-slot t + 4)
            tf32_split(z[b][3] * c1, zh[b][3], zl[b][3]);   // (row g + 8, k-slot t + 4)
        }
        // ---- Xhat = Zc V^T + mean, written over X
#pragma unroll 2
        for (int j0 = 0; j0 < p; j0 += 8) {
            const int jn = j0 + g;
            float o[4] = {0.f, 0.f, 0.f, 0.f}, ox[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                float2 v2 = make_float2(0.f, 0.f);
                if (jn < p) v2 = *reinterpret_cast<const float2 *>(Vt + jn * MR + 8 * b + 2 * t);
                uint32_t bh[2], bl[2];
                tf32_split(v2.x, bh[0], bl[0]);
                tf32_split(v2.y, bh[1], bl[1]);
                mma_tf32(ox, zl[b], bh);
                mma_tf32(o, zh[b], bh);
                mma_tf32(ox, zh[b], bl);
            }
            o[0] += ox[0]; o[1] += ox[1]; o[2] += ox[2]; o[3] += ox[3];
            const int jc = j0 + 2 * t;
            if (jc < p) {
                const float mj = mean[jc];
                if (ok0) x0[jc] = o[0] + mj;
                if (ok1) x1[jc] = o[2] + mj;
            }
            if (jc + 1 < p) {
                const float mj = mean[jc + 1];
                if (ok0) x0[jc + 1] = o[1] + mj;
                if (ok1) x1[jc + 1] = o[3] + mj;
            }
        }
    }
}

// Gram trick back-mapping: thread j computes row j of Vt:  Vt[j][r] = sum_nn Yc[nn][j] * Ut[nn][r] / sqrt(n * lambda_r)
template <int MB>
__device__ __forceinline__ void gram_map(float *Vt, const float *Ut, const float *lam, const float *mean,
                                         const float *src, const int *pb, bool fused, int rstride, int n, int p, int m,
                                         int coff, int tid) {
    if (tid >= p) return;
    float acc[8 * MB];
#pragma unroll
    for (int r = 0; r < 8 * MB; ++r) acc[r] = 0.f;
    const float mj = mean[tid];
    const float *q = src + coff;
    for (int nn = 0; nn < n; ++nn) {
        const float y = q[fused ? (long long)pb[nn] : (long long)nn * rstride] - mj;
        const float4 *ut = reinterpret_cast<const float4 *>(Ut + nn * MR);
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const float4 f = ut[2 * b], h = ut[2 * b + 1];
            const float2 yd = make_float2(y, y);
            const float2 z0 = __ffma2_rn(yd, make_float2(f.x, f.y), make_float2(acc[8 * b + 0], acc[8 * b + 1]));
            const float2 z1 = __ffma2_rn(yd, make_float2(f.z, f.w), make_float2(acc[8 * b + 2], acc[8 * b + 3]));
            const float2 z2 = __ffma2_rn(yd, make_float2(h.x, h.y), make_float2(acc[8 * b + 4], acc[8 * b + 5]));
            const float2 z3 = __ffma2_rn(yd, make_float2(h.z, h.w), make_float2(acc[8 * b + 6], acc[8 * b + 7]));
            acc[8 * b + 0] = z0.x; acc[8 * b + 1] = z0.y; acc[8 * b + 2] = z1.x; acc[8 * b + 3] = z1.y;
            acc[8 * b + 4] = z2.x; acc[8 * b + 5] = z2.y; acc[8 * b + 6] = z3.x; acc[8 * b + 7] = z3.y;
        }
    }
    float4 *vt = reinterpret_cast<float4 *>(Vt + tid * MR);
#pragma unroll
    for (int r = 0; r < 8 * MB; ++r) acc[r] = (r < m) ? acc[r] * rsqrtf((float)n * lam[r]) : 0.f;
#pragma unroll
    for (int b = 0; b < MR / 8; ++b) {
        if (b < MB) {
            vt[2 * b] = make_float4(acc[8 * (b < MB ? b : 0) + 0], acc[8 * (b < MB ? b : 0) + 1], acc[8 * (b < MB ? b : 0) + 2], acc[8 * (b < MB ? b : 0) + 3]);
            vt[2 * b + 1] = make_float4(acc[8 * (b < MB ? b : 0) + 4], acc[8 * (b < MB ? b : 0) + 5], acc[8 * (b < MB ? b : 0) + 6], acc[8 * (b < MB ? b : 0) + 7]);
        } else {
            vt[2 * b] = make_float4(0.f, 0.f, 0.f, 0.f);
            vt[2 * b + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Split path, first kernel: centre + covariance + Householder tridiagonalisation with the matrix held
// in REGISTERS (QD = compile-time dimension of the eigenproblem; thread t < QD owns one full row).
// One CTA of 128 threads per (group, channel) problem; 168 registers => 3 CTAs per SM, which is why
// this stage is its own kernel: the latency-bound phases 2-5 keep their 5 CTAs per SM in bayes_kernel.
// Output per problem in the workspace (floats): d[LDQ] e[LDQ] tau[LDQ] mean[LDQ] reflectors[nref],
// the reflectors packed exactly like the shared-memory path packs them.
//
// The rows are held index-reversed, b_t[j] = A[QD-1-t][QD-1-j], and columns are eliminated from the
// last one down, which is the forward elimination of A (same d, e, tau and reflectors as the
// shared-memory path up to rounding): the trailing matrix shrinks towards thread 0 / register 0, so
// whole warps retire (QD = 98: 2.06 warps active on average) and the unrolled column loops are entered
// through a fall-through switch at the first live block of 16 columns.
//
// Step for column c (trailing size c; x = column c, r = column c-1 of the trailing matrix, both in shared memory,
// element j written by the thread that owns row j):
//   y       = B x~               (x~ = x[0..c-1]; thread c takes part: y_c = |x~|^2 = alpha^2 + ssq)
//   beta    = -sign(alpha) sqrt(y_c), tau = (beta - alpha)/beta, v = (x~ - beta e_{c-1}) / (alpha - beta)
//   B v     = (y - beta r) / (alpha - beta)                     -- no second pass over the matrix
//   v^T B v = (x~^T y - 2 beta y_{c-1} + beta^2 r_{c-1}) / (alpha - beta)^2   -- one block reduction
//   w       = tau B v - (tau^2 v^T B v / 2) v ;   B <- B - v w^T - w v^T
// After the reduction every thread also knows w_{c-1} and w_{c-2} (from y_{c-1}, y_{c-2}), so thread j forms ITS
// elements of the next step's x and r (columns c-1 and c-2 after the update) before the update and publishes them
// together with v_j and w_j: TWO barriers per step and 5 scalar STS per thread (a barrier drains the pending
// shared-memory stores).  The update of step c and the matvec of step c-1 are ONE sweep over the registers: per
// 4 columns 3 broadcast LDS.128 (v, w, next x) + 6 FFMA2.  Column c-2 of a row is a dynamically indexed register:
// a 4-column window copy of the registers (re-read every 4th step through a switch, updated like the registers
// in between) provides it with a 4-way select.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    uint32_t a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float2 lo, float2 hi) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_newton(float x) {   // 1/x: MUFU.RCP + one Newton step, no special cases
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-x, r, 2.f);
}

// The elimination runs in PHASES, each its own kernel: a phase holds the NR x NR trailing matrix (NR rows = NR threads
// used, NR columns = NR registers per thread) and eliminates columns NR-1 .. CEND; unless CEND == 2 it then hands
// the CEND x CEND trailing matrix to the next phase through the workspace.  The register budget (hence the CTAs
// per SM) follows the trailing size, which is what the latency-bound loop needs.  QDG = dimension of the whole
// problem: step index k = QDG-1-c and the packed reflector offsets are global.
__host__ __device__ constexpr int refl_off(int qdg, int k) { return k * (qdg - 1) - (k * (k - 1)) / 2; }
// Shared scratch of tridiag_regs (floats): 2 x {xs, rs} ping-pong, vs, ws, 2 x 12 scalars, d, e, tau, packed reflectors.
template <int QDG, int NR, int CEND> __host__ __device__ constexpr int tridiag_scratch_floats() {
    return 6 * (((NR + 3) & ~3) + 4) + 24 + 3 * ((NR - CEND + 5) & ~3) +
           ((refl_off(QDG, QDG - CEND) - refl_off(QDG, QDG - NR) + 3) & ~3);
}

// NRC = 4-column chunks of a row held in registers; chunks NRC .. NCH-1 (the columns that die first) live in this
// thread's shared-memory row at byte address aBs: fewer registers => more CTAs per SM for the widest phase.
template <int QDG, int NR, int CEND, int NT, int NRC = (NR + 3) / 4, int OPITCH = ((QDG + 3) & ~3), int OREFL = 4 * ((QDG + 3) & ~3)>
__device__ __forceinline__ void tridiag_regs(float2 (&b)[2 * NRC], float *sv, float *out, float *trail, int tid, uint32_t aBs = 0) {
    constexpr int QD = NR;
    constexpr int LDQ = (QD + 3) & ~3;
    constexpr int LDG = (QDG + 3) & ~3;
    constexpr int NCH = LDQ / 4;
    constexpr int NB2 = (NCH + 1) / 2;             // blocks of 2 chunks = 8 columns
    constexpr int VL = LDQ + 4;
    constexpr int K0 = QDG - NR, K1 = QDG - 1 - CEND;          // steps of this phase: k = K0 .. K1
    constexpr int KN = (K1 - K0 + 1 + 2 + 3) & ~3;             // + the two trailing entries of the last phase
    constexpr int R0 = refl_off(QDG, K0), R1 = refl_off(QDG, K1 + 1);
    static_assert(NB2 <= 13 && QD >= 8 && CEND >= 2 && CEND < NR, "tridiag_regs: 8 <= NR <= 104");
    static_assert(NRC <= NCH && (CEND == 2 ? NRC >= 1 : CEND <= 4 * NRC), "tridiag_regs: the trailing matrix handed on must be register resident");
    // column j of this row: register (static index) or shared memory
    auto colv = [&](auto jc) -> float {
        constexpr int j = decltype(jc)::value;
        if constexpr (j / 4 < NRC) return (j & 1) ? b[j >> 1].y : b[j >> 1].x;
        else return lds32(aBs + 4 * (j - 4 * NRC));
    };
    const uint32_t s0 = smem_u32(sv);
    const uint32_t aV = s0 + 16 * VL, aW = s0 + 20 * VL, aRed = s0 + 24 * VL;   // byte addresses
    const uint32_t aD = aRed + 96, aE = aD + 4 * KN, aTau = aE + 4 * KN, aRefl = aTau + 4 * KN;
    const int lane = tid & 31, warp = tid >> 5;     // logical (rotated) warp: see warp_rotation()
    for (int j = tid; j < 6 * VL + 24; j += NT) sv[j] = 0.f;
    __syncthreads();
    constexpr int c0 = QD - 1;
    float xi, ri, yi = 0.f;
    {   // columns QD-1 and QD-2 of the untouched matrix
        const float x0 = colv(std::integral_constant<int, c0>()), r0 = colv(std::integral_constant<int, c0 - 1>());
        if (tid < c0) { sts32(s0 + 4 * tid, x0); sts32(s0 + 4 * (VL + tid), r0); }
        if (tid == c0) sts32(aRed + 24, x0);
        xi = tid < c0 ? x0 : 0.f;
        ri = r0;
    }
    int Iw = (c0 - 2) >> 2;                        // the window holds columns 4 Iw .. 4 Iw + 3 of this row
    constexpr int w0 = 4 * ((c0 - 2) >> 2);
    float win0 = colv(std::integral_constant<int, w0>()), win1 = colv(std::integral_constant<int, w0 + 1>());
    float win2 = colv(std::integral_constant<int, w0 + 2>()), win3 = colv(std::integral_constant<int, w0 + 3>());
    __syncthreads();
    if (tid <= c0) {                               // y = B x for the first column
        float2 y0 = make_float2(0.f, 0.f), y1 = y0, y2 = y0, y3 = y0;
#pragma unroll
        for (int I = 0; I < NCH; ++I) {
            const float4 x4 = lds128(s0 + 16 * I);
            float2 lo, hi;
            if (I < NRC) { lo = b[2 * (I < NRC ? I : 0)]; hi = b[2 * (I < NRC ? I : 0) + 1]; }
            else { const float4 f = lds128(aBs + 16 * (I - NRC)); lo = make_float2(f.x, f.y); hi = make_float2(f.z, f.w); }
            if (I & 1) { y2 = __ffma2_rn(lo, make_float2(x4.x, x4.y), y2); y3 = __ffma2_rn(hi, make_float2(x4.z, x4.w), y3); }
            else       { y0 = __ffma2_rn(lo, make_float2(x4.x, x4.y), y0); y1 = __ffma2_rn(hi, make_float2(x4.z, x4.w), y1); }
        }
        yi = ((y0.x + y0.y) + (y1.x + y1.y)) + ((y2.x + y2.y) + (y3.x + y3.y));
    }
    for (int c = QD - 1; c >= CEND; --c) {
        if (32 * warp > c) break;                  // this warp's rows are all eliminated: it leaves (named barriers below)
        const int k = QD - 1 - c;                  // step index inside the phase (global: K0 + k)
        const int nb2 = (c + 7) >> 3;              // live 8-column blocks
        const int nthr = 32 * ((c >> 5) + 1);      // threads of the warps still in the loop
        const uint32_t pp = k & 1;
        const uint32_t aX = s0 + pp * (8 * VL), aR = aX + 4 * VL, aXn = s0 + (pp ^ 1) * (8 * VL), aRn = aXn + 4 * VL;
        const uint32_t aRd = aRed + pp * 48, aRdn = aRed + (pp ^ 1) * 48;
        const bool active = tid < c;
        const float part = warp_sum(xi * yi);      // xi = 0 for tid >= c
        if (lane == 0) sts32(aRd + 4 * warp, part);
        if (tid == c) sts32(aRd + 16, yi);
        if (tid == c - 1) sts32(aRd + 20, yi);
        if (tid == c - 2) sts32(aRd + 36, yi);
        bar_sync_n(1, nthr);                                                          // B1
        const float4 r0 = lds128(aRd), r1 = lds128(aRd + 16);
        const float ycm2 = lds32(aRd + 36);
        const float alpha = lds32(aX + 4 * (c - 1)), bcc = lds32(aR + 4 * (c - 1));
        const float xcm2 = lds32(aX + 4 * (c - 2)), rcm2 = lds32(aR + 4 * (c - 2));
        float xBx = r0.x;                          // partial sums of the warps still in the loop, in warp order
        if (nthr > 32) xBx += r0.y;
        if (nthr > 64) xBx += r0.z;
        if (nthr > 96) xBx += r0.w;
        // y_c and alpha^2 come from two roundings of the same entries: clamp, so that a trailing block at rounding-noise
        // level takes the skip path instead of the square root of a negative number
        const float a2 = alpha * alpha;
        const float nrm2 = fmaxf(r1.x, a2), ycm1 = r1.y, dk = r1.z;
        const bool skip = (nrm2 == a2);            // nothing to annihilate: H = I (tau = 0, v = e_{c-1}, w = 0)
        const float rsq = rsqrt_approx(nrm2);
        float sq = nrm2 * rsq;
        sq = fmaf(0.5f * rsq, fmaf(-sq, sq, nrm2), sq);                                // sqrt(nrm2), one Newton step
        const float beta = skip ? alpha : -copysignf(sq, alpha);   // skip: keeps every product below finite (and e_k = alpha)
        const float tau = skip ? 0.f : (beta - alpha) * rcp_newton(beta);
        const float scale = skip ? 0.f : rcp_newton(alpha - beta);
        const float ts = tau * scale;
        const float uBu = fmaf(beta * beta, bcc, fmaf(-2.f * beta, ycm1, xBx));
        const float hs = 0.5f * ts * ts * uBu;     // tau * (v^T p) / 2
        const float vi = (tid == c - 1) ? 1.f : xi * scale;
        const float wi = fmaf(-hs, vi, ts * fmaf(-beta, ri, yi));
        const float wcm1 = fmaf(-hs, 1.f, ts * fmaf(-beta, bcc, ycm1));               // w_{c-1}  (v_{c-1} = 1)
        const float vcm2 = xcm2 * scale;
        const float wcm2 = fmaf(-hs, vcm2, ts * fmaf(-beta, rcm2, ycm2));             // w_{c-2}
        const int wsel = (c - 2) & 3;
        const float qi = wsel == 0 ? win0 : (wsel == 1 ? win1 : (wsel == 2 ? win2 : win3));   // b[c-2] of this row
        // this row's elements of columns c-1 and c-2 after the update, in the update's own operation order
        const float xnext = fmaf(-vi, wcm1, fmaf(-wi, 1.f, ri));
        const float rnext = fmaf(-vi, wcm2, fmaf(-wi, vcm2, qi));
        if (active) {
            sts32(aV + 4 * tid, vi);
            sts32(aW + 4 * tid, wi);
            sts32(aRefl + 4 * (refl_off(QDG, K0 + k) - R0 + (c - 1 - tid)), vi);
            sts32(aXn + 4 * tid, (tid < c - 1) ? xnext : 0.f);
            sts32(aRn + 4 * tid, rnext);
            if (tid == c - 1) sts32(aRdn + 24, xnext);    // B[c-1][c-1]: the next diagonal entry
        }
        if (tid == c) sts32(aXn + 4 * tid, 0.f);
        if (tid == 0) { sts32(aD + 4 * k, dk); sts32(aE + 4 * k, beta); sts32(aTau + 4 * k, tau); }
        bar_sync_n(1, nthr);                                                          // B2
        yi = 0.f;
        if (active) {
            const float2 nvi = make_float2(-vi, -vi), nwi = make_float2(-wi, -wi);
            float2 y0 = make_float2(0.f, 0.f), y1 = y0, y2 = y0, y3 = y0;
#define VNLB_SW1(J)                                                                      \
            if constexpr ((J) < NCH) {                                                  \
                constexpr int I = (J) < NCH ? (J) : 0;                                  \
                const float4 v4 = lds128(aV + 16 * I), w4 = lds128(aW + 16 * I), x4 = lds128(aXn + 16 * I);                       \
                float2 lo, hi;                                                          \
                if constexpr (I < NRC) { lo = b[2 * (I < NRC ? I : 0)]; hi = b[2 * (I < NRC ? I : 0) + 1]; }                      \
                else { const float4 f = lds128(aBs + 16 * (I - NRC)); lo = make_float2(f.x, f.y); hi = make_float2(f.z, f.w); }   \
                lo = __ffma2_rn(nvi, make_float2(w4.x, w4.y), __ffma2_rn(nwi, make_float2(v4.x, v4.y), lo));                      \
                hi = __ffma2_rn(nvi, make_float2(w4.z, w4.w), __ffma2_rn(nwi, make_float2(v4.z, v4.w), hi));                      \
                if constexpr (I < NRC) { b[2 * (I < NRC ? I : 0)] = lo; b[2 * (I < NRC ? I : 0) + 1] = hi; }                      \
                else sts128(aBs + 16 * (I - NRC), lo, hi);                              \
                if (I & 1) { y2 = __ffma2_rn(lo, make_float2(x4.x, x4.y), y2); y3 = __ffma2_rn(hi, make_float2(x4.z, x4.w), y3); } \
                else       { y0 = __ffma2_rn(lo, make_float2(x4.x, x4.y), y0); y1 = __ffma2_rn(hi, make_float2(x4.z, x4.w), y1); } \
            }
#define VNLB_SWB(G) case (G) + 1: { VNLB_SW1(2 * (G) + 1) VNLB_SW1(2 * (G)) }
            switch (nb2) {
                VNLB_SWB(12) VNLB_SWB(11) VNLB_SWB(10) VNLB_SWB(9) VNLB_SWB(8) VNLB_SWB(7) VNLB_SWB(6)
                VNLB_SWB(5) VNLB_SWB(4) VNLB_SWB(3) VNLB_SWB(2) VNLB_SWB(1) VNLB_SWB(0)
                default: break;
            }
#undef VNLB_SWB
#undef VNLB_SW1
            yi = ((y0.x + y0.y) + (y1.x + y1.y)) + ((y2.x + y2.y) + (y3.x + y3.y));
            {   // the window copy gets the same update
                const float4 v4 = lds128(aV + 16 * Iw), w4 = lds128(aW + 16 * Iw);
                win0 = fmaf(-vi, w4.x, fmaf(-wi, v4.x, win0)); win1 = fmaf(-vi, w4.y, fmaf(-wi, v4.y, win1));
                win2 = fmaf(-vi, w4.z, fmaf(-wi, v4.z, win2)); win3 = fmaf(-vi, w4.w, fmaf(-wi, v4.w, win3));
            }
        }
        if (c >= 3 && ((c - 3) >> 2) != Iw) {      // next step needs column c-3: re-read the window from the registers
            Iw = (c - 3) >> 2;
            if (Iw >= NRC) {
                const float4 f = lds128(aBs + 16 * (Iw - NRC));
                win0 = f.x; win1 = f.y; win2 = f.z; win3 = f.w;
            } else {
                switch (Iw) {
#define VNLB_W(J) case (J): if constexpr ((J) < NRC) { win0 = b[2 * ((J) < NRC ? (J) : 0)].x; win1 = b[2 * ((J) < NRC ? (J) : 0)].y; win2 = b[2 * ((J) < NRC ? (J) : 0) + 1].x; win3 = b[2 * ((J) < NRC ? (J) : 0) + 1].y; } break;
                    VNLB_W(0) VNLB_W(1) VNLB_W(2) VNLB_W(3) VNLB_W(4) VNLB_W(5) VNLB_W(6) VNLB_W(7) VNLB_W(8) VNLB_W(9) VNLB_W(10) VNLB_W(11) VNLB_W(12)
                    VNLB_W(13) VNLB_W(14) VNLB_W(15) VNLB_W(16) VNLB_W(17) VNLB_W(18) VNLB_W(19) VNLB_W(20) VNLB_W(21) VNLB_W(22) VNLB_W(23) VNLB_W(24) VNLB_W(25)
#undef VNLB_W
                    default: break;
                }
            }
        }
        xi = (tid < c - 1) ? xnext : 0.f;
        ri = rnext;
    }
    if constexpr (CEND == 2) {   // the last 2 x 2 block: B[1][1], B[1][0], B[0][0]
        if (tid == 1) { sts32(aD + 4 * (QD - 2), b[0].y); sts32(aE + 4 * (QD - 2), b[0].x); sts32(aTau + 4 * (QD - 2), 0.f); }
        if (tid == 0) { sts32(aD + 4 * (QD - 1), b[0].x); sts32(aE + 4 * (QD - 1), 0.f); sts32(aTau + 4 * (QD - 1), 0.f); }
    } else {                     // trailing CEND x CEND matrix for the next phase: row t at trail + t * round4(CEND)
        constexpr int LDT = (CEND + 3) & ~3;
        if (tid < CEND) {
#pragma unroll
            for (int I = 0; I < LDT / 4; ++I)
                reinterpret_cast<float4 *>(trail + tid * LDT)[I] = make_float4(b[2 * I].x, b[2 * I].y, b[2 * I + 1].x, b[2 * I + 1].y);
        }
    }
    __syncthreads();
    // flush this phase's (d, e, tau | reflectors) to the workspace
    constexpr int NK = K1 - K0 + 1 + (CEND == 2 ? 2 : 0);
    const float *sd = sv + 6 * VL + 24;
    for (int idx = threadIdx.x; idx < NK; idx += NT) {
        out[K0 + idx] = sd[idx];
        out[OPITCH + K0 + idx] = sd[KN + idx];
        out[2 * OPITCH + K0 + idx] = sd[2 * KN + idx];
    }
    for (int idx = threadIdx.x; idx < R1 - R0; idx += NT) out[OREFL + R0 + idx] = sd[3 * KN + idx];
    (void)LDG;
}

// ---------------------------------------------------------------------------------------------
// TWO ROWS PER THREAD, one warp per problem (NR <= 64): lane l owns rows l and l + 32 of the trailing matrix.  The sweep
// of tridiag_regs is bound by the shared-memory data pipe: per 4 columns a warp loads 3 broadcast LDS.128 (v, w, next x)
// for 6 FFMA2 of ONE row per lane (measured: 0.42 LDS.128 against 1.72 FFMA2 per cycle and SM, 2 : 1 load bound).  With
// two rows per lane the same three loads feed 12 FFMA2 and a problem needs half the warps, so the phase is balanced
// between the two pipes; and since one warp holds the whole problem, the two named barriers of a step become
// __syncwarp(), the block reduction becomes a shuffle tree and the scalars of rows c, c-1, c-2 come from their
// owners by shuffle instead of through shared memory.  Same arithmetic per element as tridiag_regs (operation order
// of the update and of the raw-column trick unchanged); only the order of the partial sums of a mat-vec differs.
// Shared scratch (floats): 2 x xs (ping-pong), vs, ws, then d, e, tau (KN each) and the packed reflectors.
template <int QDG, int NR, int CEND> __host__ __device__ constexpr int tridiag2_scratch_floats() {
    return 4 * (((NR + 3) & ~3) + 4) + 3 * ((NR - CEND + 5) & ~3) + ((refl_off(QDG, QDG - CEND) - refl_off(QDG, QDG - NR) + 3) & ~3);
}

template <int QDG, int NR, int CEND, int OPITCH, int OREFL>
__device__ __forceinline__ void tridiag_regs2(float2 (&bl)[2 * ((NR + 3) / 4)], float2 (&bh)[2 * ((NR + 3) / 4)], float *sv,
                                              float *out, float *trail, int lane) {
    constexpr int LDQ = (NR + 3) & ~3, NCH = LDQ / 4, VL = LDQ + 4;
    constexpr int K0 = QDG - NR, K1 = QDG - 1 - CEND;
    constexpr int KN = (K1 - K0 + 1 + 2 + 3) & ~3;
    constexpr int R0 = refl_off(QDG, K0), R1 = refl_off(QDG, K1 + 1);
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NR > 32 && NR <= 64 && CEND >= 32 && CEND < NR && CEND % 4 == 0, "tridiag_regs2: 32 < NR <= 64, the rest handed on is 32 x 32 or larger");
    auto coll = [&](auto jc) -> float { constexpr int j = decltype(jc)::value; return (j & 1) ? bl[j >> 1].y : bl[j >> 1].x; };
    auto colh = [&](auto jc) -> float { constexpr int j = decltype(jc)::value; return (j & 1) ? bh[j >> 1].y : bh[j >> 1].x; };
    const uint32_t s0 = smem_u32(sv);
    const uint32_t aV = s0 + 8 * VL, aW = s0 + 12 * VL;                       // byte addresses
    const uint32_t aD = s0 + 16 * VL, aE = aD + 4 * KN, aTau = aE + 4 * KN, aRefl = aTau + 4 * KN;
    const int i0 = lane, i1 = lane + 32;
    for (int j = lane; j < 4 * VL; j += 32) sv[j] = 0.f;
    __syncwarp();
    constexpr int c0 = NR - 1;
    // row c's owner: lane c & 31, its upper row when c >= 32 (c is warp uniform)
    auto from_row = [&](int row, float lo, float hi) -> float { return __shfl_sync(FULL, row >= 32 ? hi : lo, row & 31); };
    float xl, xh, rl, rh, yl = 0.f, yh = 0.f, dk;
    {
        const float x0l = coll(std::integral_constant<int, c0>()), x0h = colh(std::integral_constant<int, c0>());
        rl = coll(std::integral_constant<int, c0 - 1>());
        rh = colh(std::integral_constant<int, c0 - 1>());
        if (i0 < c0) sts32(s0 + 4 * i0, x0l);
        if (i1 < c0) sts32(s0 + 4 * i1, x0h);
        dk = from_row(c0, x0l, x0h);                                          // B[c0][c0]
        xl = i0 < c0 ? x0l : 0.f;
        xh = i1 < c0 ? x0h : 0.f;
    }
    int Iw = (c0 - 2) >> 2;                                                    // the windows hold columns 4 Iw .. 4 Iw + 3 of both rows
    constexpr int w0 = 4 * ((c0 - 2) >> 2);
    float wl0 = coll(std::integral_constant<int, w0>()), wl1 = coll(std::integral_constant<int, w0 + 1>());
    float wl2 = coll(std::integral_constant<int, w0 + 2>()), wl3 = coll(std::integral_constant<int, w0 + 3>());
    float wh0 = colh(std::integral_constant<int, w0>()), wh1 = colh(std::integral_constant<int, w0 + 1>());
    float wh2 = colh(std::integral_constant<int, w0 + 2>()), wh3 = colh(std::integral_constant<int, w0 + 3>());
    __syncwarp();
    {   // y = B x for the first column, both rows
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
        for (int I = 0; I < NCH; ++I) {
            const float4 x4 = lds128(s0 + 16 * I);
            a0 = __ffma2_rn(bl[2 * I], make_float2(x4.x, x4.y), a0); a1 = __ffma2_rn(bl[2 * I + 1], make_float2(x4.z, x4.w), a1);
            b0 = __ffma2_rn(bh[2 * I], make_float2(x4.x, x4.y), b0); b1 = __ffma2_rn(bh[2 * I + 1], make_float2(x4.z, x4.w), b1);
        }
        yl = (a0.x + a0.y) + (a1.x + a1.y);
        yh = (b0.x + b0.y) + (b1.x + b1.y);
    }
    for (int c = NR - 1; c >= CEND; --c) {
        const int k = NR - 1 - c;
        const int nb2 = (c + 7) >> 3;
        const uint32_t pp = k & 1;
        const uint32_t aXn = s0 + (pp ^ 1) * (4 * VL);
        __syncwarp();                                                          // the previous sweep's reads of v, w, x are done
        const float xBx = warp_sum(fmaf(xh, yh, xl * yl));                     // x = 0 on rows >= c
        const float yc = from_row(c, yl, yh), ycm1 = from_row(c - 1, yl, yh), ycm2 = from_row(c - 2, yl, yh);
        const float alpha = from_row(c - 1, xl, xh), bcc = from_row(c - 1, rl, rh);
        const float xcm2 = from_row(c - 2, xl, xh), rcm2 = from_row(c - 2, rl, rh);
        const float a2 = alpha * alpha;
        const float nrm2 = fmaxf(yc, a2);
        const bool skip = (nrm2 == a2);
        const float rsq = rsqrt_approx(nrm2);
        float sq = nrm2 * rsq;
        sq = fmaf(0.5f * rsq, fmaf(-sq, sq, nrm2), sq);
        const float beta = skip ? alpha : -copysignf(sq, alpha);
        const float tau = skip ? 0.f : (beta - alpha) * rcp_newton(beta);
        const float scale = skip ? 0.f : rcp_newton(alpha - beta);
        const float ts = tau * scale;
        const float uBu = fmaf(beta * beta, bcc, fmaf(-2.f * beta, ycm1, xBx));
        const float hs = 0.5f * ts * ts * uBu;
        const float wcm1 = fmaf(-hs, 1.f, ts * fmaf(-beta, bcc, ycm1));
        const float vcm2 = xcm2 * scale;
        const float wcm2 = fmaf(-hs, vcm2, ts * fmaf(-beta, rcm2, ycm2));
        const int wsel = (c - 2) & 3;
        const bool actl = i0 < c, acth = i1 < c;
        float vl = (i0 == c - 1) ? 1.f : xl * scale, vh = (i1 == c - 1) ? 1.f : xh * scale;
        float wl = fmaf(-hs, vl, ts * fmaf(-beta, rl, yl)), wh = fmaf(-hs, vh, ts * fmaf(-beta, rh, yh));
        if (!actl) { vl = 0.f; wl = 0.f; }                                     // dead rows: the update is a no-op
        if (!acth) { vh = 0.f; wh = 0.f; }
        const float ql = wsel == 0 ? wl0 : (wsel == 1 ? wl1 : (wsel == 2 ? wl2 : wl3));
        const float qh = wsel == 0 ? wh0 : (wsel == 1 ? wh1 : (wsel == 2 ? wh2 : wh3));
        const float xnl = fmaf(-vl, wcm1, fmaf(-wl, 1.f, rl)), xnh = fmaf(-vh, wcm1, fmaf(-wh, 1.f, rh));
        const float rnl = fmaf(-vl, wcm2, fmaf(-wl, vcm2, ql)), rnh = fmaf(-vh, wcm2, fmaf(-wh, vcm2, qh));
        if (actl) {
            sts32(aV + 4 * i0, vl); sts32(aW + 4 * i0, wl);
            sts32(aRefl + 4 * (refl_off(QDG, K0 + k) - R0 + (c - 1 - i0)), vl);
            sts32(aXn + 4 * i0, (i0 < c - 1) ? xnl : 0.f);
        }
        if (acth) {
            sts32(aV + 4 * i1, vh); sts32(aW + 4 * i1, wh);
            sts32(aRefl + 4 * (refl_off(QDG, K0 + k) - R0 + (c - 1 - i1)), vh);
            sts32(aXn + 4 * i1, (i1 < c - 1) ? xnh : 0.f);
        }
        if (i0 == c) sts32(aXn + 4 * i0, 0.f);
        if (i1 == c) sts32(aXn + 4 * i1, 0.f);
        if (lane == 0) { sts32(aD + 4 * k, dk); sts32(aE + 4 * k, beta); sts32(aTau + 4 * k, tau); }
        dk = from_row(c - 1, xnl, xnh);                                        // B[c-1][c-1] after the update: the next diagonal entry
        __syncwarp();
        {
            const float2 nvl = make_float2(-vl, -vl), nwl = make_float2(-wl, -wl);
            const float2 nvh = make_float2(-vh, -vh), nwh = make_float2(-wh, -wh);
            float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#define VNLB_S2(J)                                                                       \
            if constexpr ((J) < NCH) {                                                  \
                constexpr int I = (J) < NCH ? (J) : 0;                                  \
                const float4 v4 = lds128(aV + 16 * I), w4 = lds128(aW + 16 * I), x4 = lds128(aXn + 16 * I);                       \
                const float2 v01 = make_float2(v4.x, v4.y), v23 = make_float2(v4.z, v4.w);                                          \
                const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);                                          \
                const float2 x01 = make_float2(x4.x, x4.y), x23 = make_float2(x4.z, x4.w);                                          \
                bl[2 * I] = __ffma2_rn(nvl, w01, __ffma2_rn(nwl, v01, bl[2 * I]));                                                  \
                bl[2 * I + 1] = __ffma2_rn(nvl, w23, __ffma2_rn(nwl, v23, bl[2 * I + 1]));                                          \
                bh[2 * I] = __ffma2_rn(nvh, w01, __ffma2_rn(nwh, v01, bh[2 * I]));                                                  \
                bh[2 * I + 1] = __ffma2_rn(nvh, w23, __ffma2_rn(nwh, v23, bh[2 * I + 1]));                                          \
                a0 = __ffma2_rn(bl[2 * I], x01, a0); a1 = __ffma2_rn(bl[2 * I + 1], x23, a1);                                       \
                b0 = __ffma2_rn(bh[2 * I], x01, b0); b1 = __ffma2_rn(bh[2 * I + 1], x23, b1);                                       \
            }
#define VNLB_S2B(G) case (G) + 1: { VNLB_S2(2 * (G) + 1) VNLB_S2(2 * (G)) }
            switch (nb2) {
                VNLB_S2B(7) VNLB_S2B(6) VNLB_S2B(5) VNLB_S2B(4) VNLB_S2B(3) VNLB_S2B(2) VNLB_S2B(1) VNLB_S2B(0)
                default: break;
            }
#undef VNLB_S2B
#undef VNLB_S2
            yl = (a0.x + a0.y) + (a1.x + a1.y);
            yh = (b0.x + b0.y) + (b1.x + b1.y);
            {   // the window copies get the same update
                const float4 v4 = lds128(aV + 16 * Iw), w4 = lds128(aW + 16 * Iw);
                wl0 = fmaf(-vl, w4.x, fmaf(-wl, v4.x, wl0)); wl1 = fmaf(-vl, w4.y, fmaf(-wl, v4.y, wl1));
                wl2 = fmaf(-vl, w4.z, fmaf(-wl, v4.z, wl2)); wl3 = fmaf(-vl, w4.w, fmaf(-wl, v4.w, wl3));
                wh0 = fmaf(-vh, w4.x, fmaf(-wh, v4.x, wh0)); wh1 = fmaf(-vh, w4.y, fmaf(-wh, v4.y, wh1));
                wh2 = fmaf(-vh, w4.z, fmaf(-wh, v4.z, wh2)); wh3 = fmaf(-vh, w4.w, fmaf(-wh, v4.w, wh3));
            }
        }
        if (((c - 3) >> 2) != Iw) {                                            // next step needs column c-3: re-read the windows from the registers
            Iw = (c - 3) >> 2;
            switch (Iw) {
#define VNLB_W2(J) case (J): if constexpr ((J) < NCH) { constexpr int I = (J) < NCH ? (J) : 0; wl0 = bl[2 * I].x; wl1 = bl[2 * I].y; wl2 = bl[2 * I + 1].x; wl3 = bl[2 * I + 1].y; \
                                                        wh0 = bh[2 * I].x; wh1 = bh[2 * I].y; wh2 = bh[2 * I + 1].x; wh3 = bh[2 * I + 1].y; } break;
                VNLB_W2(0) VNLB_W2(1) VNLB_W2(2) VNLB_W2(3) VNLB_W2(4) VNLB_W2(5) VNLB_W2(6) VNLB_W2(7) VNLB_W2(8) VNLB_W2(9) VNLB_W2(10) VNLB_W2(11)
                VNLB_W2(12) VNLB_W2(13) VNLB_W2(14) VNLB_W2(15)
#undef VNLB_W2
                default: break;
            }
        }
        xl = (i0 < c - 1) ? xnl : 0.f;
        xh = (i1 < c - 1) ? xnh : 0.f;
        rl = rnl;
        rh = rnh;
    }
    {   // trailing CEND x CEND matrix for the next phase: row t at trail + t * CEND (CEND = 32: the lower rows only)
        static_assert(CEND == 32, "tridiag_regs2 hands a 32 x 32 matrix on");
#pragma unroll
        for (int I = 0; I < CEND / 4; ++I)
            reinterpret_cast<float4 *>(trail + lane * CEND)[I] = make_float4(bl[2 * I].x, bl[2 * I].y, bl[2 * I + 1].x, bl[2 * I + 1].y);
    }
    __syncwarp();
    constexpr int NK = K1 - K0 + 1;
    const float *sd = sv + 4 * VL;
    for (int idx = lane; idx < NK; idx += 32) {
        out[K0 + idx] = sd[idx];
        out[OPITCH + K0 + idx] = sd[KN + idx];
        out[2 * OPITCH + K0 + idx] = sd[2 * KN + idx];
    }
    for (int idx = lane; idx < R1 - R0; idx += 32) out[OREFL + R0 + idx] = sd[3 * KN + idx];
}

// ---------------------------------------------------------------------------------------------
// COLUMNS SPLIT OVER THE WARPS, several rows per lane (the widest phase, 96 -> 64).  tridiag_regs is bound by the
// shared-memory data pipe: every warp loads the broadcast operands (v, w, next x) of ALL live columns for ONE row per
// lane (3 LDS.128 per 6 FFMA2; measured 70 % pipe utilisation at 35 % FMA).  Here warp h of 4 holds the 4-column chunks
// I = h (mod 4) of the NR x NR trailing matrix for ALL rows, NS = NR / 32 rows per lane (rows l, l + 32, l + 64): the
// three loads of a chunk feed 6 FFMA2 per row slot (18 for NS = 3), and each warp only loads its own quarter of the
// columns -- 12 x less shared-memory traffic per FMA.  The per-row scalars of a step (x, r, y, v, w, the next x and r,
// the 4-column window that supplies column c - 2) are kept REDUNDANTLY by every warp in the layout of tridiag_regs2
// (shuffles fetch the rows c, c-1, c-2); what crosses warps are the partial sums of the mat-vec (part[warp][row], summed
// in warp order by everybody after barrier B1), the published v / w / next x (warp h publishes row slot h, barrier B2)
// and, every fourth step, the new window chunk from the warp that owns it.  Same arithmetic per element as
// tridiag_regs2; only the grouping of the mat-vec partial sums differs.
// Shared scratch (floats): xs[VL] vs[VL] ws[VL] part[4][NR] win[NR][4] d[KN] e[KN] tau[KN] reflectors.
// -------------------------------------------------------------- P1. tridiag_cols (column-split elimination)
template <int QDG, int NR, int CEND> __host__ __device__ constexpr int tridiag_cols_scratch_floats() {
    return 6 * (NR + 4) + 8 * NR + 8 + 3 * ((NR - CEND + 3) & ~3) +
           ((refl_off(QDG, QDG - CEND) - refl_off(QDG, QDG - NR) + 3) & ~3);
}

// Second version: a warp keeps the per-row state (x, r, window) only of the row slot it PUBLISHES (warp h < 3: rows
// l + 32 h); everything another warp needs of a row comes from shared memory -- v_i, w_i of its own rows for the sweep
// (6 loads), the scalars of rows c, c-1, c-2 (broadcast loads of part / xs / rs), and x~^T B x~ as the sum of the four
// warps' partial sums (each warp dots ITS partial mat-vec with the next x during the sweep).  xs / rs are ping-pong
// buffers: a step reads the current column while the publishers write the next one.
template <int QDG, int NR, int CEND, int OPITCH, int OREFL>
__device__ __forceinline__ void tridiag_cols(const float *Bm, int ldb, float *sv, float *out, float *trail, int lane, int h) {
    constexpr int NS = NR / 32, NCH = NR / 4, NL = NCH / 4, VL = NR + 4;
    constexpr int K0 = QDG - NR, K1 = QDG - 1 - CEND;
    constexpr int KN = (K1 - K0 + 1 + 3) & ~3;
    constexpr int R0 = refl_off(QDG, K0), R1 = refl_off(QDG, K1 + 1);
    static_assert(NS == 3 && NCH % 4 == 0 && CEND % 32 == 0 && CEND >= 32 && CEND < NR, "tridiag_cols: 96 rows in three slots, 4 warps");
    float *xs0 = sv, *rs0 = sv + 2 * VL, *vs = sv + 4 * VL, *ws = sv + 5 * VL, *part = sv + 6 * VL, *win = part + 4 * NR;
    float *xbx = win + 4 * NR;                     // [0..3] partial x~^T B x~ per warp, [4] the next diagonal entry
    float *sd = xbx + 8, *se = sd + KN, *st = se + KN, *srf = st + KN;
    const uint32_t aV = smem_u32(vs), aW = smem_u32(ws);
    int row[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) row[s] = lane + 32 * s;
    float2 b[NS][2 * NL];                          // this lane's rows, this warp's column chunks
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            const float4 f = *reinterpret_cast<const float4 *>(Bm + row[s] * ldb + 4 * (4 * t + h));
            b[s][2 * t] = make_float2(f.x, f.y);
            b[s][2 * t + 1] = make_float2(f.z, f.w);
        }
    const bool own = h < NS;                       // warps 0..2 publish row slot h; warp 3 writes d, e, tau
    const int io = lane + 32 * (own ? h : 0);
    constexpr int c0 = NR - 1;
    int Iw = (c0 - 2) >> 2;
    float x = 0.f, r = 0.f, wq0 = 0.f, wq1 = 0.f, wq2 = 0.f, wq3 = 0.f;
    if (own) {
        const float x0 = Bm[io * ldb + c0];
        r = Bm[io * ldb + c0 - 1];
        x = io < c0 ? x0 : 0.f;
        const float4 f = *reinterpret_cast<const float4 *>(Bm + io * ldb + 4 * Iw);
        wq0 = f.x; wq1 = f.y; wq2 = f.z; wq3 = f.w;
    }
    const float dk0 = Bm[c0 * ldb + c0];
    for (int j = 32 * h + lane; j < 6 * VL; j += 128) sv[j] = 0.f;
    __syncthreads();
    if (own) { xs0[io] = x; rs0[io] = r; }
    if (h == 3 && lane == 0) xbx[4] = dk0;
    __syncthreads();
    // partial mat-vec of this warp's chunks for every row slot + its share of x~^T B x~
    auto sweep_tail = [&](const float2 (&acc)[NS][2], const float *xcur) {
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float ps = (acc[s][0].x + acc[s][0].y) + (acc[s][1].x + acc[s][1].y);
            part[4 * row[s] + h] = ps;
            dot = fmaf(xcur[row[s]], ps, dot);
        }
        dot = warp_sum(dot);
        if (lane == 0) xbx[h] = dot;
    };
    {
        float2 acc[NS][2];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = make_float2(0.f, 0.f);
        const uint32_t aX = smem_u32(xs0);
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            const float4 x4 = lds128(aX + 16 * (4 * t + h));
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s][0] = __ffma2_rn(b[s][2 * t], make_float2(x4.x, x4.y), acc[s][0]);
                acc[s][1] = __ffma2_rn(b[s][2 * t + 1], make_float2(x4.z, x4.w), acc[s][1]);
            }
        }
        sweep_tail(acc, xs0);
    }
    bool winload = false;
    int pp = 0;
    for (int c = NR - 1; c >= CEND; --c) {
        const int k = NR - 1 - c;
        const float *xc = xs0 + pp * VL, *rc = rs0 + pp * VL;
        float *xn_s = xs0 + (pp ^ 1) * VL, *rn_s = rs0 + (pp ^ 1) * VL;
        __syncthreads();                                                       // B1: partial sums (and a new window) are published
        if (winload) {
            if (own) { const float4 f = *reinterpret_cast<const float4 *>(win + 4 * io); wq0 = f.x; wq1 = f.y; wq2 = f.z; wq3 = f.w; }
            winload = false;
        }
        const float4 pc = *reinterpret_cast<const float4 *>(part + 4 * c);
        const float4 pc1 = *reinterpret_cast<const float4 *>(part + 4 * (c - 1));
        const float4 pc2 = *reinterpret_cast<const float4 *>(part + 4 * (c - 2));
        const float4 xb = *reinterpret_cast<const float4 *>(xbx);
        const float yc = (pc.x + pc.y) + (pc.z + pc.w), ycm1 = (pc1.x + pc1.y) + (pc1.z + pc1.w), ycm2 = (pc2.x + pc2.y) + (pc2.z + pc2.w);
        const float xBx = (xb.x + xb.y) + (xb.z + xb.w);
        const float alpha = xc[c - 1], xcm2 = xc[c - 2], bcc = rc[c - 1], rcm2 = rc[c - 2];
        const float dk = xbx[4];
        const float a2 = alpha * alpha;
        const float nrm2 = fmaxf(yc, a2);
        const bool skip = (nrm2 == a2);
        const float rsq = rsqrt_approx(nrm2);
        float sq = nrm2 * rsq;
        sq = fmaf(0.5f * rsq, fmaf(-sq, sq, nrm2), sq);
        const float beta = skip ? alpha : -copysignf(sq, alpha);
        const float tau = skip ? 0.f : (beta - alpha) * rcp_newton(beta);
        const float scale = skip ? 0.f : rcp_newton(alpha - beta);
        const float ts = tau * scale;
        const float uBu = fmaf(beta * beta, bcc, fmaf(-2.f * beta, ycm1, xBx));
        const float hs = 0.5f * ts * ts * uBu;
        const float wcm1 = fmaf(-hs, 1.f, ts * fmaf(-beta, bcc, ycm1));
        const float vcm2 = xcm2 * scale;
        const float wcm2 = fmaf(-hs, vcm2, ts * fmaf(-beta, rcm2, ycm2));
        float vo = 0.f, wo = 0.f, xno = 0.f, rno = 0.f;
        if (own) {                                                             // this warp's row slot: v, w, next x and r
            const float4 po = *reinterpret_cast<const float4 *>(part + 4 * io);
            const float y = (po.x + po.y) + (po.z + po.w);
            const bool act = io < c;
            vo = (io == c - 1) ? 1.f : x * scale;
            wo = fmaf(-hs, vo, ts * fmaf(-beta, r, y));
            if (!act) { vo = 0.f; wo = 0.f; }                                  // dead rows: the update is a no-op
            const int wsel = (c - 2) & 3;
            const float q = wsel == 0 ? wq0 : (wsel == 1 ? wq1 : (wsel == 2 ? wq2 : wq3));
            xno = fmaf(-vo, wcm1, fmaf(-wo, 1.f, r));
            rno = fmaf(-vo, wcm2, fmaf(-wo, vcm2, q));
            vs[io] = vo;
            ws[io] = wo;
            if (act) srf[refl_off(QDG, K0 + k) - R0 + (c - 1 - io)] = vo;
            xn_s[io] = (io < c - 1) ? xno : 0.f;
            rn_s[io] = rno;
            if (io == c - 1) xbx[5] = xno;                                     // B[c-1][c-1] after the update: the next diagonal entry
        } else if (lane == 0) {
            sd[k] = dk; se[k] = beta; st[k] = tau;
        }
        __syncthreads();                                                       // B2: v, w and the next x / r are published
        if (h == 3 && lane == 0) xbx[4] = xbx[5];
        {
            float2 nv[NS], nw[NS], acc[NS][2];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const float v_s = vs[row[s]], w_s = ws[row[s]];
                nv[s] = make_float2(-v_s, -v_s);
                nw[s] = make_float2(-w_s, -w_s);
                acc[s][0] = acc[s][1] = make_float2(0.f, 0.f);
            }
            const uint32_t aX = smem_u32(xn_s);
            const int nlive = (c + 3) >> 2;                                    // chunks holding a column < c
            const int tl = nlive > h ? (nlive - h + 3) >> 2 : 0;               // ... of this warp
#define VNLB_C1(T)                                                                       \
            {                                                                           \
                constexpr int t = (T) < NL ? (T) : 0;                                   \
                const int I = 4 * t + h;                                                \
                const float4 v4 = lds128(aV + 16 * I), w4 = lds128(aW + 16 * I), x4 = lds128(aX + 16 * I);                        \
                const float2 v01 = make_float2(v4.x, v4.y), v23 = make_float2(v4.z, v4.w);                                          \
                const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);                                          \
                const float2 x01 = make_float2(x4.x, x4.y), x23 = make_float2(x4.z, x4.w);                                          \
                _Pragma("unroll") for (int s = 0; s < NS; ++s) {                        \
                    b[s][2 * t] = __ffma2_rn(nv[s], w01, __ffma2_rn(nw[s], v01, b[s][2 * t]));                                      \
                    b[s][2 * t + 1] = __ffma2_rn(nv[s], w23, __ffma2_rn(nw[s], v23, b[s][2 * t + 1]));                              \
                    acc[s][0] = __ffma2_rn(b[s][2 * t], x01, acc[s][0]);                \
                    acc[s][1] = __ffma2_rn(b[s][2 * t + 1], x23, acc[s][1]);            \
                }                                                                       \
            }
            static_assert(NL == 6, "tridiag_cols: the sweep is written for 6 chunks per warp");
            switch (tl) {
                case 6: VNLB_C1(5)
                case 5: VNLB_C1(4)
                case 4: VNLB_C1(3)
                case 3: VNLB_C1(2)
                case 2: VNLB_C1(1)
                case 1: VNLB_C1(0)
                default: break;
            }
#undef VNLB_C1
            if (own) {   // the window copy of this warp's row slot gets the same update
                const float4 v4 = lds128(aV + 16 * Iw), w4 = lds128(aW + 16 * Iw);
                wq0 = fmaf(-vo, w4.x, fmaf(-wo, v4.x, wq0)); wq1 = fmaf(-vo, w4.y, fmaf(-wo, v4.y, wq1));
                wq2 = fmaf(-vo, w4.z, fmaf(-wo, v4.z, wq2)); wq3 = fmaf(-vo, w4.w, fmaf(-wo, v4.w, wq3));
            }
            sweep_tail(acc, xn_s);
        }
        if (((c - 3) >> 2) != Iw) {                                            // next step needs column c-3: its owner publishes the chunk
            Iw = (c - 3) >> 2;
            if ((Iw & 3) == h) {
                switch (Iw >> 2) {
#define VNLB_CW(T) case (T): { constexpr int t = (T) < NL ? (T) : 0; _Pragma("unroll") for (int s = 0; s < NS; ++s)                \
                        *reinterpret_cast<float4 *>(win + 4 * row[s]) = make_float4(b[s][2 * t].x, b[s][2 * t].y, b[s][2 * t + 1].x, b[s][2 * t + 1].y); } break;
                    VNLB_CW(0) VNLB_CW(1) VNLB_CW(2) VNLB_CW(3) VNLB_CW(4) VNLB_CW(5)
#undef VNLB_CW
                    default: break;
                }
            }
            winload = true;
        }
        if (own) {
            x = (io < c - 1) ? xno : 0.f;
            r = rno;
        }
        pp ^= 1;
    }
    {   // trailing CEND x CEND matrix for the next phase: row i at trail + i * CEND
#pragma unroll
        for (int s = 0; s < CEND / 32; ++s)
#pragma unroll
            for (int t = 0; t < CEND / 16; ++t)
                *reinterpret_cast<float4 *>(trail + row[s] * CEND + 4 * (4 * t + h)) =
                    make_float4(b[s][2 * t].x, b[s][2 * t].y, b[s][2 * t + 1].x, b[s][2 * t + 1].y);
    }
    __syncthreads();
    constexpr int NK = K1 - K0 + 1;
    const int tid = 32 * h + lane;
    for (int idx = tid; idx < NK; idx += 128) {
        out[K0 + idx] = sd[idx];
        out[OPITCH + K0 + idx] = se[idx];
        out[2 * OPITCH + K0 + idx] = st[idx];
    }
    for (int idx = tid; idx < R1 - R0; idx += 128) out[OREFL + R0 + idx] = srf[idx];
}

// -------------------------------------------------------------- P2. tridiag_head_smem (first two steps)
// The first QDG - NR Householder steps on the symmetric matrix M (shared memory, pitch ld, index-reversed like
// everywhere here): plain three-pass steps (v; p = M v; rank-2 update), one thread per row -- two steps of 98 before
// tridiag_cols takes the 96 x 96 rest.  Same d / e / tau / reflector conventions as tridiag_regs.
template <int QDG, int NR, int OPITCH, int OREFL>
__device__ __forceinline__ void tridiag_head_smem(float *M, int ld, float *scr, float *out, int tid) {
    float *vs = scr, *ws = scr + ld, *red = scr + 2 * ld;
    int phase = 0;
    // block sum in LOGICAL warp order (tid = 32 * logical warp + lane): the result must not depend on warp_rotation()
    auto block_sum = [&](float v, float *rd, int &ph) -> float {
        v = warp_sum(v);
        float *r = rd + 4 * ph;
        if ((tid & 31) == 0) r[tid >> 5] = v;
        __syncthreads();
        ph ^= 1;
        return (r[0] + r[1]) + (r[2] + r[3]);
    };
    for (int k = 0; k < QDG - NR; ++k) {
        const int c = QDG - 1 - k;
        const bool act = tid < c;
        const float x = act ? M[tid * ld + c] : 0.f;
        const float alpha = M[(c - 1) * ld + c], dk = M[c * ld + c];
        const float ssq = block_sum((act && tid != c - 1) ? x * x : 0.f, red, phase);
        float *refl = out + OREFL + refl_off(QDG, k);
        if (ssq == 0.f) {                          // nothing to annihilate: H = I
            if (tid == 0) { out[k] = dk; out[OPITCH + k] = alpha; out[2 * OPITCH + k] = 0.f; }
            if (act) refl[c - 1 - tid] = (tid == c - 1) ? 1.f : 0.f;
            continue;
        }
        const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, ssq)), alpha);
        const float tau = (beta - alpha) / beta;
        const float scale = 1.f / (alpha - beta);
        const float v = act ? ((tid == c - 1) ? 1.f : x * scale) : 0.f;
        if (tid < ld) vs[tid] = v;
        if (act) refl[c - 1 - tid] = v;
        if (tid == 0) { out[k] = dk; out[OPITCH + k] = beta; out[2 * OPITCH + k] = tau; }
        __syncthreads();
        float p = 0.f;
        if (act) {
            const float *mr = M + tid * ld;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
            for (int j = 0; j < c; j += 4) {       // vs[j] = 0 for j >= c; the pitch covers the last chunk
                const float4 m4 = *reinterpret_cast<const float4 *>(mr + j), v4 = *reinterpret_cast<const float4 *>(vs + j);
                p0 = fmaf(m4.x, v4.x, p0); p1 = fmaf(m4.y, v4.y, p1); p2 = fmaf(m4.z, v4.z, p2); p3 = fmaf(m4.w, v4.w, p3);
            }
            p = (p0 + p1) + (p2 + p3);
        }
        const float sdot = block_sum(act ? p * v : 0.f, red, phase);
        const float w = act ? fmaf(-0.5f * tau * tau * sdot, v, tau * p) : 0.f;
        if (tid < ld) ws[tid] = w;
        __syncthreads();
        if (act) {
            float *mr = M + tid * ld;
            for (int j = 0; j < c; j += 4) {
                float4 m4 = *reinterpret_cast<float4 *>(mr + j);
                const float4 v4 = *reinterpret_cast<const float4 *>(vs + j), w4 = *reinterpret_cast<const float4 *>(ws + j);
                m4.x = fmaf(-v, w4.x, fmaf(-w, v4.x, m4.x)); m4.y = fmaf(-v, w4.y, fmaf(-w, v4.y, m4.y));
                m4.z = fmaf(-v, w4.z, fmaf(-w, v4.z, m4.z)); m4.w = fmaf(-v, w4.w, fmaf(-w, v4.w, m4.w));
                *reinterpret_cast<float4 *>(mr + j) = m4;
            }
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------- P3. kernels of the split path
// Workspace per problem (floats): d[LDG] e[LDG] tau[LDG] mean[LDG] reflectors[nref] trailing matrix[NR2 x NR2];
// tau[LDG-1] doubles as the "problem is valid" flag between the kernels of the split path.
constexpr int SPLIT_NR2 = 64;                      // trailing size handed from phase 1 to phase 2
template <int QD> __host__ __device__ constexpr int split_trail_off() { return 4 * ((QD + 3) & ~3) + ((((QD - 1) * QD / 2) + 3) & ~3); }
constexpr int SPLIT_NR3 = 32;                      // trailing size handed from phase 2 to phase 3
template <int QD> __host__ __device__ constexpr int split_trail2_off() { return split_trail_off<QD>() + SPLIT_NR2 * SPLIT_NR2; }
template <int QD> __host__ __device__ constexpr int split_ws_stride() { return split_trail2_off<QD>() + SPLIT_NR3 * SPLIT_NR3; }

// NRC: 4-column chunks of a row kept in registers (the rest of the row lives in shared memory until those columns are
// eliminated): NRC = NCH is everything in registers (166 registers, 3 CTAs per SM), NRC = SPLIT_NR2 / 4 keeps exactly
// the columns that survive this phase (<= 128 registers, 4 CTAs per SM).
template <bool FUSED, int QD, int NRC>
__global__ void __launch_bounds__(TT, NRC < ((QD + 3) / 4) ? 4 : 3) cov_tridiag_kernel(const BayesArgs a) {
    constexpr int LDQ = (QD + 3) & ~3, NCH = LDQ / 4;
    extern __shared__ __align__(16) float sm[];
    const VnlbBayesParams &P = a.P;
    const int n = P.k, ps = P.ps, ps2 = ps * ps, C = P.c;
    const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + warp_rotation()) & (TT / 32 - 1), tid = 32 * warp + lane;
    const int g = blockIdx.x / C, ch = blockIdx.x - g * C;
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    const bool valid_row = !a.inds || row_valid_block(a.inds + (long long)g * n, n);
    if (threadIdx.x == 0) wsp[3 * LDQ - 1] = valid_row ? 1.f : 0.f;   // flag for tridiag_tail_kernel
    if (!valid_row) return;
    float *Y = sm;                                   // Y[n][LDQ]: patches, columns reversed (column j = patch element QD-1-j)
    const int ybody = max(max(n, LDQ) * LDQ, tridiag_scratch_floats<QD, QD, SPLIT_NR2>());   // patches, then the matrix, then the scratch
    int *pb = (int *)(sm + ybody + 8);               // fused: offset of the patch corner in the image (8 floats of slack: tile reads)
    float *sv = sm;                                  // the tridiagonalisation's vectors re-use the head of Y
    const int rstride = P.pt * C * ps2;
    const long long HW = (long long)a.H * a.W, CHW = HW * C;
    if (FUSED) {
        int bad = 0;
        for (int nn = tid; nn < n; nn += TT) {
            int t, y, x;
            decode_ind(a.inds[(long long)g * n + nn], a.H, a.W, C, t, y, x);
            bad |= (t < 0 || t + P.pt > a.T || y + ps > a.H || x + ps > a.W);
            pb[nn] = (int)((long long)t * CHW + (long long)y * a.W + x);
        }
        if (__syncthreads_or(bad)) {                 // malformed index: the group is skipped (bayes_kernel does the same)
            if (threadIdx.x == 0) wsp[3 * LDQ - 1] = 0.f;
            return;
        }
    }
    auto col_off = [&](int j) -> int {
        const int dt = j / ps2, r = j - dt * ps2;
        if (FUSED) { const int dy = r / ps, dx = r - dy * ps; return (int)(dt * CHW + ch * HW + (long long)dy * a.W + dx); }
        return (dt * C + ch) * ps2 + r;
    };
    const float *src = FUSED ? (P.cov_from_basic ? a.img_basic : a.img_noisy)
                             : (P.cov_from_basic ? a.pbasic : a.pnoisy) + (long long)g * n * rstride;
    int co[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) co[q] = col_off(min(lane + 32 * q, QD - 1));
    // ---- stage all n patches (every load independent: one exposed latency)
    constexpr int SU = 5;                            // patches in flight per warp (20 independent loads per lane)
    for (int n0 = warp; n0 < n; n0 += SU * (TT / 32)) {
        float vals[SU][4];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = min(n0 + u * (TT / 32), n - 1);
            const float *q = src + (FUSED ? (long long)pb[nn] : (long long)nn * rstride);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) vals[u][qq] = (lane + 32 * qq < QD) ? q[co[qq]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = n0 + u * (TT / 32);
            if (nn < n) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const int j = lane + 32 * qq;
                    if (j < QD) Y[nn * LDQ + (QD - 1 - j)] = vals[u][qq];
                    else if (j < LDQ) Y[nn * LDQ + j] = 0.f;                         // zero pad columns QD..LDQ-1
                }
            }
        }
    }
    __syncthreads();
    // ---- centre (same summation order as bayes_kernel: 4 interleaved partial sums)
    const float inv_n = 1.f / (float)n;
    if (tid < LDQ) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int nn = 0;
        for (; nn + 3 < n; nn += 4) {
            s0 += Y[nn * LDQ + tid]; s1 += Y[(nn + 1) * LDQ + tid];
            s2 += Y[(nn + 2) * LDQ + tid]; s3 += Y[(nn + 3) * LDQ + tid];
        }
        for (; nn < n; ++nn) s0 += Y[nn * LDQ + tid];
        const float mj = ((s0 + s1) + (s2 + s3)) * inv_n;
        for (nn = 0; nn < n; ++nn) Y[nn * LDQ + tid] -= mj;
        if (tid < QD) wsp[3 * LDQ + (QD - 1 - tid)] = mj; else wsp[3 * LDQ + tid] = 0.f;
    }
    __syncthreads();
    // ---- covariance: 8 x 8 register tiles of the lower triangle, accumulated over the patches in order (entries
    //      bit-identical to bayes_kernel).  The kernel is bound by the shared-memory data pipe (ncu: 73-79 % busy), so the
    //      covariance is formed where a loaded operand feeds most FMAs (4 LDS.128 per 32 FFMA2), then mirrored through
    //      shared memory into the row-per-thread register layout of the tridiagonalisation.
    static_assert(NRC == NCH, "cov_tridiag_kernel: rows fully in registers");
    constexpr int NT8 = (LDQ + 7) / 8, NTRI = NT8 * (NT8 + 1) / 2;
    static_assert(NTRI <= TT, "one 8 x 8 tile per thread");
    int ti = -1, tj = 0;
    if (tid < NTRI) {
        ti = (int)((sqrtf(8.f * tid + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > tid) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
        tj = tid - ti * (ti + 1) / 2;
    }
    {
        // Tiles with tj & 4 load (and keep) their two 4-column halves in swapped order: the column-part loads of 8 consecutive
        // tiles then fall into 8 different 16-byte bank groups (2 tj, 2 tj + 2, ... alone cover only the even ones: ncu showed
        // 52 % excess wavefronts on these loads, 18 % of the kernel's shared-memory traffic).
        const int sw = (tj >> 2) & 1;
        float2 acc[8][4];
#pragma unroll
        for (int aa = 0; aa < 8; ++aa)
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = make_float2(0.f, 0.f);
        if (ti >= 0) {
            const float *ra = Y + 8 * ti, *rb = Y + 8 * tj;      // (columns >= LDQ of the last tile read the next row: never stored)
            for (int nn = 0; nn < n; ++nn) {
                const float4 a0 = *reinterpret_cast<const float4 *>(ra + nn * LDQ), a1 = *reinterpret_cast<const float4 *>(ra + nn * LDQ + 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(rb + nn * LDQ + 4 * sw), b1 = *reinterpret_cast<const float4 *>(rb + nn * LDQ + 4 - 4 * sw);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float2 c0 = make_float2(b0.x, b0.y), c1 = make_float2(b0.z, b0.w), c2 = make_float2(b1.x, b1.y), c3 = make_float2(b1.z, b1.w);
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const float2 ad = make_float2(av[aa], av[aa]);
                    acc[aa][0] = __ffma2_rn(ad, c0, acc[aa][0]); acc[aa][1] = __ffma2_rn(ad, c1, acc[aa][1]);
                    acc[aa][2] = __ffma2_rn(ad, c2, acc[aa][2]); acc[aa][3] = __ffma2_rn(ad, c3, acc[aa][3]);
                }
            }
        }
        __syncthreads();                             // Y is dead: its place takes the full symmetric matrix A[LDQ][LDQ]
        if (ti >= 0) {
            float *A = Y;
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {         // rows 8 ti + aa, columns 8 tj .. 8 tj + 7
                const int i = 8 * ti + aa;
                if (i < LDQ) {
                    if (8 * tj + 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 * sw) = make_float4(acc[aa][0].x * inv_n, acc[aa][0].y * inv_n, acc[aa][1].x * inv_n, acc[aa][1].y * inv_n);
                    if (8 * tj + 4 - 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 - 4 * sw) = make_float4(acc[aa][2].x * inv_n, acc[aa][2].y * inv_n, acc[aa][3].x * inv_n, acc[aa][3].y * inv_n);
                }
            }
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {         // mirrored: rows 8 tj + bb, columns 8 ti .. 8 ti + 7
                const int j = 8 * tj + bb;
                if (j < LDQ) {
                    float cv[8];
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) {
                        const float2 pr = sw ? acc[aa][(bb >> 1) ^ 2] : acc[aa][bb >> 1];   // tile column bb -> register pair
                        cv[aa] = ((bb & 1) ? pr.y : pr.x) * inv_n;
                    }
                    *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    if (8 * ti + 4 < LDQ) *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
                }
            }
        }
    }
    __syncthreads();
    float2 b[2 * NCH];                               // row t of B = C reversed
    float dg = 0.f;
    {
        const float *Ar = Y + min(tid, QD - 1) * LDQ;
        const float live = tid < QD ? 1.f : 0.f;
#pragma unroll
        for (int jj = 0; jj < NCH; ++jj) {
            const float4 f = *reinterpret_cast<const float4 *>(Ar + 4 * jj);
            b[2 * jj] = make_float2(f.x * live, f.y * live);
            b[2 * jj + 1] = make_float2(f.z * live, f.w * live);
        }
        dg = Ar[min(tid, QD - 1)] * live;            // diagonal entry (already divided by n)
    }
    if (a.dbg_mat) {                                 // parity hook: the covariance in natural (un-reversed) index order
        float *o = a.dbg_mat + (size_t)blockIdx.x * QD * QD;
        for (int idx = threadIdx.x; idx < QD * QD; idx += TT) {
            const int i = idx / QD, j = idx - i * QD;
            o[idx] = Y[(QD - 1 - i) * LDQ + (QD - 1 - j)];
        }
    }
    __syncthreads();                                 // A is dead
    if (a.rank_var) {                                // rank_var = mean over channels of trace(C) (bayes_est.py:39-40)
        float tr = warp_sum(dg);
        float *red = (float *)(pb + ((n + 3) & ~3));
        if (lane == 0) red[warp] = tr;
        __syncthreads();
        if (tid == 0) atomicAdd(&a.rank_var[g], ((red[0] + red[1]) + (red[2] + red[3])) / (float)C);
    }
    float *bs_row = nullptr;
    tridiag_regs<QD, QD, SPLIT_NR2, TT, NRC>(b, sv, wsp, wsp + split_trail_off<QD>(), tid, smem_u32(bs_row));
}

// cov_tridiag_kernel with the elimination 98 -> 64 done by tridiag_head_smem (two steps) + tridiag_cols (columns split over
// the warps, three rows per lane) instead of tridiag_regs: the matrix stays in shared memory until the warps have
// taken their register tiles.  Same workspace contents (to rounding).
constexpr int COLS_NR = 96;
template <bool FUSED, int QD>
__global__ void __launch_bounds__(TT, 4) cov_tridiag4_kernel(const BayesArgs a) {
    constexpr int NRC = (QD + 3) / 4;
    constexpr int LDQ = (QD + 3) & ~3, NCH = LDQ / 4;
    extern __shared__ __align__(16) float sm[];
    const VnlbBayesParams &P = a.P;
    const int n = P.k, ps = P.ps, ps2 = ps * ps, C = P.c;
    const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + warp_rotation()) & (TT / 32 - 1), tid = 32 * warp + lane;
    const int g = blockIdx.x / C, ch = blockIdx.x - g * C;
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    const bool valid_row = !a.inds || row_valid_block(a.inds + (long long)g * n, n);
    if (threadIdx.x == 0) wsp[3 * LDQ - 1] = valid_row ? 1.f : 0.f;   // flag for tridiag_tail_kernel
    if (!valid_row) return;
    float *Y = sm;                                   // Y[n][LDQ]: patches, columns reversed (column j = patch element QD-1-j)
    const int ybody = max(n, LDQ) * LDQ + tridiag_cols_scratch_floats<QD, COLS_NR, SPLIT_NR2>();   // patches / the matrix, then the scratch
    int *pb = (int *)(sm + ybody + 8);               // fused: offset of the patch corner in the image (8 floats of slack: tile reads)
    float *sv = sm + max(n, LDQ) * LDQ;              // scratch of the tridiagonalisation, behind the matrix
    const int rstride = P.pt * C * ps2;
    const long long HW = (long long)a.H * a.W, CHW = HW * C;
    if (FUSED) {
        int bad = 0;
        for (int nn = tid; nn < n; nn += TT) {
            int t, y, x;
            decode_ind(a.inds[(long long)g * n + nn], a.H, a.W, C, t, y, x);
            bad |= (t < 0 || t + P.pt > a.T || y + ps > a.H || x + ps > a.W);
            pb[nn] = (int)((long long)t * CHW + (long long)y * a.W + x);
        }
        if (__syncthreads_or(bad)) {                 // malformed index: the group is skipped (bayes_kernel does the same)
            if (threadIdx.x == 0) wsp[3 * LDQ - 1] = 0.f;
            return;
        }
    }
    auto col_off = [&](int j) -> int {
        const int dt = j / ps2, r = j - dt * ps2;
        if (FUSED) { const int dy = r / ps, dx = r - dy * ps; return (int)(dt * CHW + ch * HW + (long long)dy * a.W + dx); }
        return (dt * C + ch) * ps2 + r;
    };
    const float *src = FUSED ? (P.cov_from_basic ? a.img_basic : a.img_noisy)
                             : (P.cov_from_basic ? a.pbasic : a.pnoisy) + (long long)g * n * rstride;
    int co[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) co[q] = col_off(min(lane + 32 * q, QD - 1));
    // -------------------------------------------------------------- P4. cov4: stage the patches
    // ---- stage all n patches (every load independent: one exposed latency)
    constexpr int SU = 5;                            // patches in flight per warp (20 independent loads per lane)
    for (int n0 = warp; n0 < n; n0 += SU * (TT / 32)) {
        float vals[SU][4];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = min(n0 + u * (TT / 32), n - 1);
            const float *q = src + (FUSED ? (long long)pb[nn] : (long long)nn * rstride);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) vals[u][qq] = (lane + 32 * qq < QD) ? q[co[qq]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = n0 + u * (TT / 32);
            if (nn < n) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const int j = lane + 32 * qq;
                    if (j < QD) Y[nn * LDQ + (QD - 1 - j)] = vals[u][qq];
                    else if (j < LDQ) Y[nn * LDQ + j] = 0.f;                         // zero pad columns QD..LDQ-1
                }
            }
        }
    }
    __syncthreads();
    // -------------------------------------------------------------- P5. cov4: centre
    // ---- centre (same summation order as bayes_kernel: 4 interleaved partial sums)
    const float inv_n = 1.f / (float)n;
    if (tid < LDQ) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int nn = 0;
        for (; nn + 3 < n; nn += 4) {
            s0 += Y[nn * LDQ + tid]; s1 += Y[(nn + 1) * LDQ + tid];
            s2 += Y[(nn + 2) * LDQ + tid]; s3 += Y[(nn + 3) * LDQ + tid];
        }
        for (; nn < n; ++nn) s0 += Y[nn * LDQ + tid];
        const float mj = ((s0 + s1) + (s2 + s3)) * inv_n;
        for (nn = 0; nn < n; ++nn) Y[nn * LDQ + tid] -= mj;
        if (tid < QD) wsp[3 * LDQ + (QD - 1 - tid)] = mj; else wsp[3 * LDQ + tid] = 0.f;
    }
    __syncthreads();
    // -------------------------------------------------------------- P6. cov4: covariance tiles + mirror
    // ---- covariance: 8 x 8 register tiles of the lower triangle, accumulated over the patches in order (entries
    //      bit-identical to bayes_kernel).  The kernel is bound by the shared-memory data pipe (ncu: 73-79 % busy), so the
    //      covariance is formed where a loaded operand feeds most FMAs (4 LDS.128 per 32 FFMA2), then mirrored through
    //      shared memory into the row-per-thread register layout of the tridiagonalisation.
    static_assert(NRC == NCH, "cov_tridiag_kernel: rows fully in registers");
    constexpr int NT8 = (LDQ + 7) / 8, NTRI = NT8 * (NT8 + 1) / 2;
    static_assert(NTRI <= TT, "one 8 x 8 tile per thread");
    int ti = -1, tj = 0;
    if (tid < NTRI) {
        ti = (int)((sqrtf(8.f * tid + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > tid) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
        tj = tid - ti * (ti + 1) / 2;
    }
    {
        // Tiles with tj & 4 load (and keep) their two 4-column halves in swapped order: the column-part loads of 8 consecutive
        // tiles then fall into 8 different 16-byte bank groups (2 tj, 2 tj + 2, ... alone cover only the even ones: ncu showed
        // 52 % excess wavefronts on these loads, 18 % of the kernel's shared-memory traffic).
        const int sw = (tj >> 2) & 1;
        float2 acc[8][4];
#pragma unroll
        for (int aa = 0; aa < 8; ++aa)
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = make_float2(0.f, 0.f);
        if (ti >= 0) {
            const float *ra = Y + 8 * ti, *rb = Y + 8 * tj;      // (columns >= LDQ of the last tile read the next row: never stored)
            for (int nn = 0; nn < n; ++nn) {
                const float4 a0 = *reinterpret_cast<const float4 *>(ra + nn * LDQ), a1 = *reinterpret_cast<const float4 *>(ra + nn * LDQ + 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(rb + nn * LDQ + 4 * sw), b1 = *reinterpret_cast<const float4 *>(rb + nn * LDQ + 4 - 4 * sw);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float2 c0 = make_float2(b0.x, b0.y), c1 = make_float2(b0.z, b0.w), c2 = make_float2(b1.x, b1.y), c3 = make_float2(b1.z, b1.w);
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const float2 ad = make_float2(av[aa], av[aa]);
                    acc[aa][0] = __ffma2_rn(ad, c0, acc[aa][0]); acc[aa][1] = __ffma2_rn(ad, c1, acc[aa][1]);
                    acc[aa][2] = __ffma2_rn(ad, c2, acc[aa][2]); acc[aa][3] = __ffma2_rn(ad, c3, acc[aa][3]);
                }
            }
        }
        __syncthreads();                             // Y is dead: its place takes the full symmetric matrix A[LDQ][LDQ]
        if (ti >= 0) {
            float *A = Y;
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {         // rows 8 ti + aa, columns 8 tj .. 8 tj + 7
                const int i = 8 * ti + aa;
                if (i < LDQ) {
                    if (8 * tj + 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 * sw) = make_float4(acc[aa][0].x * inv_n, acc[aa][0].y * inv_n, acc[aa][1].x * inv_n, acc[aa][1].y * inv_n);
                    if (8 * tj + 4 - 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 - 4 * sw) = make_float4(acc[aa][2].x * inv_n, acc[aa][2].y * inv_n, acc[aa][3].x * inv_n, acc[aa][3].y * inv_n);
                }
            }
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {         // mirrored: rows 8 tj + bb, columns 8 ti .. 8 ti + 7
                const int j = 8 * tj + bb;
                if (j < LDQ) {
                    float cv[8];
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) {
                        const float2 pr = sw ? acc[aa][(bb >> 1) ^ 2] : acc[aa][bb >> 1];   // tile column bb -> register pair
                        cv[aa] = ((bb & 1) ? pr.y : pr.x) * inv_n;
                    }
                    *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    if (8 * ti + 4 < LDQ) *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
                }
            }
        }
    }
    __syncthreads();
    // -------------------------------------------------------------- P7. cov4: trace, then the elimination calls
    float dg = (tid < QD) ? Y[tid * LDQ + tid] : 0.f;   // diagonal entry (already divided by n)
    if (a.dbg_mat) {                                 // parity hook: the covariance in natural (un-reversed) index order
        float *o = a.dbg_mat + (size_t)blockIdx.x * QD * QD;
        for (int idx = threadIdx.x; idx < QD * QD; idx += TT) {
            const int i = idx / QD, j = idx - i * QD;
            o[idx] = Y[(QD - 1 - i) * LDQ + (QD - 1 - j)];
        }
    }
    if (a.rank_var) {                                // rank_var = mean over channels of trace(C) (bayes_est.py:39-40)
        float tr = warp_sum(dg);
        float *red = (float *)(pb + ((n + 3) & ~3));
        if (lane == 0) red[warp] = tr;
        __syncthreads();
        if (tid == 0) atomicAdd(&a.rank_var[g], ((red[0] + red[1]) + (red[2] + red[3])) / (float)C);
    }
    __syncthreads();
    tridiag_head_smem<QD, COLS_NR, LDQ, 4 * LDQ>(Y, LDQ, sv, wsp, tid);
    tridiag_cols<QD, COLS_NR, SPLIT_NR2, LDQ, 4 * LDQ>(Y, LDQ, sv, wsp, wsp + split_trail_off<QD>(), lane, warp);
}

// -------------------------------------------------------------- P8. later kernels
// Split path, later phases of the tridiagonalisation: the NR x NR trailing matrix comes from the workspace (float
// offset TIN, row pitch NR), columns NR-1 .. CEND are eliminated by NT = round32(NR) threads, and unless CEND == 2 the
// CEND x CEND rest goes back to the workspace at TOUT.  Registers ~ NR + 62: 64 threads x 126 registers => 8 CTAs per SM
// for NR = 64, 32 threads x 94 => 16+ CTAs per SM for NR = 32 -- occupancy follows the shrinking problem.
template <int QDG, int NR, int CEND, int NT, int OCC, int OPITCH, int OREFL, int TIN, int TOUT>
__global__ void __launch_bounds__(NT, OCC) tridiag_tail_kernel(const BayesArgs a) {
    constexpr int NCH = NR / 4;
    static_assert(NR % 4 == 0 && NR <= NT && NT % 32 == 0 && NT <= 64, "tridiag_tail_kernel: one row per thread, at most 2 warps");
    extern __shared__ __align__(16) float sm[];
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    if (wsp[3 * OPITCH - 1] == 0.f) return;           // group skipped by the first kernel
    const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + warp_rotation()) & (NT / 32 - 1), tid = 32 * warp + lane;
    const float4 *tr = reinterpret_cast<const float4 *>(wsp + TIN + min(tid, NR - 1) * NR);
    float2 b[2 * NCH];
#pragma unroll
    for (int I = 0; I < NCH; ++I) {
        const float4 f = tr[I];
        b[2 * I] = make_float2(f.x, f.y);
        b[2 * I + 1] = make_float2(f.z, f.w);
    }
    tridiag_regs<QDG, NR, CEND, NT, NCH, OPITCH, OREFL>(b, sm, wsp, wsp + TOUT, tid);
}


// Two-rows-per-thread version of a tail phase (tridiag_regs2): ONE warp per problem, the NR x NR trailing matrix (row pitch
// NR at float offset TIN of the problem's workspace) -> columns NR-1 .. 32 eliminated -> the 32 x 32 rest at TOUT.
template <int QDG, int NR, int OCC, int OPITCH, int OREFL, int TIN, int TOUT>
__global__ void __launch_bounds__(32, OCC) tridiag_tail2_kernel(const BayesArgs a) {
    constexpr int NCH = (NR + 3) / 4;
    static_assert(NR % 4 == 0, "tridiag_tail2_kernel: NR a multiple of 4");
    extern __shared__ __align__(16) float sm[];
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    if (wsp[3 * OPITCH - 1] == 0.f) return;           // group skipped by the first kernel
    const int lane = threadIdx.x;
    float2 bl[2 * NCH], bh[2 * NCH];
    const float4 *trl = reinterpret_cast<const float4 *>(wsp + TIN + lane * NR);
    const float4 *trh = reinterpret_cast<const float4 *>(wsp + TIN + min(lane + 32, NR - 1) * NR);
    const float live = lane + 32 < NR ? 1.f : 0.f;
#pragma unroll
    for (int I = 0; I < NCH; ++I) {
        const float4 f = trl[I], g = trh[I];
        bl[2 * I] = make_float2(f.x, f.y);
        bl[2 * I + 1] = make_float2(f.z, f.w);
        bh[2 * I] = make_float2(g.x * live, g.y * live);
        bh[2 * I + 1] = make_float2(g.z * live, g.w * live);
    }
    tridiag_regs2<QDG, NR, 32, OPITCH, OREFL>(bl, bh, sm, wsp, wsp + TOUT, lane);
}

// Split path of the Gram variant (step 2: n = 60 patches < p = 98 elements): centre + Gram matrix G = Yc Yc^T / n +
// its whole tridiagonalisation, one CTA of 64 threads per (group, channel) problem, 8 CTAs per SM.  Same structure as
// cov_tridiag_kernel with the roles of patches and patch elements swapped: the patches are staged TRANSPOSED,
// Yt[j][t] = element j of patch QD-1-t, so that thread t ends up with row t of the reversed Gram matrix.
// Workspace per problem (floats): d[64] e[64] tau[64] (tau[63] = valid flag) mean[LD] reflectors[nref].
constexpr int GRAM_PITCH = 64;
template <int QD> __host__ __device__ constexpr int gram_trail_off(int LD) { return 3 * GRAM_PITCH + LD + ((((QD - 1) * QD / 2) + 3) & ~3); }
template <int QD> __host__ __device__ constexpr int gram_full_off(int LD) { return gram_trail_off<QD>(LD) + SPLIT_NR3 * SPLIT_NR3; }   // the whole QD x QD Gram matrix (HANDOFF)
template <int QD> __host__ __device__ constexpr int gram_ws_stride(int LD) { return gram_full_off<QD>(LD) + QD * QD; }

// HANDOFF: the kernel stops after the Gram matrix and hands it (reversed order, row pitch QD) to tridiag_tail2_kernel
// through the workspace: the elimination 60 -> 32 then runs with two rows per thread, one warp per problem.
template <bool FUSED, int QD, bool HANDOFF>
__global__ void __launch_bounds__(64, 8) gram_tridiag_kernel(const BayesArgs a) {
    constexpr int NT = 64, LDQ = (QD + 3) & ~3, NCH = LDQ / 4;
    static_assert(QD <= 64 && QD % 4 == 0, "gram_tridiag_kernel: one row per thread of 2 warps");
    extern __shared__ __align__(16) float sm[];
    const VnlbBayesParams &P = a.P;
    const TriLayout &L = a.L;
    const int n = P.k, ps = P.ps, ps2 = ps * ps, C = P.c, p = L.p, LD = L.LD;
    const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + warp_rotation()) & 1, tid = 32 * warp + lane;
    const int g = blockIdx.x / C, ch = blockIdx.x - g * C;
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    const bool valid_row = !a.inds || row_valid_block(a.inds + (long long)g * n, n);
    if (threadIdx.x == 0) wsp[3 * GRAM_PITCH - 1] = valid_row ? 1.f : 0.f;
    if (!valid_row) return;
    // shared memory: Yt[p][LDQ] (+8 floats of slack for the tile reads), later A[LDQ][LDQ] and the scratch of tridiag_regs; pb[n]
    const int ybody = max(max(p * LDQ, LDQ * LDQ), tridiag_scratch_floats<QD, QD, SPLIT_NR3>());
    float *Yt = sm;
    int *pb = (int *)(sm + ybody + 8);
    const int rstride = P.pt * C * ps2;
    const long long HW = (long long)a.H * a.W, CHW = HW * C;
    if (FUSED) {
        int bad = 0;
        for (int nn = tid; nn < n; nn += NT) {
            int t, y, x;
            decode_ind(a.inds[(long long)g * n + nn], a.H, a.W, C, t, y, x);
            bad |= (t < 0 || t + P.pt > a.T || y + ps > a.H || x + ps > a.W);
            pb[nn] = (int)((long long)t * CHW + (long long)y * a.W + x);
        }
        if (__syncthreads_or(bad)) {
            if (threadIdx.x == 0) wsp[3 * GRAM_PITCH - 1] = 0.f;
            return;
        }
    }
    auto col_off = [&](int j) -> int {
        const int dt = j / ps2, r = j - dt * ps2;
        if (FUSED) { const int dy = r / ps, dx = r - dy * ps; return (int)(dt * CHW + ch * HW + (long long)dy * a.W + dx); }
        return (dt * C + ch) * ps2 + r;
    };
    const float *src = FUSED ? (P.cov_from_basic ? a.img_basic : a.img_noisy)
                             : (P.cov_from_basic ? a.pbasic : a.pnoisy) + (long long)g * n * rstride;
    int co[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) co[q] = col_off(min(lane + 32 * q, p - 1));
    // ---- stage the patches transposed and reversed: Yt[j][QD-1-nn] = patch nn, element j (5 patches in flight per warp)
    constexpr int SU = 5;
    for (int n0 = warp; n0 < LDQ; n0 += SU * (NT / 32)) {
        float vals[SU][4];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = min(n0 + u * (NT / 32), n - 1);
            const float *q = src + (FUSED ? (long long)pb[nn] : (long long)nn * rstride);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) vals[u][qq] = q[co[qq]];
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int nn = n0 + u * (NT / 32);
            if (nn < LDQ) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const int j = lane + 32 * qq;
                    if (j < p) Yt[j * LDQ + (LDQ - 1 - nn)] = nn < n ? vals[u][qq] : 0.f;   // patches n..LDQ-1: zero pad (n == QD == LDQ here)
                }
            }
        }
    }
    __syncthreads();
    // ---- centre every patch element (same summation order as bayes_kernel: 4 interleaved partial sums over the patches)
    const float inv_n = 1.f / (float)n;
    for (int j = tid; j < LD; j += NT) {
        float mj = 0.f;
        if (j < p) {
            float *row = Yt + j * LDQ + (LDQ - 1);          // patch nn at row[-nn]
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int nn = 0;
            for (; nn + 3 < n; nn += 4) { s0 += row[-nn]; s1 += row[-(nn + 1)]; s2 += row[-(nn + 2)]; s3 += row[-(nn + 3)]; }
            for (; nn < n; ++nn) s0 += row[-nn];
            mj = ((s0 + s1) + (s2 + s3)) * inv_n;
            for (nn = 0; nn < n; ++nn) row[-nn] -= mj;
        }
        wsp[3 * GRAM_PITCH + j] = mj;
    }
    __syncthreads();
    // ---- Gram matrix in 8 x 8 register tiles of the lower triangle (one tile per thread), summed over the elements in order
    constexpr int NT8 = (LDQ + 7) / 8, NTRI = NT8 * (NT8 + 1) / 2;
    static_assert(NTRI <= NT, "one 8 x 8 tile per thread");
    int ti = -1, tj = 0;
    if (tid < NTRI) {
        ti = (int)((sqrtf(8.f * tid + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > tid) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
        tj = tid - ti * (ti + 1) / 2;
    }
    {
        // Tiles with tj & 4 load (and keep) their two 4-column halves in swapped order: the column-part loads of 8 consecutive
        // tiles then fall into 8 different 16-byte bank groups (2 tj, 2 tj + 2, ... alone cover only the even ones: ncu showed
        // 52 % excess wavefronts on these loads, 18 % of the kernel's shared-memory traffic).
        const int sw = (tj >> 2) & 1;
        float2 acc[8][4];
#pragma unroll
        for (int aa = 0; aa < 8; ++aa)
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = make_float2(0.f, 0.f);
        if (ti >= 0) {
            const float *ra = Yt + 8 * ti, *rb = Yt + 8 * tj;
            for (int j = 0; j < p; ++j) {
                const float4 a0 = *reinterpret_cast<const float4 *>(ra + j * LDQ), a1 = *reinterpret_cast<const float4 *>(ra + j * LDQ + 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(rb + j * LDQ + 4 * sw), b1 = *reinterpret_cast<const float4 *>(rb + j * LDQ + 4 - 4 * sw);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float2 c0 = make_float2(b0.x, b0.y), c1 = make_float2(b0.z, b0.w), c2 = make_float2(b1.x, b1.y), c3 = make_float2(b1.z, b1.w);
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const float2 ad = make_float2(av[aa], av[aa]);
                    acc[aa][0] = __ffma2_rn(ad, c0, acc[aa][0]); acc[aa][1] = __ffma2_rn(ad, c1, acc[aa][1]);
                    acc[aa][2] = __ffma2_rn(ad, c2, acc[aa][2]); acc[aa][3] = __ffma2_rn(ad, c3, acc[aa][3]);
                }
            }
        }
        __syncthreads();                             // Yt is dead: its place takes the full symmetric matrix A[LDQ][LDQ]
        if (ti >= 0) {
            float *A = Yt;
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const int i = 8 * ti + aa;
                if (i < LDQ) {
                    if (8 * tj + 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 * sw) = make_float4(acc[aa][0].x * inv_n, acc[aa][0].y * inv_n, acc[aa][1].x * inv_n, acc[aa][1].y * inv_n);
                    if (8 * tj + 4 - 4 * sw < LDQ)
                        *reinterpret_cast<float4 *>(A + i * LDQ + 8 * tj + 4 - 4 * sw) = make_float4(acc[aa][2].x * inv_n, acc[aa][2].y * inv_n, acc[aa][3].x * inv_n, acc[aa][3].y * inv_n);
                }
            }
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
                const int j = 8 * tj + bb;
                if (j < LDQ) {
                    float cv[8];
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) {
                        const float2 pr = sw ? acc[aa][(bb >> 1) ^ 2] : acc[aa][bb >> 1];   // tile column bb -> register pair
                        cv[aa] = ((bb & 1) ? pr.y : pr.x) * inv_n;
                    }
                    *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    if (8 * ti + 4 < LDQ) *reinterpret_cast<float4 *>(A + j * LDQ + 8 * ti + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
                }
            }
        }
    }
    __syncthreads();
    if constexpr (HANDOFF) {
        static_assert(LDQ == QD, "HANDOFF: QD a multiple of 4");
        float4 *o = reinterpret_cast<float4 *>(wsp + gram_full_off<QD>(100));
        const float4 *A4 = reinterpret_cast<const float4 *>(Yt);
        for (int idx = threadIdx.x; idx < QD * QD / 4; idx += NT) o[idx] = A4[idx];
        if (a.dbg_mat) {                                 // parity hook: the Gram matrix in natural (un-reversed) patch order
            float *od = a.dbg_mat + (size_t)blockIdx.x * QD * QD;
            for (int idx = threadIdx.x; idx < QD * QD; idx += NT) {
                const int i = idx / QD, j = idx - i * QD;
                od[idx] = Yt[(QD - 1 - i) * LDQ + (QD - 1 - j)];
            }
        }
        if (a.rank_var) {                                // rank_var = mean over channels of trace(G) (bayes_est.py:39-40)
            float tr = warp_sum(tid < QD ? Yt[tid * LDQ + tid] : 0.f);
            __syncthreads();                             // (Yt is read above; red lives behind pb)
            float *red = (float *)(pb + ((n + 3) & ~3));
            if (lane == 0) red[warp] = tr;
            __syncthreads();
            if (tid == 0) atomicAdd(&a.rank_var[g], (red[0] + red[1]) / (float)C);
        }
        return;
    }
    float2 b[2 * NCH];
    float dg;
    {
        const float *Ar = Yt + min(tid, QD - 1) * LDQ;
        const float live = tid < QD ? 1.f : 0.f;
#pragma unroll
        for (int jj = 0; jj < NCH; ++jj) {
            const float4 f = *reinterpret_cast<const float4 *>(Ar + 4 * jj);
            b[2 * jj] = make_float2(f.x * live, f.y * live);
            b[2 * jj + 1] = make_float2(f.z * live, f.w * live);
        }
        dg = Ar[min(tid, QD - 1)] * live;
    }
    if (a.dbg_mat) {                                 // parity hook: the Gram matrix in natural (un-reversed) patch order
        float *o = a.dbg_mat + (size_t)blockIdx.x * QD * QD;
        for (int idx = threadIdx.x; idx < QD * QD; idx += NT) {
            const int i = idx / QD, j = idx - i * QD;
            o[idx] = Yt[(QD - 1 - i) * LDQ + (QD - 1 - j)];
        }
    }
    __syncthreads();                                 // A is dead
    if (a.rank_var) {                                // rank_var = mean over channels of trace(C) = trace(G) (bayes_est.py:39-40)
        float tr = warp_sum(dg);
        float *red = (float *)(pb + ((n + 3) & ~3));
        if (lane == 0) red[warp] = tr;
        __syncthreads();
        if (tid == 0) atomicAdd(&a.rank_var[g], (red[0] + red[1]) / (float)C);
    }
    tridiag_regs<QD, QD, SPLIT_NR3, NT, NCH, GRAM_PITCH, 3 * GRAM_PITCH + 100>(b, sm, wsp, wsp + gram_trail_off<QD>(100), tid);   // columns QD-1 .. 32; the rest in tridiag_tail_kernel
}

// SPLIT: phases 0-1 were done by cov_tridiag_kernel; (d, e, tau, mean, packed reflectors) come from the workspace.
// MMA: the Wiener filter on the tensor cores (filter_chunk_mma); its own instantiation, so that the default kernels carry
// none of its code (registers, instruction cache)
template <bool FUSED, bool GRAM, bool SPLIT, bool MMA = false>
__global__ void __launch_bounds__(TT, GRAM ? 7 : 5) bayes_kernel(const BayesArgs a) {
    constexpr int NT = GRAM ? 1 : 3;               // 4x4 tiles of the (covariance | Gram) matrix per thread
    extern __shared__ __align__(16) float sm[];
    const VnlbBayesParams &P = a.P;
    const TriLayout &L = a.L;
    const int n = P.k, ps = P.ps, ps2 = ps * ps, p = L.p, LD = L.LD, XS = L.XS, C = P.c;
    const int qd = L.q, LDq = L.LDq, ZSq = L.ZSq;   // eigenproblem dimension (p, or n with the Gram trick)
    const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + warp_rotation()) & (TT / 32 - 1), tid = 32 * warp + lane;
    const int g = FUSED ? blockIdx.x : blockIdx.x / C;
    if (a.inds && !row_valid_block(a.inds + (long long)g * n, n)) return;

    float *R = sm + L.oR, *d = sm + L.oD, *e = sm + L.oE, *e2 = sm + L.oE2, *taus = sm + L.oTau;
    float *v = sm + L.oV, *w = sm + L.oW, *mean = sm + L.oMean, *lam = sm + L.oLam, *coef = sm + L.oCoef;
    float *blo = sm + L.oLo, *bhi = sm + L.oHi;
    int *nlo = (int *)(sm + L.oNlo), *nhi = (int *)(sm + L.oNhi);
    float *red = sm + L.oRed;
    int *pb = (int *)(sm + L.oPb);                 // fused: offset of the patch corner in the image
    int *wb = pb + ((n + 3) & ~3);                 // fused: offset of the patch corner in `weights`
    int phase = 0;

    const bool step2 = P.step == 1;
    const float inv_n = 1.f / (float)n;
    const int rstride = P.pt * C * ps2;            // stack mode: floats between consecutive patches
    const long long HW = (long long)a.H * a.W, CHW = HW * C;
    const long long gbase = (long long)g * n * rstride;
    bool is_flat = false;

    if (FUSED) {
        int bad = 0;
        for (int nn = tid; nn < n; nn += TT) {
            int t, y, x;
            decode_ind(a.inds[(long long)g * n + nn], a.H, a.W, C, t, y, x);
            bad |= (t < 0 || t + P.pt > a.T || y + ps > a.H || x + ps > a.W);
            pb[nn] = (int)((long long)t * CHW + (long long)y * a.W + x);
            wb[nn] = (int)((long long)t * HW + (long long)y * a.W + x);
        }
        if (__syncthreads_or(bad)) return;         // malformed index: the group is skipped
    } else if (step2 && a.flat) {
        is_flat = a.flat[g] != 0;
    }
    // offset of patch element j inside a patch: stack layout or image layout
    auto col_off = [&](int j, int ch) -> int {
        const int dt = j / ps2, r = j - dt * ps2;
        if (FUSED) { const int dy = r / ps, dx = r - dy * ps; return (int)(dt * CHW + ch * HW + (long long)dy * a.W + dx); }
        return (dt * C + ch) * ps2 + r;
    };
    auto row_off = [&](int nn) -> long long { return FUSED ? (long long)pb[nn] : (long long)nn * rstride; };
    const float *base_noisy = FUSED ? a.img_noisy : a.pnoisy + gbase;
    const float *base_basic = FUSED ? a.img_basic : (a.pbasic ? a.pbasic + gbase : nullptr);

    if (FUSED && step2) {
        // exec_flat_areas (lib/vnlb/utils/flat_areas.py:16-34) on the gathered noisy patches
        float var_sum = 0.f;
        const int Zn = n * p;
        for (int ch = 0; ch < C; ++ch) {
            float s = 0.f, s2 = 0.f;
            if (tid < p) {
                const float *q = base_noisy + col_off(tid, ch);
                int nn = 0;
                for (; nn + 7 < n; nn += 8) {          // 8 loads in flight, accumulated in the original order
                    float x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = q[pb[nn + u]];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { s += x[u]; s2 = fmaf(x[u], x[u], s2); }
                }
                for (; nn < n; ++nn) { const float x = q[pb[nn]]; s += x; s2 = fmaf(x, x, s2); }
            }
            const float ts = block_sum(s, red, phase), ts2 = block_sum(s2, red, phase);
            var_sum += (ts2 - ts * ts / (float)Zn) / (float)(Zn - 1);
        }
        is_flat = (var_sum / (float)C) < a.flat_thresh;
    }

    const int ch_begin = FUSED ? 0 : (int)(blockIdx.x - g * C);
    const int ch_end = FUSED ? C : ch_begin + 1;
    for (int ch = ch_begin; ch < ch_end; ++ch) {
        int co[4];                                 // offsets of this lane's columns j = lane + 32 q
#pragma unroll
        for (int q = 0; q < 4; ++q) co[q] = col_off(min(lane + 32 * q, p - 1), ch);
        const float *src = P.cov_from_basic ? base_basic : base_noisy;
        __syncthreads();

        if constexpr (SPLIT) {
            // (d, e, tau, mean | packed reflectors) of problem (g, ch), written by cov_tridiag_kernel
            const float4 *wsp = reinterpret_cast<const float4 *>(a.ws + (size_t)(g * C + ch) * a.ws_stride);
            const int pitch4 = a.ws_pitch >> 2;         // d, e, tau: `ws_pitch` floats each; then mean[LD]; then the reflectors
            for (int idx = tid; idx < 3 * pitch4; idx += TT) {
                const float4 f = wsp[idx];
                const int which = idx / pitch4, off = 4 * (idx - which * pitch4);
                float *dst = which == 0 ? d : (which == 1 ? e : taus);
                *reinterpret_cast<float4 *>(dst + off) = f;
            }
            for (int idx = tid; idx < (LD >> 2); idx += TT) reinterpret_cast<float4 *>(mean)[idx] = wsp[3 * pitch4 + idx];
            const int nref4 = (((qd - 1) * qd / 2 + 3) & ~3) >> 2;
            for (int idx = tid; idx < nref4; idx += TT) reinterpret_cast<float4 *>(R)[idx] = wsp[3 * pitch4 + (LD >> 2) + idx];
            __syncthreads();
        } else {
        // -------------------------------------------------------------- 0. centre + covariance
        for (int j = tid; j < LD; j += TT) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (j < p) {
                const float *q = src + col_off(j, ch);
                int nn = 0;
                for (; nn + 11 < n; nn += 12) {         // 12 loads in flight, same 4 interleaved partial sums
                    float x[12];
#pragma unroll
                    for (int u = 0; u < 12; ++u) x[u] = q[row_off(nn + u)];
#pragma unroll
                    for (int u = 0; u < 12; u += 4) { s0 += x[u]; s1 += x[u + 1]; s2 += x[u + 2]; s3 += x[u + 3]; }
                }
                for (; nn + 3 < n; nn += 4) {
                    s0 += q[row_off(nn)]; s1 += q[row_off(nn + 1)];
                    s2 += q[row_off(nn + 2)]; s3 += q[row_off(nn + 3)];
                }
                for (; nn < n; ++nn) s0 += q[row_off(nn)];
            }
            mean[j] = ((s0 + s1) + (s2 + s3)) * inv_n;
        }
        for (int j = tid; j < LDq; j += TT) { v[j] = 0.f; w[j] = 0.f; }
        // lower-triangular 4x4 tiles of the qd x qd matrix (covariance or Gram) owned by this thread
        const int ntile = LDq >> 2, ntri = ntile * (ntile + 1) / 2;
        int ti[NT], tj[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int idx = tid + t * TT;
            int aa = -1, bb = 0;
            if (idx < ntri) {
                aa = (int)((sqrtf(8.f * idx + 1.f) - 1.f) * 0.5f);
                while (aa * (aa + 1) / 2 > idx) --aa;
                while ((aa + 1) * (aa + 2) / 2 <= idx) ++aa;
                bb = idx - aa * (aa + 1) / 2;
            }
            ti[t] = aa;
            tj[t] = bb;
        }
        float acc[NT][16];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[t][q] = 0.f;
        __syncthreads();
        if constexpr (!GRAM) {
            // C = Y^T Y: patches staged 32 at a time as rows Y[nn][0..LD)
            for (int c0 = 0; c0 < n; c0 += CH) {
                const int rows = min(CH, n - c0);
                for (int nn = warp; nn < rows; nn += TT / 32) {
                    const float *q = src + row_off(c0 + nn);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int j = lane + 32 * qq;
                        if (j < LD) R[nn * LD + j] = (j < p) ? q[co[qq]] - mean[j] : 0.f;
                    }
                }
                __syncthreads();
                for (int nn = 0; nn < rows; ++nn) {
                    const float *row = R + nn * LD;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        if (ti[t] >= 0) {
                            const float4 xi = *reinterpret_cast<const float4 *>(row + 4 * ti[t]);
                            const float4 xj = *reinterpret_cast<const float4 *>(row + 4 * tj[t]);
                            const float a4[4] = {xi.x, xi.y, xi.z, xi.w};
                            const float2 blo = make_float2(xj.x, xj.y), bhi = make_float2(xj.z, xj.w);
#pragma unroll
                            for (int aa = 0; aa < 4; ++aa) {   // packed FP32 FMAs (FFMA2): bit-identical to 4 scalar FMAs
                                const float2 ad = make_float2(a4[aa], a4[aa]);
                                const float2 r0 = __ffma2_rn(ad, blo, make_float2(acc[t][aa * 4 + 0], acc[t][aa * 4 + 1]));
                                const float2 r1 = __ffma2_rn(ad, bhi, make_float2(acc[t][aa * 4 + 2], acc[t][aa * 4 + 3]));
                                acc[t][aa * 4 + 0] = r0.x; acc[t][aa * 4 + 1] = r0.y;
                                acc[t][aa * 4 + 2] = r1.x; acc[t][aa * 4 + 3] = r1.y;
                            }
                        }
                    }
                }
                __syncthreads();
            }
        } else {
            // G = Y Y^T: 32 patch-element columns at a time, staged transposed as Yt[col][0..LDq)
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int ncol = min(32, p - 32 * qq);
                if (ncol <= 0) break;
                const int j = lane + 32 * qq;
                {   // all of this lane's loads of the chunk in flight (LDq <= 60: at most 15 per lane)
                    float vals[16];
                    const float mj = mean[min(j, p - 1)];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int nn = warp + u * (TT / 32);
                        vals[u] = (j < p && nn < n) ? src[row_off(nn) + co[qq]] - mj : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int nn = warp + u * (TT / 32);
                        if (nn < LDq) R[lane * LDq + nn] = vals[u];
                    }
                }
                __syncthreads();
                for (int l = 0; l < ncol; ++l) {
                    const float *row = R + l * LDq;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        if (ti[t] >= 0) {
                            const float4 xi = *reinterpret_cast<const float4 *>(row + 4 * ti[t]);
                            const float4 xj = *reinterpret_cast<const float4 *>(row + 4 * tj[t]);
                            const float a4[4] = {xi.x, xi.y, xi.z, xi.w};
                            const float2 blo = make_float2(xj.x, xj.y), bhi = make_float2(xj.z, xj.w);
#pragma unroll
                            for (int aa = 0; aa < 4; ++aa) {   // packed FP32 FMAs (FFMA2): bit-identical to 4 scalar FMAs
                                const float2 ad = make_float2(a4[aa], a4[aa]);
                                const float2 r0 = __ffma2_rn(ad, blo, make_float2(acc[t][aa * 4 + 0], acc[t][aa * 4 + 1]));
                                const float2 r1 = __ffma2_rn(ad, bhi, make_float2(acc[t][aa * 4 + 2], acc[t][aa * 4 + 3]));
                                acc[t][aa * 4 + 0] = r0.x; acc[t][aa * 4 + 1] = r0.y;
                                acc[t][aa * 4 + 2] = r1.x; acc[t][aa * 4 + 3] = r1.y;
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        float *A = R;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (ti[t] >= 0)
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        const float cval = acc[t][aa * 4 + bb] * inv_n;
                        A[(4 * ti[t] + aa) * LDq + 4 * tj[t] + bb] = cval;
                        A[(4 * tj[t] + bb) * LDq + 4 * ti[t] + aa] = cval;
                    }
        __syncthreads();
        {   // rank_var = mean over channels of trace(C) (bayes_est.py:39-40)
            const float tr = block_sum(tid < qd ? A[tid * LDq + tid] : 0.f, red, phase);
            if (a.rank_var && tid == 0) atomicAdd(&a.rank_var[g], tr / (float)C);
        }
        if (a.dbg_mat) {                             // parity hook
            float *o = a.dbg_mat + (size_t)(g * C + ch) * qd * qd;
            for (int idx = threadIdx.x; idx < qd * qd; idx += TT) { const int i = idx / qd; o[idx] = A[i * LDq + (idx - i * qd)]; }
        }

        // -------------------------------------------------------------- 1. tridiagonalisation
        for (int k = 0; k < qd - 2; ++k) {
            const int m = qd - 1 - k;
            const bool active = tid < m;
            const int i = k + 1 + tid;
            const float x = active ? A[k * LDq + i] : 0.f;
            const float dk = A[k * LDq + k];
            const float alpha = A[k * LDq + k + 1];
            const float ssq = block_sum((active && tid > 0) ? x * x : 0.f, red, phase);
            if (ssq == 0.f) {  // nothing to annihilate
                if (tid == 0) { d[k] = dk; e[k] = alpha; taus[k] = 0.f; v[k] = 0.f; w[k] = 0.f; }
                continue;
            }
            const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, ssq)), alpha);
            const float tau = (beta - alpha) / beta;
            const float scale = 1.f / (alpha - beta);
            const float vi = (tid == 0) ? 1.f : x * scale;
            // reflector k goes, packed, over the dead rows 0..k (every thread has read row k: block_sum synced)
            if (active) { v[i] = vi; R[k * (qd - 1) - (k * (k - 1)) / 2 + tid] = vi; }
            if (tid == 0) { v[k] = 0.f; w[k] = 0.f; d[k] = dk; e[k] = beta; taus[k] = tau; }
            __syncthreads();
            const int jb = (k + 1) & ~3;
            float yi = 0.f;
            if (active) {
                const float *ar = A + i * LDq;
                float2 y01 = make_float2(0.f, 0.f), y23 = make_float2(0.f, 0.f);   // packed FP32 FMAs (FFMA2, sm_100)
                int j = jb;
                for (; j + 4 < LDq; j += 8) {
                    const float4 a4 = *reinterpret_cast<const float4 *>(ar + j);
                    const float4 b4 = *reinterpret_cast<const float4 *>(ar + j + 4);
                    const float4 v4 = *reinterpret_cast<const float4 *>(v + j);
                    const float4 u4 = *reinterpret_cast<const float4 *>(v + j + 4);
                    y01 = __ffma2_rn(make_float2(a4.x, a4.y), make_float2(v4.x, v4.y), y01);
                    y23 = __ffma2_rn(make_float2(a4.z, a4.w), make_float2(v4.z, v4.w), y23);
                    y01 = __ffma2_rn(make_float2(b4.x, b4.y), make_float2(u4.x, u4.y), y01);
                    y23 = __ffma2_rn(make_float2(b4.z, b4.w), make_float2(u4.z, u4.w), y23);
                }
                if (j < LDq) {
                    const float4 a4 = *reinterpret_cast<const float4 *>(ar + j);
                    const float4 v4 = *reinterpret_cast<const float4 *>(v + j);
                    y01 = __ffma2_rn(make_float2(a4.x, a4.y), make_float2(v4.x, v4.y), y01);
                    y23 = __ffma2_rn(make_float2(a4.z, a4.w), make_float2(v4.z, v4.w), y23);
                }
                yi = (y01.x + y01.y) + (y23.x + y23.y);
            }
            float wi = tau * yi;
            const float s = block_sum(active ? wi * vi : 0.f, red, phase);
            wi = fmaf(-0.5f * tau * s, vi, wi);
            if (active) w[i] = wi;
            __syncthreads();
            if (active) {
                float *ar = A + i * LDq;
                const float2 nvi = make_float2(-vi, -vi), nwi = make_float2(-wi, -wi);
                int j = jb;
                for (; j + 4 < LDq; j += 8) {
                    float4 a4 = *reinterpret_cast<float4 *>(ar + j);
                    float4 b4 = *reinterpret_cast<float4 *>(ar + j + 4);
                    const float4 v4 = *reinterpret_cast<const float4 *>(v + j), w4 = *reinterpret_cast<const float4 *>(w + j);
                    const float4 u4 = *reinterpret_cast<const float4 *>(v + j + 4), z4 = *reinterpret_cast<const float4 *>(w + j + 4);
                    { const float2 lo = __ffma2_rn(nvi, make_float2(w4.x, w4.y), __ffma2_rn(nwi, make_float2(v4.x, v4.y), make_float2(a4.x, a4.y)));
                      const float2 hi = __ffma2_rn(nvi, make_float2(w4.z, w4.w), __ffma2_rn(nwi, make_float2(v4.z, v4.w), make_float2(a4.z, a4.w)));
                      a4 = make_float4(lo.x, lo.y, hi.x, hi.y); }
                    { const float2 lo = __ffma2_rn(nvi, make_float2(z4.x, z4.y), __ffma2_rn(nwi, make_float2(u4.x, u4.y), make_float2(b4.x, b4.y)));
                      const float2 hi = __ffma2_rn(nvi, make_float2(z4.z, z4.w), __ffma2_rn(nwi, make_float2(u4.z, u4.w), make_float2(b4.z, b4.w)));
                      b4 = make_float4(lo.x, lo.y, hi.x, hi.y); }
                    *reinterpret_cast<float4 *>(ar + j) = a4;
                    *reinterpret_cast<float4 *>(ar + j + 4) = b4;
                }
                if (j < LDq) {
                    float4 a4 = *reinterpret_cast<float4 *>(ar + j);
                    const float4 v4 = *reinterpret_cast<const float4 *>(v + j), w4 = *reinterpret_cast<const float4 *>(w + j);
                    { const float2 lo = __ffma2_rn(nvi, make_float2(w4.x, w4.y), __ffma2_rn(nwi, make_float2(v4.x, v4.y), make_float2(a4.x, a4.y)));
                      const float2 hi = __ffma2_rn(nvi, make_float2(w4.z, w4.w), __ffma2_rn(nwi, make_float2(v4.z, v4.w), make_float2(a4.z, a4.w)));
                      a4 = make_float4(lo.x, lo.y, hi.x, hi.y); }
                    *reinterpret_cast<float4 *>(ar + j) = a4;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            d[qd - 2] = A[(qd - 2) * LDq + qd - 2];
            e[qd - 2] = A[(qd - 2) * LDq + qd - 1];
            taus[qd - 2] = 0.f;
            d[qd - 1] = A[(qd - 1) * LDq + qd - 1];
            e[qd - 1] = 0.f;
        }
        __syncthreads();
        }  // !SPLIT

        // -------------------------------------------------------------- 2. eigenvalues above the threshold
        // e2[i] = e[i-1]^2 is the coupling that enters pivot i of the Sturm sequence
        float gl = -3.4e38f, em = 0.f;
        for (int j = tid; j < qd; j += TT) {
            const float el = j > 0 ? e[j - 1] : 0.f, er = j < qd - 1 ? e[j] : 0.f;
            e2[j] = el * el;
            gl = fmaxf(gl, d[j] + fabsf(el) + fabsf(er));
            em = fmaxf(em, el * el);
        }
        const float gmax0 = block_max(gl, red, phase);
        const float emax2 = block_max(em, red, phase);
        const float pivmin = fmaxf(1e-30f, emax2 * 1e-12f);
        const float tau_eff = fmaf(P.thresh, P.sigma2, P.sigmab2);   // eigenvalue > tau_eff <=> coefficient != 0
        const float gmax = gmax0 + 1e-6f * fabsf(gmax0) + pivmin;

        // Sturm count #eigenvalues < x of the tridiagonal matrix, 4 pivots per pair of LDS.128
        auto sturm_count = [&](float x) -> int {
            float q = 1.f;
            int cnt = 0;
            int j = 0;
            for (; j + 3 < qd; j += 4) {
                const float4 d4 = *reinterpret_cast<const float4 *>(d + j), e4 = *reinterpret_cast<const float4 *>(e2 + j);
                float t = (d4.x - x) - __fdividef(e4.x, q);
                if (fabsf(t) < pivmin) t = -pivmin;
                cnt += t < 0.f;
                float u = (d4.y - x) - __fdividef(e4.y, t);
                if (fabsf(u) < pivmin) u = -pivmin;
                cnt += u < 0.f;
                t = (d4.z - x) - __fdividef(e4.z, u);
                if (fabsf(t) < pivmin) t = -pivmin;
                cnt += t < 0.f;
                q = (d4.w - x) - __fdividef(e4.w, t);
                if (fabsf(q) < pivmin) q = -pivmin;
                cnt += q < 0.f;
            }
            for (; j < qd; ++j) {
                float t = (d[j] - x) - __fdividef(e2[j], q);
                if (fabsf(t) < pivmin) t = -pivmin;
                q = t;
                cnt += t < 0.f;
            }
            return cnt;
        };
        // Round 0: 128 DISTINCT section points over [tau_eff, gmax] (thread 0 sits exactly on the threshold), one Sturm chain
        // each: the count at the threshold gives the number m of eigenpairs to compute, the other counts bracket every one
        // of them to 1/128 of the interval at once.
        int m;
        int *scnt = (int *)(sm + L.oCnt);
        {
            const float x0 = fmaf(gmax - tau_eff, (float)tid / (float)TT, tau_eff);
            scnt[tid] = sturm_count(x0);
            __syncthreads();
            m = min(qd - scnt[0], P.rank);
            if (gmax <= tau_eff) m = 0;
        }
        float *Z = R + L.oZ;
        if (m > 0) {
            if (tid < m) {
                const int idx = qd - 1 - tid;               // ascending index of the tid-th largest eigenvalue
                int tlo = 0, thi = TT;
                for (int t = 0; t < TT; ++t) {
                    const int c = scnt[t];
                    if (c <= idx) tlo = max(tlo, t); else thi = min(thi, t);
                }
                blo[tid] = fmaf(gmax - tau_eff, (float)tlo / (float)TT, tau_eff);
                bhi[tid] = thi < TT ? fmaf(gmax - tau_eff, (float)thi / (float)TT, tau_eff) : gmax;
                if (blo[tid] > bhi[tid]) { const float mid = 0.5f * (blo[tid] + bhi[tid]); blo[tid] = mid; bhi[tid] = mid; }
            }
            __syncthreads();
            const int G = (TT * ILP) / m;                 // section points per eigenvalue and round
            int rounds = 0;
            { float res = (float)TT; while (res < 1.7e7f && rounds < 14) { res *= (float)(G + 1); ++rounds; } }  // 2^24 sections in all
            for (int r = 0; r < rounds; ++r) {
                for (int j = tid; j < m; j += TT) { nlo[j] = __float_as_int(blo[j]); nhi[j] = __float_as_int(bhi[j]); }
                __syncthreads();
                static_assert(ILP == 1, "one Sturm chain per thread");
                const int pt = tid;
                const int j = pt / G, qi = pt - j * G;
                const int ej = j < m ? j : -1;
                const float lo = blo[min(j, m - 1)], hi = bhi[min(j, m - 1)];
                const float x = fmaf(hi - lo, (float)(qi + 1) / (float)(G + 1), lo);
                const int cnt = sturm_count(x);
                if (ej >= 0) {
                    const int idx = qd - 1 - ej;          // ascending index of the ej-th largest eigenvalue
                    if (cnt <= idx) atomicMax(&nlo[ej], __float_as_int(x));   // x is a lower bound
                    else atomicMin(&nhi[ej], __float_as_int(x));              // x is an upper bound
                }
                __syncthreads();
                for (int jj = tid; jj < m; jj += TT) {
                    float l2 = __int_as_float(nlo[jj]), h2 = __int_as_float(nhi[jj]);
                    if (l2 > h2) { const float mid = 0.5f * (l2 + h2); l2 = mid; h2 = mid; }
                    blo[jj] = l2; bhi[jj] = h2;
                }
                __syncthreads();
            }
            for (int j = tid; j < m; j += TT) {
                const float l = 0.5f * (blo[j] + bhi[j]);
                lam[j] = l;
                const float ls = l - fminf(l, P.sigmab2);                                   // bayes_est.py:129-138
                coef[j] = (ls > P.thresh * P.sigma2) ? 1.f / (1.f + P.sigma2 / ls) : 0.f;   // bayes_est.py:140-144
            }
            __syncthreads();

            // ---------------------------------------------------------- 3. eigenvectors of T (twisted factorisation)
            // g_work: the warps that hold no eigenpair form the Gram entries of the reflector blocks (phase 4): g_lj = v_l^T v_j
            auto g_work = [&](int first) {
                const int nidle = TT - first, nblk = (qd - 2) >> 2;
                float *Tf = sm + L.oT;
                for (int idx = tid - first; idx < 6 * nblk; idx += nidle) {
                    const int blk = idx / 6, pr = idx - 6 * blk;
                    const int jj = pr < 1 ? 1 : (pr < 3 ? 2 : 3), ll = pr - (jj == 1 ? 0 : (jj == 2 ? 1 : 3));
                    const int kl = 4 * blk + ll, kj = 4 * blk + jj;
                    const float *vl = R + (kl * (qd - 1) - (kl * (kl - 1)) / 2) - (kl + 1);
                    const float *vj = R + (kj * (qd - 1) - (kj * (kj - 1)) / 2) - (kj + 1);
                    float g0 = 0.f, g1 = 0.f;
                    int i = kj + 1;
                    for (; i + 1 < qd; i += 2) { g0 = fmaf(vl[i], vj[i], g0); g1 = fmaf(vl[i + 1], vj[i + 1], g1); }
                    if (i < qd) g0 = fmaf(vl[i], vj[i], g0);
                    Tf[10 * blk + pr] = g0 + g1;
                }
            };
            if (2 * m <= MR) {
                // Two threads per eigenpair: thread r runs the forward pivots D+ into row r of Z while thread m + r runs
                // the backward pivots D- into the spare row m + r -- the two 98-long division chains side by side and no
                // second backward pass; then both find the twist index, and each solves its side of the eigenvector.
                const bool fwd = tid < m, bwd = !fwd && tid < 2 * m;
                const int ev = fwd ? tid : (bwd ? tid - m : 0);
                float *z = Z + ev * ZSq, *zb = Z + (m + ev) * ZSq;
                const float l = lam[ev];
                const float pm = fmaxf(pivmin, 1e-10f * fabsf(l));
                if (fwd) {
                    float dp = d[0] - l;
                    if (fabsf(dp) < pm) dp = -pm;
                    z[0] = dp;
                    for (int i = 0; i < qd - 1; ++i) {
                        dp = (d[i + 1] - l) - __fdividef(e2[i + 1], dp);
                        if (fabsf(dp) < pm) dp = -pm;
                        z[i + 1] = dp;
                    }
                } else if (bwd) {
                    float dm = d[qd - 1] - l;
                    if (fabsf(dm) < pm) dm = -pm;
                    zb[qd - 1] = dm;
                    for (int i = qd - 2; i >= 0; --i) {
                        dm = (d[i] - l) - __fdividef(e2[i + 1], dm);
                        if (fabsf(dm) < pm) dm = -pm;
                        zb[i] = dm;
                    }
                } else if (32 * warp >= 2 * m) {
                    g_work(32 * ((2 * m + 31) >> 5));
                }
                __syncthreads();
                float nrm = 0.f;
                int r = qd - 1;
                // twist index r = argmin |gamma_i|, gamma_i = D+_i + D-_i - (d_i - l): scanned from the top, ties to the higher index;
                // the pair splits the scan (upper half / lower half) and combines through shared memory
                float *sbest = sm + L.oCnt;                // round-0 Sturm counts are dead: [0, MR) upper minima, [MR, 2 MR) lower minima
                const int mid = qd >> 1;
                if (fwd) {
                    float best = fabsf(z[qd - 1]);
#pragma unroll 4
                    for (int i = qd - 2; i >= mid; --i) {
                        const float gam = fabsf(z[i] + zb[i] - (d[i] - l));
                        if (gam < best) { best = gam; r = i; }
                    }
                    sbest[ev] = best; nlo[ev] = r;
                } else if (bwd) {
                    float best = 3.4e38f;
                    r = 0;
#pragma unroll 4
                    for (int i = mid - 1; i >= 0; --i) {
                        const float gam = fabsf(z[i] + zb[i] - (d[i] - l));
                        if (gam < best) { best = gam; r = i; }
                    }
                    sbest[MR + ev] = best; nhi[ev] = r;
                }
                __syncthreads();                           // every D+ / D- has been read before the solves overwrite Z
                if (fwd || bwd) r = (sbest[MR + ev] < sbest[ev]) ? nhi[ev] : nlo[ev];
                if (fwd) {                                 // z_r = 1; upwards with D+
                    float zi = 1.f;
                    nrm = 1.f;
#pragma unroll 4
                    for (int i = r - 1; i >= 0; --i) {   // (the ratio does not depend on zi: loads and divisions of 4 iterations overlap)
                        zi = -__fdividef(e[i], z[i]) * zi;
                        z[i] = zi;
                        nrm = fmaf(zi, zi, nrm);
                    }
                    z[r] = 1.f;
                    blo[ev] = nrm;
                } else if (bwd) {                          // downwards with D-
                    float zi = 1.f;
#pragma unroll 4
                    for (int i = r; i < qd - 1; ++i) {
                        zi = -__fdividef(e[i], zb[i + 1]) * zi;
                        z[i + 1] = zi;
                        nrm = fmaf(zi, zi, nrm);
                    }
                    bhi[ev] = nrm;
                }
                __syncthreads();
                if (fwd || bwd) {
                    nrm = blo[ev] + bhi[ev];
                    const float sc = rsqrtf(nrm);
                    const int i0 = fwd ? 0 : r + 1, i1 = fwd ? r + 1 : qd;
                    if (nrm < 3.0e38f && sc > 0.f) {
#pragma unroll 4
                        for (int i = i0; i < i1; ++i) z[i] *= sc;
                    } else {                             // overflow guard (never seen): fall back to the twist basis vector
                        for (int i = i0; i < i1; ++i) z[i] = (i == r) ? 1.f : 0.f;
                    }
                }
            } else
            if (tid < m) {
                float *z = Z + tid * ZSq;
                const float l = lam[tid];
                const float pm = fmaxf(pivmin, 1e-10f * fabsf(l));
                float dp = d[0] - l;                       // forward pivots D+ (kept in z)
                if (fabsf(dp) < pm) dp = -pm;
                z[0] = dp;
                for (int i = 0; i < qd - 1; ++i) {
                    dp = (d[i + 1] - l) - __fdividef(e2[i + 1], dp);   // 2-ulp division: the chain's latency is the phase's cost
                    if (fabsf(dp) < pm) dp = -pm;
                    z[i + 1] = dp;
                }
                float dm = d[qd - 1] - l;                   // backward pivots D-: twist index r = argmin |gamma_i|
                if (fabsf(dm) < pm) dm = -pm;
                float best = fabsf(z[qd - 1]);
                int r = qd - 1;
                for (int i = qd - 2; i >= 0; --i) {
                    dm = (d[i] - l) - __fdividef(e2[i + 1], dm);
                    if (fabsf(dm) < pm) dm = -pm;
                    const float gam = fabsf(z[i] + dm - (d[i] - l));
                    if (gam < best) { best = gam; r = i; }
                }
                dm = d[qd - 1] - l;                         // second pass stores D-_i for i > r
                if (fabsf(dm) < pm) dm = -pm;
                if (r < qd - 1) z[qd - 1] = dm;
                for (int i = qd - 2; i > r; --i) {
                    dm = (d[i] - l) - __fdividef(e2[i + 1], dm);
                    if (fabsf(dm) < pm) dm = -pm;
                    z[i] = dm;
                }
                float zi = 1.f, nrm = 1.f;                 // z_r = 1; upwards with D+, downwards with D-
#pragma unroll 4
                for (int i = r - 1; i >= 0; --i) {
                    zi = -__fdividef(e[i], z[i]) * zi;
                    z[i] = zi;
                    nrm = fmaf(zi, zi, nrm);
                }
                zi = 1.f;
#pragma unroll 4
                for (int i = r; i < qd - 1; ++i) {
                    zi = -__fdividef(e[i], z[i + 1]) * zi;
                    z[i + 1] = zi;
                    nrm = fmaf(zi, zi, nrm);
                }
                z[r] = 1.f;
                const float sc = rsqrtf(nrm);
                if (nrm < 3.0e38f && sc > 0.f) {
#pragma unroll 4
                    for (int i = 0; i < qd; ++i) z[i] *= sc;
                } else {                                 // overflow guard (never seen): fall back to the twist basis vector
                    for (int i = 0; i < qd; ++i) z[i] = (i == r) ? 1.f : 0.f;
                }
            } else if (32 * warp >= m) {
                g_work(32 * ((m + 31) >> 5));
            }
            __syncthreads();

            // ---------------------------------------------------------- 4. back-transformation  z <- H_0 ... H_{qd-3} z
            // Blocks of 4 reflectors in compact-WY form, H_a H_{a+1} H_{a+2} H_{a+3} = I - V T V^T (T upper triangular,
            // LAPACK dlarft forward/columnwise): 4 dot products reduced together, one small triangular product, one
            // fused update -- a quarter of the dependent shuffle chains of reflector-by-reflector application.
            // The Gram entries g_lj = v_l^T v_j were formed by the idle warps during the twisted factorisation.
            {
                const int nrefl = qd - 2, nblk = nrefl >> 2, rem = nrefl - 4 * nblk;   // blocks cover k = 0 .. 4 nblk - 1
                float *Tf = sm + L.oT;
                if (tid < nblk) {
                    float *g = Tf + 10 * tid;
                    const float g01 = g[0], g02 = g[1], g12 = g[2], g03 = g[3], g13 = g[4], g23 = g[5];
                    const float t0 = taus[4 * tid], t1 = taus[4 * tid + 1], t2 = taus[4 * tid + 2], t3 = taus[4 * tid + 3];
                    const float T01 = -t1 * (t0 * g01);
                    const float T02 = -t2 * fmaf(T01, g12, t0 * g02), T12 = -t2 * (t1 * g12);
                    const float T03 = -t3 * fmaf(T02, g23, fmaf(T01, g13, t0 * g03)), T13 = -t3 * fmaf(T12, g23, t1 * g13), T23 = -t3 * (t2 * g23);
                    g[0] = t0; g[1] = T01; g[2] = T02; g[3] = T03; g[4] = t1; g[5] = T12; g[6] = T13; g[7] = t2; g[8] = T23; g[9] = t3;
                }
                __syncthreads();
                int S = 2;                                   // lanes per eigenvector (power of two, S*m <= 128)
                while (S * 2 * m <= TT && S < 32) S *= 2;
                const int vec = tid / S, part = tid - vec * S;
                const bool on = vec < m;
                float *z = Z + min(vec, m - 1) * ZSq;
                for (int k = qd - 3; k >= 4 * nblk; --k) {   // the (at most 3) shortest reflectors, one by one
                    const float tau = taus[k];
                    if (tau == 0.f) continue;
                    const float *vr = R + (k * (qd - 1) - (k * (k - 1)) / 2) - (k + 1);   // vr[i], i = k+1..qd-1
                    float s = 0.f;
                    for (int i = k + 1 + part; i < qd; i += S) s = fmaf(vr[i], z[i], s);
                    for (int dlt = S >> 1; dlt > 0; dlt >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dlt);
                    s *= tau;
                    if (on)
                        for (int ii = k + 1 + part; ii < qd; ii += S) z[ii] = fmaf(-s, vr[ii], z[ii]);
                    __syncwarp();
                }
                (void)rem;
                for (int bk = nblk - 1; bk >= 0; --bk) {
                    const int a0 = 4 * bk;
                    const float *v0 = R + (a0 * (qd - 1) - (a0 * (a0 - 1)) / 2) - (a0 + 1);       // v_{a0}[i], i >= a0+1
                    const float *v1 = v0 + (qd - 1 - a0) - 1;                                          // v_{a0+1}[i], i >= a0+2
                    const float *v2 = v1 + (qd - 2 - a0) - 1;
                    const float *v3 = v2 + (qd - 3 - a0) - 1;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    for (int i = a0 + 1 + part; i < qd; i += S) {
                        const float zi = z[i];
                        s0 = fmaf(v0[i], zi, s0);
                        if (i > a0 + 1) s1 = fmaf(v1[i], zi, s1);
                        if (i > a0 + 2) s2 = fmaf(v2[i], zi, s2);
                        if (i > a0 + 3) s3 = fmaf(v3[i], zi, s3);
                    }
                    for (int dlt = S >> 1; dlt > 0; dlt >>= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, dlt); s1 += __shfl_xor_sync(0xffffffffu, s1, dlt);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, dlt); s3 += __shfl_xor_sync(0xffffffffu, s3, dlt);
                    }
                    const float *T = Tf + 10 * bk;
                    const float u3 = T[9] * s3;
                    const float u2 = fmaf(T[8], s3, T[7] * s2);
                    const float u1 = fmaf(T[6], s3, fmaf(T[5], s2, T[4] * s1));
                    const float u0 = fmaf(T[3], s3, fmaf(T[2], s2, fmaf(T[1], s1, T[0] * s0)));
                    if (on)
                        for (int i = a0 + 1 + part; i < qd; i += S) {
                            float zi = fmaf(-u0, v0[i], z[i]);
                            if (i > a0 + 1) zi = fmaf(-u1, v1[i], zi);
                            if (i > a0 + 2) zi = fmaf(-u2, v2[i], zi);
                            if (i > a0 + 3) zi = fmaf(-u3, v3[i], zi);
                            z[i] = zi;
                        }
                    __syncwarp();
                }
            }
            __syncthreads();
            if constexpr (!GRAM) {
                // eigenvectors re-laid as Vt[j][r] (r contiguous, pitch MR): 4 eigenpairs per broadcast LDS.128
                float *Vt = R + L.oVt;
                // (the filter reads the first 8 * ceil(m / 8) columns of a row: columns m .. that bound are zeroed, the rest is never read)
                const int mpad = (m + 7) & ~7;
                for (int r = warp; r < mpad; r += TT / 32)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int j = lane + 32 * qq;
                        if (j < p) Vt[j * MR + r] = r < m ? Z[r * ZSq + j] : 0.f;
                    }
            } else {
                // Gram trick: Z[r][nn] -> Ut[nn][r] in place (through registers), then
                // v_r = Y^T u_r / sqrt(n lambda_r) straight into Vt[j][r]
                float *Ut = Z;
                float tz[(MR + TT / 32 - 1) / (TT / 32)][2];
#pragma unroll
                for (int rr = 0; rr < (MR + TT / 32 - 1) / (TT / 32); ++rr) {
                    const int r = warp + rr * (TT / 32);
                    tz[rr][0] = (r < m && lane < n) ? Z[r * ZSq + lane] : 0.f;
                    tz[rr][1] = (r < m && lane + 32 < n) ? Z[r * ZSq + lane + 32] : 0.f;
                }
                __syncthreads();
#pragma unroll
                for (int rr = 0; rr < (MR + TT / 32 - 1) / (TT / 32); ++rr) {
                    const int r = warp + rr * (TT / 32);
                    if (r < MR) {
                        if (lane < n) Ut[lane * MR + r] = tz[rr][0];
                        if (lane + 32 < n) Ut[(lane + 32) * MR + r] = tz[rr][1];
                    }
                }
                __syncthreads();
                float *Vt = R + L.oVt;
                switch ((m + 7) >> 3) {
                    case 1: gram_map<1>(Vt, Ut, lam, mean, src, pb, FUSED, rstride, n, p, m, tid < p ? col_off(tid, ch) : 0, tid); break;
                    case 2: gram_map<2>(Vt, Ut, lam, mean, src, pb, FUSED, rstride, n, p, m, tid < p ? col_off(tid, ch) : 0, tid); break;
                    case 3: gram_map<3>(Vt, Ut, lam, mean, src, pb, FUSED, rstride, n, p, m, tid < p ? col_off(tid, ch) : 0, tid); break;
                    case 4: gram_map<4>(Vt, Ut, lam, mean, src, pb, FUSED, rstride, n, p, m, tid < p ? col_off(tid, ch) : 0, tid); break;
                    default: gram_map<5>(Vt, Ut, lam, mean, src, pb, FUSED, rstride, n, p, m, tid < p ? col_off(tid, ch) : 0, tid); break;
                }
            }
        }
        __syncthreads();

        if (a.dbg_lam && threadIdx.x < MR) {             // parity hook: eigenvalues above the threshold and their coefficients
            const size_t o = (size_t)(g * C + ch) * MR + threadIdx.x;
            const bool live = (int)threadIdx.x < m;
            a.dbg_lam[o] = live ? lam[threadIdx.x] : 0.f;
            if (a.dbg_coef) a.dbg_coef[o] = live ? coef[threadIdx.x] : 0.f;
            if (a.dbg_m && threadIdx.x == 0) a.dbg_m[g * C + ch] = m;
        }
        // -------------------------------------------------------------- 5. Wiener filter of the noisy patches
        float *X = R + L.oX;                             // X[xrows][XS], XS odd => conflict-free rows
        const float *Vt = R + L.oVt;
        if (P.cov_from_basic || FUSED) {
#pragma unroll
            for (int q = 0; q < 4; ++q) co[q] = col_off(min(lane + 32 * q, p - 1), ch);
        }

        // centre of the noisy patches: mean_n(noisy), or cbasic for a flat group in step 2 (bayes_est.py:88-104)
        if (P.cov_from_basic ? !is_flat : false) {
            for (int j = tid; j < p; j += TT) {
                const float *q = base_noisy + col_off(j, ch);
                float s0 = 0.f, s1 = 0.f;
                int nn = 0;
                for (; nn + 11 < n; nn += 12) {         // 12 loads in flight, same 2 interleaved partial sums
                    float x[12];
#pragma unroll
                    for (int u = 0; u < 12; ++u) x[u] = q[row_off(nn + u)];
#pragma unroll
                    for (int u = 0; u < 12; u += 2) { s0 += x[u]; s1 += x[u + 1]; }
                }
                for (; nn + 1 < n; nn += 2) { s0 += q[row_off(nn)]; s1 += q[row_off(nn + 1)]; }
                if (nn < n) s0 += q[row_off(nn)];
                mean[j] = (s0 + s1) * inv_n;
            }
        } else if (!P.cov_from_basic && is_flat) {
            for (int j = tid; j < p; j += TT) {
                const float *q = base_basic + col_off(j, ch);
                float s0 = 0.f;
                for (int nn = 0; nn < n; ++nn) s0 += q[row_off(nn)];
                mean[j] = s0 * inv_n;
            }
        }
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += L.xrows) {
            const int rows = min(L.xrows, n - c0);
            {   // gather with 4 patches (16 independent loads per lane) in flight
                float mj[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) mj[qq] = mean[min(lane + 32 * qq, p - 1)];
                for (int n0 = warp; n0 < rows; n0 += 4 * (TT / 32)) {
                    float vals[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float *q = base_noisy + row_off(c0 + min(n0 + u * (TT / 32), rows - 1));
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) vals[u][qq] = q[co[qq]];       // co[] is clamped to column p-1
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nn = n0 + u * (TT / 32);
                        if (nn < rows) {
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                const int j = lane + 32 * qq;
                                if (j < p) X[nn * XS + j] = vals[u][qq] - mj[qq];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            bool filtered = false;
            if constexpr (MMA) {
                if (m > 0) {
                    switch ((m + 7) >> 3) {
                        case 1: filter_chunk_mma<1>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                        case 2: filter_chunk_mma<2>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                        case 3: filter_chunk_mma<3>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                        case 4: filter_chunk_mma<4>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                        default: filter_chunk_mma<5>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                    }
                    filtered = true;
                }
            }
            if (!filtered)
            switch ((m + 7) >> 3) {
                case 0: filter_chunk<0>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                case 1: filter_chunk<1>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                case 2: filter_chunk<2>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                case 3: filter_chunk<3>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                case 4: filter_chunk<4>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
                default: filter_chunk<5>(X, Vt, coef, mean, rows, p, XS, m, tid); break;
            }
            __syncthreads();
            int wo[4] = {0, 0, 0, 0};                    // fused: offsets of this lane's patch elements in `weights`
            if (FUSED && ch == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = min(lane + 32 * q, p - 1), dt = j / ps2, r = j - dt * ps2, dy = r / ps, dx = r - dy * ps;
                    wo[q] = (int)(dt * HW + (long long)dy * a.W + dx);
                }
            }
            for (int nn = warp; nn < rows; nn += TT / 32) {
                if (FUSED) {   // agg_patches (lib/vnlb/agg/comp_agg.py:106-138): scatter with float atomics
                    float *q = a.deno + pb[c0 + nn];
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int j = lane + 32 * qq;
                        if (j < p) atomicAdd(q + co[qq], X[nn * XS + j]);
                    }
                    if (ch == 0) {
                        float *qw = a.weights + wb[c0 + nn];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const int j = lane + 32 * qq;
                            if (j < p) atomicAdd(qw + wo[qq], 1.f);
                        }
                    }
                } else {
                    float *q = a.pnoisy + gbase + row_off(c0 + nn);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int j = lane + 32 * qq;
                        if (j < p) q[co[qq]] = X[nn * XS + j];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Split path (cov_tridiag_kernel + bayes_kernel<.., SPLIT>) for the production shape of step 1 (7x7x2 patches: q = 98,
// direct covariance); VNLB_BAYES_SPLIT=0 forces the single-kernel shared-memory path.
static int g_tail2 = []() { const char *s = getenv("VNLB_TAIL2"); return (s && s[0] == '0') ? 0 : 1; }();   // VNLB_TAIL2=0: one row per thread in the 64 -> 32 phase (A/B measurements)
static int g_split = -1;   // -1: not read yet; VNLB_BAYES_SPLIT=0 in the environment or vnlb_set_bayes_split(0) disables
static bool use_split(const TriLayout &L) {
    if (g_split < 0) { const char *s = getenv("VNLB_BAYES_SPLIT"); g_split = (s && s[0] == '0') ? 0 : 1; }
    if (g_split == 0) return false;
    if (!L.gram) return L.q == 98 && (size_t)(L.n * 100 + L.n + 16) * sizeof(float) <= 72 * 1024;   // step 1: 7x7x2, direct covariance
    return L.q == 60 && L.n == 60 && L.p == 98;                                                       // step 2: 7x7x2, k = 60, Gram trick
}
int set_bayes_split(int on) {
    const int prev = g_split < 0 ? 1 : g_split;
    g_split = on ? 1 : 0;
    return prev;
}

// Workspace of the split path: the kernels of a call hand (d, e, tau, mean, packed reflectors, trailing matrices) from one
// to the next through `ws`, 12-37 KB per (group, channel) problem.  The CALLER owns it (vnlb_bayes_workspace_bytes);
// a call with more groups than the workspace holds is processed in stream-ordered chunks.
constexpr int SPLIT_CHUNK = 16384;                 // groups per chunk at most (a round of the throughput schedule)
static size_t split_bytes_per_group(const TriLayout &L, int c) {
    const int stride = L.gram ? gram_ws_stride<60>(L.LD) : split_ws_stride<98>();
    return (size_t)c * stride * sizeof(float);
}
size_t bayes_workspace_bytes(int B, const VnlbBayesParams *p) {
    const int pd = p->pt * p->ps * p->ps;
    if (B <= 0 || pd < 3 || pd > TT || p->k < 1 || p->c < 1) return 0;
    const TriLayout L = tri_layout(p->k, pd);
    if (!use_split(L)) return 0;
    return (size_t)(B < SPLIT_CHUNK ? B : SPLIT_CHUNK) * split_bytes_per_group(L, p->c);
}

// Wiener filter on the tensor cores (filter_chunk_mma): VNLB_FILTER_MMA=1 in the environment or vnlb_set_filter_mma(1).
// Parity-green and measured SLOWER than the FFMA2 filter (15.31 vs 14.40 ms and 9.21 vs 8.35 ms per 16384 groups:
// mma.sync TF32 issues at 0.5 per cycle and SM on this part and every operand needs its 3xTF32 split): off by default.
static int g_filter_mma = -1;
static int filter_mma_setting() {
    if (g_filter_mma < 0) { const char *e = getenv("VNLB_FILTER_MMA"); g_filter_mma = (e && e[0] == '1') ? 1 : 0; }
    return g_filter_mma;
}
int set_filter_mma(int on) {
    const int prev = filter_mma_setting();
    g_filter_mma = on ? 1 : 0;
    return prev;
}

template <bool FUSED>
static int launch_bayes_chunk(BayesArgs &a, int B, const char *what, cudaStream_t st);

template <bool FUSED>
static int launch_bayes_any(BayesArgs &a, int B, const char *what, void *ws, size_t ws_bytes, cudaStream_t st) {
    a.filter_mma = filter_mma_setting();
    if (!use_split(a.L)) return launch_bayes_chunk<FUSED>(a, B, what, st);
    const VnlbBayesParams &P = a.P;
    const size_t per_group = split_bytes_per_group(a.L, P.c);
    long long chunk = ws ? (long long)(ws_bytes / per_group) : 0;
    if (chunk > SPLIT_CHUNK) chunk = SPLIT_CHUNK;
    if (chunk < 1 || ((uintptr_t)ws & 15)) {
        set_error("%s: workspace missing, misaligned or smaller than one group (%zu B per group; see vnlb_bayes_workspace_bytes)",
                  what, per_group);
        return VNLB_ERR_WORKSPACE;
    }
    a.ws = (float *)ws;
    a.ws_stride = (int)(per_group / sizeof(float) / P.c);
    const long long rstride = (long long)P.pt * P.c * P.ps * P.ps, gstride = (long long)P.k * rstride;
    for (int g0 = 0; g0 < B; g0 += (int)chunk) {
        BayesArgs c = a;
        if (c.pnoisy) c.pnoisy += g0 * gstride;
        if (c.pbasic) c.pbasic += g0 * gstride;
        if (c.flat) c.flat += g0;
        if (c.inds) c.inds += (long long)g0 * P.k;
        if (c.rank_var) c.rank_var += g0;
        if (c.dbg_mat) c.dbg_mat += (size_t)g0 * P.c * a.L.q * a.L.q;
        if (c.dbg_lam) c.dbg_lam += (size_t)g0 * P.c * MR;
        if (c.dbg_coef) c.dbg_coef += (size_t)g0 * P.c * MR;
        if (c.dbg_m) c.dbg_m += (size_t)g0 * P.c;
        const int rc = launch_bayes_chunk<FUSED>(c, B - g0 < chunk ? B - g0 : (int)chunk, what, st);
        if (rc != VNLB_OK) return rc;
    }
    return VNLB_OK;
}

template <bool FUSED>
static int launch_bayes_chunk(BayesArgs &a, int B, const char *what, cudaStream_t st) {
    const VnlbBayesParams *p = &a.P;
    cudaError_t e;
    const bool split = use_split(a.L);                   // a.ws / a.ws_stride were set by launch_bayes_any
    if (split && a.L.gram) {
        constexpr int QD = 60;
        a.ws_pitch = GRAM_PITCH;
        const size_t smem = (size_t)a.L.total * sizeof(float);
        constexpr int scr = tridiag_scratch_floats<QD, QD, SPLIT_NR3>();
        int ybody = a.L.p * QD > QD * QD ? a.L.p * QD : QD * QD;
        if (ybody < scr) ybody = scr;
        const size_t smem1 = (size_t)(ybody + 8 + ((a.L.n + 3) & ~3) + 4) * sizeof(float);
        auto k1 = g_tail2 ? gram_tridiag_kernel<FUSED, QD, true> : gram_tridiag_kernel<FUSED, QD, false>;
        auto k1c = tridiag_tail_kernel<QD, SPLIT_NR3, 2, 32, 16, GRAM_PITCH, 3 * GRAM_PITCH + 100, gram_trail_off<QD>(100), 0>;
        const size_t smem1c = (size_t)tridiag_scratch_floats<QD, SPLIT_NR3, 2>() * sizeof(float);
        auto k2 = a.filter_mma ? bayes_kernel<FUSED, true, true, true> : bayes_kernel<FUSED, true, true, false>;
        e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
        k1<<<B * p->c, 64, smem1, st>>>(a);
        if (g_tail2) {      // elimination 60 -> 32 with two rows per thread, one warp per problem
            auto k1b2 = tridiag_tail2_kernel<QD, QD, 10, GRAM_PITCH, 3 * GRAM_PITCH + 100, gram_full_off<QD>(100), gram_trail_off<QD>(100)>;
            k1b2<<<B * p->c, 32, (size_t)tridiag2_scratch_floats<QD, QD, SPLIT_NR3>() * sizeof(float), st>>>(a);
        }
        k1c<<<B * p->c, 32, smem1c, st>>>(a);
        k2<<<FUSED ? B : B * p->c, TT, smem, st>>>(a);
        return check_launch(what, g_tail2 ? 4 : 3);
    }
    if (split) {
        a.L = tri_layout(a.L.n, a.L.p, true);            // no covariance matrix in the eigen/filter kernel
        a.ws_pitch = 100;
        const size_t smem = (size_t)a.L.total * sizeof(float);
        constexpr int QD = 98;
        constexpr int scr1 = tridiag_scratch_floats<QD, QD, SPLIT_NR2>();
        const int yrows = a.L.n > 100 ? a.L.n : 100;
        const int ybody = yrows * 100 > scr1 ? yrows * 100 : scr1;
        // Whole rows in registers (166 registers, 3 CTAs per SM).  The NRC < NCH variant of the kernel (64 columns in
        // registers + the dying columns in shared memory, 128 registers, 4 CTAs per SM) measured 1.5 % SLOWER: this phase is
        // bound by FFMA2 / LDS throughput, not by occupancy (profiles/r1_summary.md).
        constexpr int NCHQ = (QD + 3) / 4;
        size_t smem1 = (size_t)(ybody + 8 + ((a.L.n + 3) & ~3) + 4) * sizeof(float);
        // Elimination 98 -> 64 with the columns split over the warps (cov_tridiag4_kernel: 12 x less shared-memory traffic per
        // FMA in the sweep, 128 registers, 4 CTAs per SM): 14.44 vs 14.73 ms per 16384 groups (profiles/r2_summary.md).
        // VNLB_COV4=0 selects tridiag_regs (one row per thread) for A/B measurements.
        static const int cols4 = []() { const char *e4 = getenv("VNLB_COV4"); return (e4 && e4[0] == '0') ? 0 : 1; }();
        const size_t smem1_4 = (size_t)(yrows * 100 + tridiag_cols_scratch_floats<QD, COLS_NR, SPLIT_NR2>() + 8 + ((a.L.n + 3) & ~3) + 4) * sizeof(float);
        auto k1 = cols4 ? cov_tridiag4_kernel<FUSED, QD> : cov_tridiag_kernel<FUSED, QD, NCHQ>;
        if (cols4) smem1 = smem1_4;
        constexpr int LDG = (QD + 3) & ~3;
        auto k1b = tridiag_tail_kernel<QD, SPLIT_NR2, SPLIT_NR3, 64, 8, LDG, 4 * LDG, split_trail_off<QD>(), split_trail2_off<QD>()>;
        auto k1c = tridiag_tail_kernel<QD, SPLIT_NR3, 2, 32, 16, LDG, 4 * LDG, split_trail2_off<QD>(), 0>;
        const size_t smem1b = (size_t)tridiag_scratch_floats<QD, SPLIT_NR2, SPLIT_NR3>() * sizeof(float);
        const size_t smem1c = (size_t)tridiag_scratch_floats<QD, SPLIT_NR3, 2>() * sizeof(float);
        auto k2 = a.filter_mma ? bayes_kernel<FUSED, false, true, true> : bayes_kernel<FUSED, false, true, false>;
        e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
        k1<<<B * p->c, TT, smem1, st>>>(a);
        if (g_tail2) {      // phase 64 -> 32 with two rows per thread, one warp per problem
            auto k1b2 = tridiag_tail2_kernel<QD, SPLIT_NR2, 10, LDG, 4 * LDG, split_trail_off<QD>(), split_trail2_off<QD>()>;
            k1b2<<<B * p->c, 32, (size_t)tridiag2_scratch_floats<QD, SPLIT_NR2, SPLIT_NR3>() * sizeof(float), st>>>(a);
        } else {
            k1b<<<B * p->c, 64, smem1b, st>>>(a);
        }
        k1c<<<B * p->c, 32, smem1c, st>>>(a);
        k2<<<FUSED ? B : B * p->c, TT, smem, st>>>(a);
        return check_launch(what, 4);
    }
    const size_t smem = (size_t)a.L.total * sizeof(float);
    auto kern = a.L.gram ? bayes_kernel<FUSED, true, false> : bayes_kernel<FUSED, false, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
    kern<<<FUSED ? B : B * p->c, TT, smem, st>>>(a);
    return check_launch(what);
}

bool bayes_tridiag_supported(const VnlbBayesParams *p) {
    const int pd = p->pt * p->ps * p->ps;
    if (pd < 3 || pd > TT || p->rank > MR || p->k < 1) return false;
    const TriLayout L = tri_layout(p->k, pd);
    return (size_t)L.total * sizeof(float) <= 227 * 1024 && ((L.LDq >> 2) * ((L.LDq >> 2) + 1) / 2) <= 3 * TT;
}

int bayes_matrix_dim(const VnlbBayesParams *p, int *is_gram) {
    const TriLayout L = tri_layout(p->k, p->pt * p->ps * p->ps);
    if (is_gram) *is_gram = L.gram;
    return L.q;
}

int launch_bayes_tridiag(float *pnoisy, const float *pbasic, const unsigned char *flat, const long long *inds, int B,
                         const VnlbBayesParams *p, float *rank_var, void *ws, size_t ws_bytes, cudaStream_t st,
                         float *dbg_mat, float *dbg_lam, float *dbg_coef, int *dbg_m) {
    BayesArgs a = {};
    a.pnoisy = pnoisy; a.pbasic = pbasic; a.flat = flat; a.inds = inds; a.rank_var = rank_var;
    a.dbg_mat = dbg_mat; a.dbg_lam = dbg_lam; a.dbg_coef = dbg_coef; a.dbg_m = dbg_m;
    a.P = *p;
    a.L = tri_layout(p->k, p->pt * p->ps * p->ps);
    return launch_bayes_any<false>(a, B, "vnlb_bayes_filter(tridiag)", ws, ws_bytes, st);
}

int launch_bayes_fused(const float *img_noisy, const float *img_basic, const long long *inds, int B, int T, int H,
                       int W, const VnlbBayesParams *p, float flat_thresh, float *deno, float *weights,
                       void *ws, size_t ws_bytes, cudaStream_t st) {
    BayesArgs a = {};
    a.img_noisy = img_noisy; a.img_basic = img_basic; a.inds = inds; a.deno = deno; a.weights = weights;
    a.T = T; a.H = H; a.W = W; a.flat_thresh = flat_thresh;
    a.P = *p;
    a.L = tri_layout(p->k, p->pt * p->ps * p->ps);
    return launch_bayes_any<true>(a, B, "vnlb_bayes_aggregate_fused", ws, ws_bytes, st);
}

}  // namespace vnlb
