// bayes_tridiag.cu -- production path of the per-group Bayes estimate
// (VNLB_EIG_TRIDIAG).  Replaces bayes_est.denoise (lib/vnlb/deno/bayes_est.py:17-62)
// including torch.linalg.eigh (:122), the dominant cost of the reference.
//
// One CTA (128 threads) per (group, channel); everything between the first
// read of the patch stack and the write of the filtered stack stays in shared
// memory (A: p x LD floats, Z: 40 x (p|1) floats, ~60 KB => 3 CTAs per SM):
//
//   0. centre + covariance   C = Y^T Y / n          register-tiled 4x4 FFMA tiles
//   1. Householder tridiagonalisation C -> (d, e)   reflectors kept in the rows of A
//   2. eigenvalues above the Wiener threshold only: one Sturm count gives their
//      number m (<= rank); parallel multisection with 4 interleaved Sturm chains
//      per thread isolates them to FP32 resolution
//   3. eigenvectors of the tridiagonal matrix by twisted factorisation
//      (one thread per eigenpair, no re-orthogonalisation needed: see
//      tools/proto_tridiag.py for the accuracy study)
//   4. back-transformation by the stored reflectors
//   5. Wiener filter  Xhat = X V diag(w) V^T + mean   (8 eigenpairs per pass)
//
// Only the m eigenpairs that can receive a non-zero filter coefficient are ever
// computed: eigenvalue_j > thresh*sigma^2 + sigma_b^2 (bayes_est.py:129-144).
#include "common.cuh"

namespace vnlb {

constexpr int TT = 128;  // threads per CTA
constexpr int MR = 40;   // eigenpairs kept at most (args.rank <= MR)
constexpr int ILP = 4;   // interleaved Sturm chains per thread
constexpr int CH = 32;   // patches staged per covariance chunk

struct TriLayout {
    int p, LD, ZS, n;
    int oA, oZ, oD, oE, oE2, oTau, oV, oW, oMean, oLam, oCoef, oLo, oHi, oNlo, oNhi, oRed, total;
};

static TriLayout tri_layout(int n, int p) {
    TriLayout L;
    L.p = p; L.n = n;
    L.LD = (p + 3) & ~3;
    L.ZS = p | 1;
    int o = 0;
    {
        const int a1 = L.LD * L.LD, a2 = n * (L.LD + 1);   // A, later X[n][LD+1]
        L.oA = o; o += ((a1 > a2 ? a1 : a2) + 3) & ~3;
    }
    int zsz = MR * L.ZS;                                   // Z[MR][ZS], later Vt[p][MR]
    if (zsz < CH * L.LD) zsz = CH * L.LD;
    if (zsz < p * MR) zsz = p * MR;
    L.oZ = o; o += (zsz + 3) & ~3;
    L.oD = o; o += L.LD;
    L.oE = o; o += L.LD;
    L.oE2 = o; o += L.LD;
    L.oTau = o; o += L.LD;
    L.oV = o; o += L.LD;
    L.oW = o; o += L.LD;
    L.oMean = o; o += L.LD;
    L.oLam = o; o += MR;
    L.oCoef = o; o += MR;
    L.oLo = o; o += MR;
    L.oHi = o; o += MR;
    L.oNlo = o; o += MR;
    L.oNhi = o; o += MR;
    L.oRed = o; o += 16;
    L.total = o;
    return L;
}

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
// block-wide sum broadcast to every thread; `red` has 8 slots, `phase` alternates 0/1
__device__ __forceinline__ float block_sum(float v, float *red, int &phase) {
    v = warp_sum(v);
    float *r = red + 4 * phase;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    phase ^= 1;
    return (r[0] + r[1]) + (r[2] + r[3]);
}
__device__ __forceinline__ float block_max(float v, float *red, int &phase) {
    v = warp_max(v);
    float *r = red + 4 * phase;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    phase ^= 1;
    return fmaxf(fmaxf(r[0], r[1]), fmaxf(r[2], r[3]));
}

__global__ void __launch_bounds__(TT, 3)
bayes_tridiag_kernel(float *__restrict__ pnoisy, const float *__restrict__ pbasic,
                     const unsigned char *__restrict__ flat, const long long *__restrict__ inds, VnlbBayesParams P,
                     float *__restrict__ rank_var, TriLayout L) {
    extern __shared__ __align__(16) float sm[];
    const int n = P.k, ps2 = P.ps * P.ps, p = L.p, LD = L.LD, ZS = L.ZS, C = P.c;
    const int g = blockIdx.x / C, ch = blockIdx.x - g * C;
    const int tid = threadIdx.x;
    if (inds && !row_valid_block(inds + (long long)g * n, n)) return;

    float *A = sm + L.oA, *Z = sm + L.oZ, *d = sm + L.oD, *e = sm + L.oE, *e2 = sm + L.oE2, *taus = sm + L.oTau;
    float *v = sm + L.oV, *w = sm + L.oW, *mean = sm + L.oMean, *lam = sm + L.oLam, *coef = sm + L.oCoef;
    float *blo = sm + L.oLo, *bhi = sm + L.oHi;
    int *nlo = (int *)(sm + L.oNlo), *nhi = (int *)(sm + L.oNhi);
    float *red = sm + L.oRed;
    int phase = 0;

    const bool step2 = P.step == 1;
    const int rstride = P.pt * C * ps2;                       // floats between consecutive patches
    const long long gbase = (long long)g * n * rstride;
    auto joff = [&](int j) { const int dt = j / ps2; return (dt * C + ch) * ps2 + (j - dt * ps2); };
    const float *src = P.cov_from_basic ? pbasic + gbase : pnoisy + gbase;
    const float inv_n = 1.f / (float)n;

    // ------------------------------------------------------------------ 0. centre + covariance
    for (int j = tid; j < LD; j += TT) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (j < p) {
            const float *q = src + joff(j);
            int nn = 0;
            for (; nn + 3 < n; nn += 4) {
                s0 += q[(long long)nn * rstride];
                s1 += q[(long long)(nn + 1) * rstride];
                s2 += q[(long long)(nn + 2) * rstride];
                s3 += q[(long long)(nn + 3) * rstride];
            }
            for (; nn < n; ++nn) s0 += q[(long long)nn * rstride];
        }
        mean[j] = ((s0 + s1) + (s2 + s3)) * inv_n;
        v[j] = 0.f;
        w[j] = 0.f;
    }
    // lower-triangular 4x4 tiles of C owned by this thread
    const int ntile = LD >> 2, ntri = ntile * (ntile + 1) / 2;
    int ti[3], tj[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int idx = tid + t * TT;
        int a = -1, b = 0;
        if (idx < ntri) {
            a = (int)((sqrtf(8.f * idx + 1.f) - 1.f) * 0.5f);
            while (a * (a + 1) / 2 > idx) --a;
            while ((a + 1) * (a + 2) / 2 <= idx) ++a;
            b = idx - a * (a + 1) / 2;
        }
        ti[t] = a;
        tj[t] = b;
    }
    float acc[3][16];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[t][q] = 0.f;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int rows = min(CH, n - c0);
        for (int idx = tid; idx < rows * LD; idx += TT) {
            const int nn = idx / LD, j = idx - nn * LD;
            Z[idx] = (j < p) ? src[(long long)(c0 + nn) * rstride + joff(j)] - mean[j] : 0.f;
        }
        __syncthreads();
        for (int nn = 0; nn < rows; ++nn) {
            const float *row = Z + nn * LD;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (ti[t] >= 0) {
                    const float4 xi = *reinterpret_cast<const float4 *>(row + 4 * ti[t]);
                    const float4 xj = *reinterpret_cast<const float4 *>(row + 4 * tj[t]);
                    const float a4[4] = {xi.x, xi.y, xi.z, xi.w}, b4[4] = {xj.x, xj.y, xj.z, xj.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[t][a * 4 + b] = fmaf(a4[a], b4[b], acc[t][a * 4 + b]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < 3; ++t)
        if (ti[t] >= 0)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float cval = acc[t][a * 4 + b] * inv_n;
                    A[(4 * ti[t] + a) * LD + 4 * tj[t] + b] = cval;
                    A[(4 * tj[t] + b) * LD + 4 * ti[t] + a] = cval;
                }
    __syncthreads();
    {   // rank_var = mean over channels of trace(C) (bayes_est.py:39-40)
        const float tr = block_sum(tid < p ? A[tid * LD + tid] : 0.f, red, phase);
        if (rank_var && tid == 0) atomicAdd(&rank_var[g], tr / (float)C);
    }

    // ------------------------------------------------------------------ 1. tridiagonalisation
    for (int k = 0; k < p - 2; ++k) {
        const int m = p - 1 - k;
        const bool active = tid < m;
        const int i = k + 1 + tid;
        const float x = active ? A[k * LD + i] : 0.f;
        const float ssq = block_sum((active && tid > 0) ? x * x : 0.f, red, phase);
        const float alpha = A[k * LD + k + 1];
        if (ssq == 0.f) {  // nothing to annihilate
            if (tid == 0) { d[k] = A[k * LD + k]; e[k] = alpha; taus[k] = 0.f; v[k] = 0.f; w[k] = 0.f; }
            continue;
        }
        const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, ssq)), alpha);
        const float tau = (beta - alpha) / beta;
        const float scale = 1.f / (alpha - beta);
        const float vi = (tid == 0) ? 1.f : x * scale;
        __syncthreads();  // every thread has read row k before it is overwritten
        if (active) { v[i] = vi; A[k * LD + i] = vi; }
        if (tid == 0) { v[k] = 0.f; w[k] = 0.f; d[k] = A[k * LD + k]; e[k] = beta; taus[k] = tau; }
        __syncthreads();
        const int jb = (k + 1) & ~3;
        float yi = 0.f;
        if (active) {
            const float *ar = A + i * LD;
            float y0 = 0.f, y1 = 0.f;
            for (int j = jb; j < LD; j += 4) {
                const float4 a4 = *reinterpret_cast<const float4 *>(ar + j);
                const float4 v4 = *reinterpret_cast<const float4 *>(v + j);
                y0 = fmaf(a4.x, v4.x, y0); y1 = fmaf(a4.y, v4.y, y1);
                y0 = fmaf(a4.z, v4.z, y0); y1 = fmaf(a4.w, v4.w, y1);
            }
            yi = y0 + y1;
        }
        float wi = tau * yi;
        const float s = block_sum(active ? wi * vi : 0.f, red, phase);
        wi = fmaf(-0.5f * tau * s, vi, wi);
        if (active) w[i] = wi;
        __syncthreads();
        if (active) {
            float *ar = A + i * LD;
            for (int j = jb; j < LD; j += 4) {
                float4 a4 = *reinterpret_cast<float4 *>(ar + j);
                const float4 v4 = *reinterpret_cast<const float4 *>(v + j);
                const float4 w4 = *reinterpret_cast<const float4 *>(w + j);
                a4.x -= fmaf(vi, w4.x, wi * v4.x); a4.y -= fmaf(vi, w4.y, wi * v4.y);
                a4.z -= fmaf(vi, w4.z, wi * v4.z); a4.w -= fmaf(vi, w4.w, wi * v4.w);
                *reinterpret_cast<float4 *>(ar + j) = a4;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (p >= 2) {
            d[p - 2] = A[(p - 2) * LD + p - 2];
            e[p - 2] = A[(p - 2) * LD + p - 1];
            taus[p - 2] = 0.f;
        }
        d[p - 1] = A[(p - 1) * LD + p - 1];
        e[p - 1] = 0.f;
    }
    __syncthreads();

    // ------------------------------------------------------------------ 2. eigenvalues above the threshold
    // e2[i] = e[i-1]^2 is the coupling that enters pivot i of the Sturm sequence
    float gl = -3.4e38f, em = 0.f;
    for (int j = tid; j < p; j += TT) {
        const float el = j > 0 ? e[j - 1] : 0.f, er = j < p - 1 ? e[j] : 0.f;
        e2[j] = el * el;
        gl = fmaxf(gl, d[j] + fabsf(el) + fabsf(er));
        em = fmaxf(em, el * el);
    }
    const float gmax0 = block_max(gl, red, phase);
    const float emax2 = block_max(em, red, phase);
    const float pivmin = fmaxf(1e-30f, emax2 * 1e-12f);
    const float tau_eff = fmaf(P.thresh, P.sigma2, P.sigmab2);   // eigenvalue > tau_eff <=> coefficient != 0
    const float gmax = gmax0 + 1e-6f * fabsf(gmax0) + pivmin;

    auto sturm4 = [&](const float (&x)[ILP], int (&cnt)[ILP]) {  // #eigenvalues < x, 4 shifts at once
        float q[ILP];
#pragma unroll
        for (int c = 0; c < ILP; ++c) { q[c] = 1.f; cnt[c] = 0; }
        for (int j = 0; j < p; ++j) {
            const float dj = d[j], ej = e2[j];
#pragma unroll
            for (int c = 0; c < ILP; ++c) {
                float t = (dj - x[c]) - __fdividef(ej, q[c]);
                if (fabsf(t) < pivmin) t = -pivmin;
                q[c] = t;
                cnt[c] += t < 0.f;
            }
        }
    };
    int m;  // eigenpairs to compute
    {
        const float x[ILP] = {tau_eff, tau_eff, tau_eff, tau_eff};
        int cnt[ILP];
        sturm4(x, cnt);
        m = min(p - cnt[0], P.rank);
        if (gmax <= tau_eff) m = 0;
    }
    if (m > 0) {
        for (int j = tid; j < m; j += TT) { blo[j] = tau_eff; bhi[j] = gmax; }
        __syncthreads();
        const int G = (TT * ILP) / m;                 // section points per eigenvalue and round
        int rounds = 1;
        { float res = (float)(G + 1); while (res < 6.7e7f && rounds < 12) { res *= (float)(G + 1); ++rounds; } }
        for (int r = 0; r < rounds; ++r) {
            for (int j = tid; j < m; j += TT) { nlo[j] = __float_as_int(blo[j]); nhi[j] = __float_as_int(bhi[j]); }
            __syncthreads();
            float x[ILP];
            int ej[ILP], cnt[ILP];
#pragma unroll
            for (int c = 0; c < ILP; ++c) {
                const int pt = tid * ILP + c;
                const int j = pt / G, q = pt - j * G;
                ej[c] = j < m ? j : -1;
                const float lo = blo[min(j, m - 1)], hi = bhi[min(j, m - 1)];
                x[c] = fmaf(hi - lo, (float)(q + 1) / (float)(G + 1), lo);
            }
            sturm4(x, cnt);
#pragma unroll
            for (int c = 0; c < ILP; ++c)
                if (ej[c] >= 0) {
                    const int idx = p - 1 - ej[c];   // ascending index of the ej-th largest eigenvalue
                    if (cnt[c] <= idx) atomicMax(&nlo[ej[c]], __float_as_int(x[c]));   // x is a lower bound
                    else atomicMin(&nhi[ej[c]], __float_as_int(x[c]));                  // x is an upper bound
                }
            __syncthreads();
            for (int j = tid; j < m; j += TT) {
                float lo = __int_as_float(nlo[j]), hi = __int_as_float(nhi[j]);
                if (lo > hi) { const float mid = 0.5f * (lo + hi); lo = mid; hi = mid; }
                blo[j] = lo; bhi[j] = hi;
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

        // -------------------------------------------------------------- 3. eigenvectors of T (twisted factorisation)
        if (tid < m) {
            float *z = Z + tid * ZS;
            const float l = lam[tid];
            const float pm = fmaxf(pivmin, 1e-7f * fabsf(l) * 1e-3f);
            // forward pivots D+ (kept in z)
            float dp = d[0] - l;
            if (fabsf(dp) < pm) dp = -pm;
            z[0] = dp;
            for (int i = 0; i < p - 1; ++i) {
                dp = (d[i + 1] - l) - e2[i + 1] / dp;
                if (fabsf(dp) < pm) dp = -pm;
                z[i + 1] = dp;
            }
            // backward pivots D-: first pass finds the twist index r = argmin |gamma_i|
            float dm = d[p - 1] - l;
            if (fabsf(dm) < pm) dm = -pm;
            float best = fabsf(z[p - 1]);   // gamma_{p-1} = D+_{p-1} + D-_{p-1} - (d_{p-1} - l) = D+_{p-1}
            int r = p - 1;
            for (int i = p - 2; i >= 0; --i) {
                dm = (d[i] - l) - e2[i + 1] / dm;
                if (fabsf(dm) < pm) dm = -pm;
                const float gam = fabsf(z[i] + dm - (d[i] - l));
                if (gam < best) { best = gam; r = i; }
            }
            // second pass stores D-_i for i > r (D+_i is no longer needed there)
            dm = d[p - 1] - l;
            if (fabsf(dm) < pm) dm = -pm;
            if (r < p - 1) z[p - 1] = dm;
            for (int i = p - 2; i > r; --i) {
                dm = (d[i] - l) - e2[i + 1] / dm;
                if (fabsf(dm) < pm) dm = -pm;
                z[i] = dm;
            }
            // z_r = 1; upwards with D+, downwards with D-
            float zi = 1.f, nrm = 1.f;
            for (int i = r - 1; i >= 0; --i) {
                zi = -(e[i] / z[i]) * zi;
                z[i] = zi;
                nrm = fmaf(zi, zi, nrm);
            }
            zi = 1.f;
            for (int i = r; i < p - 1; ++i) {
                zi = -(e[i] / z[i + 1]) * zi;
                z[i + 1] = zi;
                nrm = fmaf(zi, zi, nrm);
            }
            z[r] = 1.f;
            const float sc = rsqrtf(nrm);
            for (int i = 0; i < p; ++i) z[i] *= sc;
        }
        __syncthreads();

        // -------------------------------------------------------------- 4. back-transformation  z <- H_0 ... H_{p-3} z
        {
            int S = 2;                                   // lanes per eigenvector (power of two, S*m <= 128)
            while (S * 2 * m <= TT && S < 32) S *= 2;
            if (m > TT / 2) S = 1;
            const int vec = tid / S, part = tid - vec * S;
            const bool on = vec < m;
            float *z = Z + min(vec, m - 1) * ZS;
            for (int k = p - 3; k >= 0; --k) {
                const float tau = taus[k];
                if (tau == 0.f) continue;
                const float *vr = A + k * LD;            // reflector k: vr[k+1] = 1, vr[k+2..p-1]
                float s = 0.f;
                for (int i = k + 1 + part; i < p; i += S) s = fmaf(vr[i], z[i], s);
                for (int dlt = S >> 1; dlt > 0; dlt >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dlt);
                s *= tau;
                if (on)
                    for (int i = k + 1 + part; i < p; i += S) z[i] = fmaf(-s, vr[i], z[i]);
                __syncwarp();
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ 5. Wiener filter of the noisy patches
    // eigenvectors re-laid as Vt[j][r] (r contiguous, pitch MR) so 4 eigenpairs come per broadcast LDS.128
    if (m > 0) {
        float tmp[(MR * 129 + TT - 1) / TT];
        const int tot = m * p;
#pragma unroll
        for (int q = 0; q < (MR * 129 + TT - 1) / TT; ++q) {
            const int idx = tid + q * TT;
            tmp[q] = 0.f;
            if (idx < tot) { const int r = idx / p, j = idx - r * p; tmp[q] = Z[r * ZS + j]; }
        }
        __syncthreads();
        const int mpad = (m + 7) & ~7;
        for (int idx = tid; idx < p * MR; idx += TT) Z[idx] = 0.f;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < (MR * 129 + TT - 1) / TT; ++q) {
            const int idx = tid + q * TT;
            if (idx < tot) { const int r = idx / p, j = idx - r * p; Z[j * MR + r] = tmp[q]; }
        }
        (void)mpad;
    }
    // stage the noisy patches of this channel in the A region: X[n][XS], XS odd => conflict-free rows
    const int XS = LD + 1;
    float *X = A;
    float *xg = pnoisy + gbase;
    __syncthreads();
    for (int idx = tid; idx < n * p; idx += TT) {
        const int nn = idx / p, j = idx - nn * p;
        X[nn * XS + j] = xg[(long long)nn * rstride + joff(j)];
    }
    __syncthreads();
    const bool is_flat = step2 && flat && flat[g];
    for (int j = tid; j < p; j += TT) {
        float s0 = 0.f, s1 = 0.f;
        if (!is_flat) {                                 // cnoisy = mean_n(noisy)            bayes_est.py:96
            int nn = 0;
            for (; nn + 1 < n; nn += 2) { s0 += X[nn * XS + j]; s1 += X[(nn + 1) * XS + j]; }
            if (nn < n) s0 += X[nn * XS + j];
        } else {                                        // flat group: cnoisy = cbasic       bayes_est.py:97-101
            const float *q = pbasic + gbase + joff(j);
            for (int nn = 0; nn < n; ++nn) s0 += q[(long long)nn * rstride];
        }
        mean[j] = (s0 + s1) * inv_n;
    }
    __syncthreads();
    for (int nn = tid; nn < n; nn += TT) {
        float *xr = X + nn * XS;
        // projections  zc[r] = coef_r * <x - mean, v_r>
        float zc[MR];
#pragma unroll
        for (int r = 0; r < MR; ++r) zc[r] = 0.f;
        if (m > 0) {
            for (int j = 0; j < p; ++j) {
                const float xv = xr[j] - mean[j];
                const float4 *vt = reinterpret_cast<const float4 *>(Z + j * MR);
#pragma unroll
                for (int rb = 0; rb < MR; rb += 8)
                    if (rb < m) {
                        const float4 a = vt[rb >> 2], b = vt[(rb >> 2) + 1];
                        zc[rb + 0] = fmaf(xv, a.x, zc[rb + 0]); zc[rb + 1] = fmaf(xv, a.y, zc[rb + 1]);
                        zc[rb + 2] = fmaf(xv, a.z, zc[rb + 2]); zc[rb + 3] = fmaf(xv, a.w, zc[rb + 3]);
                        zc[rb + 4] = fmaf(xv, b.x, zc[rb + 4]); zc[rb + 5] = fmaf(xv, b.y, zc[rb + 5]);
                        zc[rb + 6] = fmaf(xv, b.z, zc[rb + 6]); zc[rb + 7] = fmaf(xv, b.w, zc[rb + 7]);
                    }
            }
#pragma unroll
            for (int r = 0; r < MR; ++r) zc[r] = (r < m) ? zc[r] * coef[r] : 0.f;
        }
        // reconstruction  xhat = sum_r zc[r] v_r + mean                                     bayes_est.py:51
        for (int j = 0; j < p; ++j) {
            float o0 = 0.f, o1 = 0.f;
            if (m > 0) {
                const float4 *vt = reinterpret_cast<const float4 *>(Z + j * MR);
#pragma unroll
                for (int rb = 0; rb < MR; rb += 8)
                    if (rb < m) {
                        const float4 a = vt[rb >> 2], b = vt[(rb >> 2) + 1];
                        o0 = fmaf(zc[rb + 0], a.x, o0); o1 = fmaf(zc[rb + 1], a.y, o1);
                        o0 = fmaf(zc[rb + 2], a.z, o0); o1 = fmaf(zc[rb + 3], a.w, o1);
                        o0 = fmaf(zc[rb + 4], b.x, o0); o1 = fmaf(zc[rb + 5], b.y, o1);
                        o0 = fmaf(zc[rb + 6], b.z, o0); o1 = fmaf(zc[rb + 7], b.w, o1);
                    }
            }
            xr[j] = (o0 + o1) + mean[j];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < n * p; idx += TT) {
        const int nn = idx / p, j = idx - nn * p;
        xg[(long long)nn * rstride + joff(j)] = X[nn * XS + j];
    }
}

bool bayes_tridiag_supported(const VnlbBayesParams *p) {
    const int pd = p->pt * p->ps * p->ps;
    if (pd < 3 || pd > TT || p->rank > MR) return false;
    const TriLayout L = tri_layout(p->k, pd);
    return (size_t)L.total * sizeof(float) <= 227 * 1024 && ((L.LD >> 2) * ((L.LD >> 2) + 1) / 2) <= 3 * TT;
}

int launch_bayes_tridiag(float *pnoisy, const float *pbasic, const unsigned char *flat, const long long *inds, int B,
                         const VnlbBayesParams *p, float *rank_var, cudaStream_t st) {
    const int pd = p->pt * p->ps * p->ps;
    const TriLayout L = tri_layout(p->k, pd);
    const size_t smem = (size_t)L.total * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(bayes_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("vnlb_bayes_filter: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
    bayes_tridiag_kernel<<<B * p->c, TT, smem, st>>>(pnoisy, pbasic, flat, inds, *p, rank_var, L);
    return check_launch("vnlb_bayes_filter(tridiag)");
}

}  // namespace vnlb
