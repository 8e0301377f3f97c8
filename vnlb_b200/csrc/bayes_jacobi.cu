// bayes_jacobi.cu -- per-group Bayes estimate with a cyclic Jacobi eigensolver
// held in shared memory (VNLB_EIG_JACOBI).  Slow (O(10 p^3) flops per sweep
// set) but unconditionally accurate: it is the cross-check for the
// tridiagonal production path in bayes_tridiag.cu and serves any (ps, pt)
// whose working set fits shared memory.
//
// Replaces bayes_est.denoise (lib/vnlb/deno/bayes_est.py:17-62):
//   centre (88-110) -> covariance + eigh (112-126) -> eigenvalue shrink
//   (129-138) -> Wiener coefficients (140-144) -> low-rank filter (146-151)
//   -> re-centre (51).
// One CTA per (group, channel); channels are independent p x p problems
// (coupleChannels = False, lib/vnlb/params.py:17).
#include "common.cuh"

namespace vnlb {

constexpr int kJacThreads = 256;

struct JacLayout {
    int n, p, ld;      // ld = p + 1 (bank-conflict-free rows)
    size_t x, y, a, v, misc, total;  // float offsets
};

static JacLayout jac_layout(int n, int p) {
    JacLayout L;
    L.n = n; L.p = p; L.ld = p + 1;
    size_t o = 0;
    L.x = o; o += (size_t)n * L.ld;      // centred noisy patches  X[n][p]
    L.y = o; o += (size_t)n * L.ld;      // centred covariance input / output accumulator
    L.a = o; o += (size_t)p * L.ld;      // covariance -> diagonal; later Z[n][r]
    L.v = o; o += (size_t)p * L.ld;      // eigenvectors (columns)
    L.misc = o; o += (size_t)8 * p + 64; // means, rotations, eigenvalues, order, coeffs
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(kJacThreads)
bayes_jacobi_kernel(float *__restrict__ pnoisy, const float *__restrict__ pbasic,
                    const unsigned char *__restrict__ flat, const long long *__restrict__ inds, VnlbBayesParams P,
                    float *__restrict__ rank_var, size_t ox, size_t oy, size_t oa, size_t ov, size_t omisc) {
    extern __shared__ __align__(16) float sm[];
    const int n = P.k, ps2 = P.ps * P.ps, p = P.pt * ps2, ld = p + 1, C = P.c;
    const int g = blockIdx.x / C, ch = blockIdx.x % C;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (inds && !row_valid_block(inds + (long long)g * n, n)) return;

    float *X = sm + ox, *Y = sm + oy, *A = sm + oa, *V = sm + ov;
    float *mean_n = sm + omisc;          // [p]
    float *mean_b = mean_n + p;          // [p]
    float *rot = mean_b + p;             // [p/2][2] (c, s)
    float *lam = rot + p + 2;            // [p]
    float *coef = lam + p;               // [p]   filter coefficient per SORTED rank
    int *order = (int *)(coef + p);      // [p]   sorted rank -> column
    int *pair_ij = order + p;            // [p+2] (i, j) per pair
    __shared__ int s_rotated;
    __shared__ int s_nsel;
    __shared__ float s_tol;

    const bool step2 = P.step == 1;
    const long long gbase = (long long)g * n * P.pt * C * ps2;
    // element (nn, j) of channel ch:  j = dt*ps2 + r  ->  ((nn*pt + dt)*C + ch)*ps2 + r
    auto gaddr = [&](int nn, int j) {
        const int dt = j / ps2, r = j - dt * ps2;
        return gbase + ((long long)(nn * P.pt + dt) * C + ch) * ps2 + r;
    };
    for (int i = tid; i < n * p; i += nthr) {
        const int nn = i / p, j = i - nn * p;
        X[nn * ld + j] = pnoisy[gaddr(nn, j)];
        if (P.cov_from_basic || step2) Y[nn * ld + j] = pbasic[gaddr(nn, j)];
    }
    __syncthreads();
    // ---- centre (bayes_est.py:88-110) ----
    const bool is_flat = step2 && flat && flat[g];
    for (int j = tid; j < p; j += nthr) {
        float s = 0.f;
        for (int nn = 0; nn < n; ++nn) s += X[nn * ld + j];
        float mn = s / (float)n, mb = 0.f;
        if (P.cov_from_basic || step2) {
            float sb = 0.f;
            for (int nn = 0; nn < n; ++nn) sb += Y[nn * ld + j];
            mb = sb / (float)n;
        }
        if (is_flat) mn = mb;
        mean_n[j] = mn;
        mean_b[j] = mb;
    }
    __syncthreads();
    for (int i = tid; i < n * p; i += nthr) {
        const int nn = i / p, j = i - nn * p;
        const float xc = X[nn * ld + j] - mean_n[j];
        X[nn * ld + j] = xc;
        Y[nn * ld + j] = P.cov_from_basic ? (Y[nn * ld + j] - mean_b[j]) : xc;
    }
    __syncthreads();
    // ---- covariance  A = Y^T Y / n,  V = I ----
    for (int e = tid; e < p * p; e += nthr) {
        const int i = e / p, j = e - i * p;
        if (j <= i) {
            float s = 0.f;
            for (int nn = 0; nn < n; ++nn) s = fmaf(Y[nn * ld + i], Y[nn * ld + j], s);
            s /= (float)n;
            A[i * ld + j] = s;
            A[j * ld + i] = s;
        }
        V[i * ld + j] = (i == j) ? 1.f : 0.f;
    }
    __syncthreads();
    if (tid == 0) {  // mean over channels of trace = sum of eigenvalues (bayes_est.py:39-40)
        float tr = 0.f;
        for (int i = 0; i < p; ++i) tr += A[i * ld + i];
        if (rank_var) atomicAdd(&rank_var[g], tr / (float)C);
        s_tol = 1e-8f * tr;      // off-diagonals below the FP32 noise floor of A are not rotated
    }
    __syncthreads();
    const float tol = s_tol;
    // ---- cyclic Jacobi, round-robin ordering: p/2 disjoint rotations per step ----
    const int pe = p + (p & 1);          // odd p: one idle slot
    const int npairs = pe / 2;
    for (int sweep = 0; sweep < 30; ++sweep) {
        if (tid == 0) s_rotated = 0;
        __syncthreads();
        for (int r = 0; r < pe - 1; ++r) {
            if (tid < npairs) {
                int i, j;
                if (tid == 0) { i = pe - 1; j = r; }
                else { i = (r + tid) % (pe - 1); j = (r - tid + pe - 1) % (pe - 1); }
                if (i > j) { const int t = i; i = j; j = t; }
                float c = 1.f, s = 0.f;
                if (j < p) {
                    const float apq = A[i * ld + j], app = A[i * ld + i], aqq = A[j * ld + j];
                    if (fabsf(apq) > tol) {
                        const float theta = (aqq - app) / (2.f * apq);
                        const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
                        c = rsqrtf(fmaf(t, t, 1.f));
                        s = t * c;
                        s_rotated = 1;
                    }
                } else { j = i; }  // idle pair
                rot[2 * tid] = c; rot[2 * tid + 1] = s;
                pair_ij[2 * tid] = i; pair_ij[2 * tid + 1] = j;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int e = tid; e < p * npairs; e += nthr) {
                const int row = e % p, k = e / p;
                const int i = pair_ij[2 * k], j = pair_ij[2 * k + 1];
                if (i == j) continue;
                const float c = rot[2 * k], s = rot[2 * k + 1];
                const float ai = A[row * ld + i], aj = A[row * ld + j];
                A[row * ld + i] = c * ai - s * aj;
                A[row * ld + j] = s * ai + c * aj;
                const float vi = V[row * ld + i], vj = V[row * ld + j];
                V[row * ld + i] = c * vi - s * vj;
                V[row * ld + j] = s * vi + c * vj;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < p * npairs; e += nthr) {
                const int col = e % p, k = e / p;
                const int i = pair_ij[2 * k], j = pair_ij[2 * k + 1];
                if (i == j) continue;
                const float c = rot[2 * k], s = rot[2 * k + 1];
                const float ai = A[i * ld + col], aj = A[j * ld + col];
                A[i * ld + col] = c * ai - s * aj;
                A[j * ld + col] = s * ai + c * aj;
            }
            __syncthreads();
        }
        if (!s_rotated) break;
        __syncthreads();
    }
    // ---- sort eigenvalues descending (ties by column), shrink, Wiener coefficients ----
    for (int j = tid; j < p; j += nthr) lam[j] = A[j * ld + j];
    if (tid == 0) s_nsel = 0;
    __syncthreads();
    for (int j = tid; j < p; j += nthr) {
        const float l = lam[j];
        int rk = 0;
        for (int i = 0; i < p; ++i) rk += (lam[i] > l) || (lam[i] == l && i < j);
        order[rk] = j;
        float w = 0.f;
        if (rk < P.rank) {
            const float ls = l - fminf(l, P.sigmab2);                       // bayes_est.py:129-138
            if (ls > P.thresh * P.sigma2) w = 1.f / (1.f + P.sigma2 / ls);  // bayes_est.py:140-144
        }
        coef[rk] = w;
        if (w != 0.f) atomicMax(&s_nsel, rk + 1);
    }
    __syncthreads();
    const int nsel = s_nsel;   // ranks >= nsel have zero weight
    // ---- filter (bayes_est.py:146-151):  Z = X V_r ;  Xhat = Z (V_r diag w)^T ----
    float *Z = Y;              // [n][nsel] at row pitch ld (Y is free after the covariance)
    for (int e = tid; e < n * nsel; e += nthr) {
        const int nn = e / nsel, r = e - nn * nsel;
        const int col = order[r];
        float s = 0.f;
        for (int j = 0; j < p; ++j) s = fmaf(X[nn * ld + j], V[j * ld + col], s);
        Z[nn * ld + r] = s * coef[r];
    }
    __syncthreads();
    for (int i = tid; i < n * p; i += nthr) {
        const int nn = i / p, j = i - nn * p;
        float s = 0.f;
        for (int r = 0; r < nsel; ++r) s = fmaf(Z[nn * ld + r], V[j * ld + order[r]], s);
        pnoisy[gaddr(nn, j)] = s + mean_n[j];                               // bayes_est.py:51
    }
}

// exec_flat_areas (lib/vnlb/utils/flat_areas.py:16-34): one CTA per group
__global__ void flat_areas_kernel(const float *__restrict__ pnoisy, const long long *__restrict__ inds,
                                  unsigned char *__restrict__ flat, int K, int C, int ps2, int pt, float thresh) {
    const int g = blockIdx.x;
    if (inds && !row_valid_block(inds + (long long)g * K, K)) {
        if (threadIdx.x == 0) flat[g] = 0;
        return;
    }
    __shared__ float red[2][32];
    __shared__ float var_sum;
    if (threadIdx.x == 0) var_sum = 0.f;
    const int pdim = pt * C * ps2;
    const float *base = pnoisy + (long long)g * K * pdim;
    const int Z = K * pt * ps2;
    for (int c = 0; c < C; ++c) {
        float s = 0.f, s2 = 0.f;
        for (int i = threadIdx.x; i < Z; i += blockDim.x) {
            const int r = i % ps2, dt = (i / ps2) % pt, nn = i / (ps2 * pt);
            const float v = base[(long long)nn * pdim + (dt * C + c) * ps2 + r];
            s += v;
            s2 = fmaf(v, v, s2);
        }
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float ts = 0.f, ts2 = 0.f;
            for (int w = 0; w < (blockDim.x + 31) / 32; ++w) { ts += red[0][w]; ts2 += red[1][w]; }
            var_sum += (ts2 - ts * ts / (float)Z) / (float)(Z - 1);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) flat[g] = (var_sum / (float)C) < thresh ? 1 : 0;
}

int launch_bayes_jacobi(float *pnoisy, const float *pbasic, const unsigned char *flat, const long long *inds, int B,
                        const VnlbBayesParams *p, float *rank_var, cudaStream_t st) {
    const int pd = p->pt * p->ps * p->ps;
    const JacLayout L = jac_layout(p->k, pd);
    const size_t smem = L.total * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("vnlb_bayes_filter(jacobi): k=%d, p=%d needs %zu B of shared memory", p->k, pd, smem);
        return VNLB_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(bayes_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("vnlb_bayes_filter: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
    bayes_jacobi_kernel<<<B * p->c, kJacThreads, smem, st>>>(pnoisy, pbasic, flat, inds, *p, rank_var, L.x, L.y, L.a,
                                                            L.v, L.misc);
    return check_launch("vnlb_bayes_filter(jacobi)");
}

}  // namespace vnlb

using namespace vnlb;

extern "C" int vnlb_flat_areas(const float *pnoisy, const int64_t *inds, uint8_t *flat, int B, int K, int C, int ps,
                               int pt, float thresh, void *stream) {
    VNLB_REQUIRE(pnoisy && flat && B >= 0 && K > 0 && C > 0 && ps >= 1 && pt >= 1, "vnlb_flat_areas: bad argument");
    if (B == 0) return VNLB_OK;
    flat_areas_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pnoisy, (const long long *)inds, flat, K, C, ps * ps, pt,
                                                          thresh);
    return check_launch("vnlb_flat_areas");
}
