// bayes.cu -- C-ABI dispatch of the per-group Bayes estimate.
#include "common.cuh"

namespace vnlb {
int launch_bayes_jacobi(float *pnoisy, const float *pbasic, const unsigned char *flat, const long long *inds, int B,
                        const VnlbBayesParams *p, float *rank_var, cudaStream_t st);
int launch_bayes_tridiag(float *pnoisy, const float *pbasic, const unsigned char *flat, const long long *inds, int B,
                         const VnlbBayesParams *p, float *rank_var, void *ws, size_t ws_bytes, cudaStream_t st,
                         float *dbg_mat = nullptr, float *dbg_lam = nullptr, float *dbg_coef = nullptr, int *dbg_m = nullptr);
int bayes_matrix_dim(const VnlbBayesParams *p, int *is_gram);
size_t bayes_workspace_bytes(int B, const VnlbBayesParams *p);
bool bayes_tridiag_supported(const VnlbBayesParams *p);
int set_bayes_split(int on);
int set_filter_mma(int on);
int launch_bayes_fused(const float *img_noisy, const float *img_basic, const long long *inds, int B, int T, int H,
                       int W, const VnlbBayesParams *p, float flat_thresh, float *deno, float *weights,
                       void *ws, size_t ws_bytes, cudaStream_t st);
}

using namespace vnlb;

extern "C" size_t vnlb_bayes_workspace_bytes(int B, const VnlbBayesParams *p) {
    if (!p || p->eig_method != VNLB_EIG_TRIDIAG) return 0;
    return bayes_workspace_bytes(B, p);
}

extern "C" int vnlb_bayes_filter(float *pnoisy, const float *pbasic, const uint8_t *flat, const int64_t *inds, int B,
                                 const VnlbBayesParams *p, float *rank_var, void *ws, size_t ws_bytes, void *stream) {
    VNLB_REQUIRE(pnoisy && p && B >= 0, "vnlb_bayes_filter: bad argument");
    VNLB_REQUIRE(p->step == 0 || p->step == 1, "vnlb_bayes_filter: step must be 0 or 1");
    VNLB_REQUIRE(p->k >= 2 && p->ps >= 1 && p->pt >= 1 && p->c >= 1, "vnlb_bayes_filter: bad patch shape");
    VNLB_REQUIRE(p->rank >= 1, "vnlb_bayes_filter: rank must be >= 1");
    VNLB_REQUIRE(p->sigma2 > 0.f && p->sigmab2 >= 0.f, "vnlb_bayes_filter: sigma2 must be > 0");
    VNLB_REQUIRE(!(p->step == 1 || p->cov_from_basic) || pbasic, "vnlb_bayes_filter: basic patches required");
    if (B == 0) return VNLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (rank_var) {
        cudaError_t e = cudaMemsetAsync(rank_var, 0, sizeof(float) * B, st);
        if (e != cudaSuccess) { set_error("vnlb_bayes_filter: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
    }
    switch (p->eig_method) {
        case VNLB_EIG_TRIDIAG:
            if (bayes_tridiag_supported(p))
                return launch_bayes_tridiag(pnoisy, pbasic, flat, (const long long *)inds, B, p, rank_var, ws, ws_bytes, st);
            // shapes outside the tridiagonal kernel's envelope (p > 128 or rank > 40) use the Jacobi kernel
        case VNLB_EIG_JACOBI:
            return launch_bayes_jacobi(pnoisy, pbasic, flat, (const long long *)inds, B, p, rank_var, st);
        default:
            set_error("vnlb_bayes_filter: unknown eig_method %d", p->eig_method);
            return VNLB_ERR_BAD_ARG;
    }
}

extern "C" int vnlb_bayes_matrix_dim(const VnlbBayesParams *p, int *is_gram) {
    if (!p || !bayes_tridiag_supported(p)) return 0;
    return bayes_matrix_dim(p, is_gram);
}

extern "C" int vnlb_bayes_debug(float *pnoisy, const float *pbasic, const uint8_t *flat, const int64_t *inds, int B,
                                const VnlbBayesParams *p, float *mat, float *lam, float *coef, int32_t *m, void *ws,
                                size_t ws_bytes, void *stream) {
    VNLB_REQUIRE(pnoisy && p && B >= 0, "vnlb_bayes_debug: bad argument");
    VNLB_REQUIRE(p->step == 0 || p->step == 1, "vnlb_bayes_debug: step must be 0 or 1");
    VNLB_REQUIRE(p->k >= 2 && p->ps >= 1 && p->pt >= 1 && p->c >= 1 && p->rank >= 1, "vnlb_bayes_debug: bad patch shape");
    VNLB_REQUIRE(p->sigma2 > 0.f && p->sigmab2 >= 0.f, "vnlb_bayes_debug: sigma2 must be > 0");
    VNLB_REQUIRE(!(p->step == 1 || p->cov_from_basic) || pbasic, "vnlb_bayes_debug: basic patches required");
    if (p->eig_method != VNLB_EIG_TRIDIAG || !bayes_tridiag_supported(p)) {
        set_error("vnlb_bayes_debug: only the VNLB_EIG_TRIDIAG path exports its intermediates");
        return VNLB_ERR_UNSUPPORTED;
    }
    if (B == 0) return VNLB_OK;
    return launch_bayes_tridiag(pnoisy, pbasic, flat, (const long long *)inds, B, p, nullptr, ws, ws_bytes,
                                (cudaStream_t)stream, mat, lam, coef, (int *)m);
}

extern "C" int vnlb_set_bayes_split(int on) { return set_bayes_split(on); }
extern "C" int vnlb_set_filter_mma(int on) { return set_filter_mma(on); }

extern "C" int vnlb_bayes_fused_supported(const VnlbBayesParams *p) { return p && bayes_tridiag_supported(p) ? 1 : 0; }

extern "C" int vnlb_bayes_aggregate_fused(const float *img_noisy, const float *img_basic, const int64_t *inds, int B,
                                          int T, int C, int H, int W, const VnlbBayesParams *p, float flat_thresh,
                                          float *deno, float *weights, void *ws, size_t ws_bytes, void *stream) {
    VNLB_REQUIRE(img_noisy && inds && p && deno && weights && B >= 0, "vnlb_bayes_aggregate_fused: null pointer");
    VNLB_REQUIRE(T > 0 && C > 0 && H > 0 && W > 0 && C == p->c, "vnlb_bayes_aggregate_fused: bad shape");
    VNLB_REQUIRE((long long)T * C * H * W < (1LL << 31), "vnlb_bayes_aggregate_fused: video too large for 32-bit offsets");
    VNLB_REQUIRE(p->step == 0 || p->step == 1, "vnlb_bayes_aggregate_fused: step must be 0 or 1");
    VNLB_REQUIRE(p->sigma2 > 0.f && p->sigmab2 >= 0.f && p->rank >= 1, "vnlb_bayes_aggregate_fused: bad parameters");
    VNLB_REQUIRE(!(p->step == 1 || p->cov_from_basic) || img_basic, "vnlb_bayes_aggregate_fused: basic image required");
    if (!bayes_tridiag_supported(p)) {
        set_error("vnlb_bayes_aggregate_fused: patch shape outside the fused kernel's envelope (p <= 128, rank <= 40)");
        return VNLB_ERR_UNSUPPORTED;
    }
    if (B == 0) return VNLB_OK;
    return launch_bayes_fused(img_noisy, img_basic, (const long long *)inds, B, T, H, W, p, flat_thresh, deno, weights,
                              ws, ws_bytes, (cudaStream_t)stream);
}
