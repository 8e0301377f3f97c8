// aggregate.cu -- weighted patch aggregation (uniform weight 1 per patch).
// Replaces agg_patches -> exec_agg_simple_numba
// (lib/vnlb/agg/comp_agg.py:47-60,106-138), which the reference runs on one
// CPU thread behind whole-video D2H/H2D copies.  Here: one CTA per group,
// float atomics (RED.ADD.F32) straight into the L2-resident accumulators;
// consecutive lanes walk a patch row, so each warp request touches few
// 32-byte sectors.
#include "common.cuh"

namespace vnlb {

__global__ void __launch_bounds__(256)
aggregate_kernel(const float *__restrict__ patches, const long long *__restrict__ inds, int K,
                 float *__restrict__ deno, float *__restrict__ weights, int T, int C, int H, int W, int ps, int pt) {
    const int g = blockIdx.x;
    const long long *row = inds + (long long)g * K;
    if (!row_valid_block(row, K)) return;
    const int ps2 = ps * ps, pdim = pt * C * ps2;
    const long long HW = (long long)H * W, CHW = (long long)C * HW;
    const float *gp = patches + (long long)g * K * pdim;
    const int items = K * pt * ps2;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int r = it % ps2, dt = (it / ps2) % pt, nn = it / (ps2 * pt);
        const int dy = r / ps, dx = r - dy * ps;
        int t, y, x;
        decode_ind(row[nn], H, W, C, t, y, x);
        const int t1 = t + dt, y1 = y + dy, x1 = x + dx;
        if (t1 >= T || y1 >= H || x1 >= W) continue;
        const long long pix = (long long)y1 * W + x1;
        const float *pp = gp + (long long)nn * pdim + (long long)dt * C * ps2 + r;
        for (int c = 0; c < C; ++c) atomicAdd(deno + (long long)t1 * CHW + c * HW + pix, pp[c * ps2]);
        atomicAdd(weights + (long long)t1 * HW + pix, 1.f);
    }
}

}  // namespace vnlb

using namespace vnlb;

extern "C" int vnlb_aggregate(const float *patches, const int64_t *inds, int B, int K, float *deno, float *weights,
                              int T, int C, int H, int W, int ps, int pt, void *stream) {
    VNLB_REQUIRE(patches && inds && deno && weights, "vnlb_aggregate: null pointer");
    VNLB_REQUIRE(B >= 0 && K > 0 && T > 0 && C > 0 && H > 0 && W > 0 && ps >= 1 && pt >= 1, "vnlb_aggregate: bad shape");
    if (B == 0) return VNLB_OK;
    aggregate_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(patches, (const long long *)inds, K, deno, weights, T, C, H,
                                                         W, ps, pt);
    return check_launch("vnlb_aggregate");
}
