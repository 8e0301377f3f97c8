// pixel_ops.cu -- HBM-bound per-pixel stages: colour transforms, normalisation,
// reference-pixel mask initialisation and the "paste trick" mask update.
#include <stdarg.h>

#include "common.cuh"

namespace vnlb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ---------------------------------------------------------------------------
// colour.  Explicit _rn intrinsics forbid FMA contraction so the result is the
// reference's float32 evaluation order bit for bit (lib/vnlb/utils/color.py).
// ---------------------------------------------------------------------------
template <bool FWD>
__global__ void color_kernel(const float *__restrict__ src, float *__restrict__ dst, int T, long long HW) {
    const float w0 = 0.57735026918962584f;   // 1/sqrt(3)
    const float w1 = 0.70710678118654757f;   // 1/sqrt(2)
    const float w2f = 1.63299316185545207f;  // 2*sqrt(2)/sqrt(3)
    const float w2 = 0.81649658092772603f;   // sqrt(2)/sqrt(3)
    const float w2h = 0.40824829046386302f;  // sqrt(2)/sqrt(3)/2
    const long long n = (long long)T * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / HW, r = i - t * HW;
        const long long o = t * 3 * HW + r;
        const float a = src[o], b = src[o + HW], c = src[o + 2 * HW];
        float x, y, z;
        if (FWD) {  // color.py:52-77
            x = __fmul_rn(w0, __fadd_rn(__fadd_rn(a, b), c));
            y = __fmul_rn(w1, __fadd_rn(a, -c));
            z = __fmul_rn(w2f, __fadd_rn(__fadd_rn(__fmul_rn(.25f, a), -__fmul_rn(.5f, b)), __fmul_rn(.25f, c)));
        } else {    // color.py:31-50
            const float ya = __fmul_rn(w0, a), ub = __fmul_rn(w1, b), vh = __fmul_rn(w2h, c);
            x = __fadd_rn(__fadd_rn(ya, ub), vh);
            y = __fadd_rn(ya, -__fmul_rn(w2, c));
            z = __fadd_rn(__fadd_rn(ya, -ub), vh);
        }
        dst[o] = x;
        dst[o + HW] = y;
        dst[o + 2 * HW] = z;
    }
}

// proc_nl.py:118-125
__global__ void normalize_kernel(float *__restrict__ deno, const float *__restrict__ weights,
                                 const float *__restrict__ fill, int T, int C, long long HW) {
    const long long n = (long long)T * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / HW, r = i - t * HW;
        const float w = weights[i];
        for (int c = 0; c < C; ++c) {
            const long long o = (t * C + c) * HW + r;
            deno[o] = (w != 0.f) ? __fdiv_rn(deno[o], w) : fill[o];
        }
    }
}

// mask.py:315-358 restated as a per-pixel predicate (whole-frame case)
// Tiles (multi-GPU): the buffer holds rows [y_off, y_off + H) of a frame of H_total rows; the lattice phase, the
// "first / last row" rule and the valid range follow the GLOBAL row hi = y_off + local row; y_begin / y_end are local.
__global__ void init_mask_kernel(int8_t *__restrict__ mask, int T, int H, int W, int end_t, int end_h,
                                 int end_w, int step, int y_begin, int y_end, int y_off, int ps) {
    const long long n = (long long)T * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int wi = (int)(i % W);
        const int hl = (int)((i / W) % H);
        const int hi = hl + y_off;
        const int ti = (int)(i / ((long long)W * H));
        bool set = false;
        if (ti < end_t && hi < end_h && wi < end_w && hl >= y_begin && hl < y_end && hl + ps <= H) {
            const bool last_t = ti == end_t - 1;
            const int phase_h = last_t ? 0 : ti;
            const bool last_h = hi == end_h - 1;
            if ((hi % step) == (phase_h % step) || hi == 0 || last_h) {
                const int phase_w = last_h ? 0 : phase_h + hi / step;
                set = (wi % step) == (phase_w % step) || wi == 0 || wi == end_w - 1;
            }
        }
        mask[i] = set ? 1 : 0;
    }
}

// mask.py:37-86,104-187: one block per row of inds
__global__ void mask_update_kernel(int8_t *__restrict__ mask, const long long *__restrict__ inds, int K,
                                   int T, int C, int H, int W, int boost) {
    const long long *row = inds + (long long)blockIdx.x * K;
    if (!row_valid_block(row, K)) return;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        int t, y, x;
        decode_ind(row[i], H, W, C, t, y, x);
        if (t < 0 || t >= T) continue;
        int8_t *m = mask + (long long)t * H * W;
        m[(long long)y * W + x] = 0;
        if (boost) {
            if (x > 0) m[(long long)y * W + x - 1] = 0;
            if (x < W - 1) m[(long long)y * W + x + 1] = 0;
            if (y < H - 1) m[(long long)(y + 1) * W + x] = 0;
            if (y > 0) m[(long long)(y - 1) * W + x] = 0;
        }
    }
}

// ---------------------------------------------------------------------------
// throughput schedule: device-side replacement of mask2inds (mask.py:18-31).
// Every set pixel draws a hash-based uniform number; pixels below `thresh`
// become queries of this round.  counters[0] = set pixels seen, counters[1] =
// queries appended (may exceed cap: extras stay in the mask for a later round).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int hash3(unsigned int a, unsigned int b, unsigned int c) {
    unsigned int h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

__global__ void count_mask_kernel(const int8_t *__restrict__ mask, long long n, unsigned int *__restrict__ counters) {
    unsigned int c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c += mask[i] != 0;
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[0], c);
}

__global__ void select_queries_kernel(int8_t *__restrict__ mask, int T, int H, int W, unsigned int thresh,
                                      unsigned int seed, unsigned int round, long long *__restrict__ qinds, int cap,
                                      unsigned int *__restrict__ counters) {
    const long long n = (long long)T * H * W;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_pad = (n + 31) / 32 * 32;  // keep warps converged for the ballot
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool sel = false;
        if (i < n && mask[i] != 0) sel = hash3((unsigned)i, (unsigned)(i >> 32) + round, seed) <= thresh;
        const unsigned int b = __ballot_sync(0xffffffffu, sel);
        if (b) {
            const int lane = threadIdx.x & 31;
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&counters[1], __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) {
                const unsigned int slot = base + __popc(b & ((1u << lane) - 1));
                if (slot < (unsigned)cap) {
                    const int x = (int)(i % W), y = (int)((i / W) % H), t = (int)(i / ((long long)W * H));
                    qinds[3 * (long long)slot] = t;
                    qinds[3 * (long long)slot + 1] = y;
                    qinds[3 * (long long)slot + 2] = x;
                    mask[i] = 0;  // a drawn pixel is consumed even if its row turns out invalid
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Device-controlled round (vnlb_round_draw): the draw probability is computed ON THE DEVICE from the live count of
// masked pixels, and the round index lives in device memory, so the kernel sequence of a round has no host-computed
// scalar left -- it can be captured once in a CUDA graph and replayed for every round of a step.
// state[0] = rounds drawn so far in this step (the current round is state[0] - 1 after round_begin_kernel).
// ---------------------------------------------------------------------------
__global__ void round_begin_kernel(unsigned int *state, unsigned int *counters) {
    counters[0] = 0u;
    counters[1] = 0u;
    state[0] += 1u;
}

__device__ __forceinline__ unsigned int draw_threshold(unsigned int remaining, float frac, int qmin, int rows) {
    // same rule as the host loop of schedule.py: target = clamp(remaining * frac, qmin, rows'), expected draw 0.97 target
    const int cap_target = max(1, (int)(((long long)rows - 256) * 4 / 5));      // rows = 1.25 target + 256
    const int target = min(cap_target, max(qmin, (int)((float)remaining * frac)));
    if (remaining <= (unsigned int)target) return 0xffffffffu;
    const double prob = (double)target / (double)remaining * 0.97;
    return (unsigned int)(prob * 4294967295.0);
}

__global__ void select_queries_dev_kernel(int8_t *__restrict__ mask, int T, int H, int W, float frac, int qmin, int rows,
                                          unsigned int seed, const unsigned int *__restrict__ state,
                                          long long *__restrict__ qinds, unsigned int *__restrict__ counters) {
    const unsigned int round = state[0] - 1u;
    const unsigned int thresh = draw_threshold(counters[0], frac, qmin, rows);
    const long long n = (long long)T * H * W;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_pad = (n + 31) / 32 * 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool sel = false;
        if (i < n && mask[i] != 0) sel = hash3((unsigned)i, (unsigned)(i >> 32) + round, seed) <= thresh;
        const unsigned int b = __ballot_sync(0xffffffffu, sel);
        if (b) {
            const int lane = threadIdx.x & 31;
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&counters[1], __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) {
                const unsigned int slot = base + __popc(b & ((1u << lane) - 1));
                if (slot < (unsigned)rows) {
                    const int x = (int)(i % W), y = (int)((i / W) % H), t = (int)(i / ((long long)W * H));
                    qinds[3 * (long long)slot] = t;
                    qinds[3 * (long long)slot + 1] = y;
                    qinds[3 * (long long)slot + 2] = x;
                    mask[i] = 0;
                }
            }
        }
    }
}

// rows of qinds beyond the number drawn become invalid queries (t = -1): lets every kernel of a round run
// with a fixed grid of `cap` rows, so the host never has to wait for the round size
__global__ void pad_queries_kernel(long long *__restrict__ qinds, const unsigned int *__restrict__ counters, int cap) {
    const unsigned int n = min(counters[1], (unsigned)cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
        if ((unsigned)i >= n) { qinds[3 * (long long)i] = -1; qinds[3 * (long long)i + 1] = -1; qinds[3 * (long long)i + 2] = -1; }
}

// ---------------------------------------------------------------------------
// throughput schedule: greedy conflict resolution INSIDE a round.  The reference processes its reference pixels
// sequentially in sub-batches of 128 and every sub-batch clears what it found before the next one is drawn
// (search.py:38-64), so a pixel covered by an earlier group is never processed.  A round of this schedule draws
// thousands of pixels at once; without this step two drawn pixels that cover each other are both processed
// (+9..13 % groups, profiles/r1b).  Every row gets a PRIORITY, a 15-bit hash of its reference pixel (the row order
// itself comes from atomics and differs from run to run; the hash keeps the schedule deterministic).  Pass 1: every
// valid row stamps owner[pixel] = min(owner, key(round, priority)) at each pixel its group would clear (found
// patches + the 4 boost neighbours).  Pass 2: a row is dropped (its indices set to -1) when the stamp on its own
// reference pixel carries a SMALLER priority of the same round (equal priorities never drop each other); its
// pixel goes back into the mask, so if the group that covered it is dropped as well it is drawn again later.
// key = (65535 - round) << 15 | priority: later rounds always win the atomicMin over stale stamps, no clearing needed.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int row_priority(const long long *q, int H, int W, unsigned int round) {
    const long long pix = (q[0] * H + q[1]) * W + q[2];
    return hash3((unsigned)pix, (unsigned)(pix >> 32) ^ 0x51ED27u, round) & 0x7fffu;
}

__global__ void round_stamp_kernel(const long long *__restrict__ qinds, const long long *__restrict__ inds, int K,
                                   unsigned int *__restrict__ owner, unsigned int round_key, unsigned int round,
                                   const unsigned int *__restrict__ state, int T, int C, int H, int W, int boost) {
    if (state) { round = state[0] - 1u; round_key = (65535u - round) << 15; }   // device-controlled round
    const long long *row = inds + (long long)blockIdx.x * K;
    if (!row_valid_block(row, K)) return;
    const unsigned int key = round_key | row_priority(qinds + 3 * (long long)blockIdx.x, H, W, round);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        int t, y, x;
        decode_ind(row[i], H, W, C, t, y, x);
        if (t < 0 || t >= T) continue;
        unsigned int *o = owner + (long long)t * H * W;
        atomicMin(o + (long long)y * W + x, key);
        if (boost) {
            if (x > 0) atomicMin(o + (long long)y * W + x - 1, key);
            if (x < W - 1) atomicMin(o + (long long)y * W + x + 1, key);
            if (y < H - 1) atomicMin(o + (long long)(y + 1) * W + x, key);
            if (y > 0) atomicMin(o + (long long)(y - 1) * W + x, key);
        }
    }
}

__global__ void round_drop_kernel(const long long *__restrict__ qinds, long long *__restrict__ inds, int B, int K,
                                  const unsigned int *__restrict__ owner, unsigned int round_key, unsigned int round,
                                  const unsigned int *__restrict__ state, int8_t *__restrict__ mask, int T, int H, int W,
                                  unsigned int *__restrict__ dropped) {
    if (state) { round = state[0] - 1u; round_key = (65535u - round) << 15; }
    const int j = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= B) return;
    const long long *q = qinds + 3 * (long long)j;
    const long long t = q[0], y = q[1], x = q[2];
    if (t < 0 || t >= T || y < 0 || y >= H || x < 0 || x >= W) return;
    const long long pix = (t * H + y) * W + x;
    const unsigned int o = owner[pix];
    if ((o & ~0x7fffu) != round_key || (o & 0x7fffu) >= row_priority(q, H, W, round)) return;   // not covered by a higher-priority row of this round
    long long *row = inds + (long long)j * K;
    if (row[0] < 0) return;                                                     // already invalid
    for (int i = lane; i < K; i += 32) row[i] = -1;
    if (lane == 0) { mask[pix] = 1; atomicAdd(dropped, 1u); }
}

}  // namespace vnlb

using namespace vnlb;

extern "C" int vnlb_round_dedup(const int64_t *qinds, int64_t *inds, int B, int K, uint32_t *owner, uint32_t round,
                                int8_t *mask, int T, int C, int H, int W, int boost, uint32_t *dropped, void *stream) {
    VNLB_REQUIRE(qinds && inds && owner && mask && dropped && B >= 0 && K > 0, "vnlb_round_dedup: bad argument");
    VNLB_REQUIRE(T > 0 && C > 0 && H > 0 && W > 0, "vnlb_round_dedup: bad shape");
    VNLB_REQUIRE(B <= 32768 && round < 65535u, "vnlb_round_dedup: at most 32768 rows per round and 65535 rounds");
    if (B == 0) return VNLB_OK;
    const unsigned int key = (65535u - round) << 15;
    round_stamp_kernel<<<B, 128, 0, (cudaStream_t)stream>>>((const long long *)qinds, (const long long *)inds, K, owner, key,
                                                            round, nullptr, T, C, H, W, boost);
    round_drop_kernel<<<div_up(B, 8), 256, 0, (cudaStream_t)stream>>>((const long long *)qinds, (long long *)inds, B, K,
                                                                     owner, key, round, nullptr, mask, T, H, W, dropped);
    return check_launch("vnlb_round_dedup", 2);
}

extern "C" int vnlb_round_dedup_dev(const int64_t *qinds, int64_t *inds, int B, int K, uint32_t *owner,
                                    const uint32_t *state, int8_t *mask, int T, int C, int H, int W, int boost,
                                    uint32_t *dropped, void *stream) {
    VNLB_REQUIRE(qinds && inds && owner && state && mask && dropped && B >= 0 && K > 0, "vnlb_round_dedup_dev: bad argument");
    VNLB_REQUIRE(T > 0 && C > 0 && H > 0 && W > 0 && B <= 32768, "vnlb_round_dedup_dev: bad shape");
    if (B == 0) return VNLB_OK;
    round_stamp_kernel<<<B, 128, 0, (cudaStream_t)stream>>>((const long long *)qinds, (const long long *)inds, K, owner, 0u,
                                                            0u, state, T, C, H, W, boost);
    round_drop_kernel<<<div_up(B, 8), 256, 0, (cudaStream_t)stream>>>((const long long *)qinds, (long long *)inds, B, K,
                                                                     owner, 0u, 0u, state, mask, T, H, W, dropped);
    return check_launch("vnlb_round_dedup_dev", 2);
}

extern "C" const char *vnlb_last_error(void) { return g_err; }
extern "C" int vnlb_version(void) { return 101; }
namespace vnlb { unsigned long long g_kernel_launches = 0; }
extern "C" unsigned long long vnlb_kernel_launches(void) { return vnlb::g_kernel_launches; }

static int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    long long cap = (long long)num_sms() * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int vnlb_rgb2yuv(const float *rgb, float *yuv, int T, int C, int H, int W, void *stream) {
    VNLB_REQUIRE(rgb && yuv && T > 0 && H > 0 && W > 0, "vnlb_rgb2yuv: bad argument");
    VNLB_REQUIRE(C == 3, "vnlb_rgb2yuv: C must be 3 (got %d)", C);
    const long long HW = (long long)H * W;
    color_kernel<true><<<grid_for(T * HW, 256), 256, 0, (cudaStream_t)stream>>>(rgb, yuv, T, HW);
    return check_launch("vnlb_rgb2yuv");
}

extern "C" int vnlb_yuv2rgb(const float *yuv, float *rgb, int T, int C, int H, int W, void *stream) {
    VNLB_REQUIRE(rgb && yuv && T > 0 && H > 0 && W > 0, "vnlb_yuv2rgb: bad argument");
    VNLB_REQUIRE(C == 3, "vnlb_yuv2rgb: C must be 3 (got %d)", C);
    const long long HW = (long long)H * W;
    color_kernel<false><<<grid_for(T * HW, 256), 256, 0, (cudaStream_t)stream>>>(yuv, rgb, T, HW);
    return check_launch("vnlb_yuv2rgb");
}

extern "C" int vnlb_normalize(float *deno, const float *weights, const float *fill, int T, int C, int H,
                              int W, void *stream) {
    VNLB_REQUIRE(deno && weights && fill && T > 0 && C > 0 && H > 0 && W > 0, "vnlb_normalize: bad argument");
    const long long HW = (long long)H * W;
    normalize_kernel<<<grid_for(T * HW, 256), 256, 0, (cudaStream_t)stream>>>(deno, weights, fill, T, C, HW);
    return check_launch("vnlb_normalize");
}

extern "C" int vnlb_init_mask_tile(int8_t *mask, int T, int H, int W, int ps, int pt, int proc_step, int y_begin,
                                   int y_end, int y_offset, int H_total, void *stream) {
    VNLB_REQUIRE(mask && T > 0 && H > 0 && W > 0, "vnlb_init_mask: bad argument");
    VNLB_REQUIRE(ps >= 1 && pt >= 1 && proc_step >= 1, "vnlb_init_mask: bad patch size / step");
    VNLB_REQUIRE(T >= pt && H >= ps && W >= ps, "vnlb_init_mask: video smaller than one patch");
    VNLB_REQUIRE(y_offset >= 0 && y_offset + H <= H_total, "vnlb_init_mask: tile rows [%d, %d) outside the frame of %d rows",
                 y_offset, y_offset + H, H_total);
    const long long n = (long long)T * H * W;
    init_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(mask, T, H, W, T - pt + 1, H_total - ps + 1,
                                                                        W - ps + 1, proc_step, y_begin, y_end, y_offset, ps);
    return check_launch("vnlb_init_mask");
}

extern "C" int vnlb_init_mask(int8_t *mask, int T, int H, int W, int ps, int pt, int proc_step, int y_begin,
                              int y_end, void *stream) {
    return vnlb_init_mask_tile(mask, T, H, W, ps, pt, proc_step, y_begin, y_end, 0, H, stream);
}

extern "C" int vnlb_count_mask(const int8_t *mask, int T, int H, int W, uint32_t *counters, void *stream) {
    VNLB_REQUIRE(mask && counters && T > 0 && H > 0 && W > 0, "vnlb_count_mask: bad argument");
    const long long n = (long long)T * H * W;
    count_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(mask, n, counters);
    return check_launch("vnlb_count_mask");
}

extern "C" int vnlb_select_queries(int8_t *mask, int T, int H, int W, double prob, uint32_t seed, uint32_t round,
                                   int64_t *qinds, int cap, uint32_t *counters, void *stream) {
    VNLB_REQUIRE(mask && qinds && counters && T > 0 && H > 0 && W > 0 && cap > 0, "vnlb_select_queries: bad argument");
    VNLB_REQUIRE(prob >= 0.0, "vnlb_select_queries: prob must be >= 0");
    const unsigned int thresh = prob >= 1.0 ? 0xffffffffu : (unsigned int)(prob * 4294967295.0);
    const long long n = (long long)T * H * W;
    select_queries_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(mask, T, H, W, thresh, seed, round,
                                                                             (long long *)qinds, cap, counters);
    return check_launch("vnlb_select_queries");
}

extern "C" int vnlb_pad_queries(int64_t *qinds, const uint32_t *counters, int cap, void *stream) {
    VNLB_REQUIRE(qinds && counters && cap > 0, "vnlb_pad_queries: bad argument");
    pad_queries_kernel<<<div_up(cap, 256), 256, 0, (cudaStream_t)stream>>>((long long *)qinds, counters, cap);
    return check_launch("vnlb_pad_queries");
}

extern "C" int vnlb_round_draw(int8_t *mask, int T, int H, int W, double frac, int qmin, int rows, uint32_t seed,
                               uint32_t *state, int64_t *qinds, uint32_t *counters, void *stream) {
    VNLB_REQUIRE(mask && state && qinds && counters && T > 0 && H > 0 && W > 0, "vnlb_round_draw: bad argument");
    VNLB_REQUIRE(frac > 0.0 && qmin >= 1 && rows >= 1, "vnlb_round_draw: frac > 0, qmin >= 1, rows >= 1");
    const long long n = (long long)T * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    round_begin_kernel<<<1, 1, 0, st>>>(state, counters);
    count_mask_kernel<<<grid_for(n, 256), 256, 0, st>>>(mask, n, counters);
    select_queries_dev_kernel<<<grid_for(n, 256), 256, 0, st>>>(mask, T, H, W, (float)frac, qmin, rows, seed, state,
                                                               (long long *)qinds, counters);
    pad_queries_kernel<<<div_up(rows, 256), 256, 0, st>>>((long long *)qinds, counters, rows);
    return check_launch("vnlb_round_draw", 4);
}

extern "C" int vnlb_mask_update(int8_t *mask, const int64_t *inds, int B, int K, int T, int C, int H, int W,
                                int boost, void *stream) {
    VNLB_REQUIRE(mask && inds && B >= 0 && K > 0 && T > 0 && C > 0 && H > 0 && W > 0,
                 "vnlb_mask_update: bad argument");
    if (B == 0) return VNLB_OK;
    mask_update_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(mask, (const long long *)inds, K, T, C, H, W, boost);
    return check_launch("vnlb_mask_update");
}
