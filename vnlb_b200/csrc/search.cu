// search.cu -- non-local similarity search (top-k of squared-L2 patch distances
// over a flow-guided space-time window) and the patch gather.
//
// Replaces vpss.exec_sim_search_burst / vpss.fill_patches
// (reference call sites lib/vnlb/search/search.py:86-98).
//
// Canonical arithmetic (shared bit for bit with oracle/vnlb_oracle.c):
//   dist = 0; for c, ht, hy, hx:  d = q - cand;  dist = fmaf(d, d, dist)
//   top-k ascending by (dist, candidate enumeration order frame -> y -> x)
//
// One CTA per query.  Three distance paths produce identical bits:
//   * generic  : any (ps, pt, w_s), candidates strided over threads, operands
//                read through L1/L2;
//   * tiled    : ps=7, pt=2, w_s=27, any number of frames: frame tiles staged in
//                shared memory with cp.async in chunks of 3 frames, each lane owns
//                a column of 9 candidates, the query plane lives in registers;
//   * quad     : the same shape with at most 13 frames (the production setting):
//                all frames of a phase resident, a lane owns 4 x 9 candidates for
//                the whole query, distances stay in registers (search_quad_kernel).
// Selection: quad kernel -- two histograms over the register-resident distances,
// survivors ranked by counting (select_topk_regs); other kernels and the fallback
// -- sample pivot or 4-pass radix select on the distances in shared memory, tie
// resolution by enumeration order, bitonic sort of the survivors (select_topk).
#include <stdlib.h>

#include "common.cuh"

namespace vnlb {

constexpr int kMaxFrames = 64;      // nWt_f + nWt_b + 1
constexpr int kSearchThreads = 256;

struct FrameWin {  // candidate window of one frame
    int t, x0, y0, nx, ny, off;
};

struct alignas(16) SearchShared {
    FrameWin fw[kMaxFrames];
    int nfr;
    int ncand;
    unsigned int hist[256];
    unsigned int sel_prefix;
    unsigned int sel_kk;
    unsigned int sel_neq;
    int sel_last;
    int sel_count;
    unsigned int red_min[16];
    unsigned int red_max[16];
    unsigned int hist2[256];  // quad kernel: second (linear) histogram of the register-resident selection
    int sel_bin;
    unsigned int sel_less;
    int shared_win;   // quad kernel: every frame has the same window and the frames are consecutive
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ void spatial_range(int c, int L, int ps, int w_s, int mode, int &a0, int &n) {
    const int half = (w_s - 1) / 2;
    int lo, hi;
    if (mode == VNLB_WINDOW_SHIFT) {
        const int shift = min(0, c - half) + max(0, c + half - L + ps);
        lo = max(0, c - half - shift);
        hi = min(L - ps, c + half - shift);
    } else {
        lo = max(0, c - half);
        hi = min(L - ps, c + half);
    }
    a0 = lo;
    n = max(0, hi - lo + 1);
}

// Thread 0: temporal range, flow trajectory, per-frame windows (oracle: temporal_range,
// trajectory, spatial_range).
__device__ void build_windows(SearchShared &S, int t0, int y0, int x0, int T, int H, int W,
                              const float *__restrict__ fflow, const float *__restrict__ bflow,
                              const VnlbSearchParams &p) {
    int r0, r1;
    if (p.window_mode == VNLB_WINDOW_SHIFT) {
        const int shift = min(0, t0 - p.nWt_b) + max(0, t0 + p.nWt_f - T + p.pt);
        r0 = max(0, t0 - p.nWt_b - shift);
        r1 = min(T - p.pt, t0 + p.nWt_f - shift);
    } else {
        r0 = max(0, t0 - p.nWt_b);
        r1 = min(T - p.pt, t0 + p.nWt_f);
    }
    const int nfr = r1 - r0 + 1;
    const long long HW = (long long)H * W;
    S.fw[t0 - r0].x0 = x0;  // temporarily holds the trajectory centre
    S.fw[t0 - r0].y0 = y0;
    for (int qt = t0 + 1; qt <= r1; ++qt) {
        int px = S.fw[qt - 1 - r0].x0, py = S.fw[qt - 1 - r0].y0;
        if (fflow) {
            const float dx = fflow[((long long)(qt - 1) * 2 + 0) * HW + (long long)py * W + px];
            const float dy = fflow[((long long)(qt - 1) * 2 + 1) * HW + (long long)py * W + px];
            px = clampi((int)roundf((float)px + dx), 0, W - 1);
            py = clampi((int)roundf((float)py + dy), 0, H - 1);
        }
        S.fw[qt - r0].x0 = px;
        S.fw[qt - r0].y0 = py;
    }
    for (int qt = t0 - 1; qt >= r0; --qt) {
        int px = S.fw[qt + 1 - r0].x0, py = S.fw[qt + 1 - r0].y0;
        if (bflow) {
            const float dx = bflow[((long long)(qt + 1) * 2 + 0) * HW + (long long)py * W + px];
            const float dy = bflow[((long long)(qt + 1) * 2 + 1) * HW + (long long)py * W + px];
            px = clampi((int)roundf((float)px + dx), 0, W - 1);
            py = clampi((int)roundf((float)py + dy), 0, H - 1);
        }
        S.fw[qt - r0].x0 = px;
        S.fw[qt - r0].y0 = py;
    }
    int off = 0;
    for (int f = 0; f < nfr; ++f) {
        int ax, nx, ay, ny;
        spatial_range(S.fw[f].x0, W, p.ps, p.w_s, p.window_mode, ax, nx);
        spatial_range(S.fw[f].y0, H, p.ps, p.w_s, p.window_mode, ay, ny);
        S.fw[f].t = r0 + f;
        S.fw[f].x0 = ax;
        S.fw[f].y0 = ay;
        S.fw[f].nx = nx;
        S.fw[f].ny = ny;
        S.fw[f].off = off;
        off += nx * ny;
    }
    S.nfr = nfr;
    S.ncand = off;
}

// Warp 0 of the quad kernel (at most 32 frames): lane 0 follows the forward flow, lane 1 the backward flow (two
// independent chains of dependent loads), then lane f computes the window of frame f and a warp scan gives the
// candidate offsets.  Same results as build_windows; also sets S.shared_win.
__device__ void build_windows_warp(SearchShared &S, int t0, int y0, int x0, int T, int H, int W,
                                   const float *__restrict__ fflow, const float *__restrict__ bflow,
                                   const VnlbSearchParams &p) {
    const int lane = threadIdx.x & 31;
    int r0, r1;
    if (p.window_mode == VNLB_WINDOW_SHIFT) {
        const int shift = min(0, t0 - p.nWt_b) + max(0, t0 + p.nWt_f - T + p.pt);
        r0 = max(0, t0 - p.nWt_b - shift);
        r1 = min(T - p.pt, t0 + p.nWt_f - shift);
    } else {
        r0 = max(0, t0 - p.nWt_b);
        r1 = min(T - p.pt, t0 + p.nWt_f);
    }
    const int nfr = r1 - r0 + 1;
    const long long HW = (long long)H * W;
    if (lane == 0) {
        int px = x0, py = y0;
        S.fw[t0 - r0].x0 = px;
        S.fw[t0 - r0].y0 = py;
        for (int qt = t0 + 1; qt <= r1; ++qt) {
            if (fflow) {
                const float dx = fflow[((long long)(qt - 1) * 2 + 0) * HW + (long long)py * W + px];
                const float dy = fflow[((long long)(qt - 1) * 2 + 1) * HW + (long long)py * W + px];
                px = clampi((int)roundf((float)px + dx), 0, W - 1);
                py = clampi((int)roundf((float)py + dy), 0, H - 1);
            }
            S.fw[qt - r0].x0 = px;
            S.fw[qt - r0].y0 = py;
        }
    } else if (lane == 1) {
        int px = x0, py = y0;
        for (int qt = t0 - 1; qt >= r0; --qt) {
            if (bflow) {
                const float dx = bflow[((long long)(qt + 1) * 2 + 0) * HW + (long long)py * W + px];
                const float dy = bflow[((long long)(qt + 1) * 2 + 1) * HW + (long long)py * W + px];
                px = clampi((int)roundf((float)px + dx), 0, W - 1);
                py = clampi((int)roundf((float)py + dy), 0, H - 1);
            }
            S.fw[qt - r0].x0 = px;
            S.fw[qt - r0].y0 = py;
        }
    }
    __syncwarp();
    int ax = 0, nx = 0, ay = 0, ny = 0;
    if (lane < nfr) {
        spatial_range(S.fw[lane].x0, W, p.ps, p.w_s, p.window_mode, ax, nx);
        spatial_range(S.fw[lane].y0, H, p.ps, p.w_s, p.window_mode, ay, ny);
    }
    const int cnt = lane < nfr ? nx * ny : 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    __syncwarp();
    if (lane < nfr) {
        FrameWin w;
        w.t = r0 + lane; w.x0 = ax; w.y0 = ay; w.nx = nx; w.ny = ny; w.off = incl - cnt;
        S.fw[lane] = w;
    }
    const int ax0 = __shfl_sync(0xffffffffu, ax, 0), ay0 = __shfl_sync(0xffffffffu, ay, 0);
    const int nx0 = __shfl_sync(0xffffffffu, nx, 0), ny0 = __shfl_sync(0xffffffffu, ny, 0);
    const bool same = lane >= nfr || (ax == ax0 && ay == ay0 && nx == nx0 && ny == ny0);
    const bool sw = __all_sync(0xffffffffu, same);
    if (lane == nfr - 1) S.ncand = incl;
    if (lane == 0) { S.nfr = nfr; S.shared_win = sw ? 1 : 0; }
}

// candidate enumeration order -> (frame slot, qy, qx)
__device__ __forceinline__ void cand_coords(const SearchShared &S, int cand, int &f, int &qy, int &qx) {
    f = 0;
    while (f + 1 < S.nfr && S.fw[f + 1].off <= cand) ++f;
    const int local = cand - S.fw[f].off;
    const int r = local / S.fw[f].nx;
    qy = S.fw[f].y0 + r;
    qx = S.fw[f].x0 + (local - r * S.fw[f].nx);
}

// ---------------------------------------------------------------------------
// selection: the m = min(k, ncand) smallest (dist, order) pairs, sorted ascending.
// dist[] (shared) holds the distance of every candidate in enumeration order;
// keys[] (shared, kKeyCap 64-bit entries) is scratch; tmin/tmax are this
// thread's smallest/largest distance bits (any subset; reduced here).
//
// Fast path: one histogram over 256 bins spread across the ACTUAL range of the
// distance bits (little atomic contention) finds the bin holding the m-th
// smallest; everything up to and including that bin (usually 100-300 keys) is
// bitonic-sorted on (distance bits, enumeration order).  If that set exceeds
// kKeyCap (heavy ties, e.g. a constant image) the exact 4-pass radix select
// with ballot-scan tie resolution takes over.  Both give the same answer.
// ---------------------------------------------------------------------------
constexpr int kKeyCap = 512;
constexpr int kRankCap = 160;   // quad kernel: survivors ranked by counting instead of sorted

__device__ void bitonic_sort_keys(unsigned long long *keys, int P) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P / 2; i += nthr) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
}

// rows of the output: the m smallest keys (sorted) as (distance, index); the rest of the k slots invalid
__device__ void write_topk(const SearchShared &S, const unsigned long long *keys, int m, int k, float *__restrict__ out_vals,
                           long long *__restrict__ out_inds, int C, int H, int W) {
    const long long CHW = (long long)C * H * W;
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
        if (r < m) {
            const unsigned long long key = keys[r];
            int f, qy, qx;
            cand_coords(S, (int)(key & 0xffffffffu), f, qy, qx);
            out_vals[r] = __uint_as_float((unsigned)(key >> 32));
            out_inds[r] = (long long)S.fw[f].t * CHW + (long long)qy * W + qx;
        } else {
            out_vals[r] = __int_as_float(0x7f800000);
            out_inds[r] = -1;
        }
    }
}

__device__ void select_topk(SearchShared &S, const float *dist, unsigned long long *keys, int k, unsigned int tmin,
                            unsigned int tmax, float *__restrict__ out_vals, long long *__restrict__ out_inds, int C,
                            int H, int W) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int ncand = S.ncand;
    const int m = min(k, ncand);
    int P = 1;
    while (P < m) P <<= 1;
    if (m > 0) {
        // ---- fast path: pivot from a sorted sample of kKeyCap candidates ----
        if (tid == 0) S.sel_count = 0;
        bool fast = false;
        unsigned long long pivot = ~0ull;
        if (ncand <= kKeyCap) {                  // few candidates: sort them all
            for (int i = tid; i < ncand; i += nthr) keys[i] = ((unsigned long long)__float_as_uint(dist[i]) << 32) | (unsigned)i;
            P = 1;
            while (P < ncand) P <<= 1;
            for (int i = ncand + tid; i < P; i += nthr) keys[i] = ~0ull;
            __syncthreads();
            fast = true;
        } else {
            for (int j = tid; j < kKeyCap; j += nthr) {
                const int i = (int)(((long long)j * ncand) / kKeyCap);
                keys[j] = ((unsigned long long)__float_as_uint(dist[i]) << 32) | (unsigned)i;
            }
            __syncthreads();
            bitonic_sort_keys(keys, kKeyCap);
            // expected #candidates <= pivot: ncand*(rk+1)/kKeyCap ~ 2 m + margin, relative spread ~ 1/sqrt(rk)
            const int rk = min(kKeyCap - 1, (int)((2.0f * (float)m * (float)kKeyCap) / (float)ncand) + 6);
            pivot = keys[rk];
            __syncthreads();
            for (int i = tid; i < ncand; i += nthr) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(dist[i]) << 32) | (unsigned)i;
                if (key <= pivot) {
                    const int slot = atomicAdd(&S.sel_count, 1);
                    if (slot < kKeyCap) keys[slot] = key;
                }
            }
            __syncthreads();
            const int total = S.sel_count;
            if (total >= m && total <= kKeyCap) {
                P = 1;
                while (P < total) P <<= 1;
                for (int i = total + tid; i < P; i += nthr) keys[i] = ~0ull;
                fast = true;
            }
            __syncthreads();
        }
        (void)tmin; (void)tmax;
        if (fast) {
            // keys[0..P) hold a superset of the m smallest; sorted below
        } else {
            // ---- exact 4-pass radix select on the distance bits ----
            if (tid == 0) { S.sel_prefix = 0u; S.sel_kk = (unsigned)m; S.sel_count = 0; }
            P = 1;
            while (P < m) P <<= 1;
            unsigned int known = 0u;
            for (int shift = 24; shift >= 0; shift -= 8) {
                __syncthreads();
                for (int i = tid; i < 256; i += nthr) S.hist[i] = 0u;
                __syncthreads();
                const unsigned int prefix = S.sel_prefix;
                for (int i = tid; i < ncand; i += nthr) {
                    const unsigned int u = __float_as_uint(dist[i]);
                    if ((u & known) == prefix) atomicAdd(&S.hist[(u >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (tid < 32) {
                    unsigned int loc[8], s = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { loc[j] = S.hist[tid * 8 + j]; s += loc[j]; }
                    unsigned int incl = s;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (tid >= d) incl += v;
                    }
                    const unsigned int excl = incl - s;
                    const unsigned int kk = S.sel_kk;
                    __syncwarp();
                    if (kk > excl && kk <= incl) {  // exactly one lane
                        unsigned int run = excl;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (kk > run && kk <= run + loc[j]) {
                                S.sel_prefix = prefix | ((unsigned)(tid * 8 + j) << shift);
                                S.sel_kk = kk - run;
                                S.sel_neq = loc[j];
                            }
                            run += loc[j];
                        }
                    }
                }
                known |= 255u << shift;
            }
            __syncthreads();
            const unsigned int vk = S.sel_prefix;  // bits of the m-th smallest distance
            // among the sel_neq candidates equal to vk keep the sel_kk first in enumeration order
            if (tid < 32) {
                int last = ncand - 1;
                if (S.sel_neq > S.sel_kk) {
                    unsigned int need = S.sel_kk;
                    for (int base = 0; base < ncand; base += 32) {
                        const int i = base + tid;
                        const bool eq = i < ncand && __float_as_uint(dist[i]) == vk;
                        const unsigned int bb = __ballot_sync(0xffffffffu, eq);
                        const unsigned int cnt = __popc(bb);
                        if (cnt >= need) {  // the need-th set bit of bb
                            unsigned int b2 = bb;
                            for (unsigned int j = 1; j < need; ++j) b2 &= b2 - 1;
                            last = base + __ffs(b2) - 1;
                            break;
                        }
                        need -= cnt;
                    }
                }
                if (tid == 0) S.sel_last = last;
            }
            __syncthreads();
            const int last = S.sel_last;
            for (int i = tid; i < ncand; i += nthr) {
                const unsigned int u = __float_as_uint(dist[i]);
                if (u < vk || (u == vk && i <= last)) {
                    const int slot = atomicAdd(&S.sel_count, 1);
                    keys[slot] = ((unsigned long long)u << 32) | (unsigned)i;
                }
            }
            for (int i = m + tid; i < P; i += nthr) keys[i] = ~0ull;
            __syncthreads();
        }
        bitonic_sort_keys(keys, P);
    }
    write_topk(S, keys, m, k, out_vals, out_inds, C, H, W);
}

// ---------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSearchThreads)
search_generic_kernel(const float *__restrict__ img, int T, int C, int H, int W,
                      const long long *__restrict__ qinds, const float *__restrict__ fflow,
                      const float *__restrict__ bflow, VnlbSearchParams p, int P, int ncand_max,
                      float *__restrict__ vals, long long *__restrict__ inds) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SearchShared &S = *reinterpret_cast<SearchShared *>(smem_raw);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(SearchShared));
    float *dist = reinterpret_cast<float *>(keys + P);
    float *qpatch = dist + ncand_max;
    unsigned int tmin = 0xffffffffu, tmax = 0u;

    const int q = blockIdx.x;
    const int t0 = (int)qinds[3 * q], y0 = (int)qinds[3 * q + 1], x0 = (int)qinds[3 * q + 2];
    float *ov = vals + (long long)q * p.k;
    long long *oi = inds + (long long)q * p.k;
    const bool ok = t0 >= 0 && t0 <= T - p.pt && y0 >= 0 && y0 <= H - p.ps && x0 >= 0 && x0 <= W - p.ps;
    if (!ok) {  // malformed query: row stays invalid
        for (int r = threadIdx.x; r < p.k; r += blockDim.x) {
            ov[r] = __int_as_float(0x7f800000);
            oi[r] = -1;
        }
        return;
    }
    if (threadIdx.x == 0) build_windows(S, t0, y0, x0, T, H, W, fflow, bflow, p);
    const long long HW = (long long)H * W, CHW = (long long)C * HW;
    const int ps = p.ps, pt = p.pt, dc = p.dist_chnls;
    const int pdim = dc * pt * ps * ps;
    for (int i = threadIdx.x; i < pdim; i += blockDim.x) {
        const int hx = i % ps, hy = (i / ps) % ps, ht = (i / (ps * ps)) % pt, c = i / (ps * ps * pt);
        qpatch[i] = img[(long long)(t0 + ht) * CHW + c * HW + (long long)(y0 + hy) * W + x0 + hx];
    }
    __syncthreads();
    const int ncand = S.ncand;
    for (int cand = threadIdx.x; cand < ncand; cand += blockDim.x) {
        int f, qy, qx;
        cand_coords(S, cand, f, qy, qx);
        const int qt = S.fw[f].t;
        float d2 = 0.f;
        const float *qp = qpatch;
        for (int c = 0; c < dc; ++c)
            for (int ht = 0; ht < pt; ++ht) {
                const float *b = img + (long long)(qt + ht) * CHW + c * HW + (long long)qy * W + qx;
                for (int hy = 0; hy < ps; ++hy) {
                    const float *br = b + (long long)hy * W;
                    for (int hx = 0; hx < ps; ++hx) {
                        const float d = __fsub_rn(*qp++, __ldg(br + hx));
                        d2 = __fmaf_rn(d, d, d2);
                    }
                }
            }
        dist[cand] = d2;
        tmin = min(tmin, __float_as_uint(d2));
        tmax = max(tmax, __float_as_uint(d2));
    }
    __syncthreads();
    select_topk(S, dist, keys, p.k, tmin, tmax, ov, oi, C, H, W);
}

// ---------------------------------------------------------------------------
// tiled kernel: ps = 7, pt = 2, w_s = 27.
//
// Work item = (frame slot f, strip s of 9 candidate rows, candidate column qx):
// 81 items per frame; the items of a chunk of 3 frames (243) map 1:1 onto the
// first 243 threads.  A thread keeps the 49-value query plane (one channel, one
// patch frame) in registers and slides down 15 tile rows, feeding 9 candidate
// accumulators; the per-candidate accumulation order is exactly the canonical
// (c, ht, hy, hx) one.  Tiles are 33x33 windows staged with cp.async at row
// pitch 35 and slot pitch 1169 words, which makes the flattened item -> lane
// map bank-conflict free (315*s + qx + 1169*f == item (mod 32)).
// ---------------------------------------------------------------------------
constexpr int TPS = 7, TPT = 2, TWS = 27;
constexpr int TTILE = TWS + TPS - 1;       // 33
constexpr int TPITCH = 35;
constexpr int TSLOT = 1169;                // >= 33*35 = 1155, == 17 (mod 32)
constexpr int TCHUNK = 3;                  // frames per chunk
constexpr int TSTRIP = 9;                  // candidate rows per item

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__global__ void __launch_bounds__(kSearchThreads, 3)
search_tiled_kernel(const float *__restrict__ img, int T, int C, int H, int W,
                    const long long *__restrict__ qinds, const float *__restrict__ fflow,
                    const float *__restrict__ bflow, VnlbSearchParams p, int P, int ncand_max,
                    float *__restrict__ vals, long long *__restrict__ inds) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SearchShared &S = *reinterpret_cast<SearchShared *>(smem_raw);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(SearchShared));
    float *dist = reinterpret_cast<float *>(keys + P);
    float *qpatch = dist + ncand_max;                    // [TPT][49] of the current channel (+pad)
    unsigned int tmin = 0xffffffffu, tmax = 0u;
    float *tiles = qpatch + 2 * 52;                      // [TPT][TCHUNK][TSLOT]

    const int q = blockIdx.x;
    const int t0 = (int)qinds[3 * q], y0 = (int)qinds[3 * q + 1], x0 = (int)qinds[3 * q + 2];
    float *ov = vals + (long long)q * p.k;
    long long *oi = inds + (long long)q * p.k;
    const bool ok = t0 >= 0 && t0 <= T - TPT && y0 >= 0 && y0 <= H - TPS && x0 >= 0 && x0 <= W - TPS;
    if (!ok) {
        for (int r = threadIdx.x; r < p.k; r += blockDim.x) {
            ov[r] = __int_as_float(0x7f800000);
            oi[r] = -1;
        }
        return;
    }
    if (threadIdx.x == 0) build_windows(S, t0, y0, x0, T, H, W, fflow, bflow, p);
    __syncthreads();
    const long long HW = (long long)H * W, CHW = (long long)C * HW;
    const int nfr = S.nfr, dc = p.dist_chnls;
    const int tid = threadIdx.x;
    // this thread's item inside a chunk
    const int it_f = tid / 81, it_r = tid - it_f * 81;
    const int it_s = it_r / TWS, it_x = it_r - it_s * TWS;
    const bool has_item = tid < TCHUNK * 81;

    for (int f0 = 0; f0 < nfr; f0 += TCHUNK) {
        const int nf = min(TCHUNK, nfr - f0);
        float acc[TSTRIP];
#pragma unroll
        for (int s = 0; s < TSTRIP; ++s) acc[s] = 0.f;
        const bool active = has_item && it_f < nf;
        bool shared_win = true;   // all frames of the chunk share one window and are consecutive
        {
            const FrameWin w0 = S.fw[f0];
            for (int fl = 1; fl < nf; ++fl) {
                const FrameWin w = S.fw[f0 + fl];
                shared_win = shared_win && w.x0 == w0.x0 && w.y0 == w0.y0 && w.nx == w0.nx && w.ny == w0.ny && w.t == w0.t + fl;
            }
        }
        for (int c = 0; c < dc; ++c) {
            __syncthreads();  // previous phase done with tiles / qpatch
            // stage the query planes of channel c
            for (int i = tid; i < TPT * 49; i += blockDim.x) {
                const int ht = i / 49, r = i - ht * 49, hy = r / 7, hx = r - hy * 7;
                qpatch[ht * 52 + r] = img[(long long)(t0 + ht) * CHW + c * HW + (long long)(y0 + hy) * W + x0 + hx];
            }
            // stage the tiles.  Zero flow (or a rigid trajectory): every frame of the chunk has the same
            // window, so slot s <- frame t0c + s and the item (frame fl, plane ht) reads slot fl + ht
            // (nf + 1 tiles instead of 2 nf).  Otherwise slot (ht, fl) <- frame fw[f0+fl].t + ht at the
            // window of fw[f0+fl].  One warp per tile row, one lane per column.
            {
                const int lane = tid & 31, wrp = tid >> 5;
                const int nslots = shared_win ? nf + 1 : nf * TPT;
                for (int sl = 0; sl < nslots; ++sl) {
                    const int fl = shared_win ? 0 : (sl >> 1), ht = shared_win ? sl : (sl & 1);
                    const FrameWin w = S.fw[f0 + fl];
                    const int rows = w.ny + TPS - 1, cols = w.nx + TPS - 1;
                    const float *src = img + (long long)(w.t + ht) * CHW + c * HW + (long long)(w.y0 + wrp) * W + w.x0;
                    float *dst = tiles + (shared_win ? sl : (sl & 1) * TCHUNK + fl) * TSLOT + wrp * TPITCH;
                    for (int r = wrp; r < rows; r += kSearchThreads / 32) {
                        if (lane < cols) cp_async4(dst + lane, src + lane);
                        if (lane == 0 && cols > 32) cp_async4(dst + 32, src + 32);
                        src += (long long)(kSearchThreads / 32) * W;
                        dst += (kSearchThreads / 32) * TPITCH;
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
            if (active) {
                const FrameWin w = S.fw[f0 + it_f];
                if (it_x < w.nx && it_s * TSTRIP < w.ny) {
#pragma unroll 1
                    for (int ht = 0; ht < TPT; ++ht) {
                        float qv[49];
#pragma unroll
                        for (int i = 0; i < 49; ++i) qv[i] = qpatch[ht * 52 + i];
                        const float *tp = tiles + (shared_win ? it_f + ht : ht * TCHUNK + it_f) * TSLOT + (it_s * TSTRIP) * TPITCH + it_x;
#pragma unroll
                        for (int rho = 0; rho < TSTRIP + TPS - 1; ++rho) {
                            float v[7];
#pragma unroll
                            for (int hx = 0; hx < 7; ++hx) v[hx] = tp[rho * TPITCH + hx];
#pragma unroll
                            for (int s = 0; s < TSTRIP; ++s) {
                                const int hy = rho - s;  // compile-time after unrolling
                                if (hy >= 0 && hy < TPS) {
#pragma unroll
                                    for (int hx = 0; hx < 7; ++hx) {
                                        const float d = __fsub_rn(qv[hy * 7 + hx], v[hx]);
                                        acc[s] = __fmaf_rn(d, d, acc[s]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (active) {
            const FrameWin w = S.fw[f0 + it_f];
            if (it_x < w.nx) {
#pragma unroll
                for (int s = 0; s < TSTRIP; ++s) {
                    const int r = it_s * TSTRIP + s;
                    if (r < w.ny) {
                        dist[w.off + r * w.nx + it_x] = acc[s];
                        tmin = min(tmin, __float_as_uint(acc[s]));
                        tmax = max(tmax, __float_as_uint(acc[s]));
                    }
                }
            }
        }
    }
    __syncthreads();
    select_topk(S, dist, keys, p.k, tmin, tmax, ov, oi, C, H, W);
}

// ---------------------------------------------------------------------------
// quad kernel: ps = 7, pt = 2, w_s = 27, at most 13 frames in the temporal window (the production shape).
//
// Work item = (frame f, strip s of 9 candidate rows, quad j of 4 candidate columns): 21 items per frame, all frames of
// the temporal window at once -> 273 of 288 threads busy for 13 frames (the CTA is launched with 21 * frames threads,
// rounded up to whole warps).  A thread owns its 36 candidates for the WHOLE query: the accumulators stay in registers
// over the (channel, patch frame) phases, so the distances never round-trip through shared memory between phases and
// the per-candidate accumulation order is exactly the canonical (c, ht, hy, hx) one.  Per phase the thread keeps the
// 49-value query plane in registers (13 broadcast LDS.128) and slides down 15 tile rows with 2 LDS.128 + 1 LDS.64 each,
// feeding 4 x 9 accumulators: 58 shared-memory loads per 3528 FSUB/FFMA (the 1-column kernel above: 154 per 882).
// Tiles are 36 x 33 boxes (row pitch 36 floats, 16-byte aligned quads), one per frame, ALL frames of a phase resident,
// staged with 4-byte cp.async (one warp per tile, lane = column).  TMA (cp.async.bulk.tensor) was built and measured
// first: the instruction needs the box START 16-byte aligned in global memory (x0 % 4 == 0 for float32; any other
// corner raises an illegal-instruction fault -- tools/experimental/tma_probe.cu), and a window corner is arbitrary;
// aligning the box instead costs an eighth quad of candidate columns per strip (+11 % FSUB/FFMA issue), more than
// the 4-byte staging (5 % of the issue slots).
// Zero flow / rigid trajectory (all frames share one window): frames+1 tiles are staged once per channel and serve
// both patch frames.  The distances are then written to shared memory (aliasing the tiles) and selected as above.
// ---------------------------------------------------------------------------
constexpr int QROWS = 9, QCOLS = 4, QQUADS = 7, QSTRIPS = 3;
constexpr int QITEMS = QQUADS * QSTRIPS;   // 21 items per frame
constexpr int QTW = 36;                    // tile row pitch (floats) = TMA box width
constexpr int QTH = TTILE;                 // 33 tile rows
constexpr int QSLOT = 1204;                // floats per tile slot: 33 * 36 = 1188, padded to 301 x 16 bytes (301 = 5 mod 8: with the
                                           // frame as the fastest item index, 8 consecutive lanes read 8 different 16-byte bank groups)
constexpr int QMAXF = 13;                  // frames of the temporal window
constexpr int QTHREADS = ((QMAXF * QITEMS + 31) / 32) * 32;   // 288

// one (channel, patch frame) plane of one item: 36 candidates x 49 terms.  The loop over the patch row hy is NOT
// unrolled: its body (9 tile rows x 56 FSUB/FFMA pairs + 29 shared-memory loads, 8.5 KB of code) stays in the
// instruction cache; the fully unrolled plane (56 KB) ran at 54 % issue utilisation with `no_instruction` as its top
// stall reason (profiles/r2_search_summary.md).  qp: the query plane as [7][8] floats.
__device__ __forceinline__ void quad_row(const float *__restrict__ row, float (&v)[10]) {
    const float4 a = *reinterpret_cast<const float4 *>(row);
    const float4 b = *reinterpret_cast<const float4 *>(row + 4);
    const float2 e = *reinterpret_cast<const float2 *>(row + 8);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = e.x; v[9] = e.y;
}
__device__ __forceinline__ void quad_terms(const float (&q)[7], const float (&v)[10], float *acc4) {
#pragma unroll
    for (int hx = 0; hx < TPS; ++hx)
#pragma unroll
        for (int col = 0; col < QCOLS; ++col) {
            const float d = __fsub_rn(q[hx], v[col + hx]);
            acc4[col] = __fmaf_rn(d, d, acc4[col]);
        }
}
__device__ __forceinline__ void quad_qrow(const float *__restrict__ qp, float (&q)[7]) {
    const float4 q0 = *reinterpret_cast<const float4 *>(qp);
    const float4 q1 = *reinterpret_cast<const float4 *>(qp + 4);
    q[0] = q0.x; q[1] = q0.y; q[2] = q0.z; q[3] = q0.w; q[4] = q1.x; q[5] = q1.y; q[6] = q1.z;
}
__device__ __forceinline__ void quad_plane(const float *__restrict__ qp, const float *__restrict__ tp, float (&acc)[QROWS * QCOLS]) {
    // patch rows in blocks of two: tile row hy + rr serves candidate row rr (patch row hy) and candidate row rr - 1
    // (patch row hy + 1), so a block loads 10 tile rows instead of 18; each candidate still sees hy ascending
#pragma unroll 1
    for (int hy = 0; hy < TPS - 1; hy += 2) {
        float qa[7], qb[7];
        quad_qrow(qp + hy * 8, qa);
        quad_qrow(qp + hy * 8 + 8, qb);
        const float *row = tp + hy * QTW;
#pragma unroll
        for (int rr = 0; rr < QROWS + 1; ++rr) {
            float v[10];
            quad_row(row + rr * QTW, v);
            if (rr > 0) quad_terms(qb, v, &acc[(rr - 1) * QCOLS]);
            if (rr < QROWS) quad_terms(qa, v, &acc[rr * QCOLS]);
        }
    }
    {
        float qa[7];
        quad_qrow(qp + (TPS - 1) * 8, qa);
        const float *row = tp + (TPS - 1) * QTW;
#pragma unroll
        for (int s = 0; s < QROWS; ++s) {
            float v[10];
            quad_row(row + s * QTW, v);
            quad_terms(qa, v, &acc[s * QCOLS]);
        }
    }
}

// ---------------------------------------------------------------------------
// Selection with the distances in registers (quad kernel): two histograms instead of two sorts.
//   1. block min / max of the distance bits;
//   2. histogram of a lattice SAMPLE (two candidates per thread) over 256 bins of the bit pattern (monotone in the
//      distance, logarithmic resolution): bin b1 below which about 2 m candidates fall;
//   3. the candidates up to bin b1 (a lower set, a few hundred) are appended to a list in shared memory;
//   4. histogram of the LIST over 256 bins linear in the distance: bin b2 at which the cumulated count reaches m;
//      the entries up to b2 (a lower set of m + a few) are bitonic-sorted on (distance bits, enumeration order).
// Both filters are monotone in the distance, so the sorted set contains every candidate smaller than any excluded
// one and all of its ties: the result is exactly the reference order.  Too few / too many survivors (heavy ties, tiny
// windows) fall back to select_topk on the distances written to shared memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_find_bin(const unsigned int *hist, unsigned int need, unsigned int &cum_out) {
    // warp 0: smallest bin whose inclusive cumulated count reaches `need` (255 and the total if it never does)
    const int lane = threadIdx.x & 31;
    unsigned int loc[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; sum += loc[j]; }
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const unsigned int excl = incl - sum;
    int bin = 255 + 256;                       // "not found" sentinel, larger than any bin
    unsigned int cum = 0;
    if (need > excl && need <= incl) {
        unsigned int run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (need > run && need <= run + loc[j]) { bin = lane * 8 + j; cum = run + loc[j]; }
            run += loc[j];
        }
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    const int best = __reduce_min_sync(0xffffffffu, bin);
    const unsigned int bcum = __reduce_max_sync(0xffffffffu, cum);
    if (best > 255) { cum_out = total; return 255; }
    cum_out = bcum;
    return best;
}

// returns false (block-uniform) if the fallback has to run; keys / list: kKeyCap 64-bit slots each
__device__ bool select_topk_regs(SearchShared &S, const float (&acc)[QROWS * QCOLS], bool active, int it_f, int it_s, int it_j,
                                 unsigned long long *keys, unsigned long long *list, int k, float *__restrict__ out_vals,
                                 long long *__restrict__ out_inds, int C, int H, int W) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wrp = tid >> 5, nw = nthr >> 5;
    const int ncand = S.ncand;
    const int m = min(k, ncand);
    if (ncand <= kKeyCap) return false;        // tiny windows: the fallback sorts everything
    FrameWin w;
    w.nx = 0; w.ny = 0; w.off = 0;
    if (active) w = S.fw[it_f];
    const int r0 = it_s * QROWS, x0 = QCOLS * it_j;
    // ---- 1. block min / max
    unsigned int lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int s = 0; s < QROWS; ++s)
#pragma unroll
        for (int col = 0; col < QCOLS; ++col)
            if (r0 + s < w.ny && x0 + col < w.nx) {
                const unsigned int u = __float_as_uint(acc[s * QCOLS + col]);
                if (u) lo = min(lo, u);    // the query itself (distance 0) would stretch the logarithmic bins to one per octave
                hi = max(hi, u);
            }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { S.red_min[wrp] = lo; S.red_max[wrp] = hi; }
    for (int i = tid; i < 256; i += nthr) { S.hist[i] = 0u; S.hist2[i] = 0u; }
    if (tid == 0) { S.sel_count = 0; S.sel_neq = 0u; S.sel_kk = 0u; }
    __syncthreads();
    lo = 0xffffffffu; hi = 0u;
    for (int i = 0; i < nw; ++i) { lo = min(lo, S.red_min[i]); hi = max(hi, S.red_max[i]); }
    if (lo > hi) lo = hi;                      // every distance is zero
    const unsigned int range = hi - lo;
    const int sh = range >= 256u ? (32 - __clz(range) - 8) : 0;
    // ---- 2. sample histogram (bins of the bit pattern)
    {
        unsigned int ns = 0;
        if (r0 + 2 < w.ny && x0 + 1 < w.nx) { const unsigned int u = __float_as_uint(acc[2 * QCOLS + 1]); atomicAdd(&S.hist[u <= lo ? 0u : (u - lo) >> sh], 1u); ++ns; }
        if (r0 + 6 < w.ny && x0 + 2 < w.nx) { const unsigned int u = __float_as_uint(acc[6 * QCOLS + 2]); atomicAdd(&S.hist[u <= lo ? 0u : (u - lo) >> sh], 1u); ++ns; }
        ns = __reduce_add_sync(0xffffffffu, ns);
        if (lane == 0 && ns) atomicAdd(&S.sel_neq, ns);
    }
    __syncthreads();
    if (wrp == 0) {
        const unsigned int nsamp = S.sel_neq;
        unsigned int need = (unsigned int)((2.0f * (float)m * (float)nsamp) / (float)ncand) + 6u;
        unsigned int cum;
        const int b1 = (nsamp == 0u || need >= nsamp) ? 255 : warp_find_bin(S.hist, need, cum);
        if (lane == 0) S.sel_bin = b1;
    }
    __syncthreads();
    const unsigned int b1 = (unsigned int)S.sel_bin;
    // ---- 3. list of the candidates up to bin b1: count, one atomic per warp for the slots, then write
    {
        int cnt = 0;
#pragma unroll
        for (int s = 0; s < QROWS; ++s)
#pragma unroll
            for (int col = 0; col < QCOLS; ++col)
                if (r0 + s < w.ny && x0 + col < w.nx) {
                    const unsigned int u = __float_as_uint(acc[s * QCOLS + col]);
                    cnt += (u <= lo || ((u - lo) >> sh) <= b1) ? 1 : 0;
                }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        int base = 0;
        if (lane == 31 && incl > 0) base = atomicAdd(&S.sel_count, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - cnt;
        if (cnt > 0) {
#pragma unroll
            for (int s = 0; s < QROWS; ++s)
#pragma unroll
                for (int col = 0; col < QCOLS; ++col)
                    if (r0 + s < w.ny && x0 + col < w.nx) {
                        const unsigned int u = __float_as_uint(acc[s * QCOLS + col]);
                        if (u <= lo || ((u - lo) >> sh) <= b1) {
                            if (pos < kKeyCap)
                                list[pos] = ((unsigned long long)u << 32) | (unsigned int)(w.off + (r0 + s) * w.nx + x0 + col);
                            ++pos;
                        }
                    }
        }
    }
    __syncthreads();
    const int nlist = S.sel_count;
    if (nlist < m || nlist > kKeyCap) return false;
    // ---- 4. linear histogram of the list, entries up to the bin that reaches m
    unsigned int upper = (b1 >= 255u) ? hi + 1u : min(hi + 1u, lo + ((b1 + 1u) << sh));
    const float dmin = __uint_as_float(lo);
    const float scale = 256.f / (__uint_as_float(upper) - dmin);
    for (int i = tid; i < nlist; i += nthr) {
        const float d = __uint_as_float((unsigned int)(list[i] >> 32));
        atomicAdd(&S.hist2[min(255, max(0, (int)((d - dmin) * scale)))], 1u);
    }
    __syncthreads();
    if (wrp == 0) {
        unsigned int cum;
        const int b2 = warp_find_bin(S.hist2, (unsigned int)m, cum);
        if (lane == 0) { S.sel_last = b2; S.sel_less = cum; }
    }
    __syncthreads();
    const int b2 = S.sel_last;
    const int n2 = (int)S.sel_less;            // entries with bin <= b2: m <= n2 <= nlist <= kKeyCap
    for (int i = tid; i < nlist; i += nthr) {
        const unsigned long long ent = list[i];
        const float d = __uint_as_float((unsigned int)(ent >> 32));
        if (min(255, max(0, (int)((d - dmin) * scale))) <= b2) keys[atomicAdd(&S.sel_kk, 1u)] = ent;
    }
    __syncthreads();
    if (n2 <= kRankCap) {
        // few survivors: the rank of a key = number of smaller keys (keys are distinct: they carry the enumeration
        // index), one barrier instead of the 28+ of a bitonic sort; the owner of rank r writes output row r
        const long long CHW = (long long)C * H * W;
        for (int i = tid; i < n2; i += nthr) {
            const unsigned long long mine = keys[i];
            int rank = 0;
            for (int j = 0; j < n2; ++j) rank += keys[j] < mine;
            if (rank < m) {
                int f, qy, qx;
                cand_coords(S, (int)(mine & 0xffffffffu), f, qy, qx);
                out_vals[rank] = __uint_as_float((unsigned)(mine >> 32));
                out_inds[rank] = (long long)S.fw[f].t * CHW + (long long)qy * W + qx;
            }
        }
        for (int r = m + tid; r < k; r += nthr) {
            out_vals[r] = __int_as_float(0x7f800000);
            out_inds[r] = -1;
        }
        return true;
    }
    int P = 1;
    while (P < n2) P <<= 1;
    for (int i = n2 + tid; i < P; i += nthr) keys[i] = ~0ull;
    __syncthreads();
    bitonic_sort_keys(keys, P);
    write_topk(S, keys, m, k, out_vals, out_inds, C, H, W);
    return true;
}

__global__ void __launch_bounds__(QTHREADS, 2)
search_quad_kernel(const float *__restrict__ img, int T, int C, int H, int W,
                   const long long *__restrict__ qinds, const float *__restrict__ fflow,
                   const float *__restrict__ bflow, VnlbSearchParams p, int P, int ncand_max,
                   float *__restrict__ vals, long long *__restrict__ inds) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SearchShared &S = *reinterpret_cast<SearchShared *>(smem_raw);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(SearchShared));
    float *qpatch = reinterpret_cast<float *>(keys + P);   // [dist_chnls][TPT][7][8]
    float *tiles = qpatch + p.dist_chnls * TPT * 56;       // 16-byte aligned (56 floats per plane)
    float *dist = tiles;                                    // after the last phase

    const int q = blockIdx.x, tid = threadIdx.x;
    const int t0 = (int)qinds[3 * q], y0 = (int)qinds[3 * q + 1], x0 = (int)qinds[3 * q + 2];
    float *ov = vals + (long long)q * p.k;
    long long *oi = inds + (long long)q * p.k;
    const bool ok = t0 >= 0 && t0 <= T - TPT && y0 >= 0 && y0 <= H - TPS && x0 >= 0 && x0 <= W - TPS;
    if (!ok) {
        for (int r = tid; r < p.k; r += blockDim.x) {
            ov[r] = __int_as_float(0x7f800000);
            oi[r] = -1;
        }
        return;
    }
    const long long HW = (long long)H * W, CHW = (long long)C * HW;
    const int dc = p.dist_chnls;
    if (tid < 32) build_windows_warp(S, t0, y0, x0, T, H, W, fflow, bflow, p);
    for (int i = tid; i < dc * TPT * 49; i += blockDim.x) {
        const int pl = i / 49, r = i - pl * 49, c = pl / TPT, ht = pl - c * TPT, hy = r / 7, hx = r - hy * 7;
        qpatch[pl * 56 + hy * 8 + hx] = img[(long long)(t0 + ht) * CHW + c * HW + (long long)(y0 + hy) * W + x0 + hx];
    }
    __syncthreads();
    const int nfr = S.nfr;
    const bool sw = S.shared_win != 0;
    const int nfrm = p.nWt_f + p.nWt_b + 1;    // items: frame fastest, then column quad, then strip
    const int it_r = tid / nfrm, it_f = tid - it_r * nfrm;
    const int it_s = it_r / QQUADS, it_j = it_r - it_s * QQUADS;
    const bool active = it_f < nfr && it_r < QITEMS;
    float acc[QROWS * QCOLS];
#pragma unroll
    for (int i = 0; i < QROWS * QCOLS; ++i) acc[i] = 0.f;
    const int nslots = sw ? nfr + 1 : nfr;
    // one loop over the (channel, patch frame) phases; with a shared window the tiles staged for patch frame 0 also
    // serve patch frame 1 (slot f + 1)
#pragma unroll 1
    for (int ph = 0; ph < dc * TPT; ++ph) {
        const int c = ph >> 1, ht = ph & 1;
        if (!sw || ht == 0) {
            __syncthreads();                 // every thread is done reading the tiles of the previous phase
            {
                // one warp per tile slot: lane = column for columns 0..31 (one copy per row and lane, two pointer
                // increments per row), then lane = row for column 32
                const int lane = tid & 31, wrp = tid >> 5, nw = blockDim.x >> 5;
                for (int sl = wrp; sl < nslots; sl += nw) {
                    const FrameWin w = S.fw[sw ? 0 : sl];
                    const int tt = sw ? w.t + sl : w.t + ht;
                    const int rows = w.ny + TPS - 1, cols = w.nx + TPS - 1;
                    const float *src = img + (long long)tt * CHW + c * HW + (long long)w.y0 * W + w.x0;
                    float *dst = tiles + sl * QSLOT;
                    if (lane < cols) {
                        const float *s2 = src + lane;
                        float *d2 = dst + lane;
                        for (int r = 0; r < rows; ++r) {
                            cp_async4(d2, s2);
                            s2 += W;
                            d2 += QTW;
                        }
                    }
                    if (cols > 32) {
                        if (lane < rows) cp_async4(dst + lane * QTW + 32, src + (long long)lane * W + 32);
                        if (lane == 0 && rows > 32) cp_async4(dst + 32 * QTW + 32, src + 32LL * W + 32);
                    }
                }
                cp_async_wait_all();
                __syncthreads();
            }
        }
        if (active)
            quad_plane(qpatch + ph * 56, tiles + (sw ? it_f + ht : it_f) * QSLOT + (it_s * QROWS) * QTW + QCOLS * it_j, acc);
    }
    __syncthreads();                         // the tiles are dead: their memory serves the selection
    unsigned long long *list = reinterpret_cast<unsigned long long *>(tiles) + (((size_t)ncand_max + 1) >> 1);   // behind `dist`
    if (select_topk_regs(S, acc, active, it_f, it_s, it_j, keys, list, p.k, ov, oi, C, H, W)) return;
    __syncthreads();                         // fallback (heavy ties, tiny windows): distances to shared memory
    if (active) {
        const FrameWin w = S.fw[it_f];
#pragma unroll
        for (int s = 0; s < QROWS; ++s) {
            const int r = it_s * QROWS + s;
#pragma unroll
            for (int col = 0; col < QCOLS; ++col) {
                const int x = QCOLS * it_j + col;
                if (r < w.ny && x < w.nx) dist[w.off + r * w.nx + x] = acc[s * QCOLS + col];
            }
        }
    }
    __syncthreads();
    select_topk(S, dist, keys, p.k, 0u, 0u, ov, oi, C, H, W);
}

// patches[b,n,dt,ch,dy,dx] = img[t+dt,ch,y+dy,x+dx]   (search.py:91-98)
__global__ void fill_patches_kernel(float *__restrict__ patches, const float *__restrict__ img,
                                    const long long *__restrict__ inds, long long BK, int T, int C, int H,
                                    int W, int ps, int pt) {
    const int pdim = pt * C * ps * ps;
    const long long HW = (long long)H * W, CHW = (long long)C * HW;
    for (long long bn = blockIdx.x; bn < BK; bn += gridDim.x) {
        const long long ind = inds[bn];
        if (ind < 0) continue;
        int t, y, x;
        decode_ind(ind, H, W, C, t, y, x);
        float *dst = patches + bn * pdim;
        for (int i = threadIdx.x; i < pdim; i += blockDim.x) {
            const int dx = i % ps, dy = (i / ps) % ps, ch = (i / (ps * ps)) % C, dt = i / (ps * ps * C);
            const int tt = t + dt, yy = y + dy, xx = x + dx;
            float v = 0.f;
            if (tt < T && yy < H && xx < W) v = img[(long long)tt * CHW + ch * HW + (long long)yy * W + xx];
            dst[i] = v;
        }
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static bool tiled_ok(const VnlbSearchParams *p) { return p->ps == TPS && p->pt == TPT && p->w_s == TWS; }

// search path: 0 / 2 = automatic (quad kernel for 7x7x2 patches, a 27x27 window and at most 13 frames, 1-column tiled
// kernel beyond 13 frames, generic kernel for other shapes), 1 = never the quad kernel
static int g_search_path = -1;
static int search_path() {
    if (g_search_path < 0) {
        const char *e = getenv("VNLB_SEARCH_PATH");
        g_search_path = e ? atoi(e) : 0;
        if (g_search_path < 0 || g_search_path > 2) g_search_path = 0;
    }
    return g_search_path;
}

}  // namespace vnlb

using namespace vnlb;

extern "C" int vnlb_set_search_path(int path) {
    const int prev = search_path();
    if (path >= 0 && path <= 2) g_search_path = path;
    return prev;
}

extern "C" size_t vnlb_search_workspace_bytes(int Q, const VnlbSearchParams *p) {
    (void)Q;
    (void)p;
    return 0;  // the search keeps its scratch in shared memory
}

extern "C" int vnlb_search_topk(const float *img, int T, int C, int H, int W, const int64_t *qinds, int Q,
                                const float *fflow, const float *bflow, const VnlbSearchParams *p, float *vals,
                                int64_t *inds, void *ws, size_t ws_bytes, void *stream) {
    (void)ws;
    (void)ws_bytes;
    VNLB_REQUIRE(img && p && (Q == 0 || (qinds && vals && inds)), "vnlb_search_topk: null pointer");
    VNLB_REQUIRE(T > 0 && C > 0 && H > 0 && W > 0 && Q >= 0, "vnlb_search_topk: bad shape");
    VNLB_REQUIRE(p->ps >= 1 && p->pt >= 1 && p->w_s >= 1 && (p->w_s & 1), "vnlb_search_topk: bad patch/window size");
    VNLB_REQUIRE(p->nWt_f >= 0 && p->nWt_b >= 0 && p->nWt_f + p->nWt_b + 1 <= kMaxFrames,
                 "vnlb_search_topk: temporal window must have 1..%d frames", kMaxFrames);
    VNLB_REQUIRE(p->k >= 1 && p->k <= 1024, "vnlb_search_topk: k must be in 1..1024 (got %d)", p->k);
    VNLB_REQUIRE(p->dist_chnls >= 1 && p->dist_chnls <= C, "vnlb_search_topk: dist_chnls must be in 1..C");
    VNLB_REQUIRE(p->window_mode == VNLB_WINDOW_SHIFT || p->window_mode == VNLB_WINDOW_CLIP,
                 "vnlb_search_topk: unknown window mode");
    VNLB_REQUIRE(T >= p->pt && H >= p->ps && W >= p->ps, "vnlb_search_topk: video smaller than one patch");
    VNLB_REQUIRE((fflow == nullptr) == (bflow == nullptr), "vnlb_search_topk: pass both flows or neither");
    if (Q == 0) return VNLB_OK;
    const int nfr = p->nWt_f + p->nWt_b + 1;
    const int ncand_max = nfr * p->w_s * p->w_s;
    const int P = next_pow2(p->k) > kKeyCap ? next_pow2(p->k) : kKeyCap;   // 64-bit key slots in shared memory
    const bool tiled = tiled_ok(p);
    const bool quad = tiled && nfr <= QMAXF && p->dist_chnls <= 8 && search_path() != 1;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (quad) {
        // [SearchShared][keys][query planes][frames+1 tile slots; after the last phase the distances + candidate list]
        const size_t tile_bytes = (size_t)(nfr + 1) * QSLOT * 4;
        const size_t sel_bytes = (((size_t)ncand_max + 1) / 2 + kKeyCap) * 8;   // distances (fallback) + candidate list
        const size_t smem = sizeof(SearchShared) + (size_t)P * 8 + (size_t)p->dist_chnls * TPT * 56 * 4 +
                            (tile_bytes > sel_bytes ? tile_bytes : sel_bytes);
        const int threads = ((nfr * QITEMS + 31) / 32) * 32;
        e = cudaFuncSetAttribute(search_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vnlb_search_topk: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
        search_quad_kernel<<<Q, threads, smem, st>>>(img, T, C, H, W, (const long long *)qinds, fflow, bflow, *p, P,
                                                    ncand_max, vals, (long long *)inds);
        return check_launch("vnlb_search_topk");
    }
    size_t smem = sizeof(SearchShared) + (size_t)P * 8 + (size_t)ncand_max * 4;
    if (tiled)
        smem += (size_t)(2 * 52 + TCHUNK * TPT * TSLOT) * 4;
    else
        smem += (size_t)p->dist_chnls * p->pt * p->ps * p->ps * 4;
    if (smem > 227 * 1024) {
        set_error("vnlb_search_topk: search window needs %zu B of shared memory (> 227 KB)", smem);
        return VNLB_ERR_UNSUPPORTED;
    }
    if (tiled) {
        e = cudaFuncSetAttribute(search_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vnlb_search_topk: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
        search_tiled_kernel<<<Q, kSearchThreads, smem, st>>>(img, T, C, H, W, (const long long *)qinds, fflow, bflow,
                                                            *p, P, ncand_max, vals, (long long *)inds);
    } else {
        e = cudaFuncSetAttribute(search_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vnlb_search_topk: %s", cudaGetErrorString(e)); return VNLB_ERR_CUDA; }
        search_generic_kernel<<<Q, kSearchThreads, smem, st>>>(img, T, C, H, W, (const long long *)qinds, fflow,
                                                              bflow, *p, P, ncand_max, vals, (long long *)inds);
    }
    return check_launch("vnlb_search_topk");
}

extern "C" int vnlb_fill_patches(float *patches, const float *img, const int64_t *inds, int B, int K, int T, int C,
                                 int H, int W, int ps, int pt, void *stream) {
    VNLB_REQUIRE(patches && img && inds && B >= 0 && K > 0, "vnlb_fill_patches: bad argument");
    VNLB_REQUIRE(T > 0 && C > 0 && H > 0 && W > 0 && ps >= 1 && pt >= 1, "vnlb_fill_patches: bad shape");
    if (B == 0) return VNLB_OK;
    const long long BK = (long long)B * K;
    const int grid = (int)(BK < 65535LL * 16 ? BK : 65535LL * 16);
    fill_patches_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(patches, img, (const long long *)inds, BK, T, C, H, W,
                                                               ps, pt);
    return check_launch("vnlb_fill_patches");
}
