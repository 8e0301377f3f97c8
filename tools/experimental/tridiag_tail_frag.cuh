// tridiag_tail_frag.cuh -- EXPERIMENT, not part of libvnlb_b200.so (round 1: measured slower than the FFMA2 sweep,
// 334 vs 230 us per 2048 groups, profiles/r1b_summary.md section 5).  Kept as a record of the mma.sync (3xTF32)
// formulation of the Householder rank-2 update.  To try it again: include this file at the end of namespace vnlb in
// vnlb_b200/csrc/bayes_tridiag.cu (it uses tridiag_scratch_floats, refl_off, BayesArgs, lds*/sts* helpers) and launch it
// in place of tridiag_tail_kernel<98: 64 -> 32> with 512 extra floats of dynamic shared memory.
// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL (vnlb_set_bayes_split(2) / VNLB_BAYES_SPLIT=2; off by default): phase 64 -> 32 with the trailing matrix held in mma ACCUMULATOR
// FRAGMENTS, the rank-2 update on the tensor cores.  Warp w owns the 16-row bands w and w+2; a band is 8 column
// tiles of m16n8k8 fragments (lane (g, t): rows g, g+8, columns 2t, 2t+1 of a tile).  Per tile and step
//     B <- B - v w^T - w v^T  =  C + [-v_hi -v_hi -v_lo -w_hi -w_hi -w_lo 0 0] . [w_hi w_lo w_hi v_hi v_lo v_hi 0 0]^T
// is ONE 3xTF32 mma (the lo*lo terms, 2^-22 relative, are dropped; tools/proto_mma_sweep.py: same accuracy as FP32);
// the B fragment comes from two scalar LDS per tile (shared by the warp's bands), the A fragment from the row scalars.
// The mat-vec of the next step reads the updated fragments (4 FFMA + one LDS.64 per tile and band, two shuffles per
// row pair at the end).  Shared-memory wavefronts per tile and 32 rows: 4 instead of 12 in tridiag_regs.
// Everything else (raw-column trick, two named barriers per step, smem layout, outputs) is tridiag_regs'.
__device__ __forceinline__ uint32_t f2tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}

template <int QDG, int OPITCH, int OREFL, int TIN, int TOUT>
__global__ void __launch_bounds__(64, 8) tridiag_tail_frag_kernel(const BayesArgs a) {
    constexpr int NR = 64, CEND = 32, NT = 64, NTL = NR / 8, NBW = 2, LDQ = 64, VL = LDQ + 4;
    constexpr int K0 = QDG - NR, K1 = QDG - 1 - CEND, KN = (K1 - K0 + 1 + 2 + 3) & ~3;
    constexpr int R0 = refl_off(QDG, K0), R1 = refl_off(QDG, K1 + 1);
    extern __shared__ __align__(16) float sm[];
    float *wsp = a.ws + (size_t)blockIdx.x * a.ws_stride;
    if (wsp[3 * OPITCH - 1] == 0.f) return;           // group skipped by the first kernel
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t s0 = smem_u32(sm);
    const uint32_t aV = s0 + 16 * VL, aW = s0 + 20 * VL, aRed = s0 + 24 * VL;
    const uint32_t aD = aRed + 96, aE = aD + 4 * KN, aTau = aE + 4 * KN, aRefl = aTau + 4 * KN;
    const uint32_t aFB = s0 + 4 * tridiag_scratch_floats<QDG, NR, CEND>();   // B-fragment table FB[col][t] = (b0, b1): 64 x 4 x 2 words, rebuilt every step
    // fragments of this warp's two bands
    float c[NBW][NTL][4];
    {
        const float *M = wsp + TIN;
#pragma unroll
        for (int q = 0; q < NBW; ++q)
#pragma unroll
            for (int j = 0; j < NTL; ++j) {
                const int row = 16 * (warp + 2 * q) + g, col = 8 * j + 2 * t;
                const float2 lo = *reinterpret_cast<const float2 *>(M + row * NR + col);
                const float2 hi = *reinterpret_cast<const float2 *>(M + (row + 8) * NR + col);
                c[q][j][0] = lo.x; c[q][j][1] = lo.y; c[q][j][2] = hi.x; c[q][j][3] = hi.y;
            }
    }
    for (int j = tid; j < 6 * VL + 24; j += NT) sm[j] = 0.f;
    for (int j = tid; j < 512; j += NT) sm[tridiag_scratch_floats<QDG, NR, CEND>() + j] = 0.f;
    __syncthreads();
    constexpr int c0 = NR - 1;
    if (t == 3) {       // columns 63 (x) and 62 (r) live in tile 7, lanes t = 3: registers 1/3 and 0/2
#pragma unroll
        for (int q = 0; q < NBW; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 16 * (warp + 2 * q) + g + 8 * h;
                const float x0 = c[q][NTL - 1][h ? 3 : 1], r0 = c[q][NTL - 1][h ? 2 : 0];
                if (i < c0) { sts32(s0 + 4 * i, x0); sts32(s0 + 4 * (VL + i), r0); }
                if (i == c0) sts32(aRed + 24, x0);
            }
    }
    int jw = (c0 - 2) >> 3;                          // window: copy of tile column jw of both bands
    float wn[NBW][4];
#pragma unroll
    for (int q = 0; q < NBW; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) wn[q][r] = c[q][NTL - 1][r];
    __syncthreads();
    float yi[NBW][2];
#pragma unroll
    for (int q = 0; q < NBW; ++q) { yi[q][0] = 0.f; yi[q][1] = 0.f; }
#pragma unroll
    for (int j = 0; j < NTL; ++j) {                  // y = B x for the first column
        const float2 x2 = lds64(s0 + 4 * (8 * j + 2 * t));
#pragma unroll
        for (int q = 0; q < NBW; ++q) {
            yi[q][0] = fmaf(c[q][j][1], x2.y, fmaf(c[q][j][0], x2.x, yi[q][0]));
            yi[q][1] = fmaf(c[q][j][3], x2.y, fmaf(c[q][j][2], x2.x, yi[q][1]));
        }
    }
#pragma unroll
    for (int q = 0; q < NBW; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            yi[q][h] += __shfl_xor_sync(0xffffffffu, yi[q][h], 1);
            yi[q][h] += __shfl_xor_sync(0xffffffffu, yi[q][h], 2);
        }
    for (int cc = NR - 1; cc >= CEND; --cc) {
        const int k = NR - 1 - cc;
        const uint32_t pp = k & 1;
        const uint32_t aX = s0 + pp * (8 * VL), aR = aX + 4 * VL, aXn = s0 + (pp ^ 1) * (8 * VL), aRn = aXn + 4 * VL;
        const uint32_t aRd = aRed + pp * 48, aRdn = aRed + (pp ^ 1) * 48;
        float xi[NBW][2], ri[NBW][2];
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < NBW; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 16 * (warp + 2 * q) + g + 8 * h;
                const bool act = i < cc;
                xi[q][h] = act ? lds32(aX + 4 * i) : 0.f;
                ri[q][h] = act ? lds32(aR + 4 * i) : 0.f;
                if (t == 0) {
                    part = fmaf(xi[q][h], yi[q][h], part);
                    if (i == cc) sts32(aRd + 16, yi[q][h]);
                    if (i == cc - 1) sts32(aRd + 20, yi[q][h]);
                    if (i == cc - 2) sts32(aRd + 36, yi[q][h]);
                }
            }
        part = warp_sum(part);
        if (lane == 0) sts32(aRd + 4 * warp, part);
        bar_sync_n(1, NT);                                                            // B1
        const float4 r0 = lds128(aRd), r1 = lds128(aRd + 16);
        const float ycm2 = lds32(aRd + 36);
        const float alpha = lds32(aX + 4 * (cc - 1)), bcc = lds32(aR + 4 * (cc - 1));
        const float xcm2 = lds32(aX + 4 * (cc - 2)), rcm2 = lds32(aR + 4 * (cc - 2));
        const float xBx = r0.x + r0.y;
        const float a2 = alpha * alpha;
        const float nrm2 = fmaxf(r1.x, a2), ycm1 = r1.y, dk = r1.z;
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
        const int cq = cc - 2, tq = (cq & 7) >> 1, odd = cq & 1;   // column c-2 sits in the window tile: lanes t = tq
        uint32_t af[NBW][4];                           // A fragments: row i -> [-v_hi -v_hi -v_lo -w_hi | -w_hi -w_lo 0 0]
#pragma unroll
        for (int q = 0; q < NBW; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 16 * (warp + 2 * q) + g + 8 * h;
                const bool act = i < cc;
                float vi = (i == cc - 1) ? 1.f : xi[q][h] * scale;
                float wi = fmaf(-hs, vi, ts * fmaf(-beta, ri[q][h], yi[q][h]));
                if (!act) { vi = 0.f; wi = 0.f; }
                const float xnext = fmaf(-vi, wcm1, fmaf(-wi, 1.f, ri[q][h]));
                if (t == 0) {
                    if (act) {
                        sts32(aV + 4 * i, vi);
                        sts32(aW + 4 * i, wi);
                        sts32(aRefl + 4 * (refl_off(QDG, K0 + k) - R0 + (cc - 1 - i)), vi);
                        sts32(aXn + 4 * i, (i < cc - 1) ? xnext : 0.f);
                        if (i == cc - 1) sts32(aRdn + 24, xnext);
                    }
                    if (i == cc) sts32(aXn + 4 * i, 0.f);
                }
                if (t == tq && act) {
                    const float qi = odd ? (h ? wn[q][3] : wn[q][1]) : (h ? wn[q][2] : wn[q][0]);
                    sts32(aRn + 4 * i, fmaf(-vi, wcm2, fmaf(-wi, vcm2, qi)));
                }
                // TF32 splits once per row and step; the 4 lanes of a row publish the 4 (b0, b1) pairs of COLUMN i
                const uint32_t vh = f2tf32(vi), wh = f2tf32(wi);
                const uint32_t vl = f2tf32(vi - __uint_as_float(vh)), wl = f2tf32(wi - __uint_as_float(wh));
                const uint32_t b0 = t == 0 ? wh : (t == 1 ? wl : (t == 2 ? wh : vh));     // k = t     of [w_hi w_lo w_hi v_hi v_lo v_hi 0 0]
                const uint32_t b1 = t == 0 ? vl : (t == 1 ? vh : 0u);                     // k = t + 4
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(aFB + 8 * (4 * i + t)), "r"(b0), "r"(b1) : "memory");
                const uint32_t sgn = 0x80000000u;          // negation is exact in TF32
                af[q][h] = (t == 0 ? vh : (t == 1 ? vh : (t == 2 ? vl : wh))) ^ sgn;      // k = t     of [-v_hi -v_hi -v_lo -w_hi -w_hi -w_lo 0 0]
                af[q][2 + h] = t == 0 ? (wh ^ sgn) : (t == 1 ? (wl ^ sgn) : 0u);          // k = t + 4
            }
        if (tid == 0) { sts32(aD + 4 * k, dk); sts32(aE + 4 * k, beta); sts32(aTau + 4 * k, tau); }
        bar_sync_n(1, NT);                                                            // B2
        float yn[NBW][2];
#pragma unroll
        for (int q = 0; q < NBW; ++q) { yn[q][0] = 0.f; yn[q][1] = 0.f; }
        const bool live0 = 16 * warp < cc, live1 = 16 * (warp + 2) < cc;
        auto bfrag = [&](int col, uint32_t &b0, uint32_t &b1) {     // B fragment of column `col` for this lane's k = t, t + 4
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(aFB + 8 * (4 * col + t)));
        };
#define VNLB_FT(J)                                                                        \
        case (J) + 1: {                                                                   \
            uint32_t b0, b1;                                                              \
            bfrag(8 * (J) + g, b0, b1);                                                   \
            const float2 x2 = lds64(aXn + 4 * (8 * (J) + 2 * t));                         \
            if (live0) {                                                                  \
                mma_tf32(c[0][J], af[0][0], af[0][1], af[0][2], af[0][3], b0, b1);        \
                yn[0][0] = fmaf(c[0][J][1], x2.y, fmaf(c[0][J][0], x2.x, yn[0][0]));      \
                yn[0][1] = fmaf(c[0][J][3], x2.y, fmaf(c[0][J][2], x2.x, yn[0][1]));      \
            }                                                                             \
            if (live1) {                                                                  \
                mma_tf32(c[1][J], af[1][0], af[1][1], af[1][2], af[1][3], b0, b1);        \
                yn[1][0] = fmaf(c[1][J][1], x2.y, fmaf(c[1][J][0], x2.x, yn[1][0]));      \
                yn[1][1] = fmaf(c[1][J][3], x2.y, fmaf(c[1][J][2], x2.x, yn[1][1]));      \
            }                                                                             \
        }
        switch ((cc + 7) >> 3) { VNLB_FT(7) VNLB_FT(6) VNLB_FT(5) VNLB_FT(4) VNLB_FT(3) VNLB_FT(2) VNLB_FT(1) VNLB_FT(0) default: break; }
#undef VNLB_FT
        {   // the window copy gets the same update
            uint32_t b0, b1;
            bfrag(8 * jw + g, b0, b1);
            mma_tf32(wn[0], af[0][0], af[0][1], af[0][2], af[0][3], b0, b1);
            mma_tf32(wn[1], af[1][0], af[1][1], af[1][2], af[1][3], b0, b1);
        }
#pragma unroll
        for (int q = 0; q < NBW; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                yn[q][h] += __shfl_xor_sync(0xffffffffu, yn[q][h], 1);
                yn[q][h] += __shfl_xor_sync(0xffffffffu, yn[q][h], 2);
                yi[q][h] = yn[q][h];
            }
        if (((cc - 3) >> 3) != jw) {                 // next step needs column c-3: re-read the window from the fragments
            jw = (cc - 3) >> 3;
            switch (jw) {
#define VNLB_FW(J) case (J): { _Pragma("unroll") for (int r = 0; r < 4; ++r) { wn[0][r] = c[0][J][r]; wn[1][r] = c[1][J][r]; } } break;
                VNLB_FW(0) VNLB_FW(1) VNLB_FW(2) VNLB_FW(3) VNLB_FW(4) VNLB_FW(5) VNLB_FW(6) VNLB_FW(7)
#undef VNLB_FW
                default: break;
            }
        }
    }
    // trailing CEND x CEND matrix for the next phase: bands 0 and 1 (q = 0 of both warps), tiles 0 .. CEND/8 - 1
    {
        float *trail = wsp + TOUT;
        const int row = 16 * warp + g;
#pragma unroll
        for (int j = 0; j < CEND / 8; ++j) {
            *reinterpret_cast<float2 *>(trail + row * CEND + 8 * j + 2 * t) = make_float2(c[0][j][0], c[0][j][1]);
            *reinterpret_cast<float2 *>(trail + (row + 8) * CEND + 8 * j + 2 * t) = make_float2(c[0][j][2], c[0][j][3]);
        }
    }
    __syncthreads();
    constexpr int NK = K1 - K0 + 1;
    const float *sd = sm + 6 * VL + 24;
    for (int idx = threadIdx.x; idx < NK; idx += NT) {
        wsp[K0 + idx] = sd[idx];
        wsp[OPITCH + K0 + idx] = sd[KN + idx];
        wsp[2 * OPITCH + K0 + idx] = sd[2 * KN + idx];
    }
    for (int idx = threadIdx.x; idx < R1 - R0; idx += NT) wsp[OREFL + R0 + idx] = sd[3 * KN + idx];
}


