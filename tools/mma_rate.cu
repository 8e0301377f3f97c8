// Round-2 microbenchmark (not part of the product): issue rates on one SM of the three instructions the Householder
// sweep is made of -- mma.sync.m16n8k8 TF32, FFMA2 and a broadcast LDS.128 -- as a function of the warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/mma_rate.cu && /tmp/mma_rate
// Output: instructions per cycle per SM (one CTA on one SM, clock64 around an unrolled loop of independent instructions).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mma(long long *out, int iters) {
    float c[8][4] = {};
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (s == 12345.678f) out[1] = 1;
}

__global__ void k_ffma2(long long *out, int iters) {
    float2 c[16];
    for (int u = 0; u < 16; ++u) c[u] = make_float2(threadIdx.x * 1e-3f, u * 1e-3f);
    const float2 x = make_float2(1.0001f, 0.9999f), y = make_float2(1e-7f, -1e-7f);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) c[u] = __ffma2_rn(c[u], x, y);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
    for (int u = 0; u < 16; ++u) s += c[u].x + c[u].y;
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (s == 12345.678f) out[1] = 1;
}

__global__ void k_lds(long long *out, int iters) {
    __shared__ float4 buf[64];
    if (threadIdx.x < 64) buf[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + 16 * ((u + i) & 63)));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;   // 4 FADD per load keep the loads live
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[1] = 1;
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int iters = 2000;
    printf("%8s %14s %14s %14s   (instructions per cycle per SM)\n", "warps", "mma.tf32.k8", "FFMA2", "LDS.128 bcast");
    for (int warps = 1; warps <= 16; warps *= 2) {
        double r[3];
        for (int which = 0; which < 3; ++which) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(d, 0, 16);
                if (which == 0) k_mma<<<1, 32 * warps>>>(d, iters);
                if (which == 1) k_ffma2<<<1, 32 * warps>>>(d, iters);
                if (which == 2) k_lds<<<1, 32 * warps>>>(d, iters);
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            }
            const double n = (double)iters * (which == 0 ? 8 : 16) * warps;
            r[which] = n / (double)h[0];
        }
        printf("%8d %14.3f %14.3f %14.3f\n", warps, r[0], r[1], r[2]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
