// TMA tile-load probe (debugging aid for search_quad_kernel): loads one box of a [Z][H][W] float tensor into shared
// memory with cp.async.bulk.tensor.3d and checks it against direct loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <bool FROM_GLOBAL>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, float *out, int bw, int bh, int x, int y, int z) {
    extern __shared__ __align__(16) unsigned char raw[];
    unsigned char *p = raw + 64;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
    p += (128u - (sa & 127u)) & 127u;
    float *tile = reinterpret_cast<float *>(p);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(raw);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bw * bh * 4) : "memory");
        const CUtensorMap *m = FROM_GLOBAL ? gmap : &tmap;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(d),
            "l"(m), "r"(b), "r"(x), "r"(y), "r"(z)
            : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D;\n"
        "bra W;\n"
        "D:\n"
        "}\n" ::"r"(b), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 64, H = 64, Z = 9;
    int bw = 36, bh = 33, x = 5, y = 7, z = 4;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    bool from_global = false;
    if (variant == 0) bw = 32;
    if (variant == 2) { x = 40; y = 40; }
    if (variant == 3) from_global = true;
    if (variant == 4) l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (variant == 5) { bw = 32; bh = 32; }
    std::vector<float> h((size_t)W * H * Z);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *dimg, *dout;
    cudaMalloc(&dimg, h.size() * 4);
    cudaMalloc(&dout, 256 * 256 * 4);
    cudaMemcpy(dimg, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    printf("variant %d entry point: %s qr=%d\n", variant, cudaGetErrorString(ge), (int)qr);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const cuuint64_t gdim[3] = {W, H, Z};
    const cuuint64_t gstr[2] = {W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)f)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dimg, gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap *gmap;
    cudaMalloc(&gmap, sizeof(map));
    cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    const int smem = 64 + 128 + bw * bh * 4;
    if (from_global)
        probe<true><<<1, 128, smem>>>(map, gmap, dout, bw, bh, x, y, z);
    else
        probe<false><<<1, 128, smem>>>(map, gmap, dout, bw, bh, x, y, z);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o((size_t)bw * bh);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < bh; ++r2)
        for (int c = 0; c < bw; ++c) {
            const int yy = y + r2, xx = x + c;
            const float want = (yy < H && xx < W) ? h[((size_t)z * H + yy) * W + xx] : 0.f;
            if (o[r2 * bw + c] != want) ++bad;
        }
    printf("variant %d: %d mismatches of %d\n", variant, bad, bw * bh);
    return bad != 0;
}
