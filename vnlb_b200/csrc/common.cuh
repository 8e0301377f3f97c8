// common.cuh -- shared helpers of libvnlb_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vnlb_b200.h"

namespace vnlb {

void set_error(const char *fmt, ...);

// kernels launched by the library so far (vnlb_kernel_launches); `kernels` = launches the call just made
extern unsigned long long g_kernel_launches;

inline int check_launch(const char *what, int kernels = 1) {
    g_kernel_launches += (unsigned long long)kernels;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return VNLB_ERR_CUDA;
    }
    return VNLB_OK;
}

#define VNLB_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            vnlb::set_error(__VA_ARGS__); \
            return VNLB_ERR_BAD_ARG;     \
        }                                \
    } while (0)

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// number of SMs of the current device (cached)
int num_sms();

// decode  ind = t*C*H*W + y*W + x
__device__ __forceinline__ void decode_ind(long long ind, int H, int W, int C, int &t, int &y, int &x) {
    const long long hw = (long long)H * W;
    t = (int)(ind / (hw * C));
    const int r = (int)(ind % hw);
    y = r / W;
    x = r - y * W;
}

// true iff all K entries of the row are != -1 (block-wide; every thread gets the answer)
__device__ __forceinline__ bool row_valid_block(const long long *row, int K) {
    int bad = 0;
    for (int i = threadIdx.x; i < K; i += blockDim.x) bad |= (row[i] == -1);
    return __syncthreads_or(bad) == 0;
}

}  // namespace vnlb
