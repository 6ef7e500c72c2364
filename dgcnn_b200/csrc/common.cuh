// Shared device/host helpers for the dgcnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgcnn_b200.h"

#define DGCNN_FULL_MASK 0xffffffffu
#define DGCNN_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define DGCNN_RETURN_IF_LAUNCH_FAILED()                 \
    do {                                                \
        if (cudaGetLastError() != cudaSuccess) return DGCNN_ERR_CUDA; \
    } while (0)

namespace dgcnn {

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// grid for a grid-stride kernel: enough CTAs for `work_items / per_cta`, capped at
// `waves` resident CTAs per SM, always >= 1
inline int grid_for(int64_t work_items, int64_t per_cta, int ctas_per_sm) {
    int64_t need = ceil_div(work_items, per_cta);
    int64_t cap = (int64_t)DGCNN_NUM_SMS * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DGCNN_FULL_MASK, v, o);
    return v;
}

// Normalisation coefficients (include/dgcnn_b200.h, K1):
//   y_i = r_i * sum_j c_j h_j     SYM: r = c = dis     RW: r = dis^2, c = 1
__device__ __forceinline__ float row_coef(float dis, int norm) {
    return norm == DGCNN_NORM_SYM ? dis : dis * dis;
}
__device__ __forceinline__ float col_coef(float dis, int norm) {
    return norm == DGCNN_NORM_SYM ? dis : 1.0f;
}

}  // namespace dgcnn
