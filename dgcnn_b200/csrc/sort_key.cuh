// Sort-key helpers shared by the SortPool kernels (sort_pool.cu, graph_stack.cu).
#pragma once
#include "common.cuh"

namespace dgcnn {

// float -> uint32 whose ASCENDING order is the contract's DESCENDING key order
__host__ __device__ __forceinline__ uint32_t descending_key_bits(float v) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(v);
#else
    union { float f; uint32_t u; } cvt; cvt.f = v; b = cvt.u;
#endif
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0u;           // NaN sorts first
    if (b == 0x80000000u) b = 0u;                              // -0.0 == +0.0
    uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u); // ascending-order map
    return ~asc;
}

constexpr uint64_t kPadComposite = ~0ull;  // sorts after every real element

__device__ __forceinline__ void bitonic_sort_block(uint64_t* buf, uint32_t p) {
    for (uint32_t size = 2; size <= p; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (p >> 1); t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
                uint32_t hi = lo + stride;
                bool up = (lo & size) == 0;
                uint64_t a = buf[lo], b = buf[hi];
                if ((a > b) == up) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

}  // namespace dgcnn
