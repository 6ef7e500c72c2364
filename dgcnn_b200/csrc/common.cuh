// Shared device/host helpers for the dgcnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/dgcnn_b200.h"

#define DGCNN_FULL_MASK 0xffffffffu
#define DGCNN_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define DGCNN_LAUNCH(kernel, grid, block, smem, st, ...) \
    dgcnn::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), st, ##__VA_ARGS__)

#define DGCNN_RETURN_IF_LAUNCH_FAILED()                 \
    do {                                                \
        if (cudaGetLastError() != cudaSuccess) return DGCNN_ERR_CUDA; \
    } while (0)

// Programmatic dependent launch.  Every kernel of the library begins with DGCNN_PDL_WAIT()
// (griddepcontrol.wait: until the PREVIOUS kernel of the stream has completed and flushed -- ordinary
// stream order for the data) and every launch goes through DGCNN_LAUNCH, which sets the programmatic
// stream-serialization attribute: the next grid is set up while its predecessor still runs.  Measured
// on the training step: -1.8 us per step inside the CUDA graph, +5 % on the eagerly launched resident
// loop.  Letting the dependents become RESIDENT early as well (griddepcontrol.launch_dependents at
// the top of every kernel, -DDGCNN_PDL_EARLY) was measured 6 % SLOWER: parked CTAs of the following
// kernels take SM resources from the side-stream kernels and from the kernel that is running.
// DGCNN_PDL=0 launches plainly (the wait is then a no-op).
#ifdef DGCNN_PDL_EARLY
#define DGCNN_PDL_WAIT() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#else
#define DGCNN_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif
// (for a kernel whose successor must NOT become resident early: the backward graph kernel takes a
// whole SM's shared memory and would lock the side-stream kernels out)
#define DGCNN_PDL_WAIT_ONLY() asm volatile("griddepcontrol.wait;" ::: "memory")

namespace dgcnn {

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* env = getenv("DGCNN_PDL");
        on = !(env && env[0] == '0');
    }
    return on != 0;
}

inline cudaLaunchAttribute pdl_attribute() {
    cudaLaunchAttribute a{};
    a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a.val.programmaticStreamSerializationAllowed = 1;
    return a;
}

template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1] = {pdl_attribute()};
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // (errors: cudaGetLastError at the call site)
}

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
