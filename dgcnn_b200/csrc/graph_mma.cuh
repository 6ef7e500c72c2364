// Tensor-core building blocks shared by the fused per-graph kernels (graph_stack_mma.cu
// forward, graph_stack_bwd_mma.cu backward): thread teams, mma.sync wrappers, adjacency
// fragments from the bitmap, fp16 hi/lo splitting.
#pragma once
#include <cuda_fp16.h>

#include "graph_stack.cuh"

namespace dgcnn {

constexpr int kWPad = 40;   // row stride (halfs) of the 32x32 weight planes: conflict-free

__host__ __device__ inline int al16(int v) { return (v + 15) & ~15; }

// A TEAM is the set of threads working on one graph: 1, 2 or 4 "quads" of 128 threads of
// the 512-thread CTA.  Big graphs get the whole CTA (16 warps, the whole SM: they are the
// critical path of the launch); small graphs run four at a time, one per quad, each with
// its own named barrier and its own slice of the dynamic shared memory.  Graphs arrive in
// descending size (gorder), so a group only ever splits, never merges.
constexpr int kCtaThreads = 512;
constexpr int kQuadThreads = 128;
constexpr int kQuads = kCtaThreads / kQuadThreads;

struct Team {
    int tid, nthreads, warp, nwarps, lane;
    int bar;                 // named barrier id (1 + first quad of the group)
    unsigned char* smem;     // the group's slice of dynamic shared memory
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
    }
};

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// two adjacency bits -> two fp16 {0,1} packed: bit 0 -> low half, bit 1 -> high half
__device__ __forceinline__ uint32_t adj_pair(uint32_t bits) {
    return ((bits & 1u) * 0x3C00u) | (((bits >> 1) & 1u) * 0x3C000000u);
}

// (x0, x1) -> packed fp16 hi parts and packed fp16 residuals; x0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void store_split(__half* hi_plane, __half* lo_plane, int idx, float v) {
    const __half h = __float2half_rn(v);
    hi_plane[idx] = h;
    lo_plane[idx] = __float2half_rn(v - __half2float(h));
}

// A fragment (rows m0+g, m0+g+8; columns kt*16 ..) of the 0/1 adjacency from the bitmap
__device__ __forceinline__ bool adj_fragment(const uint32_t* __restrict__ bm, int wpr, int row0, int kt,
                                             int t, uint32_t (&a)[4]) {
    const uint32_t w0 = bm[row0 * wpr + (kt >> 1)];
    const uint32_t w1 = bm[(row0 + 8) * wpr + (kt >> 1)];
    const int sh = ((kt & 1) << 4) + 2 * t;
    const uint32_t x0 = w0 >> sh, x1 = w1 >> sh;
    a[0] = adj_pair(x0);
    a[1] = adj_pair(x1);
    a[2] = adj_pair(x0 >> 8);
    a[3] = adj_pair(x1 >> 8);
    return __any_sync(DGCNN_FULL_MASK, (a[0] | a[1] | a[2] | a[3]) != 0u);
}

// team-wide bitonic sort of 64-bit composites (p = power of two >= 2)
__device__ __forceinline__ void bitonic_sort_team(uint64_t* buf, uint32_t p, const Team& tm) {
    for (uint32_t size = 2; size <= p; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            tm.sync();
            for (uint32_t t = tm.tid; t < (p >> 1); t += tm.nthreads) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = buf[lo], b = buf[hi];
                if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    tm.sync();
}

}  // namespace dgcnn
