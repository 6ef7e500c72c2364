// Device helpers shared by the fused per-graph kernels (graph_stack.cu forward,
// graph_stack_bwd.cu backward): shared-memory adjacency bitmap, the 8-lanes-per-row
// float4 gather with the complement trick, the 32x32 row-local projection.
#pragma once
#include "common.cuh"

namespace dgcnn {

constexpr int kHid = 32;            // hidden width, model.py:13-15
constexpr int kCat = 3 * kHid + 1;  // 97, model.py:19 (Conv1d kernel/stride 97)
constexpr int kStackMaxThreads = 512;
constexpr int kSmallF = 8;          // layer 1 aggregates first when F <= 8
constexpr int kMaxF = 128;
constexpr int kSmemBudget = 227 * 1024;


struct StackFwdParams {
    const float* x; int64_t ldx; int f;
    const int32_t* rowptr; const int32_t* col; const float* dis; const int32_t* gptr;
    const uint32_t* bitmap; const int32_t* bmoff; const int32_t* gflags;   // K0b (graph_bitmap.cu)
    const uint32_t* fragmap; const int32_t* fgoff;                          // K0b, fragment-major copy
    const int32_t* gdesc;   // K0b: {graph, first node, nodes, fgoff} per graph, descending size
    int num_graphs;
    const float* w1; const float* b1; const float* w2; const float* b2;
    const float* w3; const float* b3; const float* w4; const float* b4;
    float* xcat; int64_t ldc;
    float* pooled; int32_t* perm; int k;
    int norm; int nmax;
    const int32_t* gorder;  // optional processing order (largest graphs first), else natural
    int32_t* counter;   // work queue head, zeroed by the host wrapper
    int32_t* status;    // optional
    int64_t* trace;     // optional debug timeline, [num_graphs][16] (dgcnn_stack_fwd_set_trace)
    // SURVEY 8f N2 (tensor-core variant): the head of the dense tail fused into SortPooling.
    // conv5 = Conv1d(1,16,97,97) is a per-row 97 -> 16 linear map (model.py:19,37), so it commutes
    // with the row gather: z = W5 x_cat[node] + b5 is accumulated per NODE in the layer epilogues,
    // and after the sort ReLU + MaxPool1d(2,2) (model.py:37-38) run on the k winners.  h1 != null
    // switches it on; `pooled` may then be null (no [B, k*97] round trip at all).
    const float* w5; const float* b5;      // [16,97], [16]
    float* h1; uint8_t* arg;               // [B,16,k/2] pooled activations and the winning row (0/1, 2 = dead)
    int pairs;          // tensor-core variant: launched as clusters of two CTAs (largest graphs are split)
    int split_pct;      // split a graph whose cost exceeds this percentage of an SM's fair share
    int plain_zero;     // 1: pad `pooled` with ordinary stores instead of TMA bulk stores (sanitizer runs:
                        // compute-sanitizer initcheck does not see memory written by the async proxy)
    // Lazy adjacency maps (the one-call training step): K0b only wrote the descriptors; every team
    // expands its graph's CSR rows into the fragment-major map in shared memory itself and EXPORTS
    // it (fragmap_w + fgoff, gflags_w bit 0 = duplicate edges) for the backward kernel.
    int lazy;
    uint32_t* fragmap_w; int32_t* gflags_w;
};


__host__ __device__ inline int al4(int v) { return (v + 3) & ~3; }

inline int stack_threads_for(int nmax) { return nmax <= 64 ? 128 : (nmax <= 160 ? 256 : 512); }

inline int stack_nmax_for(int64_t max_nodes) {
    int64_t r = (max_nodes + 31) / 32 * 32;
    return (int)(r < 32 ? 32 : r);
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float f4_get(const float4& a, int i) {
    return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.w));
}

// Sum of the (pre-scaled) feature rows of N(i) U {i}, float4 layout: the caller's lane
// holds channels 4q..4q+3 of row i.  Dense rows walk the complement of the bitmap row
// and subtract from the column sum; multigraphs (dup) walk the CSR instead.
__device__ __forceinline__ float4 gather_row32(const float4* __restrict__ in4,
                                               const uint32_t* __restrict__ bm, int wpr, int n,
                                               bool dup, const int* __restrict__ rp,
                                               const int32_t* __restrict__ col_g, int base,
                                               const float* __restrict__ colsum, int i, int q,
                                               uint32_t tailmask) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!dup) {
        const uint32_t* brow = bm + i * wpr;
        int cnt = 0;                                         // neighbours + self
        for (int t = 0; t < wpr; ++t) cnt += __popc(brow[t]);
        const bool comp = 2 * cnt > n;
        for (int t = 0; t < wpr; ++t) {
            uint32_t w = brow[t];
            if (comp) {
                w = ~w;
                if (t == wpr - 1) w &= tailmask;
            }
            const float4* src = in4 + (t * 32) * 8 + q;
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                acc = f4_add(acc, src[b * 8]);
            }
        }
        if (comp) {
            const float4 s = reinterpret_cast<const float4*>(colsum)[q];
            acc = make_float4(s.x - acc.x, s.y - acc.y, s.z - acc.z, s.w - acc.w);
        }
    } else {
        acc = in4[i * 8 + q];                                  // the self loop
        for (int e = rp[i]; e < rp[i + 1]; ++e)
            acc = f4_add(acc, in4[(col_g[e] - base) * 8 + q]);
    }
    return acc;
}

// y[4q..4q+3] = init + sum_k a_k * M[k][4q..4q+3] for the row held by this 8-lane group
// (a_k lives in lane (k>>2) of the group, component k&3).  Must be executed by the
// whole warp: it shuffles with the full mask.
__device__ __forceinline__ float4 project32(const float4& acc, float4 init,
                                            const float4* __restrict__ m4, int lane, int q) {
    float4 y = init;
#pragma unroll
    for (int k = 0; k < kHid; ++k) {
        const float a = __shfl_sync(DGCNN_FULL_MASK, f4_get(acc, k & 3), (lane & 24) + (k >> 2));
        const float4 w = m4[k * 8 + q];
        y.x = fmaf(a, w.x, y.x);
        y.y = fmaf(a, w.y, y.y);
        y.z = fmaf(a, w.z, y.z);
        y.w = fmaf(a, w.w, y.w);
    }
    return y;
}

// One warp per row: s_i = sum_{j in N(i) U {i}} val[j]  (val pre-scaled by c_j)
__device__ __forceinline__ float scalar_row_sum(const float* __restrict__ val,
                                                const uint32_t* __restrict__ brow, int wpr, bool dup,
                                                const int* __restrict__ rp,
                                                const int32_t* __restrict__ col_g, int base, int i) {
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    if (!dup) {
        for (int t = 0; t < wpr; ++t)
            if ((brow[t] >> lane) & 1u) s += val[t * 32 + lane];
    } else {
        for (int e = rp[i] + lane; e < rp[i + 1]; e += 32) s += val[col_g[e] - base];
        if (lane == 0) s += val[i];
    }
    return warp_sum(s);
}

// Fetch one graph's adjacency bitmap (built once per batch by K0b, graph_bitmap.cu) into
// shared memory: np*wpr contiguous words, coalesced, several loads in flight per thread.
template <int INFLIGHT = 4>
__device__ __forceinline__ void load_bitmap(const uint32_t* __restrict__ gbm, uint32_t* __restrict__ bm,
                                            int words, int tid, int nthreads) {
    for (int i0 = tid; i0 < words; i0 += nthreads * INFLIGHT) {
        uint32_t v[INFLIGHT];
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
            const int idx = i0 + u * nthreads;
            v[u] = idx < words ? gbm[idx] : 0u;
        }
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
            const int idx = i0 + u * nthreads;
            if (idx < words) bm[idx] = v[u];
        }
    }
}

}  // namespace dgcnn
