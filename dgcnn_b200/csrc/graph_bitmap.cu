// K0b -- per-graph adjacency bitmaps for the fused per-graph kernels (KS, KSB).
//
// The fused kernels keep one graph's adjacency (with the self loop GCNConv adds,
// model.py:30-33 -> gcn_norm's add_remaining_self_loops) as a bitmap in shared memory.
// Expanding the CSR into that bitmap inside every kernel -- once in forward, once in
// backward, on the critical path of the largest graph -- cost ~15 % of their instructions,
// so it is done here ONCE per batch with the whole GPU: one warp per CSR row, bits
// combined with match/reduce inside the warp, plain stores (a row has a single owner).
// The result is 17x smaller than col[] (COLLAB-synth: 0.55 MB vs 9.4 MB), and the fused
// kernels fetch it with one coalesced copy.
//
// Layout: graph g with n_g nodes owns np_g * wpr_g words at bitmap + bmoff[g], where
// np_g = n_g rounded up to 16 (the MMA tile) and wpr_g = ceil(np_g / 32); row r, bit c set
// <=> edge c -> r (or r == c).  Padding rows/bits are zero.
// gflags[g]: bit 0 = the graph has duplicate edges (no bitmap encoding: the fused kernels
// walk the CSR for it), bit 1 = no bitmap (graph larger than max_nodes).
#include "common.cuh"

namespace dgcnn {

__device__ __forceinline__ int bitmap_words_of(int n, int max_nodes) {
    if (n <= 0 || n > max_nodes) return 0;
    const int np = (n + 15) & ~15;
    return np * ((np + 31) >> 5);
}

// exclusive scan of the per-graph word counts, one CTA
__global__ void __launch_bounds__(1024)
k0b_offsets(const int32_t* __restrict__ gptr, int num_graphs, int max_nodes,
            int32_t* __restrict__ bmoff, int32_t* __restrict__ gflags) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < num_graphs; c0 += 1024) {
        const int g = c0 + threadIdx.x;
        int v = 0;
        if (g < num_graphs) {
            const int n = gptr[g + 1] - gptr[g];
            v = bitmap_words_of(n, max_nodes);
            gflags[g] = (n > max_nodes) ? 2 : 0;
        }
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(DGCNN_FULL_MASK, winc, o);
                if (lane >= o) winc += u;
            }
            wsum[lane] = winc - w;
        }
        __syncthreads();
        const int excl = carry_s + wsum[warp] + inc - v;
        if (g < num_graphs) bmoff[g] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bmoff[num_graphs] = carry_s;
}

// one warp per node row
__global__ void __launch_bounds__(256)
k0b_fill(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
         const int32_t* __restrict__ gptr, int num_graphs, int64_t n_nodes, int max_nodes,
         const int32_t* __restrict__ bmoff, uint32_t* __restrict__ bitmap,
         int32_t* __restrict__ gflags, const int32_t* gate_word, int gate_mask) {
    if (gate_word && !(*gate_word & gate_mask)) return;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_nodes;
         i += warps) {
        // graph of node i: last g with gptr[g] <= i
        int lo = 0, hi = num_graphs;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (gptr[mid] <= i) lo = mid; else hi = mid;
        }
        const int g = lo, base = gptr[g], n = gptr[g + 1] - base;
        if (n > max_nodes) continue;
        const int np = (n + 15) & ~15, wpr = (np + 31) >> 5;
        const int r = (int)(i - base);
        uint32_t* brow = bitmap + bmoff[g] + (int64_t)r * wpr;
        const int beg = rowptr[i], end = rowptr[i + 1];
        bool dup = false;
        // a warp owns the row: accumulate one word per lane (wpr <= 32 for n <= 1024)
        uint32_t mine = 0;            // lane l holds word l of the row
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int e = c0 + lane;
            int j = -1;
            if (e < end) {
                const unsigned t = (unsigned)(col[e] - base);
                if (t < (unsigned)n) j = (int)t;       // edges leaving the graph are ignored (K0 flags them)
            }
            const int word = j >= 0 ? (j >> 5) : -1;
            const uint32_t bit = j >= 0 ? (1u << (j & 31)) : 0u;
            const uint32_t peers = __match_any_sync(DGCNN_FULL_MASK, word);
            const uint32_t val = __reduce_or_sync(peers, bit);
            if (j >= 0 && __popc(val) != __popc(peers)) dup = true;   // same bit twice in this chunk
            // hand each distinct word to the lane that owns it
#pragma unroll 1
            for (uint32_t todo = __ballot_sync(DGCNN_FULL_MASK, j >= 0 && lane == __ffs(peers) - 1); todo;
                 todo &= todo - 1) {
                const int src = __ffs(todo) - 1;
                const int w = __shfl_sync(DGCNN_FULL_MASK, word, src);
                const uint32_t v = __shfl_sync(DGCNN_FULL_MASK, val, src);
                if (lane == w) {
                    if (mine & v) dup = true;                          // bit already set by an earlier chunk
                    mine |= v;
                }
            }
        }
        if (lane == (r >> 5)) mine |= 1u << (r & 31);                  // the self loop
        if (lane < wpr) brow[lane] = mine;
        if (__any_sync(DGCNN_FULL_MASK, dup) && lane == 0) atomicOr(&gflags[g], 1);
    }
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" int64_t dgcnn_graph_bitmap_words(int64_t num_nodes, int64_t num_graphs, int64_t max_nodes) {
    if (num_nodes < 0 || num_graphs < 0 || max_nodes < 1) return 0;
    if (max_nodes > 1024) max_nodes = 1024;
    const int64_t npmax = (max_nodes + 15) / 16 * 16;
    // sum_g np_g * wpr_g <= (N + 15 B) * wpr_max
    return (num_nodes + 15 * num_graphs) * ((npmax + 31) / 32) + 32;
}

extern "C" int dgcnn_build_bitmaps(const int32_t* rowptr, const int32_t* col, const int32_t* gptr,
                                   int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                                   uint32_t* bitmap, int64_t bitmap_words, int32_t* bmoff,
                                   int32_t* gflags, const int32_t* gate_word, int32_t gate_mask,
                                   void* stream) {
    if (num_nodes < 0 || num_graphs < 0 || max_nodes < 1) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs == 0) return DGCNN_OK;
    if (!rowptr || !gptr || !bitmap || !bmoff || !gflags) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_nodes >= INT32_MAX || num_graphs >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (max_nodes > 1024) max_nodes = 1024;          // a row must fit one word per lane
    const int64_t need = dgcnn_graph_bitmap_words(num_nodes, num_graphs, max_nodes);
    if (bitmap_words < need || need >= INT32_MAX) return DGCNN_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)need, st) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    k0b_offsets<<<1, 1024, 0, st>>>(gptr, (int)num_graphs, (int)max_nodes, bmoff, gflags);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (num_nodes > 0) {
        k0b_fill<<<grid_for(num_nodes, 8, 8), 256, 0, st>>>(rowptr, col, gptr, (int)num_graphs, num_nodes,
                                                            (int)max_nodes, bmoff, bitmap, gflags,
                                                            gate_word, gate_mask);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    return DGCNN_OK;
}
