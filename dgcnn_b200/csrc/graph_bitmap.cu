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

__device__ __forceinline__ int fragmap_words_of(int n, int max_nodes) {
    if (n <= 0 || n > max_nodes) return 0;
    const int t = (n + 15) >> 4;
    return t * ((t + 3) >> 2) * 32;
}

// exclusive scans of the per-graph word counts (row bitmap | fragment map), one CTA; the two
// running sums travel packed in one 64-bit integer
__global__ void __launch_bounds__(1024)
k0b_offsets(const int32_t* __restrict__ gptr, int num_graphs, int max_nodes,
            int32_t* __restrict__ bmoff, int32_t* __restrict__ fgoff, int32_t* __restrict__ gflags,
            int32_t* __restrict__ gflags_t, const int32_t* __restrict__ gorder, int4* __restrict__ gdesc) {
    DGCNN_PDL_WAIT();
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (int c0 = 0; c0 < num_graphs; c0 += 1024) {
        const int g = c0 + threadIdx.x;
        unsigned long long v = 0ull;
        if (g < num_graphs) {
            const int n = gptr[g + 1] - gptr[g];
            v = (unsigned long long)(unsigned)bitmap_words_of(n, max_nodes) |
                ((unsigned long long)(unsigned)fragmap_words_of(n, max_nodes) << 32);
            gflags[g] = (n > max_nodes) ? 2 : 0;
            if (gflags_t) gflags_t[g] = (n > max_nodes) ? 2 : 0;
        }
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long u = __shfl_up_sync(DGCNN_FULL_MASK, winc, o);
                if (lane >= o) winc += u;
            }
            wsum[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long excl = carry_s + wsum[warp] + inc - v;
        if (g < num_graphs) {
            bmoff[g] = (int)(unsigned)(excl & 0xffffffffull);
            if (fgoff) fgoff[g] = (int)(unsigned)(excl >> 32);
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bmoff[num_graphs] = (int)(unsigned)(carry_s & 0xffffffffull);
        if (fgoff) fgoff[num_graphs] = (int)(unsigned)(carry_s >> 32);
    }
    if (gdesc && fgoff) {
        // work descriptors in processing order: one 16-byte load tells a team all it needs
        __syncthreads();
        for (int q = threadIdx.x; q < num_graphs; q += 1024) {
            const int g = gorder ? gorder[q] : q;
            const int b0 = gptr[g];
            gdesc[q] = make_int4(g, b0, gptr[g + 1] - b0, fgoff[g]);
        }
    }
}

// One warp per node row; blockIdx.y = 1 builds the bitmap of A_hat^T from the CSR by source
// (skipped on the device when K0 proved the batch symmetric: gate_word / gate_mask).
// (wpr <= 32 words per row for n <= 1024.)
__global__ void __launch_bounds__(256)
k0b_fill(const int32_t* __restrict__ rowptr0, const int32_t* __restrict__ col0,
         const int32_t* __restrict__ rowptr1, const int32_t* __restrict__ col1,
         const int32_t* __restrict__ gptr, const int64_t* __restrict__ batch,
         const int32_t* __restrict__ batch32, int num_graphs,
         int64_t n_nodes, int max_nodes, const int32_t* __restrict__ bmoff,
         uint32_t* __restrict__ bitmap0, uint32_t* __restrict__ bitmap1,
         int32_t* __restrict__ gflags0, int32_t* __restrict__ gflags1,
         const int32_t* gate_word, int gate_mask, int only_second) {
    DGCNN_PDL_WAIT();
    const bool second = only_second || blockIdx.y != 0;
    if (second && gate_word && !(*gate_word & gate_mask)) return;
    const int32_t* __restrict__ rowptr = second ? rowptr1 : rowptr0;
    const int32_t* __restrict__ col = second ? col1 : col0;
    uint32_t* __restrict__ bitmap = second ? bitmap1 : bitmap0;
    int32_t* __restrict__ gflags = second ? gflags1 : gflags0;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_nodes;
         i += warps) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        int g;
        if (batch) {
            g = (int)batch[i];
        } else if (batch32) {
            g = batch32[i];
        } else {                                     // graph of node i: last g with gptr[g] <= i
            int lo = 0, hi = num_graphs;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (gptr[mid] <= i) lo = mid; else hi = mid;
            }
            g = lo;
        }
        if ((unsigned)g >= (unsigned)num_graphs) continue;   // K0 flags a bad batch vector
        const int base = gptr[g], n = gptr[g + 1] - base;
        if (n > max_nodes || i < base || i >= (int64_t)base + n) continue;
        const int np = (n + 15) & ~15, wpr = (np + 31) >> 5;
        const int r = (int)(i - base);
        uint32_t* brow = bitmap + bmoff[g] + (int64_t)r * wpr;
        // Lanes of a 32-edge chunk that hit the same word are grouped (match), their bits OR-ed
        // (one reduction per group, all groups at once) and the group's first lane ORs the result
        // into the row -- the bitmap was zeroed by the caller, the row belongs to this warp, so the
        // atomic only merges this warp's own chunks.  Chunks carry no state: their loads overlap.
        // Duplicate edges (multigraphs) are adjacent in a K0 row (columns ascend): compare each
        // column with its left neighbour instead of waiting for what the atomic returns.
        bool dup = false;
        int prev_last = -1;                             // last column of the previous chunk
#pragma unroll 2
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int e = c0 + lane;
            int j = -1;
            if (e < end) {
                const unsigned t = (unsigned)(col[e] - base);
                if (t < (unsigned)n) j = (int)t;       // edges leaving the graph are ignored (K0 flags them)
            }
            const int left = __shfl_up_sync(DGCNN_FULL_MASK, j, 1);
            if (j >= 0 && j == (lane == 0 ? prev_last : left)) dup = true;
            prev_last = __shfl_sync(DGCNN_FULL_MASK, j, 31);
            const int word = j >= 0 ? (j >> 5) : -1;
            const uint32_t bit = j >= 0 ? (1u << (j & 31)) : 0u;
            const uint32_t peers = __match_any_sync(DGCNN_FULL_MASK, word);
            const uint32_t val = __reduce_or_sync(peers, bit);
            if (j >= 0 && lane == __ffs(peers) - 1) atomicOr(&brow[word], val);
        }
        if (lane == 0) atomicOr(&brow[r >> 5], 1u << (r & 31));       // the self loop
        if (__any_sync(DGCNN_FULL_MASK, dup) && lane == 0) atomicOr(&gflags[g], 1);
    }
}

// Fragment-major copy of the row bitmaps for the tensor-core kernels (graph_stack_mma.cu):
// graph g with T = np/16 row tiles and G = ceil(T/4) column groups owns T*G*32 words at
// fragmap + fgoff[g]; word (mt, grp, lane = 4*gq + t) holds, for the four 16x16 blocks
// kt = 4*grp + q, the lane's m16k16 A-fragment bits: pair m = 4q + i at bit m (even column)
// and bit 16 + m (odd column), i = 0: (row gq, cols 2t..), 1: (row gq+8, cols 2t..),
// 2: (row gq, cols 2t+8..), 3: (row gq+8, cols 2t+8..).  One CTA per graph (descriptor
// order), one warp per (mt, grp) unit.
__global__ void __launch_bounds__(256)
k0b_fragments(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ bmoff,
              const int4* __restrict__ gdesc, int num_graphs, int max_nodes,
              uint32_t* __restrict__ fragmap) {
    DGCNN_PDL_WAIT();
    const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    for (int q = blockIdx.x; q < num_graphs; q += gridDim.x) {
        const int4 d = gdesc[q];                     // {graph, first node, nodes, fgoff}
        const int n = d.z;
        if (n <= 0 || n > max_nodes) continue;
        const int np = (n + 15) & ~15, wpr = (np + 31) >> 5, T = np >> 4, G = (T + 3) >> 2;
        const uint32_t* bm = bitmap + bmoff[d.x];
        uint32_t* out = fragmap + d.w;
        // gridDim.y CTAs share a graph's (mt, grp) units: the largest graph is the tail of the launch
        for (int u = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); u < T * G;
             u += gridDim.y * (blockDim.x >> 5)) {
            const int mt = u / G, grp = u - mt * G;
            const uint32_t* r0 = bm + (int64_t)(mt * 16 + gq) * wpr;
            const uint32_t* r1 = r0 + 8 * wpr;
            const int w0 = 2 * grp, w1 = 2 * grp + 1;
            const unsigned long long row0 = (unsigned long long)r0[w0] |
                                            ((unsigned long long)(w1 < wpr ? r0[w1] : 0u) << 32);
            const unsigned long long row1 = (unsigned long long)r1[w0] |
                                            ((unsigned long long)(w1 < wpr ? r1[w1] : 0u) << 32);
            uint32_t o = 0u;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned long long src = (i & 1) ? row1 : row0;
                    const int c = qq * 16 + 2 * t + 8 * (i >> 1);
                    const uint32_t two = (uint32_t)(src >> c) & 3u;
                    o |= ((two & 1u) << (4 * qq + i)) | ((two >> 1) << (16 + 4 * qq + i));
                }
            out[(u << 5) + lane] = o;
        }
    }
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" int64_t dgcnn_graph_bitmap_words(int64_t num_nodes, int64_t num_graphs, int64_t max_nodes) {
    if (num_nodes < 0 || num_graphs < 0 || max_nodes < 1) return 0;
    if (max_nodes > 1024) max_nodes = 1024;
    const int64_t npmax = (max_nodes + 15) / 16 * 16;
    // sum_g np_g * wpr_g <= (N + 15 B) * wpr_max
    // (a multiple of 4 words, so that a second bitmap placed right behind stays 16-byte aligned)
    return ((num_nodes + 15 * num_graphs) * ((npmax + 31) / 32) + 32 + 3) / 4 * 4;
}

extern "C" int64_t dgcnn_graph_fragmap_words(int64_t num_nodes, int64_t num_graphs, int64_t max_nodes) {
    if (num_nodes < 0 || num_graphs < 0 || max_nodes < 1) return 0;
    if (max_nodes > 1024) max_nodes = 1024;
    const int64_t tmax = (max_nodes + 15) / 16;
    // sum_g T_g * G_g * 32 <= (N/16 + B) * G_max * 32
    return (num_nodes / 16 + num_graphs) * ((tmax + 3) / 4) * 32 + 32;
}

// lazy = 1 (internal, the one-call training step): only what the fused forward kernel cannot do for
// itself -- the offsets / work descriptors and, for a batch K0 found asymmetric, the bitmap of
// A_hat^T; the forward maps (fragmap, gflags bit 0) are then written by dgcnn_stack_fwd* itself
// (StackFwdParams::lazy) and `bitmap` stays untouched.
int dgcnn_build_bitmaps_impl(const int32_t* rowptr, const int32_t* col,
                             const int32_t* rowptr_t, const int32_t* col_t,
                             const int32_t* gptr, const int64_t* batch, const int32_t* batch32,
                             int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                             uint32_t* bitmap, uint32_t* bitmap_t, int64_t bitmap_words,
                             int32_t* bmoff, int32_t* gflags, int32_t* gflags_t,
                             uint32_t* fragmap, int64_t fragmap_words, int32_t* fgoff,
                             const int32_t* gorder, int32_t* gdesc,
                             const int32_t* gate_word, int32_t gate_mask, void* stream, int lazy) {
    if (num_nodes < 0 || num_graphs < 0 || max_nodes < 1) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs == 0) return DGCNN_OK;
    if (!rowptr || !gptr || !bitmap || !bmoff || !gflags) return DGCNN_ERR_INVALID_ARGUMENT;
    const bool transposed = bitmap_t != nullptr;
    if (transposed && (!rowptr_t || !gflags_t)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_nodes >= INT32_MAX || num_graphs >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (max_nodes > 1024) max_nodes = 1024;          // a row must fit one word per lane
    const int64_t need = dgcnn_graph_bitmap_words(num_nodes, num_graphs, max_nodes);
    if (bitmap_words < need || need >= INT32_MAX) return DGCNN_ERR_WORKSPACE;
    if (fragmap) {
        if (!fgoff) return DGCNN_ERR_INVALID_ARGUMENT;
        const int64_t fneed = dgcnn_graph_fragmap_words(num_nodes, num_graphs, max_nodes);
        if (fragmap_words < fneed || fneed >= INT32_MAX) return DGCNN_ERR_WORKSPACE;
    }
    if (gdesc && (!fragmap || ((uintptr_t)gdesc & 15))) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (lazy && (!fragmap || !gdesc)) return DGCNN_ERR_INVALID_ARGUMENT;
    // padding rows / bits must be zero; one memset when the two bitmaps are adjacent
    if (lazy) {
        if (transposed &&
            cudaMemsetAsync(bitmap_t, 0, sizeof(uint32_t) * (size_t)need, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
    } else if (transposed && bitmap_t == bitmap + need) {
        if (cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * 2 * (size_t)need, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
    } else {
        if (cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)need, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        if (transposed &&
            cudaMemsetAsync(bitmap_t, 0, sizeof(uint32_t) * (size_t)need, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
    }
    DGCNN_LAUNCH(k0b_offsets, 1, 1024, 0, st, gptr, (int)num_graphs, (int)max_nodes, bmoff,
                                    fragmap ? fgoff : nullptr, gflags, gflags_t, gorder,
                                    reinterpret_cast<int4*>(gdesc));
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (num_nodes > 0 && lazy) {
        if (transposed) {                            // (returns at once unless K0 raised the gate)
            DGCNN_LAUNCH(k0b_fill, dim3((unsigned)grid_for(num_nodes, 8, 8), 1), 256, 0, st, 
                rowptr, col, rowptr_t, col_t, gptr, batch, batch32, (int)num_graphs, num_nodes, (int)max_nodes,
                bmoff, bitmap, bitmap_t, gflags, gflags_t, gate_word, gate_mask, 1);
            DGCNN_RETURN_IF_LAUNCH_FAILED();
        }
    } else if (num_nodes > 0) {
        dim3 grid((unsigned)grid_for(num_nodes, 8, 8), transposed ? 2 : 1);
        DGCNN_LAUNCH(k0b_fill, grid, 256, 0, st, rowptr, col, rowptr_t, col_t, gptr, batch, batch32, (int)num_graphs,
                                       num_nodes, (int)max_nodes, bmoff, bitmap, bitmap_t, gflags,
                                       gflags_t, gate_word, gate_mask, 0);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        if (fragmap && gdesc) {
            DGCNN_LAUNCH(k0b_fragments, dim3((unsigned)grid_for(num_graphs, 1, 8), 4), 256, 0, st, 
                bitmap, bmoff, reinterpret_cast<const int4*>(gdesc), (int)num_graphs, (int)max_nodes,
                fragmap);
            DGCNN_RETURN_IF_LAUNCH_FAILED();
        }
    }
    return DGCNN_OK;
}

extern "C" int dgcnn_build_bitmaps(const int32_t* rowptr, const int32_t* col,
                                   const int32_t* rowptr_t, const int32_t* col_t,
                                   const int32_t* gptr, const int64_t* batch, const int32_t* batch32,
                                   int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                                   uint32_t* bitmap, uint32_t* bitmap_t, int64_t bitmap_words,
                                   int32_t* bmoff, int32_t* gflags, int32_t* gflags_t,
                                   uint32_t* fragmap, int64_t fragmap_words, int32_t* fgoff,
                                   const int32_t* gorder, int32_t* gdesc,
                                   const int32_t* gate_word, int32_t gate_mask, void* stream) {
    return dgcnn_build_bitmaps_impl(rowptr, col, rowptr_t, col_t, gptr, batch, batch32, num_nodes, num_graphs,
                                    max_nodes, bitmap, bitmap_t, bitmap_words, bmoff, gflags, gflags_t, fragmap,
                                    fragmap_words, fgoff, gorder, gdesc, gate_word, gate_mask, stream, 0);
}
