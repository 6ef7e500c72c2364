// K0 -- graph build: int64 COO edge_index + batch  ->  int32 CSR (by target and by
// source), dis = (1 + in_degree)^-1/2, per-graph node offsets.
//
// Replaces model.py:28 remove_self_loops and the gcn_norm prologue that PyG's
// GCNConv.forward re-runs in every one of model.py:30-33's four layers
// (add_remaining_self_loops -> scatter_add degree -> pow(-0.5)), plus the
// batch -> offsets part of to_dense_batch (model.py:35).  Done ONCE per batch;
// all four layers, forward and backward, share the result.
//
// Two pipelines, chosen ON THE DEVICE (no host sync, no allocation):
//
//  fast   edge_index already strictly sorted by (src,dst), loop-free and symmetric --
//         exactly what TUDataset/PyG batches look like.  Then the CSR
//         by source is the edge list itself (row pointers = run boundaries), and by
//         symmetry it is also the CSR by target.  k0_fast_build writes it in one
//         streaming pass and checks order; k0_fast_verify checks symmetry with a
//         binary search per edge and derives dis.  Any violation sets a device flag.
//
//  generic  memset degrees -> count -> 3-phase exclusive scan (+dis) -> fill (atomic
//         cursor, order arbitrary) -> per-row rank sort (ascending source id).  Always
//         launched, but every kernel returns at once unless the flag is set.
//         The row sort makes the CSR canonical: duplicates are equal values, so the
//         result -- and every later floating-point summation order -- is
//         bit-reproducible run to run although the fill uses atomics.
#include "common.cuh"

namespace dgcnn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

struct BuildWorkspace {
    int32_t* flags;  // bit 0: input is not in the fast-path form, run the generic pipeline
    int32_t* indeg;
    int32_t* outdeg;
    int32_t* bsum;   // [2][nb]
    int32_t* tmp_in;
    int32_t* tmp_out;
    size_t bytes;
    int64_t nb;
};

__host__ inline BuildWorkspace carve_build_workspace(void* base, int64_t n, int64_t e) {
    BuildWorkspace w;
    w.nb = ceil_div(n + 1, kScanTile);
    size_t off = 0;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    // flags, indeg and outdeg are adjacent so that one memset clears all three
    w.flags = reinterpret_cast<int32_t*>(take(sizeof(int32_t)));
    w.indeg = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)n));
    w.outdeg = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)n));
    w.bsum = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * 2 * (size_t)w.nb));
    w.tmp_in = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)e));
    w.tmp_out = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)e));
    w.bytes = off;
    return w;
}

// gptr[g] = first node of graph g.  batch is non-decreasing, so node i opens
// every graph in (batch[i-1], batch[i]]; the sentinel i == N closes the rest.
__device__ __forceinline__ void graph_ptr_body(const int64_t* __restrict__ batch, int64_t n,
                                               int64_t num_graphs, int32_t* __restrict__ gptr,
                                               int32_t* status, int64_t i) {
    int64_t cur = (i < n) ? batch[i] : num_graphs;
    int64_t prev = (i > 0) ? batch[i - 1] : -1;
    bool bad = (i < n) && (cur < 0 || cur >= num_graphs || cur < prev);
    if (bad) {
        if (status) atomicOr(status, DGCNN_GRAPH_BAD_BATCH);
        return;
    }
    if (prev < -1) prev = -1;
    if (prev >= num_graphs) return;
    for (int64_t g = prev + 1; g <= cur; ++g) gptr[g] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
k0_graph_ptr(const int64_t* __restrict__ batch, int64_t n, int64_t num_graphs,
             int32_t* __restrict__ gptr, int32_t* status) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride)
        graph_ptr_body(batch, n, num_graphs, gptr, status, i);
}

__global__ void __launch_bounds__(256)
k0_count(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e0,
         const int64_t* __restrict__ batch, int64_t n, int64_t num_graphs,
         int32_t* __restrict__ indeg, int32_t* __restrict__ outdeg,
         int32_t* __restrict__ gptr, int32_t* status, const int32_t* gate) {
    if (gate && !(*gate & 1)) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0 && gate && status) atomicOr(status, DGCNN_GRAPH_GENERIC);   // not proven symmetric
    for (int64_t e = tid; e < e0; e += stride) {
        int64_t s = src[e], d = dst[e];
        if ((uint64_t)s >= (uint64_t)n || (uint64_t)d >= (uint64_t)n) {
            if (status) atomicOr(status, DGCNN_GRAPH_BAD_EDGE);
            continue;
        }
        if (s == d) continue;  // remove_self_loops; the +1 in dis re-adds exactly one loop
        atomicAdd(&indeg[d], 1);
        if (outdeg) atomicAdd(&outdeg[s], 1);
    }
    if (gptr)
        for (int64_t i = tid; i <= n; i += stride)
            graph_ptr_body(batch, n, num_graphs, gptr, status, i);
}

__device__ __forceinline__ int block_sum_256(int v, int* smem /*[8]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DGCNN_FULL_MASK, v, o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) t += smem[w];
    __syncthreads();
    return t;
}

// phase 1: per-tile sums.  grid (nb, 1 or 2); y selects in- or out-degrees
__global__ void __launch_bounds__(kScanThreads)
k0_scan_reduce(const int32_t* __restrict__ indeg, const int32_t* __restrict__ outdeg, int64_t n,
               int32_t* __restrict__ bsum, int64_t nb, const int32_t* gate) {
    __shared__ int red[kScanThreads / 32];
    if (gate && !(*gate & 1)) return;
    const int32_t* deg = blockIdx.y ? outdeg : indeg;
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        int64_t i = base + q * kScanThreads + threadIdx.x;
        if (i < n) v += deg[i];
    }
    int t = block_sum_256(v, red);
    if (threadIdx.x == 0) bsum[blockIdx.y * nb + blockIdx.x] = t;
}

// phase 2: exclusive scan of the tile sums, one CTA per array
__global__ void __launch_bounds__(1024)
k0_scan_top(int32_t* __restrict__ bsum, int64_t nb, const int32_t* gate) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (gate && !(*gate & 1)) return;
    int32_t* a = bsum + blockIdx.x * nb;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < nb; c0 += 1024) {
        int64_t i = c0 + threadIdx.x;
        int v = (i < nb) ? a[i] : 0;
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
            wsum[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + wsum[warp] + inc - v;
        if (i < nb) a[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

// phase 3: per-tile exclusive scan + tile offset -> rowptr[0..n]; also dis
__global__ void __launch_bounds__(kScanThreads)
k0_scan_apply(const int32_t* __restrict__ indeg, const int32_t* __restrict__ outdeg, int64_t n,
              const int32_t* __restrict__ bsum, int64_t nb, int32_t* __restrict__ rowptr,
              int32_t* __restrict__ rowptr_t, float* __restrict__ dis, const int32_t* gate) {
    __shared__ int wsum[kScanThreads / 32];
    if (gate && !(*gate & 1)) return;
    const bool transposed = blockIdx.y != 0;
    const int32_t* deg = transposed ? outdeg : indeg;
    int32_t* out = transposed ? rowptr_t : rowptr;
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int tsum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        int64_t i = base + q;
        v[q] = (i < n) ? deg[i] : 0;
        tsum += v[q];
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) woff += (w < warp) ? wsum[w] : 0;
    int run = bsum[blockIdx.y * nb + blockIdx.x] + woff + inc - tsum;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        int64_t i = base + q;
        if (i <= n) out[i] = run;  // i == n receives the grand total
        if (!transposed && i < n) dis[i] = 1.0f / sqrtf((float)(v[q] + 1));
        run += v[q];
    }
}

// scatter sources (targets) into their rows; the degree counters double as
// reverse cursors, so they end at zero
__global__ void __launch_bounds__(256)
k0_fill(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e0, int64_t n,
        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowptr_t,
        int32_t* __restrict__ indeg, int32_t* __restrict__ outdeg,
        int32_t* __restrict__ tmp_in, int32_t* __restrict__ tmp_out, const int32_t* gate) {
    if (gate && !(*gate & 1)) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e0; e += stride) {
        int64_t s = src[e], d = dst[e];
        if ((uint64_t)s >= (uint64_t)n || (uint64_t)d >= (uint64_t)n || s == d) continue;
        int p = atomicSub(&indeg[d], 1) - 1;
        tmp_in[rowptr[d] + p] = (int32_t)s;
        if (rowptr_t) {
            int q = atomicSub(&outdeg[s], 1) - 1;
            tmp_out[rowptr_t[s] + q] = (int32_t)d;
        }
    }
}

// one warp per row: rank sort (ascending value, ties by position) from tmp to col.
// O(deg^2 / 32) shuffles per row; degrees here are tens to a few hundred.
__global__ void __launch_bounds__(256)
k0_sort_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowptr_t, int64_t n,
             const int32_t* __restrict__ tmp_in, const int32_t* __restrict__ tmp_out,
             int32_t* __restrict__ col, int32_t* __restrict__ col_t, const int32_t* gate) {
    if (gate && !(*gate & 1)) return;
    const bool transposed = blockIdx.y != 0;
    const int32_t* rp = transposed ? rowptr_t : rowptr;
    const int32_t* in = transposed ? tmp_out : tmp_in;
    int32_t* out = transposed ? col_t : col;
    int lane = threadIdx.x & 31;
    int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n;
         row += warps) {
        int beg = rp[row], deg = rp[row + 1] - beg;
        if (deg <= 0) continue;
        if (deg == 1) {
            if (lane == 0) out[beg] = in[beg];
            continue;
        }
        for (int t0 = 0; t0 < deg; t0 += 32) {
            int t = t0 + lane;
            int v = (t < deg) ? in[beg + t] : 0x7fffffff;
            int rank = 0;
            for (int c0 = 0; c0 < deg; c0 += 32) {
                int idx = c0 + lane;
                int u = (idx < deg) ? in[beg + idx] : 0x7fffffff;
                int lim = min(32, deg - c0);
                for (int s = 0; s < lim; ++s) {
                    int uu = __shfl_sync(DGCNN_FULL_MASK, u, s);
                    rank += (uu < v) || (uu == v && (c0 + s) < t);
                }
            }
            if (t < deg) out[beg + rank] = v;
        }
    }
}

// ---- fast path ---------------------------------------------------------------------
// One streaming pass: col/col_t = targets in input order, row pointers = run boundaries
// of the source column, graph offsets; flags bit 0 is raised on anything that is not a
// strictly (src,dst)-sorted, loop-free, in-range edge list.  e == e0 is the sentinel that
// closes the trailing (edge-free) rows.
__global__ void __launch_bounds__(256)
k0_fast_build(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e0,
              const int64_t* __restrict__ batch, int64_t n, int64_t num_graphs,
              int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
              int32_t* __restrict__ rowptr_t, int32_t* __restrict__ col_t,
              int32_t* __restrict__ gptr, int32_t* flags, int32_t* status) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t e = tid; e <= e0; e += stride) {
        int64_t s = n, d = 0;
        if (e < e0) {
            s = src[e];
            d = dst[e];
            if ((uint64_t)s >= (uint64_t)n || (uint64_t)d >= (uint64_t)n) {
                if (status) atomicOr(status, DGCNN_GRAPH_BAD_EDGE);
                atomicOr(flags, 1);
                continue;
            }
            col[e] = (int32_t)d;
            if (col_t) col_t[e] = (int32_t)d;
            if (s == d) atomicOr(flags, 1);
        }
        int64_t ps = -1, pd = -1;
        if (e > 0) {
            ps = src[e - 1];
            pd = dst[e - 1];
            if ((uint64_t)ps >= (uint64_t)n) { atomicOr(flags, 1); continue; }
        }
        if (e < e0 && !(ps < s || (ps == s && pd < d))) atomicOr(flags, 1);   // not strictly sorted
        if (ps < s) {
            for (int64_t r = ps + 1; r <= s; ++r) {
                rowptr[r] = (int32_t)e;
                if (rowptr_t) rowptr_t[r] = (int32_t)e;
            }
        }
    }
    if (gptr)
        for (int64_t i = tid; i <= n; i += stride)
            graph_ptr_body(batch, n, num_graphs, gptr, status, i);
}

// Processing order for the per-graph kernels: graphs by DESCENDING size (longest first
// keeps the tail of the dynamic work queue short).  One CTA, rank sort over <= 4096 sizes.
constexpr int kMaxOrderGraphs = 4096;
__global__ void __launch_bounds__(1024)
k0_graph_order(const int32_t* __restrict__ gptr, int num_graphs, int32_t* __restrict__ gorder) {
    __shared__ int sizes[kMaxOrderGraphs];
    if (num_graphs > kMaxOrderGraphs) {
        for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) gorder[g] = g;
        return;
    }
    for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) sizes[g] = gptr[g + 1] - gptr[g];
    __syncthreads();
    for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) {
        const int mine = sizes[g];
        int rank = 0;
        for (int h = 0; h < num_graphs; ++h) {
            const int other = sizes[h];
            rank += (other > mine) || (other == mine && h < g);
        }
        gorder[rank] = g;
    }
}

// symmetry: every edge (s,d) must find s in row d (rows are sorted: binary search);
// dis from the row lengths (in-degree == out-degree once symmetric)
__global__ void __launch_bounds__(256)
k0_fast_verify(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e0, int64_t n,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               float* __restrict__ dis, int32_t* flags) {
    if (*flags & 1) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < n; i += stride)
        dis[i] = 1.0f / sqrtf((float)(rowptr[i + 1] - rowptr[i] + 1));
    for (int64_t e = tid; e < e0; e += stride) {
        const int32_t s = (int32_t)src[e];
        const int64_t d = dst[e];
        int lo = rowptr[d], hi = rowptr[d + 1];
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (col[mid] < s) lo = mid + 1; else hi = mid;
        }
        if (lo >= rowptr[d + 1] || col[lo] != s) { atomicOr(flags, 1); return; }
    }
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_build_graph_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    if (num_nodes < 0 || num_edges < 0) return 0;
    return carve_build_workspace(nullptr, num_nodes, num_edges).bytes + 256;
}

extern "C" int dgcnn_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                               int32_t* gptr, int32_t* status, void* stream) {
    if (num_nodes < 0 || num_graphs < 0 || !gptr || (num_nodes > 0 && !batch))
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    k0_graph_ptr<<<grid_for(num_nodes + 1, 256, 8), 256, 0, st>>>(batch, num_nodes, num_graphs, gptr,
                                                                 status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_build_graph(const int64_t* edge_index, int64_t num_edges, const int64_t* batch,
                                 int64_t num_nodes, int64_t num_graphs, int32_t* rowptr,
                                 int32_t* col, int32_t* rowptr_t, int32_t* col_t, float* dis,
                                 int32_t* gptr, int32_t* gorder, int32_t* status, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    const int64_t n = num_nodes, e0 = num_edges;
    if (n < 0 || e0 < 0 || num_graphs < 0 || !rowptr || !dis) return DGCNN_ERR_INVALID_ARGUMENT;
    if (e0 > 0 && (!edge_index || !col)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (rowptr_t && e0 > 0 && !col_t) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!rowptr_t && col_t) return DGCNN_ERR_INVALID_ARGUMENT;
    if (gptr && n > 0 && !batch) return DGCNN_ERR_INVALID_ARGUMENT;
    if (gorder && !gptr) return DGCNN_ERR_INVALID_ARGUMENT;
    if (n >= INT32_MAX || e0 >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < dgcnn_build_graph_workspace_bytes(n, e0))
        return DGCNN_ERR_WORKSPACE;
    const bool transposed = rowptr_t != nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uintptr_t aligned = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    BuildWorkspace w = carve_build_workspace(reinterpret_cast<void*>(aligned), n, e0);
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + e0;

    size_t clear_span = (size_t)((char*)w.outdeg - (char*)w.flags) + sizeof(int32_t) * (size_t)n;
    if (cudaMemsetAsync(w.flags, 0, clear_span, st) != cudaSuccess) return DGCNN_ERR_CUDA;

    // fast path (sorted + symmetric input), verified on the device
    int64_t work = e0 + 1 > n + 1 ? e0 + 1 : n + 1;
    k0_fast_build<<<grid_for(work, 256, 8), 256, 0, st>>>(src, dst, e0, batch, n, num_graphs, rowptr,
                                                          col, rowptr_t, col_t, gptr, w.flags, status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    k0_fast_verify<<<grid_for(work, 256, 8), 256, 0, st>>>(src, dst, e0, n, rowptr, col, dis, w.flags);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (gorder && num_graphs > 0) {
        k0_graph_order<<<1, 1024, 0, st>>>(gptr, (int)num_graphs, gorder);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }

    // generic path: every kernel returns immediately unless the flag was raised
    const int32_t* gate = w.flags;
    k0_count<<<grid_for(work, 256, 8), 256, 0, st>>>(src, dst, e0, batch, n, num_graphs, w.indeg,
                                                     transposed ? w.outdeg : nullptr, nullptr, status,
                                                     gate);
    DGCNN_RETURN_IF_LAUNCH_FAILED();

    dim3 scan_grid((unsigned)w.nb, transposed ? 2 : 1);
    k0_scan_reduce<<<scan_grid, kScanThreads, 0, st>>>(w.indeg, w.outdeg, n, w.bsum, w.nb, gate);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    k0_scan_top<<<transposed ? 2 : 1, 1024, 0, st>>>(w.bsum, w.nb, gate);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    k0_scan_apply<<<scan_grid, kScanThreads, 0, st>>>(w.indeg, w.outdeg, n, w.bsum, w.nb, rowptr,
                                                      rowptr_t, dis, gate);
    DGCNN_RETURN_IF_LAUNCH_FAILED();

    if (e0 > 0) {
        k0_fill<<<grid_for(e0, 256, 8), 256, 0, st>>>(src, dst, e0, n, rowptr, rowptr_t, w.indeg,
                                                      w.outdeg, w.tmp_in, w.tmp_out, gate);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        dim3 sort_grid((unsigned)grid_for(n, 8, 8), transposed ? 2 : 1);
        k0_sort_rows<<<sort_grid, 256, 0, st>>>(rowptr, rowptr_t, n, w.tmp_in, w.tmp_out, col, col_t,
                                                gate);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    return DGCNN_OK;
}
