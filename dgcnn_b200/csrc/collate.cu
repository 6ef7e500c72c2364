// N1 -- device-resident data set + collate on the GPU.
//
// Replaces, per training step,
//   train.py:108-109  DataLoader(...) -> Batch.from_data_list (host concatenation, node-offset
//                     fix-up of edge_index, `batch`, `ptr`)
//   train.py:36       sample.to(device)  (16 bytes of int64 indices per edge over PCIe: what
//                     bounds the end-to-end rate of the host-fed step)
//   model.py:28 + gcn_norm prologue (K0) for that batch
// by a gather: the whole data set lives in HBM as ONE canonical CSR (built once by K0 over all
// graphs, so loops are dropped, duplicates kept, rows sorted and the symmetry verdict known);
// a batch is a list of graph ids.  Graphs never share nodes or edges, therefore the CSR of a
// batch is the concatenation of the per-graph CSR segments with two offsets fixed up:
//
//   gptr[b]              = sum_{b' < b} n(ids[b'])
//   rowptr[gptr[b] + i]  = ds.rowptr[sn0 + i] - se0 + eoff[b]       sn0 = ds.gptr[ids[b]]
//   col[eoff[b] + j]     = ds.col[se0 + j]    - sn0 + gptr[b]       se0 = ds.rowptr[sn0]
//   dis, x               = copies of the graph's rows               eoff = prefix sum of e(ids[.])
//   batch[gptr[b] + i]   = b,   y[b] = ds.y[ids[b]]
//
// which is bit for bit what K0 produces from the host-collated batch (tests/test_gpu_parity.py).
// Two launches: n1_plan (one CTA: offsets, graph order, labels, validation) and n1_gather
// (grid-stride, HBM-bound: 4 B read + 4 B written per edge instead of 16 B over PCIe + K0).
#include "common.cuh"

namespace dgcnn {

constexpr int kPlanThreads = 1024;
constexpr int kPlanOrderGraphs = 4096;     // rank sort in shared memory up to here (as K0)
constexpr int kGatherThreads = 256;
constexpr int kGatherSmemGraphs = 2048;    // offset tables in shared memory up to here

struct CollateWorkspace {
    int32_t* eoff;   // [B+1] first batch edge of batch graph b
    int32_t* sn0;    // [B]   first data-set node of graph ids[b]
    int32_t* se0;    // [B]   first data-set edge of graph ids[b]
    int32_t* ok;     // [1]   1 when ids and the caller's totals are consistent (n1_gather runs)
    size_t bytes;
};

__host__ inline CollateWorkspace carve_collate_workspace(void* base, int64_t b) {
    CollateWorkspace w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    w.eoff = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)(b + 1)));
    w.sn0 = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)(b > 0 ? b : 1)));
    w.se0 = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)(b > 0 ? b : 1)));
    w.ok = reinterpret_cast<int32_t*>(take(sizeof(int32_t)));
    w.bytes = off;
    return w;
}

// exclusive scan of one int per thread over the CTA (kPlanThreads = 32 warps); returns the
// exclusive prefix, *total = sum over the CTA.  warp_sums: int[33] in shared memory.
__device__ __forceinline__ int plan_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = warp_sums[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(DGCNN_FULL_MASK, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;                 // exclusive prefix of the warp sums
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    const int excl = warp_sums[warp] + inc - v;
    *total = warp_sums[32];
    __syncthreads();                                // the next scan reuses warp_sums
    return excl;
}

// One CTA.  Offsets of every batch graph (chunked block scan with a running carry), the
// processing order of the graphs (descending size, ties by index: the rule of k0_finalize),
// labels, and validation of ids and of the caller's totals.
__global__ void __launch_bounds__(kPlanThreads)
n1_plan(const int32_t* __restrict__ ds_gptr, const int32_t* __restrict__ ds_rowptr,
        const int32_t* __restrict__ ds_rowptr_t, const int64_t* __restrict__ ds_y, int64_t ds_graphs,
        int generic, const int32_t* __restrict__ ids, int num_graphs, int64_t num_nodes, int64_t num_edges,
        int32_t* __restrict__ gptr, int32_t* __restrict__ eoff, int32_t* __restrict__ sn0,
        int32_t* __restrict__ se0, int32_t* __restrict__ ok, int32_t* __restrict__ gorder,
        int64_t* __restrict__ y, int32_t* status) {
    __shared__ int warp_sums[33];
    __shared__ int sizes[kPlanOrderGraphs];
    const int B = num_graphs;
    long long carry_n = 0, carry_e = 0;
    bool bad = false;
    for (int base = 0; base < B; base += kPlanThreads) {
        const int b = base + threadIdx.x;
        int n = 0, e = 0, first_node = 0, first_edge = 0;
        if (b < B) {
            const int64_t g = ids[b];
            if (g < 0 || g >= ds_graphs) {
                bad = true;
            } else {
                first_node = ds_gptr[g];
                const int last_node = ds_gptr[g + 1];
                first_edge = ds_rowptr[first_node];
                const int last_edge = ds_rowptr[last_node];
                n = last_node - first_node;
                e = last_edge - first_edge;
                // edges never leave their graph, so both CSRs put a graph's edges in the same span
                if (ds_rowptr_t && (ds_rowptr_t[first_node] != first_edge || ds_rowptr_t[last_node] != last_edge))
                    bad = true;
                if (n < 0 || e < 0) { bad = true; n = 0; e = 0; }
                if (y) y[b] = ds_y ? ds_y[g] : 0;
            }
            sn0[b] = first_node;
            se0[b] = first_edge;
            if (b < kPlanOrderGraphs) sizes[b] = n;
        }
        int tn, te;
        const int xn = plan_scan(n, warp_sums, &tn);
        const int xe = plan_scan(e, warp_sums, &te);
        if (b < B) {
            gptr[b] = (int32_t)(carry_n + xn);
            eoff[b] = (int32_t)(carry_e + xe);
        }
        carry_n += tn;
        carry_e += te;
    }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        // the caller's totals size every output buffer; n1_gather only runs when they are right
        gptr[B] = (int32_t)num_nodes;
        eoff[B] = (int32_t)num_edges;
        const bool consistent = !any_bad && carry_n == num_nodes && carry_e == num_edges;
        int s = 0;
        if (!consistent) s |= DGCNN_GRAPH_BAD_BATCH;
        if (generic) s |= DGCNN_GRAPH_GENERIC;
        if (s && status) atomicOr(status, s);
        *ok = consistent ? 1 : 0;
    }
    if (gorder) {
        if (B > kPlanOrderGraphs) {
            for (int g = threadIdx.x; g < B; g += kPlanThreads) gorder[g] = g;
        } else {
            for (int g = threadIdx.x; g < B; g += kPlanThreads) {
                const int mine = sizes[g];
                int rank = 0;
                for (int h = 0; h < B; ++h) {
                    const int other = sizes[h];
                    rank += (other > mine) || (other == mine && h < g);
                }
                gorder[rank] = g;
            }
        }
    }
}

// largest b in [0, count) with table[b] <= v (table[0] == 0 <= v): with equal neighbours
// (empty graphs) this is the LAST of them, i.e. the graph that really owns element v
__device__ __forceinline__ int owner_of(const int32_t* table, int count, int v) {
    int lo = 0, hi = count;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (table[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

struct GatherArgs {
    // data set
    const float* ds_x; int64_t ds_ldx; int32_t num_features;
    const int32_t* ds_rowptr; const int32_t* ds_col;
    const int32_t* ds_rowptr_t; const int32_t* ds_col_t;     // NULL: symmetric, use rowptr/col
    const float* ds_dis;
    // plan
    const int32_t* gptr; const int32_t* eoff; const int32_t* sn0; const int32_t* se0; const int32_t* ok;
    int32_t num_graphs; int32_t num_nodes; int32_t num_edges;
    // batch
    float* x; int64_t ldx; int32_t* batch32;
    int32_t* rowptr; int32_t* col; int32_t* rowptr_t; int32_t* col_t; float* dis;
    int32_t* status;
};

__global__ void __launch_bounds__(kGatherThreads)
n1_gather(const GatherArgs a) {
    extern __shared__ int32_t tables[];
    const int B = a.num_graphs, N = a.num_nodes, E = a.num_edges;
    if (*a.ok == 0) return;                          // inconsistent ids / totals: flagged by n1_plan
    const int32_t *gptr = a.gptr, *eoff = a.eoff, *sn0 = a.sn0, *se0 = a.se0;
    if (B <= kGatherSmemGraphs) {                    // offset tables into shared memory
        int32_t* s_gptr = tables;
        int32_t* s_eoff = s_gptr + (B + 1);
        int32_t* s_sn0 = s_eoff + (B + 1);
        int32_t* s_se0 = s_sn0 + B;
        for (int i = threadIdx.x; i <= B; i += kGatherThreads) {
            s_gptr[i] = gptr[i];
            s_eoff[i] = eoff[i];
            if (i < B) { s_sn0[i] = sn0[i]; s_se0[i] = se0[i]; }
        }
        __syncthreads();
        gptr = s_gptr; eoff = s_eoff; sn0 = s_sn0; se0 = s_se0;
    }
    const int64_t stride = (int64_t)gridDim.x * kGatherThreads;
    const int64_t tid = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;
    bool bad = false;

    // ---- edges: four consecutive batch edges per thread, one 16-byte store per array ----
    const bool two = a.col_t != nullptr && a.col_t != a.col;
    const int32_t* src_t = a.ds_col_t ? a.ds_col_t : a.ds_col;
    const bool vec = ((reinterpret_cast<uintptr_t>(a.col) | (two ? reinterpret_cast<uintptr_t>(a.col_t) : 0)) & 15) == 0;
    const int64_t quads = ((int64_t)E + 3) >> 2;
    for (int64_t q = tid; q < quads; q += stride) {
        const int j0 = (int)(q << 2);
        int b = owner_of(eoff, B, j0);
        int32_t v[4] = {0, 0, 0, 0}, vt[4] = {0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + r;
            if (j >= E) break;
            while (b + 1 < B && j >= eoff[b + 1]) ++b;
            const int first = gptr[b], nodes = gptr[b + 1] - first;
            const int from = se0[b] + (j - eoff[b]);
            const int base = sn0[b];
            const int c = a.ds_col[from] - base;
            bad |= (unsigned)c >= (unsigned)nodes;
            v[r] = c + first;
            if (two) {
                const int ct = src_t[from] - base;
                bad |= (unsigned)ct >= (unsigned)nodes;
                vt[r] = ct + first;
            }
        }
        if (vec && j0 + 3 < E) {
            *reinterpret_cast<int4*>(a.col + j0) = make_int4(v[0], v[1], v[2], v[3]);
            if (two) *reinterpret_cast<int4*>(a.col_t + j0) = make_int4(vt[0], vt[1], vt[2], vt[3]);
        } else {
            for (int r = 0; r < 4 && j0 + r < E; ++r) {
                a.col[j0 + r] = v[r];
                if (two) a.col_t[j0 + r] = vt[r];
            }
        }
    }

    // ---- nodes: row pointers, dis, graph id; i == N closes the last row ----
    const bool two_rp = a.rowptr_t != nullptr && a.rowptr_t != a.rowptr;
    const int32_t* rp_t = a.ds_rowptr_t ? a.ds_rowptr_t : a.ds_rowptr;
    for (int64_t i = tid; i <= N; i += stride) {
        if (i == N) {
            a.rowptr[N] = E;
            if (two_rp) a.rowptr_t[N] = E;
            break;
        }
        const int b = owner_of(gptr, B, (int)i);
        const int from = sn0[b] + ((int)i - gptr[b]);
        const int shift = eoff[b] - se0[b];
        a.rowptr[i] = a.ds_rowptr[from] + shift;
        if (two_rp) a.rowptr_t[i] = rp_t[from] + shift;
        a.dis[i] = a.ds_dis[from];
        if (a.batch32) a.batch32[i] = b;
    }

    // ---- features: element-wise so that wide rows stay coalesced ----
    if (a.x) {
        const int F = a.num_features;
        const int64_t total = (int64_t)N * F;
        for (int64_t t = tid; t < total; t += stride) {
            const int i = (int)(t / F), f = (int)(t - (int64_t)i * F);
            const int b = owner_of(gptr, B, i);
            const int64_t from = sn0[b] + (i - gptr[b]);
            a.x[(int64_t)i * a.ldx + f] = a.ds_x[from * a.ds_ldx + f];
        }
    }
    if (bad && a.status) atomicOr(a.status, DGCNN_GRAPH_BAD_EDGE);
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_collate_workspace_bytes(int64_t num_graphs) {
    if (num_graphs < 0) return 0;
    return carve_collate_workspace(nullptr, num_graphs).bytes + 256;
}

extern "C" int dgcnn_collate(const dgcnn_dataset* ds, const int32_t* ids, int64_t num_graphs,
                             int64_t num_nodes, int64_t num_edges, float* x, int64_t ldx,
                             int32_t* batch32, int64_t* y, int32_t* rowptr, int32_t* col,
                             int32_t* rowptr_t, int32_t* col_t, float* dis, int32_t* gptr,
                             int32_t* gorder, int32_t* status, void* workspace, size_t workspace_bytes,
                             void* stream) {
    const int64_t B = num_graphs, N = num_nodes, E = num_edges;
    if (!ds || !ids || B < 1 || N < 0 || E < 0) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!ds->gptr || !ds->rowptr || !ds->dis || ds->num_graphs < 1 || ds->num_features < 1)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (ds->num_edges > 0 && !ds->col) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!ds->symmetric && (!ds->rowptr_t || (ds->num_edges > 0 && !ds->col_t))) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!rowptr || !dis || !gptr || (E > 0 && !col)) return DGCNN_ERR_INVALID_ARGUMENT;
    if ((rowptr_t == nullptr) != (col_t == nullptr) && E > 0) return DGCNN_ERR_INVALID_ARGUMENT;
    if (x && (!ds->x || ldx < ds->num_features || ds->ldx < ds->num_features)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (y && !ds->y) return DGCNN_ERR_INVALID_ARGUMENT;
    if (N >= INT32_MAX || E >= INT32_MAX || B >= INT32_MAX || ds->num_nodes >= INT32_MAX ||
        ds->num_edges >= INT32_MAX)
        return DGCNN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < dgcnn_collate_workspace_bytes(B)) return DGCNN_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uintptr_t aligned = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    CollateWorkspace w = carve_collate_workspace(reinterpret_cast<void*>(aligned), B);
    const bool generic = !ds->symmetric;

    n1_plan<<<1, kPlanThreads, 0, st>>>(ds->gptr, ds->rowptr, generic ? ds->rowptr_t : nullptr, ds->y,
                                        ds->num_graphs, generic ? 1 : 0, ids, (int)B, N, E, gptr, w.eoff,
                                        w.sn0, w.se0, w.ok, gorder, y, status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();

    GatherArgs a;
    a.ds_x = x ? ds->x : nullptr; a.ds_ldx = ds->ldx; a.num_features = ds->num_features;
    a.ds_rowptr = ds->rowptr; a.ds_col = ds->col;
    a.ds_rowptr_t = generic ? ds->rowptr_t : nullptr; a.ds_col_t = generic ? ds->col_t : nullptr;
    a.ds_dis = ds->dis;
    a.gptr = gptr; a.eoff = w.eoff; a.sn0 = w.sn0; a.se0 = w.se0; a.ok = w.ok;
    a.num_graphs = (int32_t)B; a.num_nodes = (int32_t)N; a.num_edges = (int32_t)E;
    a.x = x; a.ldx = ldx; a.batch32 = batch32;
    a.rowptr = rowptr; a.col = col; a.rowptr_t = rowptr_t; a.col_t = col_t; a.dis = dis;
    a.status = status;
    const int64_t work = ((E + 3) >> 2) > N + 1 ? ((E + 3) >> 2) : N + 1;
    const size_t smem = B <= kGatherSmemGraphs ? sizeof(int32_t) * (size_t)(4 * B + 2) : 0;
    n1_gather<<<grid_for(work, kGatherThreads, 8), kGatherThreads, smem, st>>>(a);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
