// N1 -- device-resident data set + collate on the GPU.
//
// Replaces, per training step,
//   train.py:108-109  DataLoader(...) -> Batch.from_data_list (host concatenation, node-offset
//                     fix-up of edge_index, `batch`, `ptr`)
//   train.py:36       sample.to(device)  (16 bytes of int64 indices per edge over PCIe: what
//                     bounds the end-to-end rate of the host-fed step)
//   model.py:28 + gcn_norm prologue (K0) and the adjacency bitmaps (K0b) for that batch
// by a gather: the whole data set lives in HBM as ONE canonical CSR (built once by K0 over all
// graphs, so loops are dropped, duplicates kept, rows sorted and the symmetry verdict known)
// plus K0b's per-graph bitmaps / fragment maps; a batch is a list of graph ids.  Graphs never
// share nodes or edges and K0b's blocks are relative to the graph's first node, therefore the
// graph structures of a batch are the concatenation of the per-graph segments with offsets
// fixed up:
//
//   gptr[b]              = sum_{b' < b} n(ids[b'])
//   rowptr[gptr[b] + i]  = ds.rowptr[sn0 + i] - se0 + eoff[b]       sn0 = ds.gptr[ids[b]]
//   col[eoff[b] + j]     = ds.col[se0 + j]    - sn0 + gptr[b]       se0 = ds.rowptr[sn0]
//   dis, x               = copies of the graph's rows               eoff = prefix sum of e(ids[.])
//   batch[gptr[b] + i]   = b,   y[b] = ds.y[ids[b]]
//   bitmap[bmoff[b] + w] = ds.bitmap[ds.bmoff[ids[b]] + w]          bmoff, fgoff = prefix sums of
//   fragmap[fgoff[b] + w]= ds.fragmap[ds.fgoff[ids[b]] + w]         the per-graph word counts
//   gorder = graphs by descending size (ties by index), gdesc[q] = {g, gptr[g], n_g, fgoff[g]}
//
// which is bit for bit what K0 + K0b produce from the host-collated batch
// (tests/test_gpu_resident.py).  ONE launch for batches of up to 1024 graphs: every CTA derives
// the offset tables itself (ids -> sizes -> block scan, a few microseconds of redundant work
// that saves a dependent launch), then all CTAs stride over one unified index space of edges,
// nodes, feature elements and map words, so the phases overlap instead of queueing.  HBM-bound
// copy: ~8 B per edge + the maps, instead of 16 B per edge over PCIe + K0 + K0b.  Larger
// batches take a one-CTA plan kernel first and read the tables from global memory.
#include "common.cuh"

namespace dgcnn {

constexpr int kGatherThreads = 256;
constexpr int kFusedGraphs = 1024;         // tables derived per CTA in shared memory up to here
constexpr int kPlanThreads = 1024;
constexpr int kOrderGraphs = 4096;         // rank sort up to here, identity order beyond (as K0)
constexpr int kMapMaxNodes = 1024;         // K0b's cap: larger graphs own no bitmap
constexpr int kEdgesPerItem = 8;           // edges one work item of n1_gather copies

__device__ __forceinline__ int map_words_bitmap(int n) {
    if (n <= 0 || n > kMapMaxNodes) return 0;
    const int np = (n + 15) & ~15;
    return np * ((np + 31) >> 5);
}
__device__ __forceinline__ int map_words_fragmap(int n) {
    if (n <= 0 || n > kMapMaxNodes) return 0;
    const int t = (n + 15) >> 4;
    return t * ((t + 3) >> 2) * 32;
}

// The offset tables of one batch, in shared memory (fused launch) or in the workspace.
struct Tables {
    int32_t* gptr;    // [B+1] first batch node of batch graph b
    int32_t* eoff;    // [B+1] first batch edge
    int32_t* bmoff;   // [B+1] first bitmap word
    int32_t* fgoff;   // [B+1] first fragment-map word
    int32_t* sn0;     // [B]   first data-set node of graph ids[b]
    int32_t* se0;     // [B]   first data-set edge of graph ids[b]
};
__host__ __device__ inline size_t table_ints(int64_t b) { return (size_t)(6 * b + 4); }
__host__ __device__ inline Tables carve_tables(int32_t* base, int64_t b) {
    Tables t;
    t.gptr = base;
    t.eoff = t.gptr + (b + 1);
    t.bmoff = t.eoff + (b + 1);
    t.fgoff = t.bmoff + (b + 1);
    t.sn0 = t.fgoff + (b + 1);
    t.se0 = t.sn0 + b;
    return t;
}

struct CollateArgs {
    dgcnn_dataset ds;
    dgcnn_batch_graph out;
    const int32_t* ids;
    int32_t num_graphs, num_nodes, num_edges;
    int32_t generic;                 // data set not symmetric: second CSR / bitmap in use
    int32_t want_maps;               // gather K0b's outputs too
    int32_t* status;
    int32_t* ws_tables;              // global tables (plan kernel) or NULL (derive per CTA)
    int32_t* ws_ok;                  // [1] verdict of the plan kernel
};

// exclusive block scan of two packed 64-bit sums (lo | hi << 32 each); red: u64[2][33] in
// shared memory.  Returns the exclusive prefixes, totals through tot0/tot1.
__device__ __forceinline__ void scan2(unsigned long long v0, unsigned long long v1,
                                      unsigned long long (*red)[33], unsigned long long* x0,
                                      unsigned long long* x1, unsigned long long* tot0,
                                      unsigned long long* tot1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = (blockDim.x + 31) >> 5;
    unsigned long long i0 = v0, i1 = v1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(DGCNN_FULL_MASK, i0, o);
        const unsigned long long b = __shfl_up_sync(DGCNN_FULL_MASK, i1, o);
        if (lane >= o) { i0 += a; i1 += b; }
    }
    if (lane == 31) { red[0][warp] = i0; red[1][warp] = i1; }
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w0 = lane < warps ? red[0][lane] : 0ull;
        const unsigned long long w1 = lane < warps ? red[1][lane] : 0ull;
        unsigned long long c0 = w0, c1 = w1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long a = __shfl_up_sync(DGCNN_FULL_MASK, c0, o);
            const unsigned long long b = __shfl_up_sync(DGCNN_FULL_MASK, c1, o);
            if (lane >= o) { c0 += a; c1 += b; }
        }
        red[0][lane] = c0 - w0;                     // exclusive prefix of the warp sums
        red[1][lane] = c1 - w1;
        if (lane == 31) { red[0][32] = c0; red[1][32] = c1; }
    }
    __syncthreads();
    *x0 = red[0][warp] + i0 - v0;
    *x1 = red[1][warp] + i1 - v1;
    *tot0 = red[0][32];
    *tot1 = red[1][32];
    __syncthreads();
}

// extent of batch graph b: ids[b] -> the data set's {first node, nodes, first edge, edges}
// record (dgcnn_dataset_prepare wrote and validated it): two dependent loads per graph.
// false when the id is outside the data set.
__device__ __forceinline__ bool graph_extent(const CollateArgs& a, int b, int* n, int* e, int* first_node,
                                             int* first_edge) {
    const int64_t g = a.ids[b];
    *n = 0; *e = 0; *first_node = 0; *first_edge = 0;
    if (g < 0 || g >= a.ds.num_graphs) return false;
    const int4 x = reinterpret_cast<const int4*>(a.ds.gext)[g];
    *first_node = x.x; *n = x.y; *first_edge = x.z; *e = x.w;
    return true;
}

// Block-wide: fill the tables for all B graphs (thread t owns a contiguous run of graphs:
// sums, one packed scan, then the running prefixes).  Returns true when every id is valid and
// the totals equal the caller's num_nodes / num_edges, which size every output buffer.
__device__ bool build_tables(const CollateArgs& a, const Tables& t, unsigned long long (*red)[33]) {
    const int B = a.num_graphs;
    const int per = (B + (int)blockDim.x - 1) / (int)blockDim.x;
    const int b0 = min(B, (int)threadIdx.x * per), b1 = min(B, b0 + per);
    unsigned long long ne = 0ull, wf = 0ull;         // (nodes | edges << 32), (bitmap | fragmap words << 32)
    bool bad = false;
    for (int b = b0; b < b1; ++b) {
        int n, e, fn, fe;
        bad |= !graph_extent(a, b, &n, &e, &fn, &fe);
        ne += (unsigned long long)(unsigned)n | ((unsigned long long)(unsigned)e << 32);
        wf += (unsigned long long)(unsigned)map_words_bitmap(n) |
              ((unsigned long long)(unsigned)map_words_fragmap(n) << 32);
    }
    unsigned long long xne, xwf, tne, twf;
    scan2(ne, wf, red, &xne, &xwf, &tne, &twf);
    for (int b = b0; b < b1; ++b) {
        int n, e, fn, fe;
        graph_extent(a, b, &n, &e, &fn, &fe);
        t.gptr[b] = (int32_t)(unsigned)(xne & 0xffffffffull);
        t.eoff[b] = (int32_t)(unsigned)(xne >> 32);
        t.bmoff[b] = (int32_t)(unsigned)(xwf & 0xffffffffull);
        t.fgoff[b] = (int32_t)(unsigned)(xwf >> 32);
        t.sn0[b] = fn;
        t.se0[b] = fe;
        xne += (unsigned long long)(unsigned)n | ((unsigned long long)(unsigned)e << 32);
        xwf += (unsigned long long)(unsigned)map_words_bitmap(n) |
               ((unsigned long long)(unsigned)map_words_fragmap(n) << 32);
    }
    if (threadIdx.x == 0) {
        t.gptr[B] = a.num_nodes;                      // the caller's totals bound every loop below
        t.eoff[B] = a.num_edges;
        t.bmoff[B] = (int32_t)(unsigned)(twf & 0xffffffffull);
        t.fgoff[B] = (int32_t)(unsigned)(twf >> 32);
    }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);   // also publishes the tables to the CTA
    return !any_bad && (unsigned)(tne & 0xffffffffull) == (unsigned)a.num_nodes &&
           (unsigned)(tne >> 32) == (unsigned)a.num_edges;
}

// Plan kernel of the large-batch path: one CTA writes the tables and the verdict to the workspace.
__global__ void __launch_bounds__(kPlanThreads)
n1_plan(const CollateArgs a) {
    DGCNN_PDL_WAIT();
    __shared__ unsigned long long red[2][33];
    const Tables t = carve_tables(a.ws_tables, a.num_graphs);
    const bool ok = build_tables(a, t, red);
    if (threadIdx.x == 0) *a.ws_ok = ok ? 1 : 0;
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

// kFused: the CTA derives the tables itself, in shared memory (the compiler then knows the
// address space of every table read); otherwise n1_plan left them in the workspace.
template <bool kFused>
__global__ void __launch_bounds__(kGatherThreads, 4)
n1_gather(const CollateArgs a) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) int32_t smem_tables[];
    __shared__ unsigned long long red[2][33];
    const int B = a.num_graphs, N = a.num_nodes, E = a.num_edges;
    Tables t;
    bool ok;
    if (kFused) {
        t = carve_tables(smem_tables, B);
        ok = build_tables(a, t, red);
    } else {                                          // large batch: n1_plan ran first
        t = carve_tables(a.ws_tables, B);
        ok = *a.ws_ok != 0;
    }
    const dgcnn_dataset& ds = a.ds;
    const dgcnn_batch_graph& o = a.out;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.status) {
        int s = ok ? 0 : DGCNN_GRAPH_BAD_BATCH;
        if (a.generic) s |= DGCNN_GRAPH_GENERIC;
        if (s) atomicOr(a.status, s);
    }
    if (!ok) return;                                  // nothing is written for an inconsistent batch
    const bool maps = a.want_maps != 0;
    const int64_t stride = (int64_t)gridDim.x * kGatherThreads;
    const int64_t tid = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;

    // ---- per-graph outputs, spread over the first threads of the grid -------------------
    for (int64_t b = tid; b <= B; b += stride) {
        o.gptr[b] = t.gptr[b];
        if (maps) { o.bmoff[b] = t.bmoff[b]; o.fgoff[b] = t.fgoff[b]; }
        if (b == B) break;
        const int64_t g = a.ids[b];
        if (o.y) o.y[b] = ds.y[g];
        if (maps) {
            const int n = t.gptr[b + 1] - t.gptr[b];
            o.gflags[b] = ds.gflags[g];
            if (o.gflags_t) o.gflags_t[b] = a.generic ? ds.gflags_t[g] : (n > kMapMaxNodes ? 2 : 0);
        }
    }
    // processing order: rank of every graph among all sizes (descending, ties by index), eight
    // threads per graph; the owner of a rank also writes that slot's work descriptor
    if (o.gorder) {
        if (B > kOrderGraphs) {
            for (int64_t b = tid; b < B; b += stride) {
                o.gorder[b] = (int32_t)b;
                if (maps && o.gdesc)
                    reinterpret_cast<int4*>(o.gdesc)[b] =
                        make_int4((int)b, t.gptr[b], t.gptr[b + 1] - t.gptr[b], t.fgoff[b]);
            }
        } else {
            for (int64_t item = tid; item < ((int64_t)B + 3) / 4 * 32; item += stride) {   // whole warps
                const int b = (int)(item >> 3), part = (int)(item & 7);
                const bool live = b < B;
                const int mine = live ? t.gptr[b + 1] - t.gptr[b] : 0;
                int rank = 0;
                if (live)
                    for (int h = part; h < B; h += 8) {
                        const int other = t.gptr[h + 1] - t.gptr[h];
                        rank += (other > mine) || (other == mine && h < b);
                    }
                rank += __shfl_xor_sync(DGCNN_FULL_MASK, rank, 4);
                rank += __shfl_xor_sync(DGCNN_FULL_MASK, rank, 2);
                rank += __shfl_xor_sync(DGCNN_FULL_MASK, rank, 1);
                if (live && part == 0) {
                    o.gorder[rank] = b;
                    if (maps && o.gdesc)
                        reinterpret_cast<int4*>(o.gdesc)[rank] = make_int4(b, t.gptr[b], mine, t.fgoff[b]);
                }
            }
        }
    }

    // ---- one index space for everything that is copied: [edge octets | nodes | feature elements
    //      | bitmap quads | bitmap_t quads | fragment-map quads] -----------------------------
    const bool two = o.col_t != nullptr && o.col_t != o.col;          // a second CSR to write
    const bool two_rp = o.rowptr_t != nullptr && o.rowptr_t != o.rowptr;
    const int32_t* col_t_src = a.generic ? ds.col_t : ds.col;
    const int32_t* rowptr_t_src = a.generic ? ds.rowptr_t : ds.rowptr;
    const int F = ds.num_features;
    const int64_t n_quads = ((int64_t)E + 32 * kEdgesPerItem - 1) / (32 * kEdgesPerItem) * 32;   // whole warps
    const int64_t n_nodes = (int64_t)N + 1;
    const int64_t n_feat = o.x ? (int64_t)N * F : 0;
    const int64_t n_bm = maps ? (t.bmoff[B] >> 2) : 0;                 // word counts are multiples of 16
    const int64_t n_bmt = (maps && a.generic && o.bitmap_t) ? n_bm : 0;
    const int64_t n_fg = maps ? (t.fgoff[B] >> 2) : 0;
    const int64_t c1 = n_quads, c2 = c1 + n_nodes, c3 = c2 + n_feat, c4 = c3 + n_bm, c5 = c4 + n_bmt,
                  c6 = c5 + n_fg;
    for (int64_t u = tid; u < c6; u += stride) {
        if (u < c1) {
            // The 32 lanes of a warp own 256 consecutive batch edges, lane l the edges jb + 32 r:
            // every load / store instruction of the warp is one coalesced 128-byte access.  First
            // the (branchy) source addresses, then eight independent loads in flight, then the
            // stores.  A graph change inside a lane's run is rare (graphs own thousands of edges).
            const int jb = (int)(u >> 5) * (32 * kEdgesPerItem) + (int)(u & 31);
            int b = owner_of(t.eoff, B, min(jb, E - 1));
            int next_e = t.eoff[b + 1], delta = t.se0[b] - t.eoff[b], add = t.gptr[b] - t.sn0[b];
            int src[kEdgesPerItem], ad[kEdgesPerItem];
#pragma unroll
            for (int r = 0; r < kEdgesPerItem; ++r) {
                const int j = jb + 32 * r;
                src[r] = -1; ad[r] = 0;
                if (j < E) {
                    while (j >= next_e && b + 1 < B) {
                        ++b;
                        next_e = t.eoff[b + 1]; delta = t.se0[b] - t.eoff[b]; add = t.gptr[b] - t.sn0[b];
                    }
                    src[r] = j + delta; ad[r] = add;
                }
            }
            int c[kEdgesPerItem];
#pragma unroll
            for (int r = 0; r < kEdgesPerItem; ++r) c[r] = src[r] >= 0 ? ds.col[src[r]] : 0;
#pragma unroll
            for (int r = 0; r < kEdgesPerItem; ++r)
                if (src[r] >= 0) o.col[jb + 32 * r] = c[r] + ad[r];
            if (two) {
#pragma unroll
                for (int r = 0; r < kEdgesPerItem; ++r) c[r] = src[r] >= 0 ? col_t_src[src[r]] : 0;
#pragma unroll
                for (int r = 0; r < kEdgesPerItem; ++r)
                    if (src[r] >= 0) o.col_t[jb + 32 * r] = c[r] + ad[r];
            }
        } else if (u < c2) {                          // row pointers, dis, graph id; i == N closes the last row
            const int i = (int)(u - c1);
            if (i == N) {
                o.rowptr[N] = E;
                if (two_rp) o.rowptr_t[N] = E;
            } else {
                const int b = owner_of(t.gptr, B, i);
                const int from = t.sn0[b] + (i - t.gptr[b]);
                const int shift = t.eoff[b] - t.se0[b];
                o.rowptr[i] = ds.rowptr[from] + shift;
                if (two_rp) o.rowptr_t[i] = rowptr_t_src[from] + shift;
                o.dis[i] = ds.dis[from];
                if (o.batch32) o.batch32[i] = b;
            }
        } else if (u < c3) {                          // features, element-wise: wide rows stay coalesced
            const int64_t e = u - c2;
            const int i = e < 0x7fffffff ? (int)((unsigned)e / (unsigned)F) : (int)(e / F);
            const int f = (int)(e - (int64_t)i * F);
            const int b = owner_of(t.gptr, B, i);
            const int64_t from = t.sn0[b] + (i - t.gptr[b]);
            o.x[(int64_t)i * o.ldx + f] = ds.x[from * ds.ldx + f];
        } else if (u < c5) {                          // adjacency bitmaps (A_hat, then A_hat^T), 16 bytes at a time
            const bool second = u >= c4;
            const int w = (int)((u - (second ? c4 : c3)) << 2);
            const int b = owner_of(t.bmoff, B, w);
            const int64_t from = (int64_t)ds.bmoff[a.ids[b]] + (w - t.bmoff[b]);
            const uint32_t* src = second ? ds.bitmap_t : ds.bitmap;
            uint32_t* dst = second ? o.bitmap_t : o.bitmap;
            *reinterpret_cast<int4*>(dst + w) = *reinterpret_cast<const int4*>(src + from);
        } else {                                      // fragment maps
            const int w = (int)((u - c5) << 2);
            const int b = owner_of(t.fgoff, B, w);
            const int64_t from = (int64_t)ds.fgoff[a.ids[b]] + (w - t.fgoff[b]);
            *reinterpret_cast<int4*>(o.fragmap + w) = *reinterpret_cast<const int4*>(ds.fragmap + from);
        }
    }
}

// Set-up pass over the whole data set (once): validates what dgcnn_collate then trusts --
// graph offsets, both CSRs agreeing on every graph's edge span, no edge leaving its graph --
// and writes the per-graph extent records {first node, nodes, first edge, edges}.
__global__ void __launch_bounds__(256)
n1_prepare(const dgcnn_dataset ds, int4* __restrict__ gext, int32_t* status) {
    DGCNN_PDL_WAIT();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t G = ds.num_graphs, Nd = ds.num_nodes;
    const bool generic = !ds.symmetric;
    int bad = 0;
    if (tid == 0 && (ds.gptr[0] != 0 || ds.gptr[G] != Nd || ds.rowptr[0] != 0 || ds.rowptr[Nd] != ds.num_edges))
        bad |= DGCNN_GRAPH_BAD_BATCH;
    for (int64_t g = tid; g < G; g += stride) {
        const int n0 = ds.gptr[g], n1 = ds.gptr[g + 1];
        int4 x = make_int4(0, 0, 0, 0);
        if (n0 < 0 || n1 < n0 || n1 > Nd) {
            bad |= DGCNN_GRAPH_BAD_BATCH;
        } else {
            const int e0 = ds.rowptr[n0], e1 = ds.rowptr[n1];
            if (e0 < 0 || e1 < e0 || e1 > ds.num_edges) bad |= DGCNN_GRAPH_BAD_EDGE;
            else if (generic && (ds.rowptr_t[n0] != e0 || ds.rowptr_t[n1] != e1)) bad |= DGCNN_GRAPH_BAD_EDGE;
            else x = make_int4(n0, n1 - n0, e0, e1 - e0);
        }
        gext[g] = x;
    }
    for (int64_t i = tid; i < Nd; i += stride) {        // one thread per row: set-up code, rows are short
        int lo = 0, hi = (int)G;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ds.gptr[mid] <= i) lo = mid; else hi = mid;
        }
        const int first = ds.gptr[lo], last = ds.gptr[lo + 1];
        if (i < first || i >= last) { bad |= DGCNN_GRAPH_BAD_BATCH; continue; }
        for (int pass = 0; pass < (generic ? 2 : 1); ++pass) {
            const int32_t* rp = pass ? ds.rowptr_t : ds.rowptr;
            const int32_t* cl = pass ? ds.col_t : ds.col;
            const int beg = rp[i], end = rp[i + 1];
            if (beg < 0 || end < beg || end > ds.num_edges) { bad |= DGCNN_GRAPH_BAD_EDGE; continue; }
            for (int e = beg; e < end; ++e) {
                const int c = cl[e];
                if (c < first || c >= last) bad |= DGCNN_GRAPH_BAD_EDGE;
            }
        }
    }
    if (bad && status) atomicOr(status, bad);
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_collate_workspace_bytes(int64_t num_graphs) {
    if (num_graphs < 0) return 0;
    return sizeof(int32_t) * (table_ints(num_graphs) + 64) + 512;
}

extern "C" int dgcnn_collate(const dgcnn_dataset* ds, const int32_t* ids, int64_t num_graphs,
                             int64_t num_nodes, int64_t num_edges, const dgcnn_batch_graph* out,
                             int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    const int64_t B = num_graphs, N = num_nodes, E = num_edges;
    if (!ds || !ids || !out || B < 1 || N < 0 || E < 0) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!ds->gptr || !ds->rowptr || !ds->dis || !ds->gext || ((uintptr_t)ds->gext & 15) || ds->num_graphs < 1 ||
        ds->num_features < 1)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (ds->num_edges > 0 && !ds->col) return DGCNN_ERR_INVALID_ARGUMENT;
    const bool generic = !ds->symmetric;
    if (generic && (!ds->rowptr_t || (ds->num_edges > 0 && !ds->col_t))) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!out->rowptr || !out->dis || !out->gptr || (E > 0 && !out->col)) return DGCNN_ERR_INVALID_ARGUMENT;
    if ((out->rowptr_t == nullptr) != (out->col_t == nullptr) && E > 0) return DGCNN_ERR_INVALID_ARGUMENT;
    if (out->x && (!ds->x || out->ldx < ds->num_features || ds->ldx < ds->num_features))
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (out->y && !ds->y) return DGCNN_ERR_INVALID_ARGUMENT;
    const bool maps = out->bitmap != nullptr;
    if (maps) {                                       // K0b's outputs come as a set
        if (!out->bmoff || !out->gflags || !out->fragmap || !out->fgoff) return DGCNN_ERR_INVALID_ARGUMENT;
        if (!ds->bitmap || !ds->bmoff || !ds->gflags || !ds->fragmap || !ds->fgoff)
            return DGCNN_ERR_INVALID_ARGUMENT;
        if (out->gdesc && (!out->gorder || ((uintptr_t)out->gdesc & 15))) return DGCNN_ERR_INVALID_ARGUMENT;
        if (generic && out->bitmap_t && (!ds->bitmap_t || !ds->gflags_t || !out->gflags_t))
            return DGCNN_ERR_INVALID_ARGUMENT;
        if (((uintptr_t)out->bitmap | (uintptr_t)out->bitmap_t | (uintptr_t)out->fragmap |
             (uintptr_t)ds->bitmap | (uintptr_t)ds->bitmap_t | (uintptr_t)ds->fragmap) & 15)
            return DGCNN_ERR_INVALID_ARGUMENT;
    }
    if (N >= INT32_MAX || E >= INT32_MAX || B >= INT32_MAX || ds->num_nodes >= INT32_MAX ||
        ds->num_edges >= INT32_MAX)
        return DGCNN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < dgcnn_collate_workspace_bytes(B)) return DGCNN_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    CollateArgs a;
    a.ds = *ds;
    a.out = *out;
    a.ids = ids;
    a.num_graphs = (int32_t)B; a.num_nodes = (int32_t)N; a.num_edges = (int32_t)E;
    a.generic = generic ? 1 : 0;
    a.want_maps = maps ? 1 : 0;
    a.status = status;
    a.ws_tables = nullptr;
    a.ws_ok = nullptr;
    size_t smem = sizeof(int32_t) * table_ints(B);
    if (B > kFusedGraphs) {
        int32_t* base = reinterpret_cast<int32_t*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        a.ws_ok = base;
        a.ws_tables = base + 64;
        smem = 0;
        DGCNN_LAUNCH(n1_plan, 1, kPlanThreads, 0, st, a);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    // upper bound of the unified index space (map words: at most (N + 15 B) * 32 / 4 quads each);
    // ONE wave of CTAs (4 per SM at <= 64 registers): every CTA of the fused launch derives the
    // tables first, a second wave would pay for that again
    const int64_t work = (E + kEdgesPerItem - 1) / kEdgesPerItem + 32 + N + 1 +
                         (out->x ? N * ds->num_features : 0) + (maps ? (N + 16 * B) : 0);
    const int grid = grid_for(work, kGatherThreads, 4);
    if (B > kFusedGraphs)
        DGCNN_LAUNCH((n1_gather<false>), grid, kGatherThreads, 0, st, a);
    else
        DGCNN_LAUNCH((n1_gather<true>), grid, kGatherThreads, smem, st, a);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_dataset_prepare(const dgcnn_dataset* ds, int32_t* gext, int32_t* status, void* stream) {
    if (!ds || !gext || ((uintptr_t)gext & 15)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!ds->gptr || !ds->rowptr || ds->num_graphs < 1 || ds->num_nodes < 0 || ds->num_edges < 0)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (ds->num_edges > 0 && !ds->col) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!ds->symmetric && (!ds->rowptr_t || (ds->num_edges > 0 && !ds->col_t))) return DGCNN_ERR_INVALID_ARGUMENT;
    if (ds->num_nodes >= INT32_MAX || ds->num_edges >= INT32_MAX || ds->num_graphs >= INT32_MAX)
        return DGCNN_ERR_UNSUPPORTED;
    const int64_t work = ds->num_nodes > ds->num_graphs ? ds->num_nodes : ds->num_graphs;
    DGCNN_LAUNCH(n1_prepare, grid_for(work, 256, 8), 256, 0, static_cast<cudaStream_t>(stream), 
        *ds, reinterpret_cast<int4*>(gext), status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
