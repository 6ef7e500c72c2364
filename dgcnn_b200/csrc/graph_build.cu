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
//         streaming pass, checks order and fingerprints symmetry; k0_finalize derives
//         dis, the verdict and the graph order.  Any violation sets a device flag.
//
//  generic  memset degrees -> count -> 3-phase exclusive scan (+dis) -> fill (atomic
//         cursor, order arbitrary) -> per-row rank sort (ascending source id): ONE
//         cooperative launch with grid barriers between the phases, always issued, which
//         returns at once unless the flag is set.
//         The row sort makes the CSR canonical: duplicates are equal values, so the
//         result -- and every later floating-point summation order -- is
//         bit-reproducible run to run although the fill uses atomics.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dgcnn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

struct BuildWorkspace {
    int32_t* flags;  // bit 0: input is not in the fast-path form, run the generic pipeline
                     // (flags + 2, + 4: two 64-bit symmetry fingerprints, see k0_fast_build)
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
    w.flags = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * 8));
    w.indeg = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)n));
    w.outdeg = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)n));
    w.bsum = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * 2 * (size_t)w.nb));
    w.tmp_in = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)e));
    w.tmp_out = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)e));
    w.bytes = off;
    return w;
}

// Index arrays arrive as int64 (what the reference / PyG hands over) or, for compact host
// batches that halve the host-to-device traffic, as int32: one loader for both.
__device__ __forceinline__ int64_t ld_idx(const void* __restrict__ p, int64_t i, bool i32) {
    return i32 ? (int64_t)static_cast<const int32_t*>(p)[i] : static_cast<const int64_t*>(p)[i];
}

// gptr[g] = first node of graph g.  batch is non-decreasing, so node i opens
// every graph in (batch[i-1], batch[i]]; the sentinel i == N closes the rest.
__device__ __forceinline__ void graph_ptr_body(const void* __restrict__ batch, bool i32, int64_t n,
                                               int64_t num_graphs, int32_t* __restrict__ gptr,
                                               int32_t* status, int64_t i) {
    int64_t cur = (i < n) ? ld_idx(batch, i, i32) : num_graphs;
    int64_t prev = (i > 0) ? ld_idx(batch, i - 1, i32) : -1;
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
k0_graph_ptr(const void* __restrict__ batch, bool i32, int64_t n, int64_t num_graphs,
             int32_t* __restrict__ gptr, int32_t* status) {
    DGCNN_PDL_WAIT();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride)
        graph_ptr_body(batch, i32, n, num_graphs, gptr, status, i);
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

// ---- generic path in ONE launch -------------------------------------------------------
// count -> tile sums -> top scan -> apply (+dis) -> fill -> row sort, separated by grid-wide
// barriers (cooperative launch: every CTA is resident).  When the fast path succeeded -- the
// normal case -- the flag is clear and every thread leaves at once: one empty launch instead
// of six.
struct GenericArgs {
    const void* src; const void* dst; int64_t e0; bool i32;
    const void* batch; int64_t n; int64_t num_graphs;
    int32_t* indeg; int32_t* outdeg; int32_t* bsum; int64_t nb;
    int32_t* rowptr; int32_t* rowptr_t; float* dis;
    int32_t* tmp_in; int32_t* tmp_out; int32_t* col; int32_t* col_t;
    int32_t* status; const int32_t* gate;
};

__global__ void __launch_bounds__(kScanThreads)
k0_generic(GenericArgs a) {
    DGCNN_PDL_WAIT();
    __shared__ int red[kScanThreads / 32];
    __shared__ int carry_s;
    if (!(*a.gate & 1)) return;
    cg::grid_group grid = cg::this_grid();
    const bool transposed = a.rowptr_t != nullptr;
    const int nwhich = transposed ? 2 : 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // 1. degrees (self loops dropped: the +1 in dis re-adds exactly one loop per node)
    if (tid == 0 && a.status) atomicOr(a.status, DGCNN_GRAPH_GENERIC);      // not proven symmetric
    for (int64_t e = tid; e < a.e0; e += stride) {
        const int64_t s = ld_idx(a.src, e, a.i32), d = ld_idx(a.dst, e, a.i32);
        if ((uint64_t)s >= (uint64_t)a.n || (uint64_t)d >= (uint64_t)a.n) {
            if (a.status) atomicOr(a.status, DGCNN_GRAPH_BAD_EDGE);
            continue;
        }
        if (s == d) continue;
        atomicAdd(&a.indeg[d], 1);
        if (transposed) atomicAdd(&a.outdeg[s], 1);
    }
    grid.sync();

    // 2. per-tile sums
    for (int64_t job = blockIdx.x; job < a.nb * nwhich; job += gridDim.x) {
        const int which = (int)(job / a.nb);
        const int64_t tile = job - which * a.nb;
        const int32_t* deg = which ? a.outdeg : a.indeg;
        const int64_t base = tile * kScanTile;
        int v = 0;
#pragma unroll
        for (int q = 0; q < kScanItems; ++q) {
            const int64_t i = base + q * kScanThreads + threadIdx.x;
            if (i < a.n) v += deg[i];
        }
        const int t = block_sum_256(v, red);
        if (threadIdx.x == 0) a.bsum[which * a.nb + tile] = t;
    }
    grid.sync();

    // 3. exclusive scan of the tile sums: one CTA per array
    if ((int)blockIdx.x < nwhich) {
        int32_t* arr = a.bsum + blockIdx.x * a.nb;
        if (threadIdx.x == 0) carry_s = 0;
        __syncthreads();
        for (int64_t c0 = 0; c0 < a.nb; c0 += kScanThreads) {
            const int64_t i = c0 + threadIdx.x;
            const int v = (i < a.nb) ? arr[i] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) red[warp] = inc;
            __syncthreads();
            int woff = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; ++w) woff += (w < warp) ? red[w] : 0;
            const int excl = carry_s + woff + inc - v;
            if (i < a.nb) arr[i] = excl;
            __syncthreads();
            if (threadIdx.x == kScanThreads - 1) carry_s = excl + v;
            __syncthreads();
        }
    }
    grid.sync();

    // 4. per-tile exclusive scan + tile offset -> rowptr[0..n]; also dis
    for (int64_t job = blockIdx.x; job < a.nb * nwhich; job += gridDim.x) {
        const int which = (int)(job / a.nb);
        const int64_t tile = job - which * a.nb;
        const int32_t* deg = which ? a.outdeg : a.indeg;
        int32_t* out = which ? a.rowptr_t : a.rowptr;
        const int64_t base = tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
        int v[kScanItems];
        int tsum = 0;
#pragma unroll
        for (int q = 0; q < kScanItems; ++q) {
            const int64_t i = base + q;
            v[q] = (i < a.n) ? deg[i] : 0;
            tsum += v[q];
        }
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(DGCNN_FULL_MASK, inc, o);
            if (lane >= o) inc += u;
        }
        __syncthreads();                               // red[] may still be read by the previous job
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        int woff = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) woff += (w < warp) ? red[w] : 0;
        int run = a.bsum[which * a.nb + tile] + woff + inc - tsum;
#pragma unroll
        for (int q = 0; q < kScanItems; ++q) {
            const int64_t i = base + q;
            if (i <= a.n) out[i] = run;                // i == n receives the grand total
            if (!which && i < a.n) a.dis[i] = 1.0f / sqrtf((float)(v[q] + 1));
            run += v[q];
        }
    }
    grid.sync();

    // 5. scatter sources (targets) into their rows; the degree counters double as reverse
    //    cursors, so they end at zero
    for (int64_t e = tid; e < a.e0; e += stride) {
        const int64_t s = ld_idx(a.src, e, a.i32), d = ld_idx(a.dst, e, a.i32);
        if ((uint64_t)s >= (uint64_t)a.n || (uint64_t)d >= (uint64_t)a.n || s == d) continue;
        const int p = atomicSub(&a.indeg[d], 1) - 1;
        a.tmp_in[a.rowptr[d] + p] = (int32_t)s;
        if (transposed) {
            const int q = atomicSub(&a.outdeg[s], 1) - 1;
            a.tmp_out[a.rowptr_t[s] + q] = (int32_t)d;
        }
    }
    grid.sync();

    // 6. one warp per row: rank sort (ascending value, ties by position) from tmp to col.
    //    O(deg^2 / 32) shuffles per row; makes the CSR canonical (bit-reproducible sums)
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; job < a.n * nwhich; job += warps) {
        const int which = job >= a.n;
        const int64_t row = which ? job - a.n : job;
        const int32_t* rp = which ? a.rowptr_t : a.rowptr;
        const int32_t* in = which ? a.tmp_out : a.tmp_in;
        int32_t* out = which ? a.col_t : a.col;
        const int beg = rp[row], deg = rp[row + 1] - beg;
        if (deg <= 0) continue;
        if (deg == 1) {
            if (lane == 0) out[beg] = in[beg];
            continue;
        }
        for (int t0 = 0; t0 < deg; t0 += 32) {
            const int t = t0 + lane;
            const int v = (t < deg) ? in[beg + t] : 0x7fffffff;
            int rank = 0;
            for (int c0 = 0; c0 < deg; c0 += 32) {
                const int idx = c0 + lane;
                const int u = (idx < deg) ? in[beg + idx] : 0x7fffffff;
                const int lim = min(32, deg - c0);
                for (int s = 0; s < lim; ++s) {
                    const int uu = __shfl_sync(DGCNN_FULL_MASK, u, s);
                    rank += (uu < v) || (uu == v && (c0 + s) < t);
                }
            }
            if (t < deg) out[beg + rank] = v;
        }
    }
}

// ---- fast path ---------------------------------------------------------------------
// 32-bit mixes of an ORDERED pair; two independent ones feed the symmetry fingerprints
__device__ __forceinline__ uint32_t pair_mix(uint32_t a, uint32_t b, uint32_t k1, uint32_t k2) {
    uint32_t h = (a * k1) ^ __funnelshift_l(b * k2, b * k2, 15);
    h *= 0xC2B2AE3Du; h ^= h >> 16;
    h *= 0x27D4EB2Fu; h ^= h >> 15;
    return h;
}

// One streaming pass: col/col_t = targets in input order, row pointers = run boundaries
// of the source column, graph offsets; flags bit 0 is raised on anything that is not a
// strictly (src,dst)-sorted, loop-free, in-range edge list.  e == e0 is the sentinel that
// closes the trailing (edge-free) rows.
//
// Symmetry (needed for "CSR by source == CSR by target") is checked in the same pass by
// fingerprinting: a strictly sorted list has no duplicates, so it is symmetric iff the
// multiset of upward edges {(s,d): s < d} equals the multiset of reversed downward edges
// {(d,s): s > d}, and two independent sums
//   F_i = sum_{s<d} mix_i(s,d) - sum_{s>d} mix_i(d,s)   (mod 2^64)
// are both zero when they are; a non-symmetric list passes with probability ~2^-64.
// k0_finalize raises the flag when either sum is non-zero.  (The exact check, one binary
// search per edge, is k0_fast_verify: `exact_verify` of dgcnn_build_graph.)
__global__ void __launch_bounds__(256)
k0_fast_build(const void* __restrict__ src, const void* __restrict__ dst, int64_t e0,
              const void* __restrict__ batch, bool i32, int64_t n, int64_t num_graphs,
              int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
              int32_t* __restrict__ rowptr_t, int32_t* __restrict__ col_t,
              int32_t* __restrict__ gptr, int32_t* flags, int32_t* status) {
    DGCNN_PDL_WAIT();
    __shared__ unsigned long long red[2][8];
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long f0 = 0ull, f1 = 0ull;
    // One edge list position: its column, the fingerprint, sortedness against the previous edge
    // (ps, pd; -1 before the first edge) and the row pointers it opens.  has == false is the
    // sentinel position e == e0 that closes the last rows.
    auto position = [&](int64_t e, bool has, int64_t s, int64_t d, int64_t ps, int64_t pd) {
        if (has) {
            if ((uint64_t)s >= (uint64_t)n || (uint64_t)d >= (uint64_t)n) {
                if (status) atomicOr(status, DGCNN_GRAPH_BAD_EDGE);
                atomicOr(flags, 1);
                return;
            }
            col[e] = (int32_t)d;
            if (col_t) col_t[e] = (int32_t)d;
            if (s == d) atomicOr(flags, 1);
            // one evaluation per edge: + mix(lo, hi) for an "upward" edge (s < d), - mix(lo, hi)
            // for a "downward" one -- the same sums as  sum mix(s,d) - sum mix(d,s)  restricted
            // to the upward pairs, which is all that symmetry of a duplicate-free list needs
            const uint32_t lo = (uint32_t)(s < d ? s : d), hi = (uint32_t)(s < d ? d : s);
            const unsigned long long m0 = pair_mix(lo, hi, 0x9E3779B1u, 0x85EBCA77u);
            const unsigned long long m1 = pair_mix(lo, hi, 0x165667B1u, 0xD3A2646Du);
            if (s < d) { f0 += m0; f1 += m1; } else if (s > d) { f0 -= m0; f1 -= m1; }
        }
        if (e > 0 && (uint64_t)ps >= (uint64_t)n) { atomicOr(flags, 1); return; }
        if (has && !(ps < s || (ps == s && pd < d))) atomicOr(flags, 1);   // not strictly sorted
        if (ps < s) {                                  // e opens row s (and any edge-free rows before it)
            rowptr[s] = (int32_t)e;
            if (rowptr_t) rowptr_t[s] = (int32_t)e;
#pragma unroll 1
            for (int64_t r = ps + 1; r < s; ++r) {
                rowptr[r] = (int32_t)e;
                if (rowptr_t) rowptr_t[r] = (int32_t)e;
            }
        }
    };
    // Two consecutive positions per thread and step: one 16-byte (int32 input: 8-byte) load per
    // index row instead of two loads, the second position's predecessor comes from registers, and
    // both positions' loads are in flight together (the pass is bound by instructions and load
    // latency, not by bandwidth: profiles/r01_k0_fast_build.md).
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & (i32 ? 7 : 15)) == 0;
    for (int64_t p = tid; 2 * p <= e0; p += stride) {
        const int64_t e = 2 * p;
        const bool has_a = e < e0, has_b = e + 1 < e0;
        int64_t sa = n, da = 0, sb = n, db = 0;
        if (vec && has_b) {
            if (i32) {
                const int2 sv = reinterpret_cast<const int2*>(src)[p], dv = reinterpret_cast<const int2*>(dst)[p];
                sa = sv.x; sb = sv.y; da = dv.x; db = dv.y;
            } else {
                const longlong2 sv = reinterpret_cast<const longlong2*>(src)[p];
                const longlong2 dv = reinterpret_cast<const longlong2*>(dst)[p];
                sa = sv.x; sb = sv.y; da = dv.x; db = dv.y;
            }
        } else {
            if (has_a) { sa = ld_idx(src, e, i32); da = ld_idx(dst, e, i32); }
            if (has_b) { sb = ld_idx(src, e + 1, i32); db = ld_idx(dst, e + 1, i32); }
        }
        int64_t ps = -1, pd = -1;
        if (e > 0) { ps = ld_idx(src, e - 1, i32); pd = ld_idx(dst, e - 1, i32); }
        position(e, has_a, sa, da, ps, pd);
        if (e + 1 <= e0) position(e + 1, has_b, sb, db, sa, da);
    }
    // block-level sums of the fingerprints, one atomic pair per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(DGCNN_FULL_MASK, f0, o);
        f1 += __shfl_xor_sync(DGCNN_FULL_MASK, f1, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = f0; red[1][warp] = f1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        if (t) atomicAdd(reinterpret_cast<unsigned long long*>(flags + 2) + threadIdx.x, t);
    }
    if (gptr)
        for (int64_t i = tid; i <= n; i += stride)
            graph_ptr_body(batch, i32, n, num_graphs, gptr, status, i);
}

// After the streaming pass: dis from the row lengths (in-degree == out-degree once the
// list is symmetric), the fingerprint verdict, and the processing order of the graphs.
// Block 0 owns the verdict and the order; all blocks share the dis sweep.
constexpr int kMaxOrderGraphs = 4096;
__global__ void __launch_bounds__(1024)
k0_finalize(int64_t n, const int32_t* __restrict__ rowptr, float* __restrict__ dis, int32_t* flags,
            int check_fingerprints, const int32_t* __restrict__ gptr, int num_graphs,
            int32_t* __restrict__ gorder) {
    DGCNN_PDL_WAIT();
    __shared__ int sizes[kMaxOrderGraphs];
    if (blockIdx.x == 0 && threadIdx.x == 0 && check_fingerprints) {
        const unsigned long long* fp = reinterpret_cast<const unsigned long long*>(flags + 2);
        if (fp[0] != 0ull || fp[1] != 0ull) atomicOr(flags, 1);
    }
    // graphs by DESCENDING size, ties by index: rank sort.  The first `ob` blocks each rank a
    // slice of the graphs against all sizes, `parts` threads per graph.
    const int ob = min((int)gridDim.x, 32);
    if (gorder && num_graphs > 0 && (int)blockIdx.x < ob) {
        if (num_graphs > kMaxOrderGraphs) {
            if (blockIdx.x == 0)
                for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) gorder[g] = g;
        } else {
            for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) sizes[g] = gptr[g + 1] - gptr[g];
            __syncthreads();
            const int chunk = (num_graphs + ob - 1) / ob;
            const int g0 = blockIdx.x * chunk, g1 = min(num_graphs, g0 + chunk);
            const int cnt = max(g1 - g0, 0);
            int parts = 1;
            while (parts < 32 && cnt * parts * 2 <= (int)blockDim.x) parts <<= 1;
            const int lp = 31 - __clz(parts);
            for (int it0 = 0; it0 < cnt * parts; it0 += blockDim.x) {
                const int item = it0 + threadIdx.x;
                const int g = g0 + (item >> lp), part = item & (parts - 1);
                const bool live = g < g1;
                const int mine = live ? sizes[g] : 0;
                int rank = 0;
                if (live)
                    for (int h = part; h < num_graphs; h += parts) {
                        const int other = sizes[h];
                        rank += (other > mine) || (other == mine && h < g);
                    }
                for (int o = parts >> 1; o > 0; o >>= 1) rank += __shfl_xor_sync(DGCNN_FULL_MASK, rank, o);
                if (live && part == 0) gorder[rank] = g;
            }
        }
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dis[i] = 1.0f / sqrtf((float)(rowptr[i + 1] - rowptr[i] + 1));
}

// exact symmetry check (optional): every edge (s,d) must find s in row d (rows are sorted:
// binary search)
__global__ void __launch_bounds__(256)
k0_fast_verify(const void* __restrict__ src, const void* __restrict__ dst, bool i32, int64_t e0, int64_t n,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int32_t* flags) {
    DGCNN_PDL_WAIT();
    if (*flags & 1) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t e = tid; e < e0; e += stride) {
        const int32_t s = (int32_t)ld_idx(src, e, i32);
        const int64_t d = ld_idx(dst, e, i32);
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
    DGCNN_LAUNCH(k0_graph_ptr, grid_for(num_nodes + 1, 256, 8), 256, 0, st, batch, false, num_nodes, num_graphs, gptr,
                                                                 status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

static int build_graph_impl(const void* edge_index, bool i32, int64_t num_edges, const void* batch,
                            int64_t num_nodes, int64_t num_graphs, int32_t* rowptr,
                            int32_t* col, int32_t* rowptr_t, int32_t* col_t, float* dis,
                            int32_t* gptr, int32_t* gorder, int32_t* status,
                            int32_t exact_verify, void* workspace, size_t workspace_bytes,
                            void* stream) {
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
    const void* src = edge_index;
    const void* dst = i32 ? static_cast<const void*>(static_cast<const int32_t*>(edge_index) + e0)
                          : static_cast<const void*>(static_cast<const int64_t*>(edge_index) + e0);

    size_t clear_span = (size_t)((char*)w.outdeg - (char*)w.flags) + sizeof(int32_t) * (size_t)n;
    if (cudaMemsetAsync(w.flags, 0, clear_span, st) != cudaSuccess) return DGCNN_ERR_CUDA;

    // fast path (sorted + symmetric input), verified on the device
    int64_t work = e0 + 1 > n + 1 ? e0 + 1 : n + 1;
    DGCNN_LAUNCH(k0_fast_build, grid_for(work, 256, 8), 256, 0, st, src, dst, e0, batch, i32, n, num_graphs, rowptr,
                                                          col, rowptr_t, col_t, gptr, w.flags, status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (exact_verify) {
        DGCNN_LAUNCH(k0_fast_verify, grid_for(work, 256, 8), 256, 0, st, src, dst, i32, e0, n, rowptr, col, w.flags);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    int fin_grid = grid_for(n + 1, 1024, 1);
    if (gorder && num_graphs > 16) {                 // enough CTAs to rank the graphs in parallel
        const int want = (int)((num_graphs + 15) / 16 < 32 ? (num_graphs + 15) / 16 : 32);
        if (fin_grid < want) fin_grid = want;
    }
    DGCNN_LAUNCH(k0_finalize, fin_grid, 1024, 0, st, n, rowptr, dis, w.flags, exact_verify ? 0 : 1, gptr,
                                           gorder ? (int)num_graphs : 0, gorder);
    DGCNN_RETURN_IF_LAUNCH_FAILED();

    // generic path: one cooperative launch that returns at once unless the flag was raised
    GenericArgs ga;
    ga.src = src; ga.dst = dst; ga.e0 = e0; ga.i32 = i32; ga.batch = batch; ga.n = n; ga.num_graphs = num_graphs;
    ga.indeg = w.indeg; ga.outdeg = transposed ? w.outdeg : nullptr; ga.bsum = w.bsum; ga.nb = w.nb;
    ga.rowptr = rowptr; ga.rowptr_t = rowptr_t; ga.dis = dis;
    ga.tmp_in = w.tmp_in; ga.tmp_out = w.tmp_out; ga.col = col; ga.col_t = col_t;
    ga.status = status; ga.gate = w.flags;
    static int coop_blocks = 0;                      // co-resident CTAs per SM, queried once
    if (coop_blocks == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k0_generic, kScanThreads, 0) != cudaSuccess ||
            per_sm < 1)
            return DGCNN_ERR_CUDA;
        coop_blocks = per_sm > 4 ? 4 : per_sm;
    }
    void* args[] = {&ga};
    if (cudaLaunchCooperativeKernel((const void*)k0_generic, dim3(DGCNN_NUM_SMS * coop_blocks),
                                    dim3(kScanThreads), args, 0, st) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    return DGCNN_OK;
}

extern "C" int dgcnn_build_graph(const int64_t* edge_index, int64_t num_edges, const int64_t* batch,
                                 int64_t num_nodes, int64_t num_graphs, int32_t* rowptr,
                                 int32_t* col, int32_t* rowptr_t, int32_t* col_t, float* dis,
                                 int32_t* gptr, int32_t* gorder, int32_t* status,
                                 int32_t exact_verify, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    return build_graph_impl(edge_index, false, num_edges, batch, num_nodes, num_graphs, rowptr, col,
                            rowptr_t, col_t, dis, gptr, gorder, status, exact_verify, workspace,
                            workspace_bytes, stream);
}

extern "C" int dgcnn_build_graph_i32(const int32_t* edge_index, int64_t num_edges, const int32_t* batch,
                                     int64_t num_nodes, int64_t num_graphs, int32_t* rowptr,
                                     int32_t* col, int32_t* rowptr_t, int32_t* col_t, float* dis,
                                     int32_t* gptr, int32_t* gorder, int32_t* status,
                                     int32_t exact_verify, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    return build_graph_impl(edge_index, true, num_edges, batch, num_nodes, num_graphs, rowptr, col,
                            rowptr_t, col_t, dis, gptr, gorder, status, exact_verify, workspace,
                            workspace_bytes, stream);
}
