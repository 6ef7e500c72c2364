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
    // A graph that alone would outlast the rest of the launch is SPLIT over the two CTAs of a
    // thread-block cluster (forward kernel): each CTA owns half of the 16-row tiles, keeps a full
    // copy of the feature planes and writes its rows into both copies (distributed shared
    // memory); the team is then all warps of BOTH CTAs and sync() is the cluster barrier.
    int split = 0;           // 1: this graph is shared with the peer CTA of the cluster
    int rank = 0;            // this CTA's rank in the pair
    uint32_t mbar = 0;       // split: shared::cta address of this CTA's exchange mbarrier
    uint32_t xparity = 0;    // split: the mbarrier's current phase (carried from graph to graph)
    __device__ __forceinline__ void sync_local() const {
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
    }
    __device__ __forceinline__ void sync() const {
        if (split) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n"
                         "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        } else {
            sync_local();
        }
    }
};

// ---- distributed shared memory exchange of a graph split over a CTA pair -------------------
// Each CTA of the pair finishes its rows of a plane in its OWN shared memory, then one thread
// pushes them into the peer's copy with a bulk async copy (cp.async.bulk shared::cta ->
// shared::cluster, the TMA engine moves the bytes) that signals the PEER's mbarrier with the
// byte count; the peer's threads wait on their mbarrier.  No remote scalar stores, no cluster
// barrier per layer.  One mbarrier per CTA, one phase per exchange.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t cta_addr, uint32_t peer_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(peer_rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    }
}
// push `bytes` at cta_addr into the same place of the peer's shared memory; completion is
// signalled on the peer's mbarrier (peer_mbar = shared::cluster address)
__device__ __forceinline__ void push_to_peer(uint32_t cta_addr, uint32_t bytes, uint32_t peer_rank,
                                             uint32_t peer_mbar) {
    const uint32_t dst = map_to_peer(cta_addr, peer_rank);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(cta_addr), "r"(bytes), "r"(peer_mbar) : "memory");
}

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

// host: can `kernel` be launched as clusters of two CTAs with one CTA per SM, every pair
// co-resident in ONE wave?  (The plan's split graphs need both CTAs of their pair running.)
template <class Kernel>
static inline int cluster_pairs_fit(Kernel kernel, int threads, size_t smem) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t probe{};
    cudaLaunchAttribute pattr[1];
    pattr[0].id = cudaLaunchAttributeClusterDimension;
    pattr[0].val.clusterDim.x = 2; pattr[0].val.clusterDim.y = 1; pattr[0].val.clusterDim.z = 1;
    probe.gridDim = dim3(DGCNN_NUM_SMS & ~1); probe.blockDim = dim3(threads);
    probe.dynamicSmemBytes = smem; probe.attrs = pattr; probe.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &probe) != cudaSuccess || 2 * nclusters < (DGCNN_NUM_SMS & ~1)) {
        cudaGetLastError();
        return 0;
    }
    return 1;
}

// ---- per-CTA plan of the fused per-graph kernels (KS forward, KSB backward) ----------------
// The batch arrives in descending size (gdesc, written by K0b: {graph, first node, nodes,
// fgoff}).  Graphs that alone cost more than an SM's fair share of the batch get an SM to
// themselves; the rest is dealt to the remaining CTAs boustrophedon-wise (s, 2S-1-s, 2S+s,
// ...), so that every SM holds one graph of each size class.  Up to kMaxTeams of a CTA's
// graphs run CONCURRENTLY, each on its own warps (more warps for more row tiles), named
// barrier and slice of shared memory: the per-graph work is a chain of short latency-bound
// phases, and the only way to fill the SM is to overlap the chains of different graphs.
constexpr int kMaxTeams = 8;
struct PlanEntry { int gi, base, n, fgoff, warp0, nwarps, smem_off, pad; };

// rough latency (cycles) of one layer of a T-tile graph on w warps: rounds x (blocks + epilogue)
// (float arithmetic: the plan is a serial chain on one warp, and an integer division by a
// variable costs ~40 dependent instructions)
__device__ __forceinline__ int layer_latency(int T, int w) {
    const int rounds = __float2int_ru(__fdividef((float)T, (float)w) - 1e-4f);
    return rounds * (T * 64 + 1200);
}

// cost of a graph in "16x16 adjacency blocks": T^2 blocks per layer plus a per-row-tile
// share (projection, epilogue, sort, copy) worth ~19 blocks
__device__ __forceinline__ int graph_cost(int n) {
    const int T = (max(n, 1) + 15) >> 4;
    return T * (T + 19);
}

// One pass of the plan, executed by ONE WARP: fills s_plan[0..count) and *s_count with the
// next graphs of this CTA that fit the shared-memory budget together.  `next` = items of this
// CTA consumed so far; `excl` (in/out, lane-uniform) = number of graphs with an SM of their
// own, computed on the first pass.  need_of(np, split) = shared-memory bytes of a graph of np
// rows (split: its share when the graph is spread over a CTA pair).
//
// The SPLIT SET (cluster launches only, `pairs`): the leading `msplit` graphs are each processed
// by the two CTAs of a cluster -- those whose own layout does not fit one CTA (mandatory), and
// those that alone would outlast an SM's fair share of the batch (split_pct).  `nsplit` pairs
// (CTAs 0 .. 2 nsplit - 1) take them round-robin, one graph per pass.
template <class NeedFn>
__device__ __forceinline__ void plan_pass(const int4* __restrict__ gdesc, int B, int nsm, int sm, int next,
                                          int& excl, bool first_pass, int budget, int total_warps,
                                          NeedFn need_of, PlanEntry* s_plan, int* s_count,
                                          int& nsplit, int& msplit, bool pairs = false, int split_pct = 80) {
    const int lane = threadIdx.x & 31;
    int4 cand = make_int4(0, 0, 0, 0);                      // the eight largest graphs
    if (first_pass) {
        excl = 0;
        nsplit = 0;
        msplit = 0;
        // every CTA reads the same few KB of descriptors: pull them into L2 / L1 now, so that the
        // dependent lookup below (position known only after the split decision) is not a second
        // DRAM round trip
        for (int i = lane * 8; i < B; i += 256)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gdesc + i));
        if (lane < 8 && lane < B) cand = gdesc[lane];
        // (both loads issued before the first use: one round trip)
        // fair share of one SM ~ (B / S) x mean cost; the median stands in for the mean
        const int nmed = B > nsm ? gdesc[B >> 1].z : 0;
        const int share = (int)fminf(1.25f * (float)graph_cost(nmed) * (float)B / (float)nsm, 2.0e9f);
        int forced = 0;                                     // graphs that only fit a CTA pair
        if (pairs) {
            const int cap = max(1, min(8, nsm / 4));
            uint32_t m = __ballot_sync(DGCNN_FULL_MASK,
                                       lane < 8 && lane < B && need_of(max(16, (cand.z + 15) & ~15), false) > budget);
            forced = __ffs(~m) - 1;
            if (forced == 8) {                              // (rare: count on through the sorted list)
                for (int i0 = 8; i0 < B; i0 += 32) {
                    const int nz = i0 + lane < B ? gdesc[i0 + lane].z : 0;
                    m = __ballot_sync(DGCNN_FULL_MASK, nz > 0 && need_of(max(16, (nz + 15) & ~15), false) > budget);
                    forced += __ffs(~m) - 1;
                    if (m != 0xffffffffu) break;
                }
            }
            forced = min(forced, B);
            int by_cost = 0;
            if (B > nsm) {                                  // leading run worth two SMs each
                const int thr = (int)fminf((float)share * (float)split_pct * 0.01f, 2.0e9f);
                const uint32_t big = __ballot_sync(DGCNN_FULL_MASK, lane < 8 && lane < B && graph_cost(cand.z) > thr);
                by_cost = min(__ffs(~big) - 1, cap);
            } else {
                // fewer graphs than SMs: the spare CTAs double up on the largest graphs (>= 8 row tiles)
                const uint32_t big = __ballot_sync(DGCNN_FULL_MASK, lane < 8 && lane < B && cand.z > 112);
                by_cost = max(0, min(__ffs(~big) - 1, min(cap, nsm - B)));
            }
            msplit = max(by_cost, forced);
            nsplit = min(msplit, max(by_cost, cap));
            if (2 * nsplit >= nsm) {                       // (toy grids) no CTA left for the rest: the
                nsplit = max(1, nsm / 2);                   // pairs take every graph, pass by pass
                msplit = B;
            }
        }
        if (B > nsm && msplit < 8) {
            const uint32_t big1 = __ballot_sync(DGCNN_FULL_MASK, lane < 8 && lane < B && lane >= msplit &&
                                                                    graph_cost(cand.z) > share) >> msplit;
            excl = max(0, min(__ffs(~big1) - 1, (nsm - 2 * nsplit) / 2));   // then: an SM of their own
        }
    }
    excl = __shfl_sync(DGCNN_FULL_MASK, excl, 0);
    nsplit = __shfl_sync(DGCNN_FULL_MASK, nsplit, 0);
    msplit = __shfl_sync(DGCNN_FULL_MASK, msplit, 0);
    // lane j proposes the CTA's item next + j
    const int item = next + lane;
    int pos;
    const bool is_split = sm < 2 * nsplit;
    const int s1 = sm - 2 * nsplit, n1 = nsm - 2 * nsplit;
    if (is_split) {
        pos = lane == 0 ? (sm >> 1) + next * nsplit : B;    // one graph per pass: the pair works as one team
        if (pos >= msplit) pos = B;
    } else if (s1 < excl) {
        pos = item == 0 ? msplit + s1 : B;
    } else {
        const int s2 = s1 - excl, n2 = n1 - excl;
        pos = msplit + excl + item * n2 + ((item & 1) ? n2 - 1 - s2 : s2);
    }
    const bool valid = lane < kMaxTeams && pos < B;
    int4 d = make_int4(0, 0, 0, 0);
    if (first_pass && (is_split || s1 < excl)) {
        // a graph with SMs of its own is one of the candidates just loaded: no second round trip
        const int src = (is_split ? (sm >> 1) : msplit + s1) & 7;
        d.x = __shfl_sync(DGCNN_FULL_MASK, cand.x, src); d.y = __shfl_sync(DGCNN_FULL_MASK, cand.y, src);
        d.z = __shfl_sync(DGCNN_FULL_MASK, cand.z, src); d.w = __shfl_sync(DGCNN_FULL_MASK, cand.w, src);
        if (!valid) d = make_int4(0, 0, 0, 0);
    } else if (valid) {
        d = gdesc[pos];
    }
    const int n = d.z, np = max(16, (n + 15) & ~15), T = np >> 4;
    const int need = valid ? need_of(np, is_split) : 0;
    int incl = need;
#pragma unroll
    for (int o = 1; o < kMaxTeams; o <<= 1) {
        const int u = __shfl_up_sync(DGCNN_FULL_MASK, incl, o);
        if (lane >= o) incl += u;
    }
    // members = the leading items that fit together (the first one always does: host check)
    const uint32_t fit = __ballot_sync(DGCNN_FULL_MASK, valid && (incl <= budget || lane == 0));
    const int count = __ffs(~fit) - 1;
    const bool member = lane < count;
    // warps: one each plus a share of the rest proportional to the graph's cost (closed form),
    // then whatever is still spare goes, one at a time, to whoever has the longest layer
    int w = member ? 1 : 0;
    if (count > 0) {
        const int cost = member ? graph_cost(n) : 0;
        const int total = (int)__reduce_add_sync(DGCNN_FULL_MASK, (unsigned)cost);
        if (member)
            w = min(T, 1 + (int)((float)(total_warps - count) * (float)cost / (float)max(total, 1) - 1e-3f));
    }
    if (is_split && member) w = total_warps;               // the cluster barrier wants every warp
    const int used = (int)__reduce_add_sync(DGCNN_FULL_MASK, (unsigned)w);
    for (int spare = total_warps - used; spare > 0 && count > 0; --spare) {
        const uint32_t lat = (member && w < T) ? (uint32_t)layer_latency(T, w) : 0u;
        const uint32_t best = __reduce_max_sync(DGCNN_FULL_MASK, (lat << 5) | (uint32_t)(31 - lane));
        if ((best >> 5) == 0u) break;
        if (lane == 31 - (int)(best & 31u)) ++w;
    }
    int winc = w;
#pragma unroll
    for (int o = 1; o < kMaxTeams; o <<= 1) {
        const int u = __shfl_up_sync(DGCNN_FULL_MASK, winc, o);
        if (lane >= o) winc += u;
    }
    if (member) {
        PlanEntry e;
        e.gi = d.x; e.base = d.y; e.n = n; e.fgoff = d.w;
        e.warp0 = winc - w; e.nwarps = w; e.smem_off = incl - need; e.pad = is_split ? 1 : 0;
        s_plan[lane] = e;
    }
    if (lane == 0) *s_count = count;
}

// ---- fragment-major adjacency (K0b, graph_bitmap.cu) and the MMA helpers built on it -------
__host__ __device__ inline int frag_words(int np) {     // fragment-major adjacency of one graph
    const int t = np >> 4;
    return t * ((t + 3) >> 2) * 32;
}


__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}


struct GraphCtx {
    int n, np, T, G, S, base;
    bool dup;
    const uint32_t* fbm;            // shared: fragment-major adjacency
    const int* rp;                  // shared: local row pointers (multigraphs only)
    const int32_t* col_g;           // global: CSR columns of this graph (multigraphs only)
    const float* cs;                // shared: c_j (0 on padding)
    const float* rs;                // shared: 0.5 * r_i (0 on padding); 0.5 undoes A = {0, 2}
};

// mma.sync without `volatile`: a pure function of its operands, so that ptxas may interleave
// the MMAs of one block with the ldmatrix of the next (the loops below are latency-bound).
__device__ __forceinline__ void mma_fp16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Sum over N(i) U {i} of single-column-per-feature planes [F' <= 8][S] (hi at pl, lo at
// pl + lo_off) for the 16-row tile mt of this warp: feature f lands in column f of one 8-wide
// n-tile (C layout); value = 2 * sum (A fragments are {0, 2.0}).
__device__ __forceinline__ void block8(uint32_t w, int q, const uint32_t* __restrict__ hi32,
                                       const uint32_t* __restrict__ lo32, bool live, float (&acc)[4]) {
    uint32_t a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        a[i] = __funnelshift_l(w, w, (14 - (4 * q + i)) & 31) & 0x40004000u;
    const uint32_t h0 = live ? hi32[q * 8] : 0u, h1 = live ? hi32[q * 8 + 4] : 0u;
    const uint32_t l0 = live ? lo32[q * 8] : 0u, l1 = live ? lo32[q * 8 + 4] : 0u;
    mma_fp16(acc, a, h0, h1);
    mma_fp16(acc, a, l0, l1);
}

__device__ __forceinline__ void aggregate8(const GraphCtx& c, const __half* __restrict__ pl, int lo_off,
                                           int nf, int mt, int lane, float (&acc)[4]) {
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    const int g = lane >> 2, t = lane & 3;
    if (!c.dup) {
        const uint32_t* hi32 = reinterpret_cast<const uint32_t*>(pl + g * c.S) + t;
        const uint32_t* lo32 = reinterpret_cast<const uint32_t*>(pl + lo_off + g * c.S) + t;
        const bool live = g < nf;
        const uint32_t* fb = c.fbm + mt * c.G * 32 + lane;
        uint32_t w = fb[0];
        float acc2[4] = {0.f, 0.f, 0.f, 0.f};               // second chain: halves the MMA dependency depth
        for (int grp = 0; grp < c.G; ++grp, hi32 += 32, lo32 += 32) {
            const uint32_t wn = grp + 1 < c.G ? fb[(grp + 1) * 32] : 0u;
            const int nb = c.T - grp * 4;
            if (__any_sync(DGCNN_FULL_MASK, w != 0u)) {
                if (nb >= 4) {
                    block8(w, 0, hi32, lo32, live, acc);
                    block8(w, 1, hi32, lo32, live, acc2);
                    block8(w, 2, hi32, lo32, live, acc);
                    block8(w, 3, hi32, lo32, live, acc2);
                } else {
#pragma unroll 1
                    for (int q = 0; q < nb; ++q) block8(w, q, hi32, lo32, live, acc);
                }
            }
            w = wn;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += acc2[i];
    } else {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int row = mt * 16 + g + 8 * half;
            if (row >= c.n) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int f = 2 * t + u;
                if (f >= nf) continue;
                float s = 0.f;
                for (int e = c.rp[row] - 1; e < c.rp[row + 1]; ++e) {
                    const int j = e < c.rp[row] ? row : c.col_g[e] - c.base;
                    s += __half2float(pl[f * c.S + j]) + __half2float(pl[lo_off + f * c.S + j]);
                }
                acc[2 * half + u] = 2.f * s;
            }
        }
    }
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
