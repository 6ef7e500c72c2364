// KS, tensor-core variant: the fused forward of model.py:28-35 in ONE launch
// (remove_self_loops + 4 x tanh(GCNConv) + cat + SortAggregation), one thread team per graph.
//
// Data flow per graph (n nodes, np = n rounded up to 16, T = np/16 row tiles):
//
//   adjacency   A_hat = A + I is 0/1.  K0b ships it in FRAGMENT-MAJOR form (graph_bitmap.cu):
//               for every 16-row tile and every group of four 16-column blocks one 32-bit word
//               per lane, in which the lane's eight m16k16 A-fragment bits of each block sit
//               at bit m (low half) and 16+m (high half).  A fragment register is then
//               rotate + mask:  rotl(w, 14-m) & 0x40004000  = two fp16 values in {0, 2.0}
//               (the factor 2 is folded into the row coefficient).  8 ALU per 16x16 block.
//   features    must keep fp32 accuracy (1e-5 parity bar): every value is split h = hi + lo,
//               hi = fp16(h), lo = fp16(h - hi) (22 mantissa bits), stored [node][hi 32|lo 32|pad 8]
//               halfs in shared memory (144 B rows: conflict-free for ldmatrix and for the
//               packed half2 epilogue stores).  B fragments come from ldmatrix.x4.trans.
//               Products 0/2 x fp16 are exact, only the fp32 accumulation rounds.
//   projection  (r_i agg) @ W^T reuses the accumulator fragments as A operands (C layout of
//               two n-tiles == A layout of one k-tile), split hi/lo in registers, against
//               hi/lo planes of W (ldmatrix):  hi*hi + lo*hi + hi*lo.
//   layer 1     F <= 8: aggregate first -- one 8-wide n-tile holds all features, scaled per
//               graph by a power of two so that the split never leaves the fp16 range --
//               then an fp32 FMA projection F -> 32.  F > 8: project first (FMA), aggregate
//               on the tensor cores.
//   layer 4     32 -> 1: v = c_i (x_3 . w4) is emitted by layer 3's epilogue; one MMA column.
//   SortPool    64-bit (key, index) composites, rank sort (n <= 640) or bitonic; the k winners
//               are copied row by row (one warp per row) from x_cat (L2 hits) into `pooled`.
//   conv5       SURVEY 8f N2 (h1 != null): z[node][16] = W5 x_cat[node] + b5 accumulated in the layer
//               epilogues (project16), ReLU + pair maximum on the k winners; `pooled` optional.
//   clusters    the kernel is launched as clusters of two CTAs; the plan (graph_mma.cuh) SPLITS
//               the largest graphs over a pair -- mandatorily those that do not fit one CTA's
//               shared memory: each CTA owns half of the row tiles and pushes its rows into the
//               peer's planes with DSMEM bulk copies + mbarrier transactions (pair_exchange).
//   lazy maps   stack_fwd_mma_kernel<true> (the one-call training step): no K0b maps -- every team
//               expands its graph's CSR rows into the fragment-major map itself and exports it
//               for the backward kernel (phase 0 of process_graph).
//
// mma.sync (register fragments) is used on purpose, not tcgen05: the operands are generated
// in registers from a bitmap for graphs of ~30-500 nodes; tcgen05 needs shared-memory operand
// tiles of M >= 64/128 rows in descriptor layouts plus a TMEM round trip per tile, which
// would waste most of each tile on 75-node graphs.  The kernel is bound by issue slots and
// latency, not by the tensor pipe (profiles/).
#include <cooperative_groups.h>
#include <cstdlib>

#include "graph_mma.cuh"
#include "sort_key.cuh"

namespace cg = cooperative_groups;

namespace dgcnn {

// The forward kernel runs FWD_THREADS threads per CTA (default 640 = 20 warps at <= 96
// registers): the largest graphs of a batch bound the launch, and a 19-tile graph then does
// one round of row tiles per layer instead of two.
#ifndef DGCNN_FWD_THREADS
#define DGCNN_FWD_THREADS 640
#endif
constexpr int kFwdThreads = DGCNN_FWD_THREADS;
constexpr int kRowH = 72;                  // halfs per node row of a feature plane pair
constexpr int kRowB = kRowH * 2;           // 144 bytes
constexpr int kRankSortMax = 640;          // rank sort (one pass, no barriers) up to this many nodes

// CTA-wide region: weights, shared by all teams.  Byte offsets, 16-byte aligned.
constexpr int kC5 = 16;                    // conv5 output channels (model.py:19)
constexpr int kW5Pad = 104;                // row stride (halfs) of the W5 planes: 208 B, conflict-free ldmatrix

constexpr int kZeroBytes = 4096;          // zeroed staging buffer: source of the bulk stores that pad `pooled`

struct SharedLayout { int w2p, w3p, w1t, misc, w5p, w5x, zero, total; };

__host__ __device__ inline SharedLayout shared_layout(int f, bool conv5 = false) {
    SharedLayout L;
    int o = 0;
    L.w2p = o; o += 2 * kHid * kWPad * 2;                // [plane][cout][kWPad] fp16
    L.w3p = o; o += 2 * kHid * kWPad * 2;
    L.w1t = o; o += al16(f * kHid * 4);                  // [F][32] fp32
    L.misc = o; o += 4 * kHid * 4;                       // w4, b1, b2, b3
    L.w5p = o; o += conv5 ? 2 * kC5 * kW5Pad * 2 : 0;    // [plane][16][kW5Pad] fp16: W5[:, 0..95]
    L.w5x = o; o += conv5 ? 2 * kC5 * 4 : 0;             // W5[:, 96] and b5, fp32
    L.zero = o; o += kZeroBytes;
    L.total = o;
    return L;
}

// Per-graph region, sized by the graph's own padded node count np (multiple of 16).
struct TeamLayout {
    int PA, PB, vpl, fbm, xs, cs, rs, rp, zs, total;
    int S;      // row stride (halfs) of the single-column planes (vpl, xs): np + 8
};

// (split: the graph is spread over a CTA pair; each CTA keeps full copies of the planes but only
// the adjacency fragments of its own row tiles -- the first or last ceil(T / 2) of them)
__host__ __device__ inline TeamLayout team_layout(int f, int np, bool conv5 = false, bool split = false) {
    TeamLayout L;
    L.S = np + 8;
    int o = 0;
    L.PA = o; o += np * kRowB;
    L.PB = o; o += np * kRowB;
    L.vpl = o; o += al16(2 * L.S * 2);                   // layer-4 input: hi and lo [S]
    {
        const int T = np >> 4, own = split ? (T + 1) >> 1 : T;
        L.fbm = o; o += own * ((T + 3) >> 2) * 32 * 4;
    }
    L.xs = o; o += (f <= kSmallF) ? al16(2 * f * L.S * 2) : 0;   // layer-1 input planes [2][F][S]
    L.cs = o; o += al16(np * 4);
    L.rs = o; o += al16(np * 4);
    L.rp = o; o += al16((np + 1) * 4);
    L.zs = o; o += conv5 ? np * kC5 * 4 : 0;             // conv5 pre-activations per node, fp32 [np][16]
    L.total = o;
    return L;
}

// bytes of one quad's slice when the CTA takes (almost) all of the SM's shared memory
__host__ __device__ inline int quad_bytes(int f, bool conv5 = false) {
    return ((kSmemBudget - 1024 - shared_layout(f, conv5).total) / kQuads) & ~15;
}

// tanh for the hidden layers: (1 - t) / (1 + t), t = 2^(-2 log2(e) |v|); two MUFU ops, no
// branch.  Absolute error <= ~1.5e-7 (ex2.approx is 2^-22 relative), NaN propagates.
// The sort key (layer 4) uses tanhf.
__device__ __forceinline__ float tanh_hidden(float v) {
    float t, r;
    const float a = fabsf(v) * -2.885390081777927f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return copysignf((1.0f - t) * r, v);
}

// packed hi / lo halves of (x0, x1), x0 in the low half
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Optional per-graph timeline (debug hook dgcnn_stack_fwd_set_trace): clock64 at phase ends.
__device__ __forceinline__ int64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (int64_t)t;
}

// one 16x16 adjacency block (q-th of its word) times the 16x32 hi and lo feature blocks
__device__ __forceinline__ void block32(uint32_t w, int q, uint32_t addr, float (&acc)[4][4]) {
    uint32_t a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        a[i] = __funnelshift_l(w, w, (14 - (4 * q + i)) & 31) & 0x40004000u;
    uint32_t b[4][4];
    ldsm_x4_t(addr, b[0][0], b[0][1], b[0][2], b[0][3]);              // hi, channels 0..15
    ldsm_x4_t(addr + 32, b[1][0], b[1][1], b[1][2], b[1][3]);         // hi, channels 16..31
    ldsm_x4_t(addr + 64, b[2][0], b[2][1], b[2][2], b[2][3]);         // lo, channels 0..15
    ldsm_x4_t(addr + 96, b[3][0], b[3][1], b[3][2], b[3][3]);         // lo, channels 16..31
    mma_fp16(acc[0], a, b[0][0], b[0][1]);
    mma_fp16(acc[1], a, b[0][2], b[0][3]);
    mma_fp16(acc[2], a, b[1][0], b[1][1]);
    mma_fp16(acc[3], a, b[1][2], b[1][3]);
    mma_fp16(acc[0], a, b[2][0], b[2][1]);
    mma_fp16(acc[1], a, b[2][2], b[2][3]);
    mma_fp16(acc[2], a, b[3][0], b[3][1]);
    mma_fp16(acc[3], a, b[3][2], b[3][3]);
}

// Sum over N(i) U {i} of the 32-wide rows of a plane pair, for the 16-row tile mt of this
// warp: acc[nt][..] in the m16n8 C layout, value = 2 * sum (both paths).
// Full groups of four blocks run branch-free (one big basic block: loads and MMAs overlap);
// a group whose four blocks are all empty is skipped, which is what sparse graphs need.
__device__ __forceinline__ void aggregate32(const GraphCtx& c, const __half* __restrict__ in_pl, int mt,
                                            int lane, float (&acc)[4][4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    if (!c.dup) {
        const uint32_t lane_off = (((lane & 7) + ((lane >> 3) & 1) * 8) * kRowH + (lane >> 4) * 8) * 2;
        uint32_t addr = smem_u32(in_pl) + lane_off;
        const uint32_t* fb = c.fbm + mt * c.G * 32 + lane;
        uint32_t w = fb[0];
        for (int grp = 0; grp < c.G; ++grp, addr += 64 * kRowB) {
            const uint32_t wn = grp + 1 < c.G ? fb[(grp + 1) * 32] : 0u;
            const int nb = c.T - grp * 4;
            if (__any_sync(DGCNN_FULL_MASK, w != 0u)) {
                if (nb >= 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) block32(w, q, addr + q * 16 * kRowB, acc);
                } else {
#pragma unroll 1
                    for (int q = 0; q < nb; ++q) block32(w, q, addr + q * 16 * kRowB, acc);
                }
            }
            w = wn;
        }
    } else {
        // multigraph: duplicates have no bitmap encoding (PyG counts them): walk the CSR
        const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int row = mt * 16 + g + 8 * half;
            if (row >= c.n) continue;
            for (int e = c.rp[row] - 1; e < c.rp[row + 1]; ++e) {
                const int j = e < c.rp[row] ? row : c.col_g[e] - c.base;  // first the self loop
                const __half2* r2 = reinterpret_cast<const __half2*>(in_pl + j * kRowH);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float2 h = __half22float2(r2[nt * 4 + t]);
                    const float2 l = __half22float2(r2[16 + nt * 4 + t]);
                    if (half == 0) { acc[nt][0] += 2.f * (h.x + l.x); acc[nt][1] += 2.f * (h.y + l.y); }
                    else           { acc[nt][2] += 2.f * (h.x + l.x); acc[nt][3] += 2.f * (h.y + l.y); }
                }
            }
        }
    }
}

// y += (acc) @ W^T on the tensor cores; acc already scaled by r_i.  wp: [plane][32][kWPad].
__device__ __forceinline__ void project32(const float (&acc)[4][4], const __half* __restrict__ wp,
                                          int lane, float (&y)[4][4]) {
    const int j = lane >> 3;
    const uint32_t wbase = smem_u32(wp) +
        (uint32_t)(((j >> 1) * kHid * kWPad + (lane & 7) * kWPad + (j & 1) * 8) * 2);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        uint32_t ah[4], al[4];
        split_pair(acc[2 * kk][0], acc[2 * kk][1], ah[0], al[0]);
        split_pair(acc[2 * kk][2], acc[2 * kk][3], ah[1], al[1]);
        split_pair(acc[2 * kk + 1][0], acc[2 * kk + 1][1], ah[2], al[2]);
        split_pair(acc[2 * kk + 1][2], acc[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t h0, h1, l0, l1;
            ldsm_x4(wbase + (uint32_t)((nt * 8 * kWPad + kk * 16) * 2), h0, h1, l0, l1);
            mma_fp16(y[nt], ah, h0, h1);
            mma_fp16(y[nt], al, h0, h1);
            mma_fp16(y[nt], ah, l0, l1);
        }
    }
}

// z += x_l @ W5[:, slice]^T on the tensor cores: x_l = the layer's tanh outputs in
// the accumulator fragments (C layout of two n-tiles == A layout of one k-tile), split hi/lo;
// W5 as hi/lo planes [plane][16][kW5Pad] (columns 0..95 of conv5's [16,97] weight).
__device__ __forceinline__ void project16(const float (&y)[4][4], const __half* __restrict__ w5p, int slice,
                                          int lane, float (&z)[2][4]) {
    const int j = lane >> 3;
    const uint32_t wbase = smem_u32(w5p) +
        (uint32_t)(((j >> 1) * kC5 * kW5Pad + (lane & 7) * kW5Pad + (j & 1) * 8 + slice * kHid) * 2);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        uint32_t ah[4], al[4];
        split_pair(y[2 * kk][0], y[2 * kk][1], ah[0], al[0]);
        split_pair(y[2 * kk][2], y[2 * kk][3], ah[1], al[1]);
        split_pair(y[2 * kk + 1][0], y[2 * kk + 1][1], ah[2], al[2]);
        split_pair(y[2 * kk + 1][2], y[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            uint32_t h0, h1, l0, l1;
            ldsm_x4(wbase + (uint32_t)((nt * 8 * kW5Pad + kk * 16) * 2), h0, h1, l0, l1);
            mma_fp16(z[nt], ah, h0, h1);
            mma_fp16(z[nt], al, h0, h1);
            mma_fp16(z[nt], ah, l0, l1);
        }
    }
}

// tanh, x_l slice of x_cat to HBM, c_i * x_l as hi/lo planes for the next layer (out_pl),
// and v = c_i (x_l . w4) for layer 4 (vpl).  y holds the pre-activations (C layout).
__device__ __forceinline__ void layer_epilogue(const GraphCtx& c, float (&y)[4][4], int mt, int lane,
                                               float* __restrict__ xo, int64_t ldc, bool vec2,
                                               __half* __restrict__ out_pl, __half* __restrict__ vpl,
                                               const float* __restrict__ w4s) {
    const int g = lane >> 2, t = lane & 3;
    const int row0 = mt * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        y[nt][0] = tanh_hidden(y[nt][0]); y[nt][1] = tanh_hidden(y[nt][1]);
        y[nt][2] = tanh_hidden(y[nt][2]); y[nt][3] = tanh_hidden(y[nt][3]);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int row = half ? row1 : row0;
        if (row < c.n) {
            float* o = xo + (int64_t)row * ldc + 2 * t;
            if (vec2) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    *reinterpret_cast<float2*>(o + nt * 8) = make_float2(y[nt][2 * half], y[nt][2 * half + 1]);
            } else {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { o[nt * 8] = y[nt][2 * half]; o[nt * 8 + 1] = y[nt][2 * half + 1]; }
            }
        }
    }
    const float c0 = c.cs[row0], c1 = c.cs[row1];                         // 0 on padding rows
    if (out_pl) {
        uint32_t* o0 = reinterpret_cast<uint32_t*>(out_pl + row0 * kRowH) + t;
        uint32_t* o1 = reinterpret_cast<uint32_t*>(out_pl + row1 * kRowH) + t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t hi, lo;
            split_pair(c0 * y[nt][0], c0 * y[nt][1], hi, lo);
            o0[nt * 4] = hi; o0[16 + nt * 4] = lo;
            split_pair(c1 * y[nt][2], c1 * y[nt][3], hi, lo);
            o1[nt * 4] = hi; o1[16 + nt * 4] = lo;
        }
    }
    if (vpl) {
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float2 w = *reinterpret_cast<const float2*>(w4s + nt * 8 + 2 * t);
            p0 = fmaf(y[nt][0], w.x, fmaf(y[nt][1], w.y, p0));
            p1 = fmaf(y[nt][2], w.x, fmaf(y[nt][3], w.y, p1));
        }
        p0 += __shfl_xor_sync(DGCNN_FULL_MASK, p0, 1);
        p0 += __shfl_xor_sync(DGCNN_FULL_MASK, p0, 2);
        p1 += __shfl_xor_sync(DGCNN_FULL_MASK, p1, 1);
        p1 += __shfl_xor_sync(DGCNN_FULL_MASK, p1, 2);
        if (t == 0) {
            store_split(vpl, vpl + c.S, row0, c0 * p0);
            store_split(vpl, vpl + c.S, row1, c1 * p1);
        }
    }
}

// zeros over [beg, end) of a float array that is only 4-byte aligned (rows of 97 floats): the
// 16-byte aligned body goes out as bulk async stores (TMA engine, cp.async.bulk shared -> global)
// from a zeroed staging buffer, issued by the lanes of the team's first warp -- no store traffic
// through the LSU, nothing for the team to wait for; head and tail (< 4 floats each) are plain
// stores.  Returns true on the lanes that issued bulk stores (they call bulk_store_wait() before
// the kernel ends).
__device__ __forceinline__ bool zero_fill_bulk(float* __restrict__ base, int beg, int end, int tid,
                                               uint32_t zero_smem, int nthreads, bool plain) {
    if (beg >= end) return false;
    if (plain) {                                         // (debug: ordinary stores, see StackFwdParams)
        for (int i = beg + tid; i < end; i += nthreads) base[i] = 0.f;
        return false;
    }
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(base + beg);
    const int head = min(end - beg, (int)(((16u - (unsigned)(a0 & 15u)) & 15u) >> 2));
    const int body = ((end - beg - head) >> 2) << 4;                  // bytes, multiple of 16
    const int tail0 = beg + head + (body >> 2);
    if (tid < head) base[beg + tid] = 0.f;
    if (tid < end - tail0) base[tail0 + tid] = 0.f;
    bool issued = false;
    if (tid < 32 && tid * kZeroBytes < body) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        char* g = reinterpret_cast<char*>(base + beg + head);
        for (int off = tid * kZeroBytes; off < body; off += 32 * kZeroBytes) {
            const int bytes = min(kZeroBytes, body - off);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(g + off), "r"(zero_smem), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        issued = true;
    }
    return issued;
}
__device__ __forceinline__ void bulk_store_wait() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void load_bias(const float* __restrict__ bias, int t, float (&y)[4][4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float2 b = *reinterpret_cast<const float2*>(bias + nt * 8 + 2 * t);
        y[nt][0] = b.x; y[nt][1] = b.y; y[nt][2] = b.x; y[nt][3] = b.y;
    }
}

// model.py:28-35 for ONE graph, executed by one team.  A SPLIT graph (tm.split) is executed by
// the two CTAs of a cluster together: CTA `tm.rank` owns the row tiles [t0, t1), every plane /
// key / order store of its rows also goes into the peer's shared memory, and the team barrier
// is the cluster barrier.
template <bool LAZY>
__device__ __forceinline__ void process_graph(const StackFwdParams& p, Team& tm, const PlanEntry& e,
                                              const unsigned char* shraw) {
    const int tid = tm.tid, lane = tm.lane, warp = tm.warp;
    const int nthreads = tm.nthreads, nwarps = tm.nwarps;
    const int f = p.f;
    const bool small_f = f <= kSmallF;
    const bool conv5 = p.h1 != nullptr;                  // conv5 + ReLU + max-pool fused in
    const SharedLayout SL = shared_layout(f, conv5);
    const __half* w2p = reinterpret_cast<const __half*>(shraw + SL.w2p);
    const __half* w3p = reinterpret_cast<const __half*>(shraw + SL.w3p);
    const float* w1t = reinterpret_cast<const float*>(shraw + SL.w1t);
    const float* w4s = reinterpret_cast<const float*>(shraw + SL.misc);
    const float* b1s = w4s + kHid;                       // b1, b2, b3 contiguous
    const __half* w5p = reinterpret_cast<const __half*>(shraw + SL.w5p);
    const float* w5x = reinterpret_cast<const float*>(shraw + SL.w5x);      // W5[:, 96] | b5
    const float b4 = p.b4 ? p.b4[0] : 0.f;

    const int gi = e.gi, base = e.base, n = e.n;
    const int keep = min(n, p.k);
    float* __restrict__ pooled_g = p.pooled + (int64_t)gi * p.k * kCat;
    int32_t* __restrict__ perm_g = p.perm + (int64_t)gi * p.k;
    // the two CTAs of a split graph act as one team of 2 x nthreads threads wherever the work is
    // per element (sort, copy-out, padding): gtid / gthreads / gwarp / gwarps
    const bool split = tm.split != 0;
    const int rank = split ? tm.rank : 0, ways = split ? 2 : 1;
    const int gtid = tid + rank * nthreads, gthreads = ways * nthreads;
    const int gwarp = warp + rank * nwarps, gwarps = ways * nwarps;
    const bool tracer = p.trace && tm.tid == 0 && rank == 0;
#define KS_TRACE(slot) do { if (tracer) p.trace[(int64_t)gi * 16 + (slot)] = clock64(); } while (0)
    // rows of `pooled` past the graph's last node: zeros (PyG's fill trick), perm -1.  Plain
    // stores, issued first: they drain while the team waits for its inputs.
    bool bulk_pending = false;
    if (p.pooled) {
        const int z0 = keep * kCat, z1 = p.k * kCat, zm = split ? z0 + (((z1 - z0) >> 1) & ~3) : z1;
        bulk_pending = zero_fill_bulk(pooled_g, rank ? zm : z0, rank ? z1 : zm, tid,
                                      smem_u32(shraw + SL.zero), nthreads, p.plain_zero != 0);
    }
    for (int r = keep + gtid; r < p.k; r += gthreads) perm_g[r] = -1;
    if (n == 0) {
        if (bulk_pending) bulk_store_wait();
        if (conv5) {                                   // every row is padding: z = b5
            const int L1 = p.k >> 1;
            for (int item = tid; item < kC5 * L1; item += nthreads) {
                const float z = fmaxf(w5x[kC5 + item / L1], 0.f);
                p.h1[(int64_t)gi * kC5 * L1 + item] = z;
                p.arg[(int64_t)gi * kC5 * L1 + item] = (uint8_t)(z <= 0.f ? 2 : 0);
            }
        }
        return;
    }
    if (tracer) {
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[(int64_t)gi * 16 + 15] = ((int64_t)smid << 32) | (uint32_t)(tm.nthreads | (n << 12));
        p.trace[(int64_t)gi * 16 + 9] = global_ns();
    }
    KS_TRACE(0);

    GraphCtx c;
    c.n = n; c.np = (n + 15) & ~15; c.T = c.np >> 4; c.G = (c.T + 3) >> 2; c.base = base;
    const int np = c.np;
    const int t0 = split ? (rank ? (c.T + 1) >> 1 : 0) : 0;          // this CTA's row tiles
    const int t1 = split ? (rank ? c.T : (c.T + 1) >> 1) : c.T;
    const TeamLayout L = team_layout(f, np, conv5, split);
    c.S = L.S;
    const int S = L.S;
    unsigned char* smraw = tm.smem;
    float* zs = reinterpret_cast<float*>(smraw + L.zs);  // conv5 pre-activations per node [np][16]
    __half* PA = reinterpret_cast<__half*>(smraw + L.PA);
    __half* PB = reinterpret_cast<__half*>(smraw + L.PB);
    __half* vpl = reinterpret_cast<__half*>(smraw + L.vpl);
    // (a split graph holds only its own row tiles' fragments: indexed by tile like the full map)
    uint32_t* fbm = reinterpret_cast<uint32_t*>(smraw + L.fbm) - (split ? t0 * c.G * 32 : 0);
    __half* xs = reinterpret_cast<__half*>(smraw + L.xs);
    float* cs = reinterpret_cast<float*>(smraw + L.cs);
    float* rs = reinterpret_cast<float*>(smraw + L.rs);
    float* key = rs;                                    // x_4 overwrites r_i in place
    int* order = reinterpret_cast<int*>(cs);            // c_j is dead after layer 3
    int* rp = reinterpret_cast<int*>(smraw + L.rp);
    float* stage = reinterpret_cast<float*>(PB);        // layer-1 inputs c_j x_j, fp32 [n][f] (F <= 8)
    float* wmax = reinterpret_cast<float*>(vpl);        // per-warp max |c_j x_j|
    c.fbm = fbm; c.cs = cs; c.rs = rs; c.rp = rp;
    // split graph: rows are exchanged with bulk copies (pair_exchange below); only the sort ranks
    // are scattered remote stores
    int* order_peer = nullptr;
    uint32_t peer_mbar = 0;
    uint32_t xparity = tm.xparity;                       // (the mbarrier's phase outlives the graph)
    const uint32_t peer = (uint32_t)(rank ^ 1);
    if (split) {
        cg::cluster_group cluster = cg::this_cluster();
        order_peer = cluster.map_shared_rank(order, peer);
        peer_mbar = map_to_peer(tm.mbar, peer);
    }
    const int my_rows = (t1 - t0) * 16, peer_rows = np - my_rows, row0 = t0 * 16;
    // push my rows of up to three row-major regions (row_bytes each) into the peer's copies and
    // wait for the peer's rows; replaces the team barrier at the end of a phase
    auto pair_exchange = [&](const void* b0, int rb0, const void* b1, int rb1, const void* b2, int rb2) {
        tm.sync_local();                                 // my rows are complete in MY shared memory
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(tm.mbar, (uint32_t)(peer_rows * (rb0 + rb1 + rb2)));
            push_to_peer(smem_addr_u32(b0) + (uint32_t)(row0 * rb0), (uint32_t)(my_rows * rb0), peer, peer_mbar);
            if (rb1) push_to_peer(smem_addr_u32(b1) + (uint32_t)(row0 * rb1), (uint32_t)(my_rows * rb1), peer, peer_mbar);
            if (rb2) push_to_peer(smem_addr_u32(b2) + (uint32_t)(row0 * rb2), (uint32_t)(my_rows * rb2), peer, peer_mbar);
        }
        mbar_wait(tm.mbar, xparity);
        xparity ^= 1u;
    };

    float* xc = p.xcat + (int64_t)base * p.ldc;
    const bool vec2 = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.xcat) & 7) == 0);
    const bool pad4 = ((p.ldc & 3) == 0) && p.ldc >= 100 && ((reinterpret_cast<uintptr_t>(p.xcat) & 15) == 0);
    const int g = lane >> 2, t = lane & 3;

    // ---- phase 0, one DRAM round trip: adjacency fragments (K0b), coefficients, inputs ----
    constexpr bool lazy = LAZY;                          // (a kernel of its own: StackFwdParams::lazy)
    int* dupflag = reinterpret_cast<int*>(vpl) + 23;     // lazy: "this graph has duplicate edges" (wmax uses [0, 20))
    int e0 = 0;
    if (!lazy) {
        c.dup = (p.gflags[gi] & 1) != 0;             // multigraph: walk the CSR instead
        load_bitmap<8>(p.fragmap + e.fgoff + t0 * c.G * 32, fbm + t0 * c.G * 32, (t1 - t0) * c.G * 32, tid,
                       nthreads);                        // (only the CTA's own row tiles)
    } else {
        // no K0b maps: clear the own tiles' words and fetch the graph's row pointers; the rows are
        // expanded below, once these are in shared memory
        uint32_t* own = fbm + t0 * c.G * 32;
        for (int idx = tid; idx < (t1 - t0) * c.G * 32; idx += nthreads) own[idx] = 0u;
        e0 = p.rowptr[base];
        for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr[base + j] - e0;
        if (tid == 0) *dupflag = 0;
    }
    for (int j = tid; j < np; j += nthreads) {
        const float d = j < n ? p.dis[base + j] : 0.f;
        cs[j] = j < n ? col_coef(d, p.norm) : 0.f;
        rs[j] = j < n ? 0.5f * row_coef(d, p.norm) : 0.f;
    }
    if (small_f) {
        float m = 0.f;
        for (int idx = tid; idx < n * f; idx += nthreads) {
            const int j = idx / f, k = idx - j * f;
            const float v = col_coef(p.dis[base + j], p.norm) * p.x[(int64_t)(base + j) * p.ldx + k];
            stage[idx] = v;
            m = fmaxf(m, fabsf(v));                  // NaN-ignoring: a NaN flows through hi/lo instead
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(DGCNN_FULL_MASK, m, o));
        if (lane == 0) wmax[warp] = m;
    }
    if (!lazy && c.dup) {
        e0 = p.rowptr[base];
        for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr[base + j] - e0;
    }
    c.col_g = p.col + e0;
    if (lazy) {
        // CSR rows -> fragment-major bits (the layout k0b_fragments writes: graph_bitmap.cu): one warp
        // per own row, 32 edges per step, shared-memory atomics; word (mt, grp, lane = 4 (r & 7) + t),
        // bit m + 16 (column odd), m = 4 q + (row >= 8) + 2 (column >= 8) inside the 16 x 16 block.
        // Duplicate edges are adjacent in a K0 row: such a graph walks its CSR instead (c.dup).
        tm.sync_local();
        const int* col_g = c.col_g;
        bool dup = false;
        const int r_hi = min(t1 * 16, n);
        auto set_bit = [&](int r, int j) {
            const int kt = j >> 4, cc = j & 15;
            atomicOr(fbm + ((r >> 4) * c.G + (kt >> 2)) * 32 + 4 * (r & 7) + ((cc & 7) >> 1),
                     1u << (4 * (kt & 3) + ((r >> 3) & 1) + 2 * (cc >> 3) + 16 * (cc & 1)));
        };
        // A warp takes eight consecutive rows at a time: their edges are one contiguous run of col[],
        // read 8 x 32 at a time with every load in flight (a row at a time was one L2 round trip per
        // row); the row of an edge = how many of the group's row boundaries lie at or below it.
        constexpr int R = 8, U = 8;
        for (int r0 = t0 * 16 + warp * R; r0 < r_hi; r0 += nwarps * R) {
            const int rows = min(R, r_hi - r0);
            const int beg = rp[r0], end = rp[r0 + rows];
            int bnd[R - 1];
#pragma unroll
            for (int i = 1; i < R; ++i) bnd[i - 1] = i < rows ? rp[r0 + i] : 0x7fffffff;
            int prev_j = -1, prev_r = -1;                         // last edge of the previous chunk
            for (int c0 = beg; c0 < end; c0 += 32 * U) {
                int cv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int ee = c0 + 32 * u + lane;
                    cv[u] = ee < end ? col_g[ee] : -1;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (c0 + 32 * u >= end) break;                // (warp-uniform)
                    const int ee = c0 + 32 * u + lane;
                    int j = -1, r = r0;
                    if (ee < end) {
                        const unsigned tcol = (unsigned)(cv[u] - base);
                        if (tcol < (unsigned)n) j = (int)tcol;   // edges leaving the graph are ignored (K0 flags them)
#pragma unroll
                        for (int i = 0; i < R - 1; ++i) r += ee >= bnd[i];
                    }
                    int lj = __shfl_up_sync(DGCNN_FULL_MASK, j, 1), lr = __shfl_up_sync(DGCNN_FULL_MASK, r, 1);
                    if (lane == 0) { lj = prev_j; lr = prev_r; }
                    if (j >= 0 && j == lj && r == lr) dup = true; // duplicates are adjacent in a K0 row
                    prev_j = __shfl_sync(DGCNN_FULL_MASK, j, 31);
                    prev_r = __shfl_sync(DGCNN_FULL_MASK, r, 31);
                    if (j >= 0) set_bit(r, j);
                }
            }
            if (lane < rows) set_bit(r0 + lane, r0 + lane);       // the self loops
        }
        if (__any_sync(DGCNN_FULL_MASK, dup) && lane == 0) *dupflag = 1;
    }
    tm.sync();            // (split: also tells that the peer CTA runs -- its shared memory may be written)
    if (lazy) {
        // (split: either half may hold the duplicates; the peer's flag is final after the barrier and
        // its slot is not reused before the next exchange, which needs this CTA)
        c.dup = *dupflag != 0 || (split && *cg::this_cluster().map_shared_rank(dupflag, peer) != 0);
        // export for the backward kernel (fire and forget)
        uint32_t* out = p.fragmap_w + e.fgoff + t0 * c.G * 32;
        const uint32_t* own = fbm + t0 * c.G * 32;
        for (int idx = tid; idx < (t1 - t0) * c.G * 32; idx += nthreads) out[idx] = own[idx];
        if (c.dup && tid == 0 && rank == 0) atomicOr(p.gflags_w + gi, 1);
    }
    KS_TRACE(1);

    float pow2 = 1.f;
    if (small_f) {
        // layer-1 inputs scaled by a power of two just below their maximum, so that the
        // fp16 hi/lo split keeps 22 bits relative to the graph's own range: values in (-2, 2)
        float m = 0.f;
        for (int w = 0; w < nwarps; ++w) m = fmaxf(m, wmax[w]);
        uint32_t eb = __float_as_uint(m) & 0x7F800000u;
        if (eb == 0u || eb >= 0x7E800000u) {
            if (eb >= 0x7E800000u && tid == 0 && p.status) atomicOr(p.status, DGCNN_GRAPH_RANGE);
            eb = 0x3F800000u;
        }
        pow2 = __uint_as_float(eb);
        const float inv = __uint_as_float(0x7F000000u - eb);
        for (int idx = tid; idx < f * S; idx += nthreads) {
            const int k = idx / S, j = idx - k * S;
            store_split(xs, xs + f * S, idx, j < n ? stage[j * f + k] * inv : 0.f);
        }
    } else {
        // F > 8: project first (FMA), c_j (x_j W1^T) as planes in PB (a split graph: both CTAs
        // compute all rows -- cheaper than an exchange)
        for (int j = warp; j < np; j += nwarps) {
            float acc = 0.f;
            if (j < n) {
                const float* xr = p.x + (int64_t)(base + j) * p.ldx;
                for (int k = 0; k < f; ++k) acc = fmaf(xr[k], w1t[k * kHid + lane], acc);
                acc *= cs[j];
                if (fabsf(acc) > 6.0e4f && p.status) atomicOr(p.status, DGCNN_GRAPH_RANGE);
            }
            store_split(PB + j * kRowH, PB + j * kRowH + kHid, lane, acc);
        }
    }
    tm.sync_local();
    KS_TRACE(2);

    // ---- layers 1..3: aggregate on the tensor cores, project, tanh ------------------------
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
        const __half* in_pl = layer == 1 ? PA : PB;
        __half* out_pl = layer == 0 ? PA : (layer == 1 ? PB : nullptr);
        const __half* wp = layer == 1 ? w2p : w3p;
        const float* bias = b1s + layer * kHid;
        float* xo = xc + layer * kHid;
#pragma unroll 1
        for (int mt = t0 + warp; mt < t1; mt += nwarps) {
            float y[4][4];
            load_bias(bias, t, y);
            if (layer == 0 && small_f) {
                float a4[4];
                aggregate8(c, xs, f * S, f, mt, lane, a4);
                const float r0 = rs[mt * 16 + g] * pow2, r1 = rs[mt * 16 + g + 8] * pow2;
                a4[0] *= r0; a4[1] *= r0; a4[2] *= r1; a4[3] *= r1;
                for (int k = 0; k < f; ++k) {
                    const int src = (lane & ~3) | (k >> 1);
                    const float v0 = __shfl_sync(DGCNN_FULL_MASK, (k & 1) ? a4[1] : a4[0], src);
                    const float v1 = __shfl_sync(DGCNN_FULL_MASK, (k & 1) ? a4[3] : a4[2], src);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const float2 w = *reinterpret_cast<const float2*>(w1t + k * kHid + nt * 8 + 2 * t);
                        y[nt][0] = fmaf(v0, w.x, y[nt][0]); y[nt][1] = fmaf(v0, w.y, y[nt][1]);
                        y[nt][2] = fmaf(v1, w.x, y[nt][2]); y[nt][3] = fmaf(v1, w.y, y[nt][3]);
                    }
                }
            } else {
                float acc[4][4];
                aggregate32(c, in_pl, mt, lane, acc);
                const float r0 = rs[mt * 16 + g], r1 = rs[mt * 16 + g + 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    acc[nt][0] *= r0; acc[nt][1] *= r0; acc[nt][2] *= r1; acc[nt][3] *= r1;
                }
                if (layer == 0) {
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        y[nt][0] += acc[nt][0]; y[nt][1] += acc[nt][1];
                        y[nt][2] += acc[nt][2]; y[nt][3] += acc[nt][3];
                    }
                } else {
                    project32(acc, wp, lane, y);
                }
            }
            layer_epilogue(c, y, mt, lane, xo, p.ldc, vec2, out_pl, layer == 2 ? vpl : nullptr, w4s);
            if (conv5) {
                // y now holds x_l of this tile: z += x_l W5[:, 32 l .. 32 l + 31]^T (+ b5 first time).
                // The same thread owns the same (row, channel) entries in every layer: no barrier.
                float z[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) z[nt][0] = z[nt][1] = z[nt][2] = z[nt][3] = 0.f;
                project16(y, w5p, layer, lane, z);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int col = nt * 8 + 2 * t;
                    float2* a0 = reinterpret_cast<float2*>(zs + (mt * 16 + g) * kC5 + col);
                    float2* a1 = reinterpret_cast<float2*>(zs + (mt * 16 + g + 8) * kC5 + col);
                    float2 v0, v1;
                    if (layer == 0) {
                        const float2 b = *reinterpret_cast<const float2*>(w5x + kC5 + col);
                        v0 = b; v1 = b;
                    } else {
                        v0 = *a0; v1 = *a1;
                    }
                    v0.x += z[nt][0]; v0.y += z[nt][1]; v1.x += z[nt][2]; v1.y += z[nt][3];
                    *a0 = v0; *a1 = v1;
                }
            }
        }
        if (!split) tm.sync();
        else if (layer < 2) pair_exchange(out_pl, kRowB, nullptr, 0, nullptr, 0);
        else pair_exchange(vpl, 2, vpl + S, 2, conv5 ? zs : nullptr, conv5 ? kC5 * 4 : 0);
        KS_TRACE(3 + layer);
    }

    // ---- layer 4: 32 -> 1, already projected into v: one MMA column -------------------
    for (int mt = t0 + warp; mt < t1; mt += nwarps) {
        float a4[4];
        aggregate8(c, vpl, S, 1, mt, lane, a4);
        if (t == 0) {                                  // column 0 of the tile lives in t == 0
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = mt * 16 + g + 8 * half;
                if (row < n) {
                    const float x4 = tanhf(fmaf(rs[row], a4[2 * half], b4));
                    key[row] = x4;
                    float* o = xc + (int64_t)row * p.ldc + 3 * kHid;
                    // padded rows (ldc >= 100, 16-byte aligned): write the pad too, so that no
                    // 32-byte sector of x_cat is left partially written (a later read of such a
                    // sector has to be filled from DRAM)
                    if (pad4) *reinterpret_cast<float4*>(o) = make_float4(x4, 0.f, 0.f, 0.f);
                    else *o = x4;
                }
            }
        }
    }
    if (!split) tm.sync();
    else pair_exchange(key, 4, nullptr, 0, nullptr, 0);
    KS_TRACE(6);

    // ---- SortPool: order by x_4 descending, ties by node index -----------------------
    uint64_t* comp = reinterpret_cast<uint64_t*>(PA);
    if (n <= kRankSortMax) {
        for (int j = tid; j < n; j += nthreads)
            comp[j] = ((uint64_t)descending_key_bits(key[j]) << 32) | (uint32_t)j;
        tm.sync_local();
        // rank = number of smaller composites; `parts` lanes share one element (interleaved j)
        int parts = 1;
        while (parts < 32 && n * parts * 2 <= gthreads) parts <<= 1;
        const int lp = 31 - __clz(parts);
        for (int it0 = 0; it0 < n * parts; it0 += gthreads) {
            const int item = it0 + gtid;
            const int i = item >> lp, part = item & (parts - 1);
            const bool live = i < n;
            const uint64_t mine = live ? comp[i] : 0ull;
            int rnk = 0;
            if (live) {
                int j = part, r1 = 0, r2 = 0, r3 = 0;                // four independent count chains
                for (; j + 7 * parts < n; j += 8 * parts) {          // eight independent loads in flight
                    uint64_t o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = comp[j + u * parts];
                    rnk += (o[0] < mine) + (o[4] < mine);
                    r1 += (o[1] < mine) + (o[5] < mine);
                    r2 += (o[2] < mine) + (o[6] < mine);
                    r3 += (o[3] < mine) + (o[7] < mine);
                }
                for (; j < n; j += parts) rnk += comp[j] < mine;
                rnk += r1 + r2 + r3;
            }
            for (int o = parts >> 1; o > 0; o >>= 1) rnk += __shfl_xor_sync(DGCNN_FULL_MASK, rnk, o);
            if (live && part == 0 && rnk < keep) {
                order[rnk] = i;
                if (order_peer) order_peer[rnk] = i;
            }
        }
        tm.sync();
    } else {
        // (a split graph sorts redundantly in both CTAs: no exchange)
        const uint32_t pw = next_pow2((uint32_t)n);
        for (uint32_t j = tid; j < pw; j += nthreads)
            comp[j] = (j < (uint32_t)n) ? (((uint64_t)descending_key_bits(key[j]) << 32) | j) : ~0ull;
        Team local = tm;
        local.split = 0;
        bitonic_sort_team(comp, pw, local);
        for (int r = tid; r < keep; r += nthreads) order[r] = (int)(uint32_t)(comp[r] & 0xffffffffu);
        tm.sync_local();
    }
    KS_TRACE(7);

    // ---- the k winners, one warp per row (rows of x_cat this team just wrote: L2 hits);
    //      sixteen rows in flight per warp, loads unconditional (clamped) so they all overlap:
    //      the copy is pure L2 latency ------------------------------------------------------
    if (conv5) {
        // conv5 pre-activation of pooled row r: z[order[r]] + W5[:, 96] x_4 (padding rows: b5), then
        // ReLU and the max over the row pair (2j, 2j+1): h1[g][c][j]; arg = winning row, 2 = dead
        const int L1 = p.k >> 1;
        float* h1g = p.h1 + (int64_t)gi * kC5 * L1;
        uint8_t* argg = p.arg + (int64_t)gi * kC5 * L1;
        for (int item = gtid; item < kC5 * L1; item += gthreads) {
            const int ch = item / L1, j = item - ch * L1;
            float zr[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = 2 * j + u;
                if (r < keep) {
                    const int node = order[r];
                    zr[u] = fmaf(w5x[ch], key[node], zs[node * kC5 + ch]);
                } else {
                    zr[u] = w5x[kC5 + ch];
                }
                zr[u] = fmaxf(zr[u], 0.f);
            }
            const float m = fmaxf(zr[0], zr[1]);
            h1g[item] = m;
            argg[item] = (uint8_t)(m <= 0.f ? 2 : (zr[0] >= zr[1] ? 0 : 1));
        }
    }
    if (p.pooled) {
        const float* __restrict__ xsrc = xc;
        constexpr int R = 16;
        for (int r0 = gwarp * R; r0 < keep; r0 += gwarps * R) {
            float v[R][3], v96;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const float* src = xsrc + (int64_t)order[min(r0 + u, keep - 1)] * p.ldc;
                v[u][0] = src[lane]; v[u][1] = src[32 + lane]; v[u][2] = src[64 + lane];
            }
            v96 = xsrc[(int64_t)order[min(r0 + (lane & (R - 1)), keep - 1)] * p.ldc + 96];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                if (r0 + u < keep) {
                    float* dst = pooled_g + (r0 + u) * kCat;
                    dst[lane] = v[u][0]; dst[32 + lane] = v[u][1]; dst[64 + lane] = v[u][2];
                }
            }
            if (lane < R && r0 + lane < keep) pooled_g[(r0 + lane) * kCat + 96] = v96;
        }
    }
    for (int r = gtid; r < keep; r += gthreads) perm_g[r] = base + order[r];
    if (bulk_pending) bulk_store_wait();
    tm.xparity = xparity;
    KS_TRACE(8);
    if (tracer) p.trace[(int64_t)gi * 16 + 10] = global_ns();
#undef KS_TRACE
}

// The CTA's graphs and their teams: plan_pass() in graph_mma.cuh.
template <bool LAZY>
__global__ void __launch_bounds__(kFwdThreads, 1) stack_fwd_mma_kernel(StackFwdParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ PlanEntry s_plan[kMaxTeams];
    __shared__ int s_count;
    const int64_t cta_t0 = p.trace ? global_ns() : 0;
    const int64_t cta_c0 = p.trace ? clock64() : 0;
    int64_t cta_c1 = 0, cta_c2 = 0;
    const int f = p.f;
    const bool conv5 = p.h1 != nullptr;
    const SharedLayout SL = shared_layout(f, conv5);
    unsigned char* team_base = smraw + al16(SL.total);
    const int budget = kQuads * quad_bytes(f, conv5);
    const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsm = gridDim.x, sm = blockIdx.x, B = p.num_graphs;
    constexpr int kWarps = kFwdThreads / 32;
    const int4* gdesc = reinterpret_cast<const int4*>(p.gdesc);      // {graph, base, n, fgoff}
    int next = 0;                                    // items of this CTA consumed so far
    int excl = 0;                                    // graphs with an SM of their own (warp 0 only)
    int nsplit = 0, msplit = 0;                      // CTA pairs / graphs split over a pair (warp 0 only)
    uint32_t xparity = 0;                            // phase of the pair's exchange mbarrier
    uint32_t crank = 0;                              // rank of this CTA in its cluster
    __shared__ __align__(8) uint64_t s_pair_mbar;    // exchange barrier of a graph split over the pair
    if (p.pairs) {
        asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        if (threadIdx.x == 0) mbar_init(smem_addr_u32(&s_pair_mbar), 1);   // (visible to the peer after the
    }                                                                      //  split graph's first cluster barrier)

    for (int pass = 0;; ++pass) {
        if (warp_id == 0) {
            plan_pass(gdesc, B, nsm, sm, next, excl, pass == 0, budget, kWarps,
                      [f, conv5](int np, bool split) { return team_layout(f, np, conv5, split).total; }, s_plan,
                      &s_count, nsplit, msplit, p.pairs != 0, p.split_pct);
        } else if (pass == 0) {
            // meanwhile the other warps stage the weights: W1 transposed fp32 (F -> 32 stays on
            // the FMA pipe), W2/W3 as hi/lo fp16 planes [cout][cin] = the MMA "col" operand
            const int tid = threadIdx.x - 32, nthreads = kFwdThreads - 32;
            __half* w2p = reinterpret_cast<__half*>(smraw + SL.w2p);
            __half* w3p = reinterpret_cast<__half*>(smraw + SL.w3p);
            float* w1t = reinterpret_cast<float*>(smraw + SL.w1t);
            float* w4s = reinterpret_cast<float*>(smraw + SL.misc);
            {   // all global loads of a thread are issued before its first store: one round trip
                constexpr int R = (kHid * kHid + (kFwdThreads - 32) - 1) / (kFwdThreads - 32);
                float v2[R], v3[R], vb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int idx = tid + r * nthreads;
                    v2[r] = idx < kHid * kHid ? p.w2[idx] : 0.f;
                    v3[r] = idx < kHid * kHid ? p.w3[idx] : 0.f;
                }
                if (tid < kHid) {
                    vb[0] = p.w4[tid];
                    vb[1] = p.b1 ? p.b1[tid] : 0.f;
                    vb[2] = p.b2 ? p.b2[tid] : 0.f;
                    vb[3] = p.b3 ? p.b3[tid] : 0.f;
                }
                const float w1first = tid < f * kHid ? p.w1[tid] : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int idx = tid + r * nthreads;
                    if (idx < kHid * kHid) {
                        const int c = idx >> 5, k = idx & 31;
                        store_split(w2p, w2p + kHid * kWPad, c * kWPad + k, v2[r]);
                        store_split(w3p, w3p + kHid * kWPad, c * kWPad + k, v3[r]);
                    }
                }
                if (tid < kHid) {
                    w4s[tid] = vb[0]; w4s[kHid + tid] = vb[1];
                    w4s[2 * kHid + tid] = vb[2]; w4s[3 * kHid + tid] = vb[3];
                }
                if (tid < f * kHid) { const int c = tid / f, k = tid - c * f; w1t[k * kHid + c] = w1first; }
                for (int idx = tid + nthreads; idx < f * kHid; idx += nthreads) {
                    const int c = idx / f, k = idx - c * f;
                    w1t[k * kHid + c] = p.w1[idx];
                }
                {
                    uint32_t* zb = reinterpret_cast<uint32_t*>(smraw + SL.zero);
                    for (int idx = tid; idx < kZeroBytes / 4; idx += nthreads) zb[idx] = 0u;
                }
                if (conv5) {   // W5[:, 0..95] as hi/lo planes [16][kW5Pad]; W5[:, 96] and b5 in fp32
                    __half* w5p = reinterpret_cast<__half*>(smraw + SL.w5p);
                    float* w5x = reinterpret_cast<float*>(smraw + SL.w5x);
                    for (int idx = tid; idx < kC5 * kCat; idx += nthreads) {
                        const int c = idx / kCat, k = idx - c * kCat;
                        const float v = p.w5[idx];
                        if (k < 3 * kHid) store_split(w5p, w5p + kC5 * kW5Pad, c * kW5Pad + k, v);
                        else w5x[c] = v;
                    }
                    if (tid < kC5) w5x[kC5 + tid] = p.b5 ? p.b5[tid] : 0.f;
                }
            }
        }
        __syncthreads();                             // the plan (and, first time, the weights)
        const int count = s_count;
        if (count == 0) break;
        if (p.trace) cta_c1 = clock64();
        int mine = -1;
        for (int j = 0; j < count; ++j)
            if (warp_id >= s_plan[j].warp0 && warp_id < s_plan[j].warp0 + s_plan[j].nwarps) mine = j;
        if (mine >= 0) {
            const PlanEntry e = s_plan[mine];
            if (e.n > p.nmax) {                      // host promised this cannot happen
                if (threadIdx.x == e.warp0 * 32 && p.status) atomicOr(p.status, DGCNN_GRAPH_BAD_BATCH);
            } else {
                Team tm;
                tm.tid = threadIdx.x - e.warp0 * 32;
                tm.nthreads = e.nwarps * 32;
                tm.warp = tm.tid >> 5;
                tm.nwarps = e.nwarps;
                tm.lane = lane;
                tm.bar = 1 + mine;
                tm.smem = team_base + e.smem_off;
                tm.split = e.pad;                    // shared with the peer CTA of the cluster
                tm.rank = (int)crank;
                tm.mbar = smem_addr_u32(&s_pair_mbar);
                tm.xparity = xparity;
                if (p.trace && tm.tid == 0 && (!tm.split || crank == 0)) {
                    cta_c2 = clock64();
                    p.trace[(int64_t)e.gi * 16 + 11] = cta_t0;
                    p.trace[(int64_t)e.gi * 16 + 12] = cta_c0;       // CTA entry
                    p.trace[(int64_t)e.gi * 16 + 13] = cta_c1;       // plan + weights done
                    p.trace[(int64_t)e.gi * 16 + 14] = cta_c2;       // padding rows zeroed, team starts
                }
                process_graph<LAZY>(p, tm, e, smraw);
                xparity = tm.xparity;
            }
        }
        __syncthreads();                             // shared memory is re-carved by the next pass
        next += count;
    }
}

}  // namespace dgcnn

using namespace dgcnn;

// implemented in graph_stack.cu (FMA gather variant)
int dgcnn_stack_fwd_fma(const float* x, int64_t ldx, int32_t num_features, const int32_t* rowptr,
                        const int32_t* col, const float* dis, const int32_t* gptr, const int32_t* gorder,
                        const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                        int64_t num_nodes, int64_t num_graphs, int64_t max_nodes, const float* w1,
                        const float* b1,
                        const float* w2, const float* b2, const float* w3, const float* b3,
                        const float* w4, const float* b4, float* xcat, int64_t ldc, float* pooled,
                        int32_t* perm, int32_t k, int32_t norm, int32_t* status, int32_t* counter,
                        cudaStream_t st);
int dgcnn_stack_fwd_fma_supported(int32_t num_features, int64_t max_nodes);

// The next dgcnn_stack_fwd / dgcnn_stack_fwd_conv5 call of this thread's library state fills the
// adjacency maps itself (StackFwdParams::lazy).  Internal: train_step.cu pairs it with
// dgcnn_build_bitmaps_impl(.., lazy = 1); `fragmap` / `gflags` are then OUTPUTS of the forward call.
static int g_next_lazy = 0;
void dgcnn_stack_fwd_next_lazy(int on) { g_next_lazy = on; }

static int64_t* g_trace = nullptr;
extern "C" void dgcnn_stack_fwd_set_trace(int64_t* device_buffer) { g_trace = device_buffer; }

void dgcnn_stack_bwd_mma_configure(int pairs, int split_pct);     // graph_stack_bwd_mma.cu

// Clusters of two CTAs (one TPC): the plan splits the largest graphs of the batch over a pair
// (planes exchanged through distributed shared memory), everything else runs as before.
// DGCNN_KS_PAIRS=0 launches without clusters; DGCNN_KS_SPLIT_PCT tunes the split threshold.
static int g_pairs_ok = -1, g_split_pct = 80;           // -1: not probed, -2: requested, 0 / 1: decided
extern "C" void dgcnn_stack_fwd_configure(int32_t pairs, int32_t split_pct) {
    g_pairs_ok = pairs < 0 ? -1 : (pairs ? -2 : 0);
    if (split_pct > 0) g_split_pct = split_pct;
    dgcnn_stack_bwd_mma_configure(pairs, split_pct);
}
static int pairs_enabled() {
    if (g_pairs_ok >= 0) return g_pairs_ok;
    int ok = 1;
    if (g_pairs_ok == -1) {
        const char* env = getenv("DGCNN_KS_PAIRS");
        const char* pct = getenv("DGCNN_KS_SPLIT_PCT");
        if (pct && atoi(pct) > 0) g_split_pct = atoi(pct);
        ok = !(env && env[0] == '0');
    }
    if (ok) ok = cluster_pairs_fit(stack_fwd_mma_kernel<false>, kFwdThreads, (size_t)(kSmemBudget - 1024)) &&
                 cluster_pairs_fit(stack_fwd_mma_kernel<true>, kFwdThreads, (size_t)(kSmemBudget - 1024));
    g_pairs_ok = ok;
    return ok;
}

static int mma_supported(int32_t f, int64_t max_nodes, bool conv5 = false) {
    if (f < 1 || f > kMaxF || max_nodes < 1 || max_nodes > 1024) return 0;
    const int np = (int)((max_nodes + 15) / 16 * 16);
    const int budget = kQuads * quad_bytes(f, conv5);
    if (team_layout(f, np, conv5, false).total <= budget) return 1;
    // larger graphs only as a CTA pair (mandatory split: plan_pass)
    return team_layout(f, np, conv5, true).total <= budget && pairs_enabled() ? 1 : 0;
}

extern "C" int dgcnn_stack_fwd_supported(int32_t num_features, int64_t max_nodes) {
    return mma_supported(num_features, max_nodes);
}

extern "C" int dgcnn_stack_fwd_conv5_supported(int32_t num_features, int64_t max_nodes) {
    return mma_supported(num_features, max_nodes, true);
}

extern "C" size_t dgcnn_stack_fwd_workspace_bytes(void) { return 256; }

static int stack_fwd_impl(const float* x, int64_t ldx, int32_t num_features,
                          const int32_t* rowptr, const int32_t* col, const float* dis,
                          const int32_t* gptr, const int32_t* gorder,
                          const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                          const uint32_t* fragmap, const int32_t* fgoff, const int32_t* gdesc,
                          int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                          const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* w3, const float* b3, const float* w4, const float* b4,
                          float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                          int32_t norm, int32_t variant, int32_t* status, void* workspace,
                          size_t workspace_bytes, void* stream,
                          const float* w5, const float* b5, float* h1, uint8_t* arg) {
    const bool conv5 = h1 != nullptr;
    if (num_nodes < 0 || num_graphs < 0 || k < 1 || num_features < 1 || ldx < num_features ||
        ldc < kCat)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (conv5 && (!w5 || !b5 || !arg || k < 2 || variant != DGCNN_STACK_MMA)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    if (variant != DGCNN_STACK_MMA && variant != DGCNN_STACK_FMA) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs == 0) return DGCNN_OK;
    if (num_graphs >= INT32_MAX || num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (variant == DGCNN_STACK_MMA ? !mma_supported(num_features, max_nodes, conv5)
                                   : !dgcnn_stack_fwd_fma_supported(num_features, max_nodes))
        return DGCNN_ERR_UNSUPPORTED;
    if (!bitmap || !bmoff || !gflags) return DGCNN_ERR_INVALID_ARGUMENT;
    if (variant == DGCNN_STACK_MMA && (!fragmap || !fgoff || !gdesc)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (max_nodes > 1024) return DGCNN_ERR_UNSUPPORTED;
    if (!rowptr || !dis || !gptr || !w1 || !w2 || !w3 || !w4 || !xcat || (!pooled && !conv5) || !perm ||
        (num_nodes > 0 && !x))
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_stack_fwd_workspace_bytes()) return DGCNN_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uintptr_t aligned = ((uintptr_t)workspace + 127) & ~(uintptr_t)127;
    int32_t* counter = reinterpret_cast<int32_t*>(aligned);
    if (variant == DGCNN_STACK_FMA) {
        if (cudaMemsetAsync(counter, 0, sizeof(int32_t), st) != cudaSuccess) return DGCNN_ERR_CUDA;
        return dgcnn_stack_fwd_fma(x, ldx, num_features, rowptr, col, dis, gptr, gorder, bitmap, bmoff,
                                   gflags, num_nodes, num_graphs,
                                   max_nodes, w1, b1, w2, b2, w3, b3, w4, b4, xcat, ldc, pooled, perm, k,
                                   norm, status, counter, st);
    }

    StackFwdParams p{};
    p.x = x; p.ldx = ldx; p.f = num_features;
    p.rowptr = rowptr; p.col = col; p.dis = dis; p.gptr = gptr; p.num_graphs = (int)num_graphs;
    p.bitmap = bitmap; p.bmoff = bmoff; p.gflags = gflags;
    p.fragmap = fragmap; p.fgoff = fgoff;
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.w4 = w4; p.b4 = b4;
    p.xcat = xcat; p.ldc = ldc; p.pooled = pooled; p.perm = perm; p.k = k;
    p.norm = norm; p.nmax = (int)max_nodes;
    p.gdesc = gdesc; p.counter = counter; p.status = status;
    p.trace = g_trace;
    p.w5 = w5; p.b5 = b5; p.h1 = h1; p.arg = arg;
    p.lazy = g_next_lazy;                             // (one-call training step: dgcnn_stack_fwd_next_lazy)
    g_next_lazy = 0;
    p.fragmap_w = const_cast<uint32_t*>(fragmap); p.gflags_w = const_cast<int32_t*>(gflags);
    // one CTA per SM with (almost) all of its shared memory: 4 quad slices + the weights
    const size_t smem = (size_t)al16(shared_layout(p.f, conv5).total) + (size_t)kQuads * quad_bytes(p.f, conv5);
    auto kernel = p.lazy ? stack_fwd_mma_kernel<true> : stack_fwd_mma_kernel<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    p.pairs = pairs_enabled();
    p.split_pct = g_split_pct;
    {
        static int plain = -1;
        if (plain < 0) { const char* env = getenv("DGCNN_KS_PLAIN_ZERO"); plain = (env && env[0] == '1') ? 1 : 0; }
        p.plain_zero = plain;
    }
    int64_t grid = DGCNN_NUM_SMS;
    if (p.pairs) {
        // fewer graphs than SMs: up to 8 spare CTAs double up on the largest graphs
        if (grid > num_graphs + 8) grid = num_graphs + 8;
        grid += grid & 1;
        if (grid > DGCNN_NUM_SMS) grid = DGCNN_NUM_SMS;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1] = pdl_attribute();
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kFwdThreads);
        cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (cudaLaunchKernelEx(&cfg, kernel, p) != cudaSuccess) return DGCNN_ERR_CUDA;
    } else {
        if (grid > num_graphs) grid = num_graphs;
        DGCNN_LAUNCH(kernel, (unsigned)grid, kFwdThreads, smem, st, p);
    }
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_stack_fwd(const float* x, int64_t ldx, int32_t num_features,
                               const int32_t* rowptr, const int32_t* col, const float* dis,
                               const int32_t* gptr, const int32_t* gorder,
                               const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                               const uint32_t* fragmap, const int32_t* fgoff, const int32_t* gdesc,
                               int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                               const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* w3, const float* b3, const float* w4, const float* b4,
                               float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                               int32_t norm, int32_t variant, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
    return stack_fwd_impl(x, ldx, num_features, rowptr, col, dis, gptr, gorder, bitmap, bmoff, gflags, fragmap,
                          fgoff, gdesc, num_nodes, num_graphs, max_nodes, w1, b1, w2, b2, w3, b3, w4, b4, xcat,
                          ldc, pooled, perm, k, norm, variant, status, workspace, workspace_bytes, stream,
                          nullptr, nullptr, nullptr, nullptr);
}

// KS + the head of the dense tail (model.py:36-38): as dgcnn_stack_fwd, and
// h1[g][c][j] = max(relu(conv5(pooled)[g][c][2j]), relu(...[2j+1])), arg = the winning row (2: dead).
// `pooled` may be NULL: the [B, k*97] SortPooling output is then never materialised.
extern "C" int dgcnn_stack_fwd_conv5(const float* x, int64_t ldx, int32_t num_features,
                                     const int32_t* rowptr, const int32_t* col, const float* dis,
                                     const int32_t* gptr, const int32_t* gorder,
                                     const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                                     const uint32_t* fragmap, const int32_t* fgoff, const int32_t* gdesc,
                                     int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                                     const float* w1, const float* b1, const float* w2, const float* b2,
                                     const float* w3, const float* b3, const float* w4, const float* b4,
                                     const float* w5, const float* b5,
                                     float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                                     float* h1, uint8_t* arg, int32_t norm, int32_t* status, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!h1 || !arg || !w5 || !b5) return DGCNN_ERR_INVALID_ARGUMENT;
    return stack_fwd_impl(x, ldx, num_features, rowptr, col, dis, gptr, gorder, bitmap, bmoff, gflags, fragmap,
                          fgoff, gdesc, num_nodes, num_graphs, max_nodes, w1, b1, w2, b2, w3, b3, w4, b4, xcat,
                          ldc, pooled, perm, k, norm, DGCNN_STACK_MMA, status, workspace, workspace_bytes, stream,
                          w5, b5, h1, arg);
}
