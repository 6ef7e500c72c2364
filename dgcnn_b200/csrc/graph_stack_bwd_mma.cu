// Tensor-core version of the fused backward (autograd of model.py:28-35, train.py:40),
// the mirror image of graph_stack_mma.cu.  One TEAM of warps per graph (the per-CTA plan of
// graph_mma.cuh: up to eight graphs of a 640-thread CTA run concurrently); per layer l = 4..1:
//
//   dpre = dy * (1 - y^2)            row-local, db += column sums
//   dh   = c_i * (A_hat^T . r dpre)  mma.sync: A^T from the bitmap (exact 0/1 fp16),
//                                    r*dpre split hi/lo into fp16 planes
//   dx   = dh W (+ pooled gradient)  mma.sync straight from the accumulator fragments
//   dW  += dh^T x_in                 mma.sync: dh planes [c][node] x x_in (from x_cat, split in registers)
//
// Gradients span many orders of magnitude, so each graph picks a power-of-two scale from
// the largest pooled gradient it receives (exact to apply and to undo) before anything is
// split into fp16 hi/lo pairs; a graph whose pooled gradient is all zero is skipped.
//
// A graph that the plan SPLITS over the two CTAs of a cluster (large graphs; mandatory beyond one
// CTA's shared memory): each CTA takes half of the 16-row tiles -- the column-indexed planes are
// built whole in both, dh / adjacency / parameter-gradient sums only for the own rows, the
// gradient rows G cross through L2 between two alternating buffers under ONE cluster barrier per
// layer (bwd_process_graph).
//
// SURVEY 8f N2 (dh1 != null): the backward of conv5 + ReLU + MaxPool1d is fused in -- dz planes
// from d(h1) / arg (the graph's slab staged once in shared memory), dz W5 added where a layer's
// input gradient is assembled (dz_w5_tile), dW5 / db5 in the dW tile loop.
//
// Determinism without float atomics: the plan is a pure function of the graph sizes, a CTA adds
// its teams' parameter-gradient vectors in team order and writes ONE partial vector to HBM; a
// second kernel sums the CTAs' vectors in CTA order.  (148 x 3.8k floats = 2.2 MB.)
#include "graph_mma.cuh"

namespace dgcnn {

// threads per CTA of the backward kernel (see DGCNN_FWD_THREADS in graph_stack_mma.cu)
#ifndef DGCNN_BWD_THREADS
#define DGCNN_BWD_THREADS 640
#endif
constexpr int kBwdThreads = DGCNN_BWD_THREADS;

struct StackBwdMmaParams {
    const float* dpooled; const int32_t* perm; int k;
    const float* xcat; int64_t ldc;
    const float* x; int64_t ldx; int f;
    const int32_t* rowptr_t; const int32_t* col_t; const float* dis; const int32_t* gptr;
    const int32_t* gorder; int num_graphs;
    const int32_t* gdesc;            // K0b work descriptors {graph, first node, nodes, fgoff}
    const uint32_t* fragmap;         // K0b fragment-major A_hat (== A_hat^T when K0 proved symmetry)
    const uint32_t* bitmap; const int32_t* bmoff; const int32_t* gflags;
    const uint32_t* bitmap_t; const int32_t* bmoff_t; const int32_t* gflags_t;
    const float* w2; const float* w3; const float* w4;
    int norm; int nmax;
    float* partials;     // [num_graphs][P]
    float* gws;          // [N][32] fp32: gradient w.r.t. the current layer's output (L2-resident)
    int32_t* counter;
    int32_t* status;
    int64_t* trace;      // optional debug timeline [num_graphs][16] (dgcnn_stack_bwd_set_trace)
    // conv5 + ReLU + max-pool's backward fused in (dh1 != null; dpooled is then unused):
    // the pooled gradient is dz W5 with dz[r][c] = dh1[c][r/2] routed through `arg`, never materialised
    const float* dh1; const uint8_t* arg; const float* w5;
    int64_t num_nodes;   // gws holds two [num_nodes][32] buffers (split graphs alternate between them)
    int pairs;           // launched as clusters of two CTAs: the plan may split graphs over a pair
    int split_pct;
};
#define KSB_TRACE(slot) do { if (p.trace && tm.tid == 0 && !(tm.split && tm.rank)) p.trace[(int64_t)gi * 16 + (slot)] = clock64(); } while (0)

constexpr int kC5b = 16;                   // conv5 output channels
constexpr int kW5TPad = 24;                // row stride (halfs) of the transposed W5 planes [97][16]

struct GradOffsetsM { int w1, b1, w2, b2, w3, b3, w4, b4, w5, b5, total; };

__host__ __device__ inline GradOffsetsM grad_offsets_m(int f, bool conv5 = false) {
    GradOffsetsM g;
    int o = 0;
    g.w1 = o; o += kHid * f;
    g.b1 = o; o += kHid;
    g.w2 = o; o += kHid * kHid;
    g.b2 = o; o += kHid;
    g.w3 = o; o += kHid * kHid;
    g.b3 = o; o += kHid;
    g.w4 = o; o += kHid;
    g.b4 = o; o += 1;
    g.w5 = o; o += conv5 ? kC5b * kCat : 0;              // conv5.weight [16,97], conv5.bias [16]: the
    g.b5 = o; o += conv5 ? kC5b : 0;                     // next two tensors of the flat parameter order
    g.total = o;
    return g;
}

// CTA-wide: W2^T, W3^T hi/lo planes [k][c] (B operand of dx = dh W), w4
struct BwdShared { int w2p, w3p, w4, w5t, w5x, acc, total; };
__host__ __device__ inline BwdShared bwd_shared_layout(int f, bool conv5 = false) {
    BwdShared L;
    int o = 0;
    L.w2p = o; o += 2 * kHid * kWPad * 2;
    L.w3p = o; o += 2 * kHid * kWPad * 2;
    L.w4 = o; o += kHid * 4;
    L.w5t = o; o += conv5 ? 2 * 104 * kW5TPad * 2 : 0;   // W5^T hi/lo planes [plane][column i][channel]
    L.w5x = o; o += conv5 ? (kC5b + 4) * 4 : 0;          // W5[:, 96] fp32, then max_i sum_c |W5[c][i]|
    L.acc = o;                                           // (the CTA's running parameter-gradient sum lives in
                                                         //  its slot of `partials` in HBM / L2: 15 KB more
                                                         //  shared memory for the graphs -- one team more per
                                                         //  pass, and a second pass costs a whole ~50 us chain)
    L.total = o;
    return L;
}

// per graph (bytes).  split: the graph is spread over the two CTAs of a cluster; each CTA owns the
// first / last ceil(T / 2) row tiles and keeps the dh planes and the adjacency of THOSE rows only
// (everything indexed by column -- P, dz, coefficients, ranks -- stays whole)
struct BwdTeamLayout { int P, DH, DZ, vpl, bm, cs, rs, hv, rank, rp, red, sacc, total; int S, Sd; };
__host__ __device__ inline BwdTeamLayout bwd_team_layout(int f, int np, bool conv5 = false, bool split = false) {
    BwdTeamLayout L;
    L.S = np + 8;
    const int wpr = (np + 31) >> 5;
    const int T = np >> 4, own_t = split ? (T + 1) >> 1 : T, own_np = own_t * 16;
    L.Sd = own_np + 8;
    int o = 0;
    L.P = o; o += 2 * kHid * L.S * 2;                    // hi/lo planes of r * dpre * scale
    L.DH = o; o += 2 * kHid * L.Sd * 2;                  // hi/lo planes of dh * scale (own rows)
    L.DZ = o; o += conv5 ? 2 * kC5b * L.S * 2 : 0;       // hi/lo planes [16][S] of dz * scale (conv5 pre-activation grads)
    L.vpl = o; o += al16(2 * L.S * 2);
    {
        const int plain = own_np * wpr * 4, frag = own_t * ((T + 3) >> 2) * 32 * 4;
        L.bm = o; o += al16(plain > frag ? plain : frag);
    }
    L.cs = o; o += al16(np * 4);
    L.rs = o; o += al16(np * 4);
    L.hv = o; o += al16(np * 4);
    L.rank = o; o += al16(np * 4);
    L.rp = o; o += al16((np + 1) * 4);
    L.red = o; o += (kBwdThreads / 32) * kHid * 4;
    L.sacc = o; o += al16(grad_offsets_m(f, conv5).total * 4);
    L.total = o;
    return L;
}

__host__ __device__ inline int bwd_quad_bytes(int f, bool conv5 = false) {
    return ((kSmemBudget - 1024 - bwd_shared_layout(f, conv5).total) / kQuads) & ~15;
}

// sum red[w][c] over the team's warps (c < 32)
__device__ __forceinline__ float reduce_rows_t(const float* red, int nwarps, int c) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w * kHid + c];
    return s;
}

// SURVEY 8f N2.  Pooled gradient of one x_cat slice for the 16-row tile mt, never materialised in
// HBM:  gz[node][k] = sum_c dz[node][c] W5[c][off + k]  (scaled like dz).  A = dz planes
// [channel][node] read transposed (ldmatrix.trans), B = W5^T planes [column][channel];
// hi*hi + lo*hi + hi*lo.  C layout: gz[nt][0..1] row g, gz[nt][2..3] row g+8, columns nt*8+2t, +1.
struct Conv5Bwd {
    const __half* DZ;       // [plane][16][S]
    const __half* w5t;      // [plane][104][kW5TPad]
    bool on;
};

__device__ __forceinline__ void dz_w5_tile(const Conv5Bwd& c5, int S, int mt, int off, int lane,
                                           float (&gz)[4][4]) {
    const int g = lane >> 2, t = lane & 3, j = lane >> 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) gz[nt][0] = gz[nt][1] = gz[nt][2] = gz[nt][3] = 0.f;
    // matrix j of the x4 load: channels (j>>1)*8 .. +7 (rows), nodes mt*16 + (j&1)*8 .. +7
    const uint32_t a0 = smem_u32(c5.DZ) + (uint32_t)((((j >> 1) * 8 + (lane & 7)) * S + mt * 16 + (j & 1) * 8) * 2);
    uint32_t ah[4], al[4];
    ldsm_x4_t(a0, ah[0], ah[1], ah[2], ah[3]);
    ldsm_x4_t(a0 + (uint32_t)(kC5b * S * 2), al[0], al[1], al[2], al[3]);
    const uint32_t* wh = reinterpret_cast<const uint32_t*>(c5.w5t);
    const uint32_t* wl = reinterpret_cast<const uint32_t*>(c5.w5t + 104 * kW5TPad);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int widx = ((off + nt * 8 + g) * kW5TPad + 2 * t) >> 1;
        const uint32_t h0 = wh[widx], h1 = wh[widx + 4], l0 = wl[widx], l1 = wl[widx + 4];
        mma_fp16(gz[nt], ah, h0, h1);
        mma_fp16(gz[nt], al, h0, h1);
        mma_fp16(gz[nt], ah, l0, l1);
    }
}

// dh tile(s) = c_i * (A_hat^T . P) for every 16-row tile owned by this warp; then
//   - dh planes (still scaled) for the dW product,
//   - if wp: dx = dh W via the fragments, unscaled, + pooled gradient of the slice -> Gout
__device__ __forceinline__ void bwd_mma_layer(const __half* __restrict__ P, const __half* __restrict__ wp,
                                              __half* __restrict__ DH, int Sd, float* __restrict__ Gout,
                                              const uint32_t* __restrict__ bm, bool frag, int wpr, int n, int S,
                                              int t_lo, int t_hi, bool dup, const int* __restrict__ rp,
                                              const int32_t* __restrict__ col_g, int base,
                                              const float* __restrict__ cs, const int* __restrict__ rank,
                                              const float* __restrict__ dp, int offx, float inv_scale,
                                              const Team& tm, const Conv5Bwd& c5) {
    const int lane = tm.lane, warp = tm.warp, nwarps = tm.nwarps;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = (n + 15) >> 4;
    const uint32_t* in32[2] = {reinterpret_cast<const uint32_t*>(P),
                               reinterpret_cast<const uint32_t*>(P + kHid * S)};
    for (int mt = t_lo + warp; mt < t_hi; mt += nwarps) {      // (bm, DH: offset to the own rows by the caller)
        const int row0 = mt * 16 + g, row1 = row0 + 8;
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        if (!dup && frag) {
            // fragment-major adjacency (rotate + mask -> {0, 2.0}), B fragments by ldmatrix from
            // the [channel][node] planes; full groups of four blocks run branch-free
            const int G = (tiles + 3) >> 2, j = lane >> 3;
            uint32_t addr = smem_u32(P) + (uint32_t)((((j >> 1) * 8 + (lane & 7)) * S + (j & 1) * 8) * 2);
            const uint32_t nt2 = (uint32_t)(16 * S * 2), lo = (uint32_t)(kHid * S * 2);
            const uint32_t* fb = bm + mt * G * 32 + lane;
            uint32_t w = fb[0];
            auto block = [&](int q) {
                uint32_t a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = __funnelshift_l(w, w, (14 - (4 * q + i)) & 31) & 0x40004000u;
                const uint32_t ad = addr + (uint32_t)(q * 32);
                uint32_t b[4][4];
                ldsm_x4(ad, b[0][0], b[0][1], b[0][2], b[0][3]);                 // hi, channels 0..15
                ldsm_x4(ad + nt2, b[1][0], b[1][1], b[1][2], b[1][3]);           // hi, channels 16..31
                ldsm_x4(ad + lo, b[2][0], b[2][1], b[2][2], b[2][3]);            // lo, channels 0..15
                ldsm_x4(ad + lo + nt2, b[3][0], b[3][1], b[3][2], b[3][3]);      // lo, channels 16..31
                mma_fp16(acc[0], a, b[0][0], b[0][1]);
                mma_fp16(acc[1], a, b[0][2], b[0][3]);
                mma_fp16(acc[2], a, b[1][0], b[1][1]);
                mma_fp16(acc[3], a, b[1][2], b[1][3]);
                mma_fp16(acc[0], a, b[2][0], b[2][1]);
                mma_fp16(acc[1], a, b[2][2], b[2][3]);
                mma_fp16(acc[2], a, b[3][0], b[3][1]);
                mma_fp16(acc[3], a, b[3][2], b[3][3]);
            };
            for (int grp = 0; grp < G; ++grp, addr += 128) {
                const uint32_t wn = grp + 1 < G ? fb[(grp + 1) * 32] : 0u;
                const int nb = tiles - grp * 4;
                if (__any_sync(DGCNN_FULL_MASK, w != 0u)) {
                    if (nb >= 4) {
                        block(0); block(1); block(2); block(3);
                    } else {
#pragma unroll 1
                        for (int q = 0; q < nb; ++q) block(q);
                    }
                }
                w = wn;
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {                      // A was {0, 2}
                acc[nt][0] *= 0.5f; acc[nt][1] *= 0.5f; acc[nt][2] *= 0.5f; acc[nt][3] *= 0.5f;
            }
        } else if (!dup) {
            for (int kt = 0; kt < tiles; ++kt) {
                uint32_t a[4];
                if (!adj_fragment(bm, wpr, row0, kt, t, a)) continue;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int idx = ((nt * 8 + g) * S + kt * 16 + 2 * t) >> 1;
                        mma_f16(acc[nt], a, in32[pl][idx], in32[pl][idx + 4]);
                    }
            }
        } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = half ? row1 : row0;
                if (row >= n) continue;
                for (int e = rp[row] - 1; e < rp[row + 1]; ++e) {
                    const int j = e < rp[row] ? row : col_g[e] - base;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int idx = (nt * 8 + 2 * t + q) * S + j;
                            acc[nt][2 * half + q] += __half2float(P[idx]) + __half2float(P[kHid * S + idx]);
                        }
                }
            }
        }
        const float c0 = cs[row0], c1 = cs[row1];                        // 0 on padding rows
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            acc[nt][0] *= c0; acc[nt][1] *= c0; acc[nt][2] *= c1; acc[nt][3] *= c1;
        }
        {   // dh planes [channel][node] (scaled), the A operand of dW = dh^T x_in
            __half* oh = DH;
            __half* ol = DH + kHid * Sd;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ch = nt * 8 + 2 * t + q;
                    store_split(oh, ol, ch * Sd + row0, acc[nt][q]);
                    store_split(oh, ol, ch * Sd + row1, acc[nt][2 + q]);
                }
        }
        if (wp) {
            float y[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) y[nt][0] = y[nt][1] = y[nt][2] = y[nt][3] = 0.f;
            const uint32_t* wh = reinterpret_cast<const uint32_t*>(wp);
            const uint32_t* wl = reinterpret_cast<const uint32_t*>(wp + kHid * kWPad);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t ah[4], al[4];
                split2(acc[2 * kk][0], acc[2 * kk][1], ah[0], al[0]);
                split2(acc[2 * kk][2], acc[2 * kk][3], ah[1], al[1]);
                split2(acc[2 * kk + 1][0], acc[2 * kk + 1][1], ah[2], al[2]);
                split2(acc[2 * kk + 1][2], acc[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int widx = ((nt * 8 + g) * kWPad + kk * 16 + 2 * t) >> 1;
                    const uint32_t h0 = wh[widx], h1 = wh[widx + 4];
                    const uint32_t l0 = wl[widx], l1 = wl[widx + 4];
                    mma_f16(y[nt], ah, h0, h1);
                    mma_f16(y[nt], al, h0, h1);
                    mma_f16(y[nt], ah, l0, l1);
                }
            }
            if (c5.on) {                                 // + dz W5[:, slice] (still scaled)
                float gz[4][4];
                dz_w5_tile(c5, S, mt, offx, lane, gz);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    y[nt][0] += gz[nt][0]; y[nt][1] += gz[nt][1]; y[nt][2] += gz[nt][2]; y[nt][3] += gz[nt][3];
                }
            }
            // unscale, add the pooled gradient of x_in's slice, store as the next layer's G
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = half ? row1 : row0;
                if (row >= n) continue;
                const int r = c5.on ? -1 : rank[row];
                const float* gp = r >= 0 ? dp + r * kCat + offx + 2 * t : nullptr;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    float v0 = y[nt][2 * half] * inv_scale, v1 = y[nt][2 * half + 1] * inv_scale;
                    if (gp) { v0 += gp[nt * 8]; v1 += gp[nt * 8 + 1]; }
                    *reinterpret_cast<float2*>(Gout + row * kHid + nt * 8 + 2 * t) = make_float2(v0, v1);
                }
            }
        }
    }
}

__device__ __forceinline__ void bwd_process_graph(const StackBwdMmaParams& p, const Team& tm, int gi,
                                                  int base, int n,
                                                  int fgoff, const unsigned char* shraw,
                                                  const uint32_t* gbm, const int32_t* gbo,
                                                  const int32_t* gfl, bool frag) {
    const int tid = tm.tid, lane = tm.lane, warp = tm.warp;
    const int nthreads = tm.nthreads, nwarps = tm.nwarps;
    const int f = p.f;
    const bool conv5 = p.dh1 != nullptr;
    const GradOffsetsM GO = grad_offsets_m(f, conv5);
    const BwdShared SL = bwd_shared_layout(f, conv5);
    const __half* w2p = reinterpret_cast<const __half*>(shraw + SL.w2p);
    const __half* w3p = reinterpret_cast<const __half*>(shraw + SL.w3p);
    const float* w4s = reinterpret_cast<const float*>(shraw + SL.w4);
    const float* w5x = reinterpret_cast<const float*>(shraw + SL.w5x);    // W5[:, 96], then the column-sum bound
    const int L1 = p.k >> 1;
    const float* dh1g = conv5 ? p.dh1 + (int64_t)gi * kC5b * L1 : nullptr;
    const uint8_t* argg = conv5 ? p.arg + (int64_t)gi * kC5b * L1 : nullptr;
    // db5[c] = sum_j dh1[c][j] over the live pairs: every (c, j) feeds exactly one pooled row, real
    // or padding (a padding row's pre-activation is b5 itself)
    // (a graph split over a CTA pair: each CTA leaves the sums over ITS rows; db5 is rank 0's)
    const bool split = tm.split != 0;
    const int crank = split ? tm.rank : 0;
    // (dh1v / argv: the graph's slab, staged in shared memory when it fits -- see phase 0)
    auto conv5_bias_grad = [&](float* sacc_, const float* dh1v, const uint8_t* argv) {
        for (int c = warp; c < kC5b; c += nwarps) {
            float sb = 0.f;
            if (crank == 0)
                for (int j = lane; j < L1; j += 32) sb += argv[c * L1 + j] != 2 ? dh1v[c * L1 + j] : 0.f;
            sb = warp_sum(sb);
            if (lane == 0) sacc_[GO.b5 + c] = sb;
        }
    };
    // The team leaves this graph's parameter-gradient vector in `sacc` (every entry is
    // written exactly once below); the CTA adds the vectors of a pass in team order.
    if (n == 0) {
        float* sacc0 = reinterpret_cast<float*>(tm.smem + bwd_team_layout(f, 16, conv5, split).sacc);
        for (int idx = tid; idx < GO.total; idx += nthreads) sacc0[idx] = 0.f;
        if (conv5) {
            tm.sync_local();
            conv5_bias_grad(sacc0, dh1g, argg);
        }
        return;
    }
    const int keep = min(n, p.k);
    const int np = (n + 15) & ~15;
    const int wpr = (np + 31) >> 5;
    const int tiles = np >> 4;
    const BwdTeamLayout L = bwd_team_layout(f, np, conv5, split);
    const int S = L.S, Sd = L.Sd;
    // this CTA's row tiles [t_lo, t_hi), rows [r_lo, r_hi) of which [r_lo, n_hi) are nodes
    const int t_lo = split ? (crank ? (tiles + 1) >> 1 : 0) : 0;
    const int t_hi = split ? (crank ? tiles : (tiles + 1) >> 1) : tiles;
    const int r_lo = t_lo * 16, r_hi = t_hi * 16, n_hi = min(r_hi, n);
    unsigned char* sm = tm.smem;
    __half* DZ = reinterpret_cast<__half*>(sm + L.DZ);
    Conv5Bwd c5;
    c5.DZ = DZ; c5.w5t = reinterpret_cast<const __half*>(shraw + SL.w5t); c5.on = conv5;
    // gradient w.r.t. the current layer's output, this graph's rows.  A split graph alternates
    // between two buffers: a CTA writes its rows of the next layer's G while the peer may still
    // read the current one; one cluster barrier per layer orders the rest.
    float* G = p.gws + (int64_t)base * kHid;             // read by phase A (all rows)
    float* Gw = split ? p.gws + ((int64_t)p.num_nodes + base) * kHid : G;   // written by phase B (own rows)
    __half* P = reinterpret_cast<__half*>(sm + L.P);
    __half* DH = reinterpret_cast<__half*>(sm + L.DH) - r_lo;        // [channel][Sd], indexed by node
    __half* vpl = reinterpret_cast<__half*>(sm + L.vpl);
    // adjacency of the own rows, indexed like the whole map
    uint32_t* bm = reinterpret_cast<uint32_t*>(sm + L.bm) - (frag ? t_lo * ((tiles + 3) >> 2) * 32 : r_lo * wpr);
    float* cs = reinterpret_cast<float*>(sm + L.cs);
    float* rs = reinterpret_cast<float*>(sm + L.rs);
    float* hv = reinterpret_cast<float*>(sm + L.hv);
    int* rank = reinterpret_cast<int*>(sm + L.rank);
    int* rp = reinterpret_cast<int*>(sm + L.rp);
    float* red0 = reinterpret_cast<float*>(sm + L.red);
    float* sacc = reinterpret_cast<float*>(sm + L.sacc);

    const bool dup = (gfl[gi] & 1) != 0;
    const int e0 = dup ? p.rowptr_t[base] : 0;
    const int32_t* col_g = p.col_t + e0;
    const float* xc = p.xcat + (int64_t)base * p.ldc;
    const float* dp = p.dpooled + (int64_t)gi * p.k * kCat;
    const int32_t* perm_g = p.perm + (int64_t)gi * p.k;

    if (p.trace && tm.tid == 0) {
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[(int64_t)gi * 16 + 15] = ((int64_t)smid << 32) | (uint32_t)(tm.nthreads | (n << 12));
    }
    KSB_TRACE(0);
    // ---- phase 0: bitmap, coefficients, inverse permutation, gradient scale -----------------
    if (frag) {
        const int gw = ((tiles + 3) >> 2) * 32;
        load_bitmap<8>(p.fragmap + fgoff + t_lo * gw, bm + t_lo * gw, (t_hi - t_lo) * gw, tid, nthreads);
    } else {
        load_bitmap(gbm + gbo[gi] + r_lo * wpr, bm + r_lo * wpr, (r_hi - r_lo) * wpr, tid, nthreads);
    }
    for (int j = tid; j < np; j += nthreads) {
        const float d = j < n ? p.dis[base + j] : 0.f;
        cs[j] = j < n ? col_coef(d, p.norm) : 0.f;
        rs[j] = j < n ? row_coef(d, p.norm) : 0.f;
        rank[j] = -1;
    }
    if (dup)
        for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr_t[base + j] - e0;
    float amax = 0.f;
    // conv5: the graph's d(h1) / arg slab ([16][k/2] floats + bytes) is read three times below (the
    // scale, the dz planes, db5) through dependent indices -- with 2-3 warps per small graph that
    // was half of the graph's whole chain.  One coalesced copy into the P / dh planes' space (free
    // until layer 4), everything else from shared memory.
    const int slab = kC5b * L1;
    const bool staged = conv5 && slab * 5 <= L.DZ - L.P;
    const float* dh1v = dh1g;
    const uint8_t* argv = argg;
    if (staged) {
        uint32_t* stg = reinterpret_cast<uint32_t*>(sm + L.P);
        // (both slabs are multiples of 16 bytes per graph and 16-byte aligned: one round of 16-byte loads)
        const uint4* s4 = reinterpret_cast<const uint4*>(dh1g);
        const uint4* a4 = reinterpret_cast<const uint4*>(argg);
        uint4* d4 = reinterpret_cast<uint4*>(stg);
        const int nd = slab >> 2, na = slab >> 4, nq = nd + na;
        const bool al = ((reinterpret_cast<uintptr_t>(dh1g) | reinterpret_cast<uintptr_t>(argg)) & 15) == 0;
        if (al) {
            for (int i0 = tid; i0 < nq; i0 += nthreads * 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = i0 + u * nthreads;
                    v[u] = idx < nd ? s4[idx] : (idx < nq ? a4[idx - nd] : make_uint4(0u, 0u, 0u, 0u));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = i0 + u * nthreads;
                    if (idx < nq) d4[idx] = v[u];
                }
            }
        } else {
            load_bitmap<8>(reinterpret_cast<const uint32_t*>(dh1g), stg, slab, tid, nthreads);
            load_bitmap<8>(reinterpret_cast<const uint32_t*>(argg), stg + slab, slab >> 2, tid, nthreads);
        }
        dh1v = reinterpret_cast<const float*>(stg);
        argv = reinterpret_cast<const uint8_t*>(stg + slab);
        tm.sync_local();
    }
    auto scatter_ranks = [&]() {                         // inverse permutation of the kept rows
        for (int r = tid; r < keep; r += nthreads) {
            const int node = perm_g[r] - base;
            if ((unsigned)node < (unsigned)n) rank[node] = r;
        }
    };
    if (staged) scatter_ranks();                         // (rank[] was initialised before the barrier above)
    if (conv5) {
        // |pooled gradient| <= max |dz| * max_i sum_c |W5[c][i]|
        for (int idx = tid; idx < slab; idx += nthreads)
            if (argv[idx] != 2) amax = fmaxf(amax, fabsf(dh1v[idx]));
        amax *= w5x[kC5b];
    } else {   // sixteen loads in flight per thread: the sweep is pure memory latency
        const int total = keep * kCat;
        int idx = tid;
        for (; idx + 15 * nthreads < total; idx += 16 * nthreads) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = dp[idx + u * nthreads];
#pragma unroll
            for (int u = 0; u < 16; ++u) amax = fmaxf(amax, fabsf(v[u]));
        }
        for (; idx + 3 * nthreads < total; idx += 4 * nthreads) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = dp[idx + u * nthreads];
#pragma unroll
            for (int u = 0; u < 4; ++u) amax = fmaxf(amax, fabsf(v[u]));
        }
        for (; idx < total; idx += nthreads) amax = fmaxf(amax, fabsf(dp[idx]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(DGCNN_FULL_MASK, amax, o));
    if (lane == 0) red0[warp] = amax;
    tm.sync_local();
    if (!staged) scatter_ranks();
    amax = 0.f;
    for (int w = 0; w < nwarps; ++w) amax = fmaxf(amax, red0[w]);
    tm.sync_local();
    if (!(amax > 0.f) || !(amax < 3.0e38f)) {
        // nothing flows into this graph (or the gradient is not finite: propagate as zeros
        // would hide it, so write NaN-free zeros only for the exact-zero case)
        const float fillv = amax > 0.f ? amax * 0.f + (amax - amax) : 0.f;   // NaN if inf/nan
        for (int idx = tid; idx < GO.total; idx += nthreads) sacc[idx] = fillv;
        return;                                          // (conv5: every live dh1 entry is zero, so is db5)
    }
    // power-of-two scale: max |pooled gradient| -> about 2^6, leaving 2^9 of head room
    int ex;
    frexpf(amax, &ex);
    const float scale = ldexpf(1.f, 6 - ex), inv_scale = ldexpf(1.f, ex - 6);

    if (conv5) {
        // dz[node][c] = dh1[c][r/2] when pooled row r = rank[node] won its pair (arg), else 0; as
        // scaled hi/lo planes [c][node]
        for (int idx = tid; idx < kC5b * S; idx += nthreads) {
            const int ch = idx / S, i = idx - ch * S;
            float v = 0.f;
            const int r = i < n ? rank[i] : -1;
            if (r >= 0 && (r >> 1) < L1 && argv[ch * L1 + (r >> 1)] == (r & 1)) v = dh1v[ch * L1 + (r >> 1)];
            store_split(DZ, DZ + kC5b * S, idx, v * scale);
        }
        conv5_bias_grad(sacc, dh1v, argv);
        tm.sync_local();
    }
    KSB_TRACE(1);
    // ---- layer 4 (32 -> 1) --------------------------------------------------------------------
    float* x4s = reinterpret_cast<float*>(sm + L.DH);    // conv5: x_4 of every node (the dh planes are free until layer 3)
    {
        float dbp = 0.f;
        for (int i = tid; i < np; i += nthreads) {
            float v = 0.f;
            if (i < n) {
                const float y = xc[(int64_t)i * p.ldc + 3 * kHid];
                if (conv5) x4s[i] = y;                   // (for dW5[:, 96] below)
                float gy;
                if (conv5) {                             // dz . W5[:, 96]
                    float sg = 0.f;
#pragma unroll
                    for (int ch = 0; ch < kC5b; ++ch)
                        sg = fmaf(__half2float(DZ[ch * S + i]) + __half2float(DZ[(kC5b + ch) * S + i]), w5x[ch], sg);
                    gy = sg * inv_scale;
                } else {
                    gy = rank[i] >= 0 ? dp[rank[i] * kCat + 3 * kHid] : 0.f;
                }
                const float d = gy * (1.f - y * y);
                if (i >= r_lo && i < r_hi) dbp += d;
                v = rs[i] * d * scale;
            }
            store_split(vpl, vpl + S, i, v);
        }
        dbp = warp_sum(dbp);
        if (lane == 0) red0[warp] = dbp;
    }
    tm.sync_local();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red0[w];
        sacc[GO.b4] = s;
    }
    if (frag) {
        GraphCtx c;
        c.n = n; c.np = np; c.T = tiles; c.G = (tiles + 3) >> 2; c.S = S; c.base = base; c.dup = dup;
        c.fbm = bm; c.rp = rp; c.col_g = col_g; c.cs = cs; c.rs = rs;
        const int g = lane >> 2, t = lane & 3;
        for (int mt = t_lo + warp; mt < t_hi; mt += nwarps) {
            float a4[4];
            aggregate8(c, vpl, S, 1, mt, lane, a4);          // 2 x sum (A fragments are {0, 2})
            if (t == 0) {
                const int row0 = mt * 16 + g, row1 = row0 + 8;
                hv[row0] = cs[row0] * a4[0] * (0.5f * inv_scale);      // dh4
                hv[row1] = cs[row1] * a4[2] * (0.5f * inv_scale);
            }
        }
    } else
    {
        const int g = lane >> 2, t = lane & 3;
        const uint32_t* v32[2] = {reinterpret_cast<const uint32_t*>(vpl),
                                  reinterpret_cast<const uint32_t*>(vpl + S)};
        for (int mt = t_lo + warp; mt < t_hi; mt += nwarps) {
            const int row0 = mt * 16 + g, row1 = row0 + 8;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (!dup) {
                for (int kt = 0; kt < tiles; ++kt) {
                    uint32_t a[4];
                    if (!adj_fragment(bm, wpr, row0, kt, t, a)) continue;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        const int idx = (kt * 16 + 2 * t) >> 1;
                        const uint32_t b0 = g == 0 ? v32[pl][idx] : 0u;
                        const uint32_t b1 = g == 0 ? v32[pl][idx + 4] : 0u;
                        mma_f16(acc, a, b0, b1);
                    }
                }
            } else if (t == 0) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = half ? row1 : row0;
                    if (row >= n) continue;
                    float s = 0.f;
                    for (int e = rp[row] - 1; e < rp[row + 1]; ++e) {
                        const int j = e < rp[row] ? row : col_g[e] - base;
                        s += __half2float(vpl[j]) + __half2float(vpl[S + j]);
                    }
                    acc[2 * half] = s;
                }
            }
            if (t == 0) {
                hv[row0] = cs[row0] * acc[0] * inv_scale;      // dh4
                hv[row1] = cs[row1] * acc[2] * inv_scale;
            }
        }
    }
    tm.sync_local();
    float* xin3 = reinterpret_cast<float*>(P);           // conv5: x3 as fp32 [feature][S] (P is still free)
    const uint32_t* dz32[2] = {reinterpret_cast<const uint32_t*>(DZ),
                               reinterpret_cast<const uint32_t*>(DZ + kC5b * S)};
    if (conv5) {
        // G3[i][k] = dh4[i] w4[k] + pooled gradient of the x3 slice (dz W5[:, 64..95], unscaled),
        // from the tile's registers straight into G
        const int g = lane >> 2, t = lane & 3;
        for (int mt = t_lo + warp; mt < t_hi; mt += nwarps) {
            float gz[4][4];
            dz_w5_tile(c5, S, mt, 2 * kHid, lane, gz);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = mt * 16 + g + 8 * half;
                if (row >= n) continue;
                const float h = hv[row];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float2 w4v = *reinterpret_cast<const float2*>(w4s + nt * 8 + 2 * t);
                    *reinterpret_cast<float2*>(Gw + row * kHid + nt * 8 + 2 * t) =
                        make_float2(fmaf(h, w4v.x, gz[nt][2 * half] * inv_scale),
                                    fmaf(h, w4v.y, gz[nt][2 * half + 1] * inv_scale));
                }
            }
        }
        // dW5[c][96] = sum_i dz[i][c] x4[i]
        for (int ch = warp; ch < kC5b; ch += nwarps) {
            float sw = 0.f;
            for (int i = r_lo + lane; i < n_hi; i += 32)
                sw = fmaf(__half2float(DZ[ch * S + i]) + __half2float(DZ[(kC5b + ch) * S + i]), x4s[i], sw);
            sw = warp_sum(sw);
            if (lane == 0) sacc[GO.w5 + ch * kCat + 3 * kHid] = sw * inv_scale;
        }
    }
    {   // dW4[k] += sum_i dh4[i] x3[i][k];  G3[i][k] = dh4[i] w4[k] + pooled gradient of x3 (conv5: above)
        float dwp = 0.f;
        const float w4k = w4s[lane];
        for (int i0 = r_lo + warp; i0 < r_hi; i0 += 8 * nwarps) {     // eight rows in flight per warp
            float xv[8], gp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * nwarps;
                xv[u] = 0.f; gp[u] = 0.f;
                if (i < n_hi) {
                    xv[u] = xc[(int64_t)i * p.ldc + 2 * kHid + lane];
                    if (!conv5) {
                        const int r = rank[i];
                        if (r >= 0) gp[u] = dp[r * kCat + 2 * kHid + lane];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * nwarps;
                if (i < n_hi) {
                    const float h = hv[i];
                    dwp = fmaf(h, xv[u], dwp);
                    if (!conv5) Gw[i * kHid + lane] = fmaf(h, w4k, gp[u]);
                }
                if (conv5 && i < r_hi) xin3[lane * S + i] = xv[u];   // (zero beyond n: NaN-free MMA padding)
            }
        }
        red0[warp * kHid + lane] = dwp;
    }
    tm.sync_local();
    if (tid < kHid) sacc[GO.w4 + tid] = reduce_rows_t(red0, nwarps, tid);
    if (conv5) {
        // dW5[c][64 + k] = sum_i dz[i][c] x3[i][k]: four 16 x 8 output tiles on the tensor cores
        const int g = lane >> 2, t = lane & 3;
        for (int nk = warp; nk < 4; nk += nwarps) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* xcol = xin3 + (8 * nk + g) * S + 2 * t;
            for (int kt = t_lo; kt < t_hi; ++kt) {
                const float2 x01 = *reinterpret_cast<const float2*>(xcol + kt * 16);
                const float2 x89 = *reinterpret_cast<const float2*>(xcol + kt * 16 + 8);
                const int ia = (g * S + kt * 16 + 2 * t) >> 1;
                uint32_t ah[4], al[4];
                ah[0] = dz32[0][ia]; ah[1] = dz32[0][ia + 4 * S]; ah[2] = dz32[0][ia + 4]; ah[3] = dz32[0][ia + 4 * S + 4];
                al[0] = dz32[1][ia]; al[1] = dz32[1][ia + 4 * S]; al[2] = dz32[1][ia + 4]; al[3] = dz32[1][ia + 4 * S + 4];
                uint32_t bh0, bl0, bh1, bl1;
                split2(x01.x, x01.y, bh0, bl0);
                split2(x89.x, x89.y, bh1, bl1);
                mma_fp16(acc, ah, bh0, bh1);
                mma_fp16(acc, al, bh0, bh1);
                mma_fp16(acc, ah, bl0, bl1);
            }
            float* o = sacc + GO.w5 + g * kCat + 2 * kHid + 8 * nk + 2 * t;
            o[0] = acc[0] * inv_scale; o[1] = acc[1] * inv_scale;
            o[8 * kCat] = acc[2] * inv_scale; o[8 * kCat + 1] = acc[3] * inv_scale;
        }
    }
    tm.sync();                                           // (split: the peer's rows of G are in L2 too)
    if (split) { float* tmp = G; G = Gw; Gw = tmp; }

    KSB_TRACE(2);
    // ---- layers 3, 2, 1 -------------------------------------------------------------------------
#pragma unroll 1
    for (int layer = 3; layer >= 1; --layer) {
        const int offy = (layer - 1) * kHid;
        const int offx = (layer - 2) * kHid;             // slice of x_in (layers 3 and 2)
        // A: P planes <- r * dpre * scale, db
        {
            float dbp = 0.f;
            __half* ph = P;  __half* pl = P + kHid * S;
            for (int i0 = warp; i0 < np; i0 += 8 * nwarps) {     // eight rows in flight per warp
                float yv[8], gv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * nwarps;
                    yv[u] = 0.f; gv[u] = 0.f;
                    if (i < n) {
                        yv[u] = xc[(int64_t)i * p.ldc + offy + lane];
                        gv[u] = __ldcg(G + i * kHid + lane);     // (the peer's rows: written this launch)
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * nwarps;
                    if (i < np) {
                        float sc = 0.f;
                        if (i < n) {
                            const float d = gv[u] * (1.f - yv[u] * yv[u]);
                            if (i >= r_lo && i < r_hi) dbp += d;
                            sc = rs[i] * d * scale;
                        }
                        store_split(ph, pl, lane * S + i, sc);
                    }
                }
            }
            red0[warp * kHid + lane] = dbp;
        }
        tm.sync_local();
        if (tid < kHid) {
            const int ob = layer == 3 ? GO.b3 : (layer == 2 ? GO.b2 : GO.b1);
            sacc[ob + tid] = reduce_rows_t(red0, nwarps, tid);
        }
        KSB_TRACE(3 + 3 * (3 - layer));
        // B: dh planes, and G <- dx (layers 3, 2)
        bwd_mma_layer(P, layer == 3 ? w3p : (layer == 2 ? w2p : nullptr), DH, Sd, Gw, bm, frag, wpr, n, S, t_lo, t_hi,
                      dup, rp, col_g, base, cs, rank, dp, offx, inv_scale, tm, c5);
        tm.sync_local();
        KSB_TRACE(4 + 3 * (3 - layer));
        // C: parameter gradient of the layer
        if (layer >= 2) {
            // dW[c][k] = sum_i dh[i][c] x_in[i][k]: 8 output tiles (2 x 4), one per warp.
            // A = dh planes [c][node]; B = x_in.  The P planes are dead after the aggregation:
            // the team copies x_in (its slice of x_cat, from L2, coalesced, every load in
            // flight) into that space as fp32 [feature][S] -- one round trip instead of one per
            // k-tile -- and the MMA loop splits it hi/lo in registers.  S = np + 8 makes the
            // 8-byte fragment loads conflict-free per half-warp.
            float* xin = reinterpret_cast<float*>(P);
            for (int i0 = r_lo + warp; i0 < r_hi; i0 += 8 * nwarps) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * nwarps;
                    v[u] = i < n_hi ? xc[(int64_t)i * p.ldc + offx + lane] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * nwarps;
                    if (i < r_hi) xin[lane * S + i] = v[u];
                }
            }
            tm.sync_local();
            const int ow = layer == 3 ? GO.w3 : GO.w2;
            const int g = lane >> 2, t = lane & 3;
            const uint32_t* dh32[2] = {reinterpret_cast<const uint32_t*>(DH),
                                       reinterpret_cast<const uint32_t*>(DH + kHid * Sd)};
            // (conv5: four more tiles, dW5[c][offx + k] = sum_i dz[i][c] x_in[i][k], A = the dz planes)
            for (int tile = warp; tile < (conv5 ? 12 : 8); tile += nwarps) {
                const int mc = tile >> 2, nk = tile & 3;
                const bool zt = mc == 2;
                const uint32_t* a_hi = zt ? dz32[0] : dh32[0];
                const uint32_t* a_lo = zt ? dz32[1] : dh32[1];
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const float* xcol = xin + (8 * nk + g) * S + 2 * t;      // feature column of this lane
                const int Sa = zt ? S : Sd;                  // row stride of the A planes
                for (int kt = t_lo; kt < t_hi; ++kt) {
                    const float2 x01 = *reinterpret_cast<const float2*>(xcol + kt * 16);
                    const float2 x89 = *reinterpret_cast<const float2*>(xcol + kt * 16 + 8);
                    const int ia = (((zt ? 0 : 16 * mc) + g) * Sa + kt * 16 + 2 * t) >> 1;
                    uint32_t ah[4], al[4];
                    ah[0] = a_hi[ia]; ah[1] = a_hi[ia + 4 * Sa]; ah[2] = a_hi[ia + 4];
                    ah[3] = a_hi[ia + 4 * Sa + 4];
                    al[0] = a_lo[ia]; al[1] = a_lo[ia + 4 * Sa]; al[2] = a_lo[ia + 4];
                    al[3] = a_lo[ia + 4 * Sa + 4];
                    uint32_t bh0, bl0, bh1, bl1;
                    split2(x01.x, x01.y, bh0, bl0);
                    split2(x89.x, x89.y, bh1, bl1);
                    mma_fp16(acc, ah, bh0, bh1);
                    mma_fp16(acc, al, bh0, bh1);
                    mma_fp16(acc, ah, bl0, bl1);
                }
                const int ld = zt ? kCat : kHid;
                float* o = zt ? sacc + GO.w5 + g * kCat + offx + 8 * nk + 2 * t
                              : sacc + ow + (16 * mc + g) * kHid + 8 * nk + 2 * t;
                o[0] = acc[0] * inv_scale; o[1] = acc[1] * inv_scale;
                o[8 * ld] = acc[2] * inv_scale; o[8 * ld + 1] = acc[3] * inv_scale;
            }
        } else {
            // dW1[c][k] = sum_i dh[i][c] x0[i][k], k < F (any F): FMA, dh rebuilt from its planes.
            // With few outputs (small F) `parts` lanes share one output and split the rows.
            const __half* dhh = DH;
            const __half* dhl = DH + kHid * Sd;
            const int outs = kHid * f;
            int parts = 1;
            while (parts < 32 && outs * parts * 2 <= nthreads) parts <<= 1;
            const int lp = 31 - __clz(parts);
            for (int it0 = 0; it0 < outs * parts; it0 += nthreads) {
                const int item = it0 + tid;
                const int o = item >> lp, part = item & (parts - 1);
                const bool live = o < outs;
                const int c = o & 31, k = o >> 5;
                float a0 = 0.f;
                if (live) {
                    const float* xr = p.x + (int64_t)base * p.ldx + k;
                    int i = r_lo + part;
                    for (; i + 7 * parts < n_hi; i += 8 * parts) {       // eight loads in flight
                        float xv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) xv[u] = xr[(int64_t)(i + u * parts) * p.ldx];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ii = c * Sd + i + u * parts;
                            a0 = fmaf(__half2float(dhh[ii]) + __half2float(dhl[ii]), xv[u], a0);
                        }
                    }
                    for (; i < n_hi; i += parts) {
                        const int ii = c * Sd + i;
                        a0 = fmaf(__half2float(dhh[ii]) + __half2float(dhl[ii]), xr[(int64_t)i * p.ldx], a0);
                    }
                }
                for (int q = parts >> 1; q > 0; q >>= 1) a0 += __shfl_xor_sync(DGCNN_FULL_MASK, a0, q);
                if (live && part == 0) sacc[GO.w1 + c * f + k] = a0 * inv_scale;
            }
        }
        if (layer > 1) {
            tm.sync();                                   // (split: cluster barrier -- G's new rows are visible)
            if (split) { float* tmp = G; G = Gw; Gw = tmp; }
        } else {
            tm.sync_local();
        }
        KSB_TRACE(5 + 3 * (3 - layer));
    }

    KSB_TRACE(12);
}

__global__ void __launch_bounds__(kBwdThreads, 1) stack_bwd_mma_kernel(StackBwdMmaParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ PlanEntry s_plan[kMaxTeams];
    __shared__ int s_count;
    const int f = p.f;
    const bool conv5 = p.dh1 != nullptr;
    const BwdShared SL = bwd_shared_layout(f, conv5);
    const int gtotal = grad_offsets_m(f, conv5).total;
    float* out = p.partials + (int64_t)blockIdx.x * gtotal;      // this CTA's partial vector (running sum)
    // A_hat^T: the forward bitmap when K0 proved the batch symmetric, else the transposed one
    const bool use_t = p.status && (*p.status & DGCNN_GRAPH_GENERIC) && p.bitmap_t;
    const uint32_t* gbm = use_t ? p.bitmap_t : p.bitmap;
    const int32_t* gbo = use_t ? p.bmoff_t : p.bmoff;
    const int32_t* gfl = use_t ? p.gflags_t : p.gflags;
    const bool frag = !use_t && p.fragmap != nullptr;   // symmetric batch: A_hat^T == A_hat

    unsigned char* team_base = smraw + al16(SL.total);
    const int budget = kQuads * bwd_quad_bytes(f, conv5);
    const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsm = gridDim.x, sm = blockIdx.x, B = p.num_graphs;
    constexpr int kWarps = kBwdThreads / 32;
    const int4* gdesc = reinterpret_cast<const int4*>(p.gdesc);
    int next = 0, excl = 0, nsplit = 0, msplit = 0;
    uint32_t crank = 0;                              // rank of this CTA in its cluster
    if (p.pairs) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

    for (int pass = 0;; ++pass) {
        if (warp_id == 0) {
            plan_pass(gdesc, B, nsm, sm, next, excl, pass == 0, budget, kWarps,
                      [f, conv5](int np, bool split) { return bwd_team_layout(f, np, conv5, split).total; }, s_plan,
                      &s_count, nsplit, msplit, p.pairs != 0, p.split_pct);
        } else if (pass == 0) {
            const int tid = threadIdx.x - 32, nthreads = kBwdThreads - 32;
            __half* w2p = reinterpret_cast<__half*>(smraw + SL.w2p);
            __half* w3p = reinterpret_cast<__half*>(smraw + SL.w3p);
            float* w4s = reinterpret_cast<float*>(smraw + SL.w4);
            // transposed planes [k][c]: the "col" operand of dx[k] = sum_c dh[c] W[c][k]
            for (int idx = tid; idx < kHid * kHid; idx += nthreads) {
                const int c = idx >> 5, k = idx & 31;
                store_split(w2p, w2p + kHid * kWPad, k * kWPad + c, p.w2[idx]);
                store_split(w3p, w3p + kHid * kWPad, k * kWPad + c, p.w3[idx]);
            }
            if (tid < kHid) w4s[tid] = p.w4[tid];
            if (conv5) {
                // W5^T as hi/lo planes [column i][channel c] (the "col" operand of dz W5), W5[:, 96],
                // and the bound max_i sum_c |W5[c][i]| that sizes the per-graph gradient scale
                __half* w5t = reinterpret_cast<__half*>(smraw + SL.w5t);
                float* w5x = reinterpret_cast<float*>(smraw + SL.w5x);
                for (int idx = tid; idx < kC5b * kCat; idx += nthreads) {
                    const int c = idx / kCat, i = idx - c * kCat;
                    const float v = p.w5[idx];
                    store_split(w5t, w5t + 104 * kW5TPad, i * kW5TPad + c, v);
                    if (i == 3 * kHid) w5x[c] = v;
                }
                if (tid < 32) {
                    float m = 0.f;
                    for (int i = tid; i < kCat; i += 32) {
                        float sum = 0.f;
                        for (int c = 0; c < kC5b; ++c) sum += fabsf(p.w5[c * kCat + i]);
                        m = fmaxf(m, sum);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(DGCNN_FULL_MASK, m, o));
                    if (tid == 0) w5x[kC5b] = m;
                }
            }
        }
        __syncthreads();                             // the plan (and, first time, the weights)
        const int count = s_count;
        if (count == 0) {
            if (pass == 0)                           // a CTA without graphs still owns a row of the reduction
                for (int idx = threadIdx.x; idx < gtotal; idx += kBwdThreads) out[idx] = 0.f;
            break;
        }
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
                bwd_process_graph(p, tm, e.gi, e.base, e.n, e.fgoff, smraw, gbm, gbo, gfl, frag);
            }
        }
        __syncthreads();                             // every team's vector is complete
        // fixed order (pass by pass, team by team) on a static plan: bit-reproducible sums.  Every
        // thread owns the same elements in every pass, so the running sum needs no barrier of its own.
        for (int idx = threadIdx.x; idx < gtotal; idx += kBwdThreads) {
            float a = pass == 0 ? 0.f : out[idx];
            for (int j = 0; j < count; ++j) {
                const PlanEntry& e = s_plan[j];
                if (e.n > p.nmax) continue;
                const int np = max(16, (e.n + 15) & ~15);
                const float* sacc = reinterpret_cast<const float*>(team_base + e.smem_off +
                                                                   bwd_team_layout(f, np, conv5, e.pad != 0).sacc);
                a += sacc[idx];
            }
            out[idx] = a;
        }
        next += count;
        __syncthreads();                             // s_plan and the teams' vectors are free for the next pass
    }
}

// grads[o] = sum over graphs (deterministic: fixed partition, fixed order).  Block =
// 32 outputs x 8 graph lanes: lane y sums graphs y, y+8, ... (coalesced 128 B rows), then
// the 8 lane sums are added in order.
__global__ void __launch_bounds__(256)
stack_bwd_reduce_graphs(const float* __restrict__ partials, int parts, int total,
                        float* __restrict__ grads) {
    DGCNN_PDL_WAIT();
    __shared__ float red[8][33];
    const int ox = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + ox;
    float s0 = 0.f, s1 = 0.f;
    if (o < total) {
        int b = gy;
        for (; b + 8 < parts; b += 16) {
            s0 += partials[(int64_t)b * total + o];
            s1 += partials[(int64_t)(b + 8) * total + o];
        }
        if (b < parts) s0 += partials[(int64_t)b * total + o];
    }
    red[gy][ox] = s0 + s1;
    __syncthreads();
    if (gy == 0 && o < total) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][ox];
        grads[o] = s;
    }
}

}  // namespace dgcnn

using namespace dgcnn;

static int64_t* g_bwd_trace = nullptr;
extern "C" void dgcnn_stack_bwd_set_trace(int64_t* device_buffer) { g_bwd_trace = device_buffer; }

// Clusters of two CTAs, like the forward kernel (graph_stack_mma.cu): the plan splits the largest
// graphs of the batch over a pair.  -1: not probed yet, 0: off, 1: on.  DGCNN_KS_PAIRS=0 and
// dgcnn_stack_fwd_configure(0, ..) switch both kernels to plain launches.
static int g_bwd_pairs = -1, g_bwd_split_pct = 80;
void dgcnn_stack_bwd_mma_configure(int pairs, int split_pct) {
    g_bwd_pairs = pairs < 0 ? -1 : (pairs ? -2 : 0);     // -2: requested, still to be probed
    if (split_pct > 0) g_bwd_split_pct = split_pct;
}
static int bwd_pairs_enabled() {
    if (g_bwd_pairs >= 0) return g_bwd_pairs;
    int ok = 1;
    if (g_bwd_pairs == -1) {
        const char* env = getenv("DGCNN_KS_PAIRS");
        ok = !(env && env[0] == '0');
    }
    if (ok) ok = cluster_pairs_fit(stack_bwd_mma_kernel, kBwdThreads, (size_t)(kSmemBudget - 1024));
    g_bwd_pairs = ok;
    return ok;
}

int dgcnn_stack_bwd_mma_supported(int32_t f, int64_t max_nodes, bool conv5) {
    if (f < 1 || f > kMaxF || max_nodes < 1 || max_nodes > 1024) return 0;
    const int np = (int)((max_nodes + 15) / 16 * 16);
    const int budget = kQuads * bwd_quad_bytes(f, conv5);
    if (bwd_team_layout(f, np, conv5, false).total <= budget) return 1;
    // larger graphs only as a CTA pair (mandatory split: plan_pass)
    return bwd_team_layout(f, np, conv5, true).total <= budget && bwd_pairs_enabled() ? 1 : 0;
}

static int64_t bwd_grid(int64_t num_graphs, bool pairs) {
    int64_t grid = DGCNN_NUM_SMS;
    if (pairs) {
        if (grid > num_graphs + 8) grid = num_graphs + 8;   // spare CTAs double up on the largest graphs
        grid += grid & 1;
        if (grid > DGCNN_NUM_SMS) grid = DGCNN_NUM_SMS & ~1;
    } else if (grid > num_graphs) {
        grid = num_graphs;
    }
    return grid < 1 ? 1 : grid;
}

size_t dgcnn_stack_bwd_mma_workspace_bytes(int32_t f, int64_t num_graphs, int64_t num_nodes) {
    // (sized for the conv5 variant: one partial vector per CTA; two G buffers)
    return sizeof(float) * ((size_t)grad_offsets_m(f, true).total * (size_t)bwd_grid(num_graphs, true) +
                            2 * (size_t)(num_nodes > 0 ? num_nodes : 0) * kHid) + 1024;
}

int dgcnn_stack_bwd_mma(const float* dpooled, const int32_t* perm, int32_t k, const float* xcat,
                        int64_t ldc, const float* x, int64_t ldx, int32_t f, const int32_t* rowptr_t,
                        const int32_t* col_t, const float* dis, const int32_t* gptr,
                        const int32_t* gorder, const int32_t* gdesc, const uint32_t* fragmap,
                        const uint32_t* bitmap, const int32_t* bmoff,
                        const int32_t* gflags, const uint32_t* bitmap_t, const int32_t* bmoff_t,
                        const int32_t* gflags_t, int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                        const float* w2,
                        const float* w3, const float* w4, int32_t norm, float* grads, int32_t* status,
                        void* workspace, cudaStream_t st, const float* dh1, const uint8_t* arg,
                        const float* w5) {
    const bool conv5 = dh1 != nullptr;
    StackBwdMmaParams p{};
    p.dh1 = dh1; p.arg = arg; p.w5 = w5;
    p.dpooled = dpooled; p.perm = perm; p.k = k; p.xcat = xcat; p.ldc = ldc;
    p.x = x; p.ldx = ldx; p.f = f;
    p.rowptr_t = rowptr_t; p.col_t = col_t; p.dis = dis; p.gptr = gptr; p.gorder = gorder;
    p.gdesc = gdesc; p.fragmap = fragmap;
    p.num_graphs = (int)num_graphs;
    p.bitmap = bitmap; p.bmoff = bmoff; p.gflags = gflags;
    p.bitmap_t = bitmap_t; p.bmoff_t = bmoff_t; p.gflags_t = gflags_t;
    p.w2 = w2; p.w3 = w3; p.w4 = w4; p.norm = norm; p.nmax = (int)max_nodes;
    uintptr_t aligned = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    p.counter = reinterpret_cast<int32_t*>(aligned);
    p.partials = reinterpret_cast<float*>(aligned + 256);
    p.pairs = bwd_pairs_enabled();
    p.split_pct = g_bwd_split_pct;
    p.num_nodes = num_nodes;
    const int64_t grid = bwd_grid(num_graphs, p.pairs != 0);
    p.gws = reinterpret_cast<float*>(
        ((uintptr_t)(p.partials + (size_t)grad_offsets_m(f, conv5).total * (size_t)grid) + 255) &
        ~(uintptr_t)255);
    p.status = status;
    p.trace = g_bwd_trace;
    if (!gdesc) return DGCNN_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)al16(bwd_shared_layout(f, conv5).total) + (size_t)kQuads * bwd_quad_bytes(f, conv5);
    if (cudaFuncSetAttribute(stack_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    if (p.pairs) {
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1] = pdl_attribute();
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kBwdThreads);
        cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (cudaLaunchKernelEx(&cfg, stack_bwd_mma_kernel, p) != cudaSuccess) return DGCNN_ERR_CUDA;
    } else {
        DGCNN_LAUNCH(stack_bwd_mma_kernel, (unsigned)grid, kBwdThreads, smem, st, p);
    }
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    const int total = grad_offsets_m(f, conv5).total;
    DGCNN_LAUNCH(stack_bwd_reduce_graphs, (total + 31) / 32, 256, 0, st, p.partials, (int)grid, total, grads);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
