// Tensor-core version of the fused forward (model.py:28-35 in one launch; see
// graph_stack.cu for the data flow).  Same bitmap-in-shared-memory design, but the
// aggregations  A_hat . H  and the 32x32 projections run on the tensor cores:
//
//   * A (adjacency + self loop) is 0/1, hence EXACT in fp16.  Its m16k16 fragments are
//     expanded from the bitmap directly into registers (a few shifts and masks per
//     k-tile); all-zero 16x16 blocks are skipped, so sparse graphs stay cheap.
//   * H must keep fp32 accuracy (1e-5 parity bar): every value is split as
//     h = hi + lo with hi = fp16(h), lo = fp16(h - hi)  (22 mantissa bits; inputs are
//     tanh outputs times c_j <= 1, so no range problem), stored as two fp16 planes
//     [channel][node] in shared memory, and multiplied in two MMAs that accumulate in
//     fp32.  The products are exact; only the fp32 accumulation rounds.
//   * the projection (r_i * agg) @ W^T reuses the accumulator fragments as A operands
//     (C layout of two n-tiles == A layout of one k-tile), split hi/lo in registers,
//     against hi/lo planes of W:  hi*hi + lo*hi + hi*lo  (the lo*lo term is < 2^-22).
//
// mma.sync (register fragments) is used on purpose, not tcgen05: the operands are
// generated in registers from a bitmap for graphs of ~30-500 nodes; tcgen05 needs
// shared-memory operand tiles of M >= 64/128 rows in descriptor layouts plus a TMEM
// round trip per tile, which would waste most of each tile on 75-node graphs.
//
// Instruction count per graph drops ~5x against the FMA gather version
// (profiles/r01_stack_fwd_fma.md), which was issue-bound.
#include "graph_mma.cuh"
#include "sort_key.cuh"

namespace dgcnn {


// CTA-wide region: weights, shared by all teams.  Byte offsets, 16-byte aligned.
struct SharedLayout { int w2p, w3p, w1t, misc, total; };

__host__ __device__ inline SharedLayout shared_layout(int f) {
    SharedLayout L;
    int o = 0;
    L.w2p = o; o += 2 * kHid * kWPad * 2;
    L.w3p = o; o += 2 * kHid * kWPad * 2;
    L.w1t = o; o += al16(f * kHid * 4);
    L.misc = o; o += 4 * kHid * 4;                       // w4, b1, b2, b3
    L.total = o;
    return L;
}

// Per-graph region, sized by the graph's own padded node count np (multiple of 16).
struct TeamLayout {
    int PA, PB, vpl, bm, xs, cs, rs, rp, total;
    int S;      // plane row stride in halfs: np + 8
};

__host__ __device__ inline TeamLayout team_layout(int f, int np) {
    TeamLayout L;
    L.S = np + 8;
    const int wpr = (np + 31) >> 5;
    int o = 0;
    L.PA = o; o += 2 * kHid * L.S * 2;                   // hi and lo planes [32][S] fp16
    L.PB = o; o += 2 * kHid * L.S * 2;
    L.vpl = o; o += al16(2 * L.S * 2);                   // layer-4 input: hi and lo [S]
    L.bm = o; o += al16(np * wpr * 4);
    L.xs = o; o += (f <= kSmallF) ? al16(f * np * 4) : 0;
    L.cs = o; o += al16(np * 4);
    L.rs = o; o += al16(np * 4);
    L.rp = o; o += al16((np + 1) * 4);
    L.total = o;
    return L;
}

// bytes of one quad's slice when the CTA takes (almost) all of the SM's shared memory
__host__ __device__ inline int quad_bytes(int f) {
    return ((kSmemBudget - 1024 - shared_layout(f).total) / kQuads) & ~15;
}

// quads a graph of n nodes is given: enough warps for its 16-row tiles, enough memory
__host__ __device__ inline int quads_needed(int f, int n) {
    const int np = (n + 15) & ~15, tiles = np >> 4;
    int q = tiles <= 4 ? 1 : (tiles <= 8 ? 2 : 4);
    const int need = team_layout(f, np < 16 ? 16 : np).total, qb = quad_bytes(f);
    while (q < kQuads && need > q * qb) q <<= 1;
    return q;
}

// One 32-wide layer on the tensor cores for every 16-row tile owned by this warp.
//   in_pl   hi/lo planes [32][S] of c_j * x_{l-1}
//   PROJECT y = tanh((r_i agg) @ W^T + b) with W's hi/lo planes wp, else y = tanh(r_i agg + b)
//   out_pl  (optional) hi/lo planes of c_i * y for the next layer
//   EMIT_V  also v = c_i * (y . w4) as hi/lo fp16 into vpl (layer 4's input)
template <bool PROJECT, bool EMIT_V>
__device__ __forceinline__ void mma_layer(const __half* __restrict__ in_pl, const __half* __restrict__ wp,
                                          const float* __restrict__ bias, __half* __restrict__ out_pl,
                                          __half* __restrict__ vpl, const float* __restrict__ w4s,
                                          const uint32_t* __restrict__ bm, int wpr, int n, int S, bool dup,
                                          const int* __restrict__ rp, const int32_t* __restrict__ col_g,
                                          int base, const float* __restrict__ cs,
                                          const float* __restrict__ rs, float* __restrict__ xo,
                                          int64_t ldc, const Team& tm) {
    const int lane = tm.lane, warp = tm.warp, nwarps = tm.nwarps;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = (n + 15) >> 4;
    const uint32_t* in32[2] = {reinterpret_cast<const uint32_t*>(in_pl),
                               reinterpret_cast<const uint32_t*>(in_pl + kHid * S)};
    for (int mt = warp; mt < tiles; mt += nwarps) {
        const int row0 = mt * 16 + g, row1 = row0 + 8;
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        if (!dup) {
            for (int kt = 0; kt < tiles; ++kt) {
                uint32_t a[4];
                if (!adj_fragment(bm, wpr, row0, kt, t, a)) continue;       // empty 16x16 block
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int idx = ((nt * 8 + g) * S + kt * 16 + 2 * t) >> 1;
                        mma_f16(acc[nt], a, in32[pl][idx], in32[pl][idx + 4]);
                    }
                }
            }
        } else {
            // multigraph: duplicates have no bitmap encoding (PyG counts them): walk the CSR
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = half ? row1 : row0;
                if (row >= n) continue;
                for (int e = rp[row] - 1; e < rp[row + 1]; ++e) {
                    const int j = e < rp[row] ? row : col_g[e] - base;       // first the self loop
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int idx = (nt * 8 + 2 * t + q) * S + j;
                            acc[nt][2 * half + q] += __half2float(in_pl[idx]) +
                                                     __half2float(in_pl[kHid * S + idx]);
                        }
                }
            }
        }
        const float r0 = rs[row0], r1 = rs[row1];                           // 0 on padding rows
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            acc[nt][0] *= r0; acc[nt][1] *= r0; acc[nt][2] *= r1; acc[nt][3] *= r1;
        }
        float y[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float bx = bias[nt * 8 + 2 * t], by = bias[nt * 8 + 2 * t + 1];
            y[nt][0] = bx; y[nt][1] = by; y[nt][2] = bx; y[nt][3] = by;
        }
        if (PROJECT) {
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
        } else {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                y[nt][0] += acc[nt][0]; y[nt][1] += acc[nt][1];
                y[nt][2] += acc[nt][2]; y[nt][3] += acc[nt][3];
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            y[nt][0] = tanhf(y[nt][0]); y[nt][1] = tanhf(y[nt][1]);
            y[nt][2] = tanhf(y[nt][2]); y[nt][3] = tanhf(y[nt][3]);
        }
        // x_l to HBM (its slice of x_cat)
        if (row0 < n) {
            float* o = xo + (int64_t)row0 * ldc + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { o[nt * 8] = y[nt][0]; o[nt * 8 + 1] = y[nt][1]; }
        }
        if (row1 < n) {
            float* o = xo + (int64_t)row1 * ldc + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { o[nt * 8] = y[nt][2]; o[nt * 8 + 1] = y[nt][3]; }
        }
        const float c0 = cs[row0], c1 = cs[row1];                           // 0 on padding rows
        if (out_pl) {
            __half* oh = out_pl;
            __half* ol = out_pl + kHid * S;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ch = nt * 8 + 2 * t + q;
                    store_split(oh, ol, ch * S + row0, c0 * y[nt][q]);
                    store_split(oh, ol, ch * S + row1, c1 * y[nt][2 + q]);
                }
        }
        if (EMIT_V) {
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float wa = w4s[nt * 8 + 2 * t], wb = w4s[nt * 8 + 2 * t + 1];
                p0 = fmaf(y[nt][0], wa, fmaf(y[nt][1], wb, p0));
                p1 = fmaf(y[nt][2], wa, fmaf(y[nt][3], wb, p1));
            }
            p0 += __shfl_xor_sync(DGCNN_FULL_MASK, p0, 1);
            p0 += __shfl_xor_sync(DGCNN_FULL_MASK, p0, 2);
            p1 += __shfl_xor_sync(DGCNN_FULL_MASK, p1, 1);
            p1 += __shfl_xor_sync(DGCNN_FULL_MASK, p1, 2);
            if (t == 0) {
                store_split(vpl, vpl + S, row0, c0 * p0);
                store_split(vpl, vpl + S, row1, c1 * p1);
            }
        }
    }
}

// model.py:28-35 for ONE graph, executed by one team
__device__ __forceinline__ void process_graph(const StackFwdParams& p, const Team& tm, int gi,
                                              const unsigned char* shraw) {
    const int tid = tm.tid, lane = tm.lane, warp = tm.warp;
    const int nthreads = tm.nthreads, nwarps = tm.nwarps;
    const int f = p.f;
    const SharedLayout SL = shared_layout(f);
    const __half* w2p = reinterpret_cast<const __half*>(shraw + SL.w2p);
    const __half* w3p = reinterpret_cast<const __half*>(shraw + SL.w3p);
    const float* w1t = reinterpret_cast<const float*>(shraw + SL.w1t);
    const float* w4s = reinterpret_cast<const float*>(shraw + SL.misc);
    const float* b1s = w4s + kHid;  const float* b2s = b1s + kHid;  const float* b3s = b2s + kHid;
    const float b4 = p.b4 ? p.b4[0] : 0.f;

    const int base = p.gptr[gi];
    const int n = p.gptr[gi + 1] - base;
    const int keep = min(n, p.k);
    float* pooled_g = p.pooled + (int64_t)gi * p.k * kCat;
    int32_t* perm_g = p.perm + (int64_t)gi * p.k;
    for (int idx = keep * kCat + tid; idx < p.k * kCat; idx += nthreads) pooled_g[idx] = 0.f;
    for (int r = keep + tid; r < p.k; r += nthreads) perm_g[r] = -1;
    if (n == 0) return;

    const int np = (n + 15) & ~15;                   // rows/columns padded to the MMA tile
    const int wpr = (np + 31) >> 5;
    const TeamLayout L = team_layout(f, np);
    const int S = L.S;
    unsigned char* smraw = tm.smem;
    __half* PA = reinterpret_cast<__half*>(smraw + L.PA);
    __half* PB = reinterpret_cast<__half*>(smraw + L.PB);
    __half* vpl = reinterpret_cast<__half*>(smraw + L.vpl);
    uint32_t* bm = reinterpret_cast<uint32_t*>(smraw + L.bm);
    float* xs = reinterpret_cast<float*>(smraw + L.xs);
    float* cs = reinterpret_cast<float*>(smraw + L.cs);
    float* rs = reinterpret_cast<float*>(smraw + L.rs);
    float* key = rs;                                    // x_4 overwrites r_i in place
    int* order = reinterpret_cast<int*>(cs);            // c_j is dead after layer 3
    int* rp = reinterpret_cast<int*>(smraw + L.rp);

    const bool dup = (p.gflags[gi] & 1) != 0;        // multigraph: walk the CSR instead
    const int e0 = dup ? p.rowptr[base] : 0;
    const int32_t* col_g = p.col + e0;
    float* xc = p.xcat + (int64_t)base * p.ldc;

    // ---- phase 0: adjacency bitmap (from K0b), per-node coefficients, layer-1 input ------
    load_bitmap(p.bitmap + p.bmoff[gi], bm, np * wpr, tid, nthreads);
    for (int j = tid; j < np; j += nthreads) {
        const float d = j < n ? p.dis[base + j] : 0.f;
        cs[j] = j < n ? col_coef(d, p.norm) : 0.f;
        rs[j] = j < n ? row_coef(d, p.norm) : 0.f;
    }
    if (dup)
        for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr[base + j] - e0;
    tm.sync();

    // ---- layer 1: F -> 32 (FMA pipe: arbitrary input range, tiny work) -----------------
    if (f <= kSmallF) {
        for (int idx = tid; idx < n * f; idx += nthreads) {
            int j = idx / f, k = idx - j * f;
            xs[k * np + j] = cs[j] * p.x[(int64_t)(base + j) * p.ldx + k];
        }
        // padding columns of the output planes must be finite zeros (0 * NaN = NaN)
        for (int idx = tid; idx < (np - n) * kHid; idx += nthreads) {
            const int c = idx / (np - n), j = n + idx - c * (np - n);
            PA[c * S + j] = __float2half_rn(0.f);
            PA[kHid * S + c * S + j] = __float2half_rn(0.f);
        }
        tm.sync();
        for (int i = warp; i < n; i += nwarps) {
            const float r = rs[i];
            float acc = b1s[lane];
            for (int k = 0; k < f; ++k) {
                const float a = r * scalar_row_sum(xs + k * np, bm + i * wpr, wpr, dup, rp, col_g, base, i);
                acc = fmaf(a, w1t[k * kHid + lane], acc);
            }
            const float y = tanhf(acc);
            xc[(int64_t)i * p.ldc + lane] = y;
            store_split(PA, PA + kHid * S, lane * S + i, cs[i] * y);
        }
    } else {
        // project first: c_j * (x_j W1^T) as planes in PB, then aggregate on the tensor cores
        for (int j = warp; j < np; j += nwarps) {
            float acc = 0.f;
            if (j < n) {
                const float* xr = p.x + (int64_t)(base + j) * p.ldx;
                for (int k = 0; k < f; ++k) acc = fmaf(xr[k], w1t[k * kHid + lane], acc);
                acc *= cs[j];
                if (fabsf(acc) > 6.0e4f && p.status) atomicOr(p.status, DGCNN_GRAPH_RANGE);
            }
            store_split(PB, PB + kHid * S, lane * S + j, acc);
        }
        tm.sync();
        mma_layer<false, false>(PB, nullptr, b1s, PA, nullptr, nullptr, bm, wpr, n, S, dup, rp, col_g,
                                base, cs, rs, xc, p.ldc, tm);
    }
    tm.sync();

    // ---- layers 2 and 3 on the tensor cores ------------------------------------------
    mma_layer<true, false>(PA, w2p, b2s, PB, nullptr, nullptr, bm, wpr, n, S, dup, rp, col_g, base, cs,
                           rs, xc + kHid, p.ldc, tm);
    tm.sync();
    mma_layer<true, true>(PB, w3p, b3s, nullptr, vpl, w4s, bm, wpr, n, S, dup, rp, col_g, base, cs, rs,
                          xc + 2 * kHid, p.ldc, tm);
    tm.sync();

    // ---- layer 4: 32 -> 1, already projected into v: one 8-wide MMA column ---------------
    {
        const int g = lane >> 2, t = lane & 3;
        const int tiles = np >> 4;
        const uint32_t* v32[2] = {reinterpret_cast<const uint32_t*>(vpl),
                                  reinterpret_cast<const uint32_t*>(vpl + S)};
        for (int mt = warp; mt < tiles; mt += nwarps) {
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
            if (t == 0) {                                  // column 0 of the tile lives in t == 0
                if (row0 < n) {
                    const float x4 = tanhf(fmaf(rs[row0], acc[0], b4));
                    key[row0] = x4;
                    xc[(int64_t)row0 * p.ldc + 3 * kHid] = x4;
                }
                if (row1 < n) {
                    const float x4 = tanhf(fmaf(rs[row1], acc[2], b4));
                    key[row1] = x4;
                    xc[(int64_t)row1 * p.ldc + 3 * kHid] = x4;
                }
            }
        }
    }
    tm.sync();

    // ---- SortPool: order by x_4 descending, ties by node index -----------------------
    uint64_t* comp = reinterpret_cast<uint64_t*>(PA);
    if (n <= 256) {
        for (int j = tid; j < n; j += nthreads)
            comp[j] = ((uint64_t)descending_key_bits(key[j]) << 32) | (uint32_t)j;
        tm.sync();
        for (int i = tid; i < n; i += nthreads) {
            const uint64_t mine = comp[i];
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += comp[j] < mine;
            if (rank < keep) order[rank] = i;
        }
    } else {
        const uint32_t pw = next_pow2((uint32_t)n);
        for (uint32_t j = tid; j < pw; j += nthreads)
            comp[j] = (j < (uint32_t)n) ? (((uint64_t)descending_key_bits(key[j]) << 32) | j) : ~0ull;
        bitonic_sort_team(comp, pw, tm);
        for (int r = tid; r < keep; r += nthreads) order[r] = (int)(uint32_t)(comp[r] & 0xffffffffu);
    }
    tm.sync();

    // ---- gather the k winners (rows of x_cat this team just wrote: L2 hits) ------------
    {
        const int total = keep * kCat;
        for (int i0 = tid; i0 < total; i0 += nthreads * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + u * nthreads;
                if (idx < total) {
                    const int r = idx / kCat, c = idx - r * kCat;
                    v[u] = xc[(int64_t)order[r] * p.ldc + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + u * nthreads;
                if (idx < total) pooled_g[idx] = v[u];
            }
        }
    }
    for (int r = tid; r < keep; r += nthreads) perm_g[r] = base + order[r];
}

__global__ void __launch_bounds__(kCtaThreads, 1) stack_fwd_mma_kernel(StackFwdParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ int s_item[kQuads];
    const int f = p.f;
    const SharedLayout SL = shared_layout(f);
    {   // weights once per CTA: W1 transposed fp32 (layer 1 stays on the FMA pipe), W2/W3 as
        // hi/lo fp16 planes [cout][cin] = the MMA "col" operand of y = agg @ W^T
        const int tid = threadIdx.x, nthreads = blockDim.x;
        __half* w2p = reinterpret_cast<__half*>(smraw + SL.w2p);
        __half* w3p = reinterpret_cast<__half*>(smraw + SL.w3p);
        float* w1t = reinterpret_cast<float*>(smraw + SL.w1t);
        float* w4s = reinterpret_cast<float*>(smraw + SL.misc);
        for (int idx = tid; idx < f * kHid; idx += nthreads) {
            int c = idx / f, k = idx - c * f;
            w1t[k * kHid + c] = p.w1[idx];
        }
        for (int idx = tid; idx < kHid * kHid; idx += nthreads) {
            const int c = idx >> 5, k = idx & 31;
            store_split(w2p, w2p + kHid * kWPad, c * kWPad + k, p.w2[idx]);
            store_split(w3p, w3p + kHid * kWPad, c * kWPad + k, p.w3[idx]);
        }
        if (tid < kHid) {
            w4s[tid] = p.w4[tid];
            w4s[kHid + tid] = p.b1 ? p.b1[tid] : 0.f;
            w4s[2 * kHid + tid] = p.b2 ? p.b2[tid] : 0.f;
            w4s[3 * kHid + tid] = p.b3 ? p.b3[tid] : 0.f;
        }
        __syncthreads();
    }

    const int quad = threadIdx.x / kQuadThreads;
    const int qb = quad_bytes(f);
    unsigned char* team_base = smraw + al16(SL.total);
    const bool can_split = p.gorder != nullptr;     // descending sizes: a group only ever splits
    int first = 0, nq = kQuads;                      // my group = quads [first, first + nq)

    for (;;) {
        Team tm;
        tm.tid = threadIdx.x - first * kQuadThreads;
        tm.nthreads = nq * kQuadThreads;
        tm.warp = tm.tid >> 5;
        tm.nwarps = tm.nthreads >> 5;
        tm.lane = threadIdx.x & 31;
        tm.bar = 1 + first;
        tm.smem = team_base + (size_t)first * qb;

        if (tm.tid == 0) s_item[first] = atomicAdd(p.counter, 1);
        tm.sync();
        int q = s_item[first];
        tm.sync();                                   // everyone has read the slot
        if (q >= p.num_graphs) break;
        const int gi = p.gorder ? p.gorder[q] : q;
        const int n = p.gptr[gi + 1] - p.gptr[gi];
        if (n > p.nmax) {                            // host promised this cannot happen
            if (tm.tid == 0 && p.status) atomicOr(p.status, DGCNN_GRAPH_BAD_BATCH);
            continue;
        }
        if (can_split) {
            const int need = quads_needed(f, n);
            bool refetch = false;
            while (need < nq) {                      // the lower half keeps the graph,
                nq >>= 1;                            // the upper half fetches its own
                if (quad >= first + nq) { first += nq; refetch = true; break; }
            }
            if (refetch) continue;
            tm.tid = threadIdx.x - first * kQuadThreads;
            tm.nthreads = nq * kQuadThreads;
            tm.warp = tm.tid >> 5;
            tm.nwarps = tm.nthreads >> 5;
        }
        process_graph(p, tm, gi, smraw);
        tm.sync();                                   // the slice is reused by the next graph
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

static int mma_supported(int32_t f, int64_t max_nodes) {
    if (f < 1 || f > kMaxF || max_nodes < 1 || max_nodes > 1024) return 0;
    const int np = (int)((max_nodes + 15) / 16 * 16);
    return team_layout(f, np).total <= kQuads * quad_bytes(f) ? 1 : 0;
}

extern "C" int dgcnn_stack_fwd_supported(int32_t num_features, int64_t max_nodes) {
    return mma_supported(num_features, max_nodes);
}

extern "C" size_t dgcnn_stack_fwd_workspace_bytes(void) { return 256; }

extern "C" int dgcnn_stack_fwd(const float* x, int64_t ldx, int32_t num_features,
                               const int32_t* rowptr, const int32_t* col, const float* dis,
                               const int32_t* gptr, const int32_t* gorder,
                               const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                               int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                               const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* w3, const float* b3, const float* w4, const float* b4,
                               float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                               int32_t norm, int32_t variant, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (num_nodes < 0 || num_graphs < 0 || k < 1 || num_features < 1 || ldx < num_features ||
        ldc < kCat)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    if (variant != DGCNN_STACK_MMA && variant != DGCNN_STACK_FMA) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs == 0) return DGCNN_OK;
    if (num_graphs >= INT32_MAX || num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (variant == DGCNN_STACK_MMA ? !mma_supported(num_features, max_nodes)
                                   : !dgcnn_stack_fwd_fma_supported(num_features, max_nodes))
        return DGCNN_ERR_UNSUPPORTED;
    if (!bitmap || !bmoff || !gflags) return DGCNN_ERR_INVALID_ARGUMENT;
    if (max_nodes > 1024) return DGCNN_ERR_UNSUPPORTED;
    if (!rowptr || !dis || !gptr || !w1 || !w2 || !w3 || !w4 || !xcat || !pooled || !perm ||
        (num_nodes > 0 && !x))
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_stack_fwd_workspace_bytes()) return DGCNN_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uintptr_t aligned = ((uintptr_t)workspace + 127) & ~(uintptr_t)127;
    int32_t* counter = reinterpret_cast<int32_t*>(aligned);
    if (cudaMemsetAsync(counter, 0, sizeof(int32_t), st) != cudaSuccess) return DGCNN_ERR_CUDA;
    if (variant == DGCNN_STACK_FMA)
        return dgcnn_stack_fwd_fma(x, ldx, num_features, rowptr, col, dis, gptr, gorder, bitmap, bmoff,
                                   gflags, num_nodes, num_graphs,
                                   max_nodes, w1, b1, w2, b2, w3, b3, w4, b4, xcat, ldc, pooled, perm, k,
                                   norm, status, counter, st);

    StackFwdParams p{};
    p.x = x; p.ldx = ldx; p.f = num_features;
    p.rowptr = rowptr; p.col = col; p.dis = dis; p.gptr = gptr; p.num_graphs = (int)num_graphs;
    p.bitmap = bitmap; p.bmoff = bmoff; p.gflags = gflags;
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.w4 = w4; p.b4 = b4;
    p.xcat = xcat; p.ldc = ldc; p.pooled = pooled; p.perm = perm; p.k = k;
    p.norm = norm; p.nmax = (int)max_nodes;
    p.gorder = gorder; p.counter = counter; p.status = status;
    // one CTA per SM with (almost) all of its shared memory: 4 quad slices + the weights
    const size_t smem = (size_t)al16(shared_layout(p.f).total) + (size_t)kQuads * quad_bytes(p.f);
    if (cudaFuncSetAttribute(stack_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    int64_t grid = DGCNN_NUM_SMS;
    if (grid > num_graphs) grid = num_graphs;
    stack_fwd_mma_kernel<<<(unsigned)grid, kCtaThreads, smem, st>>>(p);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
