// K1 / K3 -- per-layer graph convolution, forward and backward, for ANY graph
// size and any channel widths up to 128 (the fused per-graph kernels in
// graph_stack.cu cover the model's fixed 32/32/32/1 stack when graphs fit a CTA).
//
// Reference call sites: model.py:30-33 torch.tanh(self.convN(x, edge_index)),
// i.e. PyG GCNConv.forward = lin -> propagate(gather, scale, scatter_add) -> +bias.
// PyG materialises three [E+N, Cout] temporaries per layer and scatters with
// atomics; here one warp owns one target row, walks its CSR segment once,
// accumulates in registers in a fixed order (deterministic), applies the
// D^-1/2 scales, the [Cin x Cout] projection, the bias and the tanh before the
// single store -- straight into the layer's column slice of the [N,97] buffer.
//
// Aggregate-first ((A_hat X) W^T instead of A_hat (X W^T)) is used because it
// lets the projection stay row-local; both orders agree to fp32 rounding.
//
// The same kernel runs the backward aggregation: with `swap_coef` the row/column
// scales trade places (A_hat^T), the "features" are dpre = dy*(1-y^2), and the
// row-local matrix is W itself (dx = dh W) instead of W^T.
#include "common.cuh"

namespace dgcnn {

constexpr int kConvThreads = 256;
constexpr int kMaxChannels = 128;

struct AggParams {
    const float* feat;   // [N, fin] rows gathered
    int64_t ldf;
    int fin;
    const int32_t* rowptr;
    const int32_t* col;
    const float* dis;
    const float* mat;    // global [cout, cin] weight
    int mat_transposed;  // 1: smem[k*fout+c] = mat[c*fin+k] (forward); 0: smem = mat (backward)
    const float* bias;   // [fout] or null
    float* agg_out;      // optional [N, fin] dense copy of the aggregated rows (dh)
    float* out;          // optional [N, fout] (ld = ldo)
    int64_t ldo;
    int fout;
    int accumulate;      // out += instead of out =
    int act;
    int norm;
    int swap_coef;       // 0: forward (A_hat), 1: backward (A_hat^T)
    int64_t n;
    // gc_aggregate_vec32 only: when gptr != null, process ONLY the rows of graphs with more than
    // `big_rows` nodes (what gc_aggregate_staged leaves behind)
    const int32_t* gptr;
    int num_graphs;
    int big_rows;
};

__device__ __forceinline__ void load_matrix(const AggParams& p, float* smat, float* sbias) {
    const int total = p.fin * p.fout;
    if (p.out) {
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            if (p.mat_transposed) {
                int c = idx / p.fin, k = idx - c * p.fin;  // mat[c][k], c < fout, k < fin
                smat[k * p.fout + c] = p.mat[idx];
            } else {
                smat[idx] = p.mat[idx];                     // mat[k][c] with k < fin, c < fout
            }
        }
        for (int c = threadIdx.x; c < p.fout; c += blockDim.x) sbias[c] = p.bias ? p.bias[c] : 0.0f;
    }
    __syncthreads();
}

// channel-parallel: lane l owns input channels l, l+32, ... (CPL of them)
template <int CPL>
__global__ void __launch_bounds__(kConvThreads) gc_aggregate_chan(AggParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ float smem[];
    float* smat = smem;
    float* sbias = smem + p.fin * p.fout;
    load_matrix(p, smat, sbias);

    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kConvThreads / 32);
    for (int64_t row = (int64_t)blockIdx.x * (kConvThreads / 32) + (threadIdx.x >> 5); row < p.n;
         row += warps) {
        const float di = p.dis[row];
        const float self_c = p.swap_coef ? row_coef(di, p.norm) : col_coef(di, p.norm);
        const float scale = p.swap_coef ? col_coef(di, p.norm) : row_coef(di, p.norm);
        float agg[CPL];
        const float* fi = p.feat + row * p.ldf;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            int k = lane + 32 * r;
            agg[r] = (k < p.fin) ? self_c * fi[k] : 0.0f;
        }
        const int beg = p.rowptr[row], end = p.rowptr[row + 1];
        for (int base = beg; base < end; base += 32) {
            int e = base + lane;
            int j = 0;
            float cj = 0.0f;
            if (e < end) {
                j = p.col[e];
                float dj = p.dis[j];
                cj = p.swap_coef ? row_coef(dj, p.norm) : col_coef(dj, p.norm);
            }
            const int cnt = min(32, end - base);
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                int jj = __shfl_sync(DGCNN_FULL_MASK, j, t);
                float cc = __shfl_sync(DGCNN_FULL_MASK, cj, t);
                const float* fj = p.feat + (int64_t)jj * p.ldf;
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    int k = lane + 32 * r;
                    if (k < p.fin) agg[r] = fmaf(cc, fj[k], agg[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            agg[r] *= scale;
            int k = lane + 32 * r;
            if (p.agg_out && k < p.fin) p.agg_out[row * p.fin + k] = agg[r];
        }
        if (!p.out) continue;
        float acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = lane + 32 * q;
            acc[q] = (c < p.fout) ? sbias[c] : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            for (int kk = 0; kk < 32; ++kk) {
                int k = 32 * r + kk;
                if (k >= p.fin) break;
                float a = __shfl_sync(DGCNN_FULL_MASK, agg[r], kk);
                const float* mrow = smat + k * p.fout;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int c = lane + 32 * q;
                    if (c < p.fout) acc[q] = fmaf(a, mrow[c], acc[q]);
                }
            }
        }
        float* orow = p.out + row * p.ldo;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = lane + 32 * q;
            if (c < p.fout) {
                float v = p.act == DGCNN_ACT_TANH ? tanhf(acc[q]) : acc[q];
                orow[c] = p.accumulate ? orow[c] + v : v;
            }
        }
    }
}

// edge-parallel: for narrow rows (fin <= FINP <= 8) the lanes split the
// neighbours instead of the channels and combine with warp-shuffle partial sums
template <int FINP>
__global__ void __launch_bounds__(kConvThreads) gc_aggregate_edge(AggParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ float smem[];
    float* smat = smem;
    float* sbias = smem + p.fin * p.fout;
    load_matrix(p, smat, sbias);

    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kConvThreads / 32);
    for (int64_t row = (int64_t)blockIdx.x * (kConvThreads / 32) + (threadIdx.x >> 5); row < p.n;
         row += warps) {
        const float di = p.dis[row];
        const float self_c = p.swap_coef ? row_coef(di, p.norm) : col_coef(di, p.norm);
        const float scale = p.swap_coef ? col_coef(di, p.norm) : row_coef(di, p.norm);
        float agg[FINP];
#pragma unroll
        for (int k = 0; k < FINP; ++k) agg[k] = 0.0f;
        const int beg = p.rowptr[row], end = p.rowptr[row + 1];
        for (int e = beg + lane; e < end; e += 32) {
            int j = p.col[e];
            float dj = p.dis[j];
            float cj = p.swap_coef ? row_coef(dj, p.norm) : col_coef(dj, p.norm);
            const float* fj = p.feat + (int64_t)j * p.ldf;
#pragma unroll
            for (int k = 0; k < FINP; ++k)
                if (k < p.fin) agg[k] = fmaf(cj, fj[k], agg[k]);
        }
        const float* fi = p.feat + row * p.ldf;
#pragma unroll
        for (int k = 0; k < FINP; ++k) {
            float s = warp_sum(agg[k]);
            agg[k] = (k < p.fin) ? scale * fmaf(self_c, fi[k], s) : 0.0f;
            if (p.agg_out && k < p.fin && lane == 0) p.agg_out[row * p.fin + k] = agg[k];
        }
        if (!p.out) continue;
        float* orow = p.out + row * p.ldo;
        for (int c = lane; c < p.fout; c += 32) {
            float acc = sbias[c];
#pragma unroll
            for (int k = 0; k < FINP; ++k)
                if (k < p.fin) acc = fmaf(agg[k], smat[k * p.fout + c], acc);
            float v = p.act == DGCNN_ACT_TANH ? tanhf(acc) : acc;
            orow[c] = p.accumulate ? orow[c] + v : v;
        }
    }
}


// ---- 32-channel rows, vectorised: the kernel for graphs of ANY size ------------------------------
// Eight lanes own one target row (lane q holds channels 4q..4q+3 as a float4), so a warp works on
// FOUR rows at once and a neighbour row is ONE 16-byte load per lane.  Per step a group takes 8
// neighbours: lane q fetches column index q of the step, the indices are shuffled to the group and
// eight independent LDG.128 are in flight per lane (32 rows of 128 B per warp) -- the per-layer path
// is bound by how many neighbour rows are in flight, not by arithmetic (the one-row-per-warp kernel
// above keeps 4).  Needs fin == 32, 16-byte aligned rows (ldf % 4 == 0) and fout <= 32; anything
// else takes the kernels above.  The summation order is the CSR order: deterministic.
constexpr int kVecRows = 4;      // rows per warp

__global__ void __launch_bounds__(kConvThreads) gc_aggregate_vec32(AggParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ float smem[];
    float* smat = smem;                    // [32][32]: smat[k*32 + c], zero for c >= fout
    float* sbias = smem + 32 * 32;
    if (p.out) {
        for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
            const int k = idx >> 5, c = idx & 31;
            float v = 0.f;
            if (!p.mat) v = k == c ? 1.f : 0.f;        // rows were projected first: identity
            else if (c < p.fout) v = p.mat_transposed ? p.mat[c * 32 + k] : p.mat[k * p.fout + c];
            smat[idx] = v;
        }
        for (int c = threadIdx.x; c < 32; c += blockDim.x) sbias[c] = (p.bias && c < p.fout) ? p.bias[c] : 0.0f;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, q = lane & 7, grp = lane >> 3;
    const unsigned gmask = 0xffu << (grp * 8);
    const int64_t warps = (int64_t)gridDim.x * (kConvThreads / 32);
    const int64_t wid = (int64_t)blockIdx.x * (kConvThreads / 32) + (threadIdx.x >> 5);
    // row ranges: the whole batch, or (big_only) one graph above the cap after the other
    const int nrange = p.gptr ? p.num_graphs : 1;
    for (int rg = 0; rg < nrange; ++rg) {
    int64_t r_lo = 0, r_hi = p.n;
    if (p.gptr) {
        r_lo = p.gptr[rg]; r_hi = p.gptr[rg + 1];
        if (r_hi - r_lo <= p.big_rows) continue;
    }
    const int64_t ntask = (r_hi - r_lo + kVecRows - 1) / kVecRows;
    for (int64_t task = wid; task < ntask; task += warps) {
        const int64_t row = r_lo + task * kVecRows + grp;
        const bool live = row < r_hi;
        const int64_t rr = live ? row : r_hi - 1;
        const float di = p.dis[rr];
        const float self_c = p.swap_coef ? row_coef(di, p.norm) : col_coef(di, p.norm);
        const float scale = p.swap_coef ? col_coef(di, p.norm) : row_coef(di, p.norm);
        const float4 fs = *reinterpret_cast<const float4*>(p.feat + rr * p.ldf + 4 * q);
        float4 agg = make_float4(self_c * fs.x, self_c * fs.y, self_c * fs.z, self_c * fs.w);
        const int beg = live ? p.rowptr[rr] : 0, end = live ? p.rowptr[rr + 1] : 0;
        // all four groups of the warp step together (the shuffles are group-local, the loop is not)
        int steps = (end - beg + 7) >> 3;
        steps = max(steps, __shfl_xor_sync(DGCNN_FULL_MASK, steps, 8));
        steps = max(steps, __shfl_xor_sync(DGCNN_FULL_MASK, steps, 16));
        int jn = beg + q < end ? p.col[beg + q] : 0;             // columns of step 0
        for (int it = 0; it < steps; ++it) {
            const int e = beg + it * 8 + q;
            const int j = jn;
            const int en = e + 8;
            jn = en < end ? p.col[en] : 0;                        // next step's columns are in flight
            float cj = 0.0f;
            if (e < end) {
                const float dj = p.dis[j];
                cj = p.swap_coef ? row_coef(dj, p.norm) : col_coef(dj, p.norm);
            }
            float4 v[8];
            float cc[8];
            const int left = end - beg - it * 8;       // neighbours of this group's row in this step
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int jj = __shfl_sync(DGCNN_FULL_MASK, j, grp * 8 + u);
                cc[u] = __shfl_sync(DGCNN_FULL_MASK, cj, grp * 8 + u);
                v[u] = u < left ? *reinterpret_cast<const float4*>(p.feat + (int64_t)jj * p.ldf + 4 * q)
                                : make_float4(0.f, 0.f, 0.f, 0.f);     // (never 0 * NaN)
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                agg.x = fmaf(cc[u], v[u].x, agg.x); agg.y = fmaf(cc[u], v[u].y, agg.y);
                agg.z = fmaf(cc[u], v[u].z, agg.z); agg.w = fmaf(cc[u], v[u].w, agg.w);
            }
        }
        agg.x *= scale; agg.y *= scale; agg.z *= scale; agg.w *= scale;
        if (p.agg_out && live) *reinterpret_cast<float4*>(p.agg_out + row * 32 + 4 * q) = agg;
        if (!p.out) continue;
        // y[4q..4q+3] = b + sum_k a_k M[k][4q..4q+3]; a_k sits in lane k>>2 of the group, component k&3
        float4 y = *reinterpret_cast<const float4*>(sbias + 4 * q);
        const float4* m4 = reinterpret_cast<const float4*>(smat);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float comp = (k & 3) == 0 ? agg.x : ((k & 3) == 1 ? agg.y : ((k & 3) == 2 ? agg.z : agg.w));
            const float a = __shfl_sync(DGCNN_FULL_MASK, comp, grp * 8 + (k >> 2));
            const float4 w = m4[k * 8 + q];
            y.x = fmaf(a, w.x, y.x); y.y = fmaf(a, w.y, y.y); y.z = fmaf(a, w.z, y.z); y.w = fmaf(a, w.w, y.w);
        }
        if (p.act == DGCNN_ACT_TANH) { y.x = tanhf(y.x); y.y = tanhf(y.y); y.z = tanhf(y.z); y.w = tanhf(y.w); }
        if (!live) continue;
        float* orow = p.out + row * p.ldo + 4 * q;
        const bool vec_ok = p.fout == 32 && ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
        if (vec_ok && !p.accumulate) {
            *reinterpret_cast<float4*>(orow) = y;
        } else {
            const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * q + u < p.fout) orow[u] = p.accumulate ? orow[u] + yy[u] : yy[u];
        }
    }
    }
    (void)gmask;
}


// ---- per-graph, neighbour rows staged in shared memory ------------------------------------------
// The vectorised kernel above fetches every neighbour row from L2: 128 B per edge and layer, 650 MB
// per layer on the power-law configuration -- the L2 gather bandwidth is its bound.  Graphs never
// share nodes, so a CTA that owns one graph can read the graph's input rows ONCE (coalesced,
// pre-scaled by c_j), keep them in shared memory (128 B per node: 1728 nodes fit one SM) and serve
// the gather from there: HBM / L2 traffic drops to one read + one write of the rows and the CSR
// columns, the gather runs at shared-memory bandwidth.  Persistent CTAs walk the graphs largest
// first (gorder); a graph that does not fit is left to gc_aggregate_vec32 in `big_only` mode.
constexpr int kStagedThreads = 1024;      // 32 warps: the gather is a chain of dependent loads per warp
constexpr int kStagedMaxRows = 1728;     // 216 KB of rows + 4.2 KB of weights

struct StagedParams {
    AggParams a;
    const int32_t* gptr;
    const int32_t* gorder;     // optional: graph ids by descending size
    int num_graphs;
    int cap_rows;              // rows the dynamic shared buffer holds
};

__global__ void __launch_bounds__(kStagedThreads) gc_aggregate_staged(StagedParams sp) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) float smem[];
    const AggParams& p = sp.a;
    float* smat = smem;                    // [32][32]
    float* sbias = smem + 32 * 32;         // [32]
    float4* rows = reinterpret_cast<float4*>(smem + 32 * 32 + 32);     // [cap_rows][8]
    if (p.out) {
        for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
            const int k = idx >> 5, c = idx & 31;
            float v = 0.f;
            if (!p.mat) v = k == c ? 1.f : 0.f;
            else if (c < p.fout) v = p.mat_transposed ? p.mat[c * 32 + k] : p.mat[k * p.fout + c];
            smat[idx] = v;
        }
        for (int c = threadIdx.x; c < 32; c += blockDim.x) sbias[c] = (p.bias && c < p.fout) ? p.bias[c] : 0.0f;
    }
    const int lane = threadIdx.x & 31, q = lane & 7, grp = lane >> 3;
    const int warp = threadIdx.x >> 5, nwarps = kStagedThreads / 32;
    const bool vec_ok = p.out && p.fout == 32 && ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    for (int gi = blockIdx.x; gi < sp.num_graphs; gi += gridDim.x) {
        const int g = sp.gorder ? sp.gorder[gi] : gi;
        const int base = sp.gptr[g], n = sp.gptr[g + 1] - base;
        if (n <= 0 || n > sp.cap_rows) continue;                  // (too large: the big_only pass takes it)
        __syncthreads();                                          // the previous graph's readers are done
        // stage c_j * x_j, eight 16-byte loads in flight per thread
        for (int i0 = threadIdx.x; i0 < n * 8; i0 += kStagedThreads * 8) {
            float4 v[8];
            float c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + u * kStagedThreads;
                const int r = min(idx, n * 8 - 1) >> 3;
                v[u] = *reinterpret_cast<const float4*>(p.feat + (int64_t)(base + r) * p.ldf + 4 * (idx & 7));
                const float d = p.dis[base + r];
                c[u] = p.swap_coef ? row_coef(d, p.norm) : col_coef(d, p.norm);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = i0 + u * kStagedThreads;
                if (idx < n * 8) rows[idx] = make_float4(c[u] * v[u].x, c[u] * v[u].y, c[u] * v[u].z, c[u] * v[u].w);
            }
        }
        __syncthreads();
        for (int r0 = warp * kVecRows; r0 < n; r0 += nwarps * kVecRows) {
            const int row = r0 + grp;
            const bool live = row < n;
            const int rr = live ? row : n - 1;
            const float di = p.dis[base + rr];
            const float scale = p.swap_coef ? col_coef(di, p.norm) : row_coef(di, p.norm);
            float4 agg = rows[rr * 8 + q];                        // the self loop, already c_i * x_i
            const int beg = live ? p.rowptr[base + rr] : 0, end = live ? p.rowptr[base + rr + 1] : 0;
            int steps = (end - beg + 7) >> 3;
            steps = max(steps, __shfl_xor_sync(DGCNN_FULL_MASK, steps, 8));
            steps = max(steps, __shfl_xor_sync(DGCNN_FULL_MASK, steps, 16));
            int jn = beg + q < end ? p.col[beg + q] - base : 0;       // columns of step 0
            for (int it = 0; it < steps; ++it) {
                const int j = jn;
                const int en = beg + (it + 1) * 8 + q;                // next step's columns: in flight
                jn = en < end ? p.col[en] - base : 0;                 // while this step's rows are added
                const int left = end - beg - it * 8;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int jj = __shfl_sync(DGCNN_FULL_MASK, j, grp * 8 + u);
                    if (u < left) {
                        const float4 v = rows[jj * 8 + q];
                        agg.x += v.x; agg.y += v.y; agg.z += v.z; agg.w += v.w;
                    }
                }
            }
            agg.x *= scale; agg.y *= scale; agg.z *= scale; agg.w *= scale;
            if (p.agg_out && live) *reinterpret_cast<float4*>(p.agg_out + (int64_t)(base + row) * 32 + 4 * q) = agg;
            if (!p.out) continue;
            float4 y = *reinterpret_cast<const float4*>(sbias + 4 * q);
            const float4* m4 = reinterpret_cast<const float4*>(smat);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float comp = (k & 3) == 0 ? agg.x : ((k & 3) == 1 ? agg.y : ((k & 3) == 2 ? agg.z : agg.w));
                const float a = __shfl_sync(DGCNN_FULL_MASK, comp, grp * 8 + (k >> 2));
                const float4 w = m4[k * 8 + q];
                y.x = fmaf(a, w.x, y.x); y.y = fmaf(a, w.y, y.y); y.z = fmaf(a, w.z, y.z); y.w = fmaf(a, w.w, y.w);
            }
            if (p.act == DGCNN_ACT_TANH) { y.x = tanhf(y.x); y.y = tanhf(y.y); y.z = tanhf(y.z); y.w = tanhf(y.w); }
            if (!live) continue;
            float* orow = p.out + (int64_t)(base + row) * p.ldo + 4 * q;
            if (vec_ok && !p.accumulate) {
                *reinterpret_cast<float4*>(orow) = y;
            } else {
                const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (4 * q + u < p.fout) orow[u] = p.accumulate ? orow[u] + yy[u] : yy[u];
            }
        }
    }
}

static bool vec32_ok(const AggParams& p) {
    return p.fin == 32 && (p.ldf & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feat) & 15) == 0 &&
           (!p.out || p.fout <= 32) && p.n > 0;
}

static int launch_aggregate(const AggParams& p, cudaStream_t st) {
    if (vec32_ok(p)) {
        const size_t smem_v = p.out ? sizeof(float) * (32 * 32 + 32) : 0;
        // one task (four rows) per warp whenever the device can hold them: the rows of a batch are
        // independent latency chains
        const int grid_v = grid_for((p.n + kVecRows - 1) / kVecRows, kConvThreads / 32, 16);
        DGCNN_LAUNCH(gc_aggregate_vec32, grid_v, kConvThreads, smem_v, st, p);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        return DGCNN_OK;
    }
    size_t smem = p.out ? sizeof(float) * ((size_t)p.fin * p.fout + p.fout) : 0;
    int grid = grid_for(p.n, kConvThreads / 32, 8);
#define DGCNN_LAUNCH_AGG(KERNEL)                                                              \
    do {                                                                                      \
        if (smem > 48 * 1024 &&                                                               \
            cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                 (int)smem) != cudaSuccess)                                   \
            return DGCNN_ERR_CUDA;                                                            \
        DGCNN_LAUNCH(KERNEL, grid, kConvThreads, smem, st, p);                                          \
    } while (0)
    if (p.fin <= 1) DGCNN_LAUNCH_AGG(gc_aggregate_edge<1>);
    else if (p.fin <= 2) DGCNN_LAUNCH_AGG(gc_aggregate_edge<2>);
    else if (p.fin <= 4) DGCNN_LAUNCH_AGG(gc_aggregate_edge<4>);
    else if (p.fin <= 8) DGCNN_LAUNCH_AGG(gc_aggregate_edge<8>);
    else if (p.fin <= 32) DGCNN_LAUNCH_AGG(gc_aggregate_chan<1>);
    else if (p.fin <= 64) DGCNN_LAUNCH_AGG(gc_aggregate_chan<2>);
    else if (p.fin <= 96) DGCNN_LAUNCH_AGG(gc_aggregate_chan<3>);
    else DGCNN_LAUNCH_AGG(gc_aggregate_chan<4>);
#undef DGCNN_LAUNCH_AGG
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

// Aggregation with the batch's graph structure known (gptr): graphs that fit one SM's shared
// memory go through gc_aggregate_staged, larger ones (max_nodes unknown or above the cap) through
// gc_aggregate_vec32 restricted to their rows.  Falls back to launch_aggregate when the rows are
// not 32 aligned channels.
static int launch_aggregate_graphs(const AggParams& p0, const int32_t* gptr, const int32_t* gorder,
                                   int64_t num_graphs, int64_t max_nodes, cudaStream_t st) {
    // one CTA per graph: worth it only with enough graphs to fill the device (D&D's 64 graphs of a
    // few hundred nodes are better spread row by row)
    if (!gptr || num_graphs < 96 || !vec32_ok(p0)) return launch_aggregate(p0, st);
    const int cap = (max_nodes > 0 && max_nodes < kStagedMaxRows) ? (int)((max_nodes + 15) / 16 * 16) : kStagedMaxRows;
    StagedParams sp{};
    sp.a = p0; sp.gptr = gptr; sp.gorder = gorder; sp.num_graphs = (int)num_graphs; sp.cap_rows = cap;
    const size_t smem = sizeof(float) * (32 * 32 + 32) + (size_t)cap * 128;
    if (cudaFuncSetAttribute(gc_aggregate_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return DGCNN_ERR_CUDA;
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm > 2) per_sm = 2;                                  // 1024 threads each
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)DGCNN_NUM_SMS * per_sm;
    if (grid > num_graphs) grid = num_graphs;
    DGCNN_LAUNCH(gc_aggregate_staged, (unsigned)grid, kStagedThreads, smem, st, sp);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (max_nodes <= 0 || max_nodes > cap) {                     // someone may not have fitted
        AggParams pb = p0;
        pb.gptr = gptr; pb.num_graphs = (int)num_graphs; pb.big_rows = cap;
        const size_t smem_v = pb.out ? sizeof(float) * (32 * 32 + 32) : 0;
        DGCNN_LAUNCH(gc_aggregate_vec32, DGCNN_NUM_SMS * 4, kConvThreads, smem_v, st, pb);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    return DGCNN_OK;
}

// ---- backward, step A: dpre = dy * act'(y), db = column sums -------------------
// block = (CX, 256/CX): x indexes the channel, y a row inside the block's slab
__global__ void __launch_bounds__(256)
gc_bwd_dpre(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy,
            int cout, int64_t n, int act, float* __restrict__ dpre, float* __restrict__ db_part) {
    DGCNN_PDL_WAIT();
    __shared__ float red[256];
    const int c = threadIdx.x;
    const int rows_per_block = blockDim.y;
    float part = 0.0f;
    for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.y; row < n;
         row += (int64_t)gridDim.x * rows_per_block) {
        if (c < cout) {
            float g = dy[row * lddy + c];
            if (act == DGCNN_ACT_TANH) {
                float v = y[row * ldy + c];
                g *= 1.0f - v * v;
            }
            dpre[row * cout + c] = g;
            part += g;
        }
    }
    if (!db_part) return;
    red[threadIdx.y * blockDim.x + c] = part;
    __syncthreads();
    if (threadIdx.y == 0 && c < cout) {
        float s = 0.0f;
        for (int r = 0; r < rows_per_block; ++r) s += red[r * blockDim.x + c];
        db_part[(int64_t)blockIdx.x * cout + c] = s;     // summed in CTA order by gc_reduce_parts
    }
}

// out[o] = sum over `parts` partial vectors, in a fixed order (no float atomics: the
// gradients are bit-reproducible).  Block = 32 outputs x 8 part lanes.
__global__ void __launch_bounds__(256)
gc_reduce_parts(const float* __restrict__ part, int parts, int total, float* __restrict__ out) {
    DGCNN_PDL_WAIT();
    __shared__ float red[8][33];
    const int ox = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + ox;
    float s0 = 0.f, s1 = 0.f;
    if (o < total) {
        int b = py;
        for (; b + 8 < parts; b += 16) {
            s0 += part[(int64_t)b * total + o];
            s1 += part[(int64_t)(b + 8) * total + o];
        }
        if (b < parts) s0 += part[(int64_t)b * total + o];
    }
    red[py][ox] = s0 + s1;
    __syncthreads();
    if (py == 0 && o < total) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][ox];
        out[o] = s;
    }
}

// ---- backward, step C: dw[c][k] = sum_i dh[i][c] * x[i][k] -----------------------
// grid (row slabs, output tiles of 1024): each CTA keeps 4 outputs per thread in
// registers while it strides over 32-row slabs staged in shared memory
constexpr int kDwRows = 32;
__global__ void __launch_bounds__(256)
gc_bwd_dw(const float* __restrict__ dh, const float* __restrict__ x, int64_t ldx, int cin, int cout,
          int64_t n, float* __restrict__ dw_part) {
    DGCNN_PDL_WAIT();
    extern __shared__ float smem[];
    float* sdh = smem;                   // [kDwRows][cout]
    float* sx = smem + kDwRows * cout;   // [kDwRows][cin]
    const int total = cin * cout;
    int oc[4], ok[4];
    float acc[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int o = blockIdx.y * 1024 + m * 256 + threadIdx.x;
        oc[m] = (o < total) ? o / cin : -1;
        ok[m] = (o < total) ? o - oc[m] * cin : 0;
        acc[m] = 0.0f;
    }
    const int64_t slabs = ceil_div(n, kDwRows);
    for (int64_t slab = blockIdx.x; slab < slabs; slab += gridDim.x) {
        const int64_t r0 = slab * kDwRows;
        const int rows = (int)min((int64_t)kDwRows, n - r0);
        for (int idx = threadIdx.x; idx < rows * cout; idx += 256) sdh[idx] = dh[r0 * cout + idx];
        for (int idx = threadIdx.x; idx < rows * cin; idx += 256) {
            int r = idx / cin, k = idx - r * cin;
            sx[idx] = x[(r0 + r) * ldx + k];
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            if (oc[m] >= 0) {
                float a = acc[m];
                for (int r = 0; r < rows; ++r) a = fmaf(sdh[r * cout + oc[m]], sx[r * cin + ok[m]], a);
                acc[m] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
        if (oc[m] >= 0) dw_part[(int64_t)blockIdx.x * total + oc[m] * cin + ok[m]] = acc[m];
}

struct BwdWorkspace {
    float* dpre;
    float* dh;
    float* db_part;   // [grid of gc_bwd_dpre][cout]
    float* dw_part;   // [grid.x of gc_bwd_dw][cin * cout]
    int grid_a, grid_c;
    size_t bytes;
};

__host__ inline BwdWorkspace carve_bwd_workspace(void* base, int64_t n, int cin, int cout) {
    BwdWorkspace w;
    const int cx = (int)next_pow2((uint32_t)cout);
    w.grid_a = grid_for(n, 256 / cx, 8);
    w.grid_c = grid_for(ceil_div(n, kDwRows), 1, 2);
    size_t each = align_up(sizeof(float) * (size_t)n * (size_t)cout, 256);
    size_t dbp = align_up(sizeof(float) * (size_t)w.grid_a * (size_t)cout, 256);
    size_t dwp = align_up(sizeof(float) * (size_t)w.grid_c * (size_t)cin * (size_t)cout, 256);
    char* p = static_cast<char*>(base);
    w.dpre = reinterpret_cast<float*>(p);
    w.dh = reinterpret_cast<float*>(p ? p + each : nullptr);
    w.db_part = reinterpret_cast<float*>(p ? p + 2 * each : nullptr);
    w.dw_part = reinterpret_cast<float*>(p ? p + 2 * each + dbp : nullptr);
    w.bytes = 2 * each + dbp + dwp;
    return w;
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" int dgcnn_graph_conv_fwd(const float* x, int64_t ldx, int32_t cin, const int32_t* rowptr,
                                    const int32_t* col, const float* dis, const float* weight,
                                    const float* bias, float* y, int64_t ldy, int32_t cout,
                                    int64_t num_nodes, int32_t norm, int32_t act, void* stream) {
    if (num_nodes < 0 || cin < 1 || cout < 1 || ldx < cin || ldy < cout)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    if (act != DGCNN_ACT_NONE && act != DGCNN_ACT_TANH) return DGCNN_ERR_INVALID_ARGUMENT;
    if (cin > kMaxChannels || cout > kMaxChannels) return DGCNN_ERR_UNSUPPORTED;
    if (num_nodes == 0) return DGCNN_OK;
    if (!x || !rowptr || !dis || !y) return DGCNN_ERR_INVALID_ARGUMENT;
    // weight == NULL: the rows were projected first (dgcnn_project_rows); only for 32 -> 32 rows
    // that the vectorised kernel can take
    if (!weight && !(cin == 32 && cout == 32 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0))
        return DGCNN_ERR_INVALID_ARGUMENT;
    AggParams p{};
    p.feat = x; p.ldf = ldx; p.fin = cin;
    p.rowptr = rowptr; p.col = col; p.dis = dis;
    p.mat = weight; p.mat_transposed = 1; p.bias = bias;
    p.agg_out = nullptr; p.out = y; p.ldo = ldy; p.fout = cout;
    p.accumulate = 0; p.act = act; p.norm = norm; p.swap_coef = 0; p.n = num_nodes;
    return launch_aggregate(p, static_cast<cudaStream_t>(stream));
}

// dgcnn_graph_conv_fwd with the batch's graph offsets: same result, but every graph that fits one
// SM's shared memory (<= 1728 nodes at 32 channels) is aggregated from rows staged there.
extern "C" int dgcnn_graph_conv_fwd_graphs(const float* x, int64_t ldx, int32_t cin, const int32_t* rowptr,
                                           const int32_t* col, const float* dis, const int32_t* gptr,
                                           const int32_t* gorder, int64_t num_graphs, int64_t max_nodes,
                                           const float* weight, const float* bias, float* y, int64_t ldy,
                                           int32_t cout, int64_t num_nodes, int32_t norm, int32_t act,
                                           void* stream) {
    if (num_nodes < 0 || cin < 1 || cout < 1 || ldx < cin || ldy < cout || num_graphs < 0)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    if (act != DGCNN_ACT_NONE && act != DGCNN_ACT_TANH) return DGCNN_ERR_INVALID_ARGUMENT;
    if (cin > kMaxChannels || cout > kMaxChannels) return DGCNN_ERR_UNSUPPORTED;
    if (num_nodes == 0) return DGCNN_OK;
    if (!x || !rowptr || !dis || !y) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!weight && !(cin == 32 && cout == 32 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0))
        return DGCNN_ERR_INVALID_ARGUMENT;
    AggParams p{};
    p.feat = x; p.ldf = ldx; p.fin = cin;
    p.rowptr = rowptr; p.col = col; p.dis = dis;
    p.mat = weight; p.mat_transposed = 1; p.bias = bias;
    p.agg_out = nullptr; p.out = y; p.ldo = ldy; p.fout = cout;
    p.accumulate = 0; p.act = act; p.norm = norm; p.swap_coef = 0; p.n = num_nodes;
    return launch_aggregate_graphs(p, gptr, gorder, num_graphs, max_nodes, static_cast<cudaStream_t>(stream));
}

extern "C" size_t dgcnn_graph_conv_bwd_workspace_bytes(int64_t num_nodes, int32_t cin, int32_t cout) {
    if (num_nodes < 0 || cout < 1 || cin < 1) return 0;
    return carve_bwd_workspace(nullptr, num_nodes, cin, cout).bytes + 256;
}

extern "C" int dgcnn_graph_conv_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy,
                                    const float* x, int64_t ldx, int32_t cin,
                                    const int32_t* rowptr_t, const int32_t* col_t, const float* dis,
                                    const float* weight, float* dx, int64_t lddx,
                                    int32_t accumulate_dx, float* dw, float* db, int32_t cout,
                                    int64_t num_nodes, int32_t norm, int32_t act, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    const int64_t n = num_nodes;
    if (n < 0 || cin < 1 || cout < 1 || ldx < cin || lddy < cout) return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    if (act != DGCNN_ACT_NONE && act != DGCNN_ACT_TANH) return DGCNN_ERR_INVALID_ARGUMENT;
    if (act == DGCNN_ACT_TANH && (!y || ldy < cout) && n > 0) return DGCNN_ERR_INVALID_ARGUMENT;
    if (cin > kMaxChannels || cout > kMaxChannels) return DGCNN_ERR_UNSUPPORTED;
    if (!dw || (dx && lddx < cin)) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cin * cout, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        if (db && cudaMemsetAsync(db, 0, sizeof(float) * (size_t)cout, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        return DGCNN_OK;
    }
    if (!dy || !x || !rowptr_t || !dis || !weight) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_graph_conv_bwd_workspace_bytes(n, cin, cout))
        return DGCNN_ERR_WORKSPACE;
    uintptr_t aligned = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    BwdWorkspace w = carve_bwd_workspace(reinterpret_cast<void*>(aligned), n, cin, cout);

    // A: dpre, db (per-CTA partial sums, added in CTA order)
    int cx = (int)next_pow2((uint32_t)cout);
    dim3 block_a(cx, 256 / cx);
    DGCNN_LAUNCH(gc_bwd_dpre, w.grid_a, block_a, 0, st, dy, lddy, y, ldy, cout, n, act, w.dpre, db ? w.db_part : nullptr);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (db) {
        DGCNN_LAUNCH(gc_reduce_parts, (cout + 31) / 32, 256, 0, st, w.db_part, w.grid_a, cout, db);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }

    // B: dh = A_hat^T dpre (kept dense for step C), dx (+)= dh W
    AggParams p{};
    p.feat = w.dpre; p.ldf = cout; p.fin = cout;
    p.rowptr = rowptr_t; p.col = col_t; p.dis = dis;
    p.mat = weight; p.mat_transposed = 0; p.bias = nullptr;
    p.agg_out = w.dh; p.out = dx; p.ldo = lddx; p.fout = cin;
    p.accumulate = accumulate_dx; p.act = DGCNN_ACT_NONE; p.norm = norm; p.swap_coef = 1; p.n = n;
    int rc = launch_aggregate(p, st);
    if (rc != DGCNN_OK) return rc;

    // C: dw = dh^T x (per-CTA partial products, added in CTA order)
    dim3 grid_c((unsigned)w.grid_c, (unsigned)ceil_div(cin * cout, 1024));
    size_t smem_c = sizeof(float) * kDwRows * (size_t)(cin + cout);
    DGCNN_LAUNCH(gc_bwd_dw, grid_c, 256, smem_c, st, w.dh, x, ldx, cin, cout, n, w.dw_part);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(gc_reduce_parts, (cin * cout + 31) / 32, 256, 0, st, w.dw_part, w.grid_c, cin * cout, dw);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
