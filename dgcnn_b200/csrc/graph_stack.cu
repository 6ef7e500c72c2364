// Fused forward of the whole hot path, model.py:28-35, in ONE launch:
//
//   x_1..x_4 = tanh(GCNConv_l(x_{l-1}))   (32/32/32/1 channels, model.py:13-16,30-33)
//   x_cat    = cat(x_1..x_4)  [N,97]       (model.py:34)
//   pooled   = SortAggregation(k)(x_cat)   (model.py:17,35)
//
// One CTA owns one graph at a time (dynamic atomic work queue over the batch).
// Graphs are independent block-diagonal components, so everything between the
// first read of x and the last write of `pooled` stays on chip:
//
//   * the graph's adjacency (with the self loop) is expanded ONCE into a bitmap in
//     shared memory, n x ceil(n/32) words, straight from the int32 CSR -- the only
//     time col[] is read for all four layers;
//   * layer inputs/outputs ping-pong between two [n,32] fp32 shared buffers; the
//     only HBM traffic per layer is the coalesced copy-out of its x_cat slice
//     (needed by backward and by the SortPool gather);
//   * 32-wide aggregations walk bitmap rows with 8 lanes per row (float4 per lane,
//     4 rows per warp): one LDS.128 moves a whole neighbour row for a row group, so
//     the gather runs at the shared-memory crossbar rate rather than the issue rate;
//   * DENSE rows (more than half of the graph adjacent -- COLLAB's near-cliques) are
//     aggregated through the COMPLEMENT: sum_{j in N(i)} h_j = S - sum_{j not in N(i)} h_j
//     with S the per-layer column sum, cutting the gathers from deg to n-1-deg;
//   * scalar layers (layer 1 when F <= 8, layer 4 always) test bitmap bits against a
//     per-node scalar with all 32 lanes (n/32 steps per row);
//   * SortPool ranks the graph's keys in shared memory (rank sort for n <= 256,
//     bitonic above) and gathers the k winning rows of x_cat (L2 hits: this CTA just
//     wrote them) into `pooled`, emitting `perm` for backward.
//
// Multigraphs (duplicate edges cannot live in a bitmap; PyG counts them) are
// detected while the bitmap is built and take a CSR-walking path in the same kernel.
// Graphs larger than the shared-memory budget are not handled here: the host picks
// the per-layer kernels (graph_conv.cu + sort_pool.cu) for such batches.
#include "graph_stack.cuh"
#include "sort_key.cuh"

namespace dgcnn {

// shared-memory carve-up, in 4-byte words; every region starts 16-byte aligned
struct StackLayout {
    int w1t, w2t, w3t, w4, b1, b2, b3, colsum, red, bufA, bufB, bm, cs, rs, v, key, order, rp, total;
};


__host__ __device__ inline StackLayout stack_layout(int f, int nmax, int nwarps) {
    StackLayout L;
    int o = 0;
    int wpr = (nmax + 31) >> 5;
    L.w1t = o; o += al4(f * kHid);
    L.w2t = o; o += kHid * kHid;
    L.w3t = o; o += kHid * kHid;
    L.w4 = o; o += kHid;
    L.b1 = o; o += kHid;
    L.b2 = o; o += kHid;
    L.b3 = o; o += kHid;
    L.colsum = o; o += kHid;
    L.red = o; o += nwarps * kHid;
    L.bufA = o; o += nmax * kHid;
    L.bufB = o; o += nmax * kHid;
    L.bm = o; o += al4(nmax * wpr);
    L.cs = o; o += al4(nmax);
    L.rs = o; o += al4(nmax);
    L.v = o; o += al4(nmax);
    L.key = L.rs;     // x_4 overwrites r_i in place (same index, same thread, layer 4)
    L.order = L.cs;   // c_j is dead once layer 3 has emitted v
    L.rp = o; o += al4(nmax + 1);
    L.total = o;
    return L;
}

// 32-wide aggregation of one layer, float4 layout (lane = 8*group + q; group -> row,
// q -> channels 4q..4q+3).  `in` rows are already scaled by c_j.
//   PROJECT: y = tanh(r_i * agg @ Wt + b)   else  y = tanh(r_i * agg + b)
//   EMIT_H4: also v[i] = c_i * (y . w4)  -- layer 4's projected, pre-scaled input
template <bool PROJECT, bool EMIT_H4>
__device__ __forceinline__ void aggregate32(const float* __restrict__ in, float* __restrict__ out,
                                            const uint32_t* __restrict__ bm, int wpr, int n, bool dup,
                                            const int* __restrict__ rp,
                                            const int32_t* __restrict__ col_g, int base,
                                            const float* __restrict__ rs, const float* __restrict__ colsum,
                                            const float* __restrict__ wt, const float* __restrict__ bias,
                                            const float* __restrict__ w4s, const float* __restrict__ cs,
                                            float* __restrict__ v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane >> 3, q = lane & 7;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    const float4* wt4 = reinterpret_cast<const float4*>(wt);
    const float4 bias4 = reinterpret_cast<const float4*>(bias)[q];
    const uint32_t tailmask = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
    for (int i0 = warp * 4; i0 < n; i0 += nwarps * 4) {
        const int i = i0 + grp;
        const bool active = i < n;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            acc = gather_row32(in4, bm, wpr, n, dup, rp, col_g, base, colsum, i, q, tailmask);
            const float r = rs[i];
            acc = make_float4(acc.x * r, acc.y * r, acc.z * r, acc.w * r);
        }
        float4 y;
        if (PROJECT) y = project32(acc, bias4, wt4, lane, q);
        else y = f4_add(acc, bias4);
        y = make_float4(tanhf(y.x), tanhf(y.y), tanhf(y.z), tanhf(y.w));
        if (active) reinterpret_cast<float4*>(out)[i * 8 + q] = y;
        if (EMIT_H4) {
            const float4 w4v = reinterpret_cast<const float4*>(w4s)[q];
            float part = y.x * w4v.x + y.y * w4v.y + y.z * w4v.z + y.w * w4v.w;
            part += __shfl_xor_sync(DGCNN_FULL_MASK, part, 1);
            part += __shfl_xor_sync(DGCNN_FULL_MASK, part, 2);
            part += __shfl_xor_sync(DGCNN_FULL_MASK, part, 4);
            if (active && q == 0) v[i] = cs[i] * part;
        }
    }
}

__global__ void __launch_bounds__(kStackMaxThreads, 2) stack_fwd_kernel(StackFwdParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_graph;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const StackLayout L = stack_layout(p.f, p.nmax, nwarps);
    float* w1t = sm + L.w1t;  float* w2t = sm + L.w2t;  float* w3t = sm + L.w3t;
    float* w4s = sm + L.w4;   float* b1s = sm + L.b1;   float* b2s = sm + L.b2;  float* b3s = sm + L.b3;
    float* colsum = sm + L.colsum;  float* red = sm + L.red;
    float* bufA = sm + L.bufA;  float* bufB = sm + L.bufB;
    uint32_t* bm = reinterpret_cast<uint32_t*>(sm + L.bm);
    float* cs = sm + L.cs;  float* rs = sm + L.rs;  float* v = sm + L.v;  float* key = sm + L.key;
    int* order = reinterpret_cast<int*>(sm + L.order);
    int* rp = reinterpret_cast<int*>(sm + L.rp);
    const int f = p.f, nmax = p.nmax;
    const float b4 = p.b4 ? p.b4[0] : 0.f;

    // weights once per CTA, transposed to k-major so that a lane reads consecutive channels
    for (int idx = tid; idx < f * kHid; idx += nthreads) {
        int c = idx / f, k = idx - c * f;
        w1t[k * kHid + c] = p.w1[idx];
    }
    for (int idx = tid; idx < kHid * kHid; idx += nthreads) {
        int c = idx >> 5, k = idx & 31;
        w2t[k * kHid + c] = p.w2[idx];
        w3t[k * kHid + c] = p.w3[idx];
    }
    if (tid < kHid) {
        w4s[tid] = p.w4[tid];
        b1s[tid] = p.b1 ? p.b1[tid] : 0.f;
        b2s[tid] = p.b2 ? p.b2[tid] : 0.f;
        b3s[tid] = p.b3 ? p.b3[tid] : 0.f;
    }
    __syncthreads();

    for (;;) {
        if (tid == 0) s_graph = atomicAdd(p.counter, 1);
        __syncthreads();
        if (s_graph >= p.num_graphs) break;
        const int g = p.gorder ? p.gorder[s_graph] : s_graph;
        const int base = p.gptr[g];
        int n = p.gptr[g + 1] - base;
        if (n > nmax) {                       // host promised this cannot happen
            if (tid == 0 && p.status) atomicOr(p.status, DGCNN_GRAPH_BAD_BATCH);
            n = 0;
        }
        const int keep = min(n, p.k);
        float* pooled_g = p.pooled + (int64_t)g * p.k * kCat;
        int32_t* perm_g = p.perm + (int64_t)g * p.k;
        // zero padding first (independent of everything else)
        for (int idx = keep * kCat + tid; idx < p.k * kCat; idx += nthreads) pooled_g[idx] = 0.f;
        for (int r = keep + tid; r < p.k; r += nthreads) perm_g[r] = -1;
        if (n == 0) { __syncthreads(); continue; }

        const int wpr = (n + 31) >> 5;                   // == ceil(round16(n) / 32), K0b's row stride
        const bool dup = (p.gflags[g] & 1) != 0;         // multigraph: walk the CSR instead
        const int e0 = dup ? p.rowptr[base] : 0;
        const int32_t* col_g = p.col + e0;
        float* xc = p.xcat + (int64_t)base * p.ldc;

        // ---- phase 0: adjacency bitmap (from K0b), per-node coefficients ----------------
        load_bitmap(p.bitmap + p.bmoff[g], bm, n * wpr, tid, nthreads);
        for (int j = tid; j < n; j += nthreads) {
            const float d = p.dis[base + j];
            cs[j] = col_coef(d, p.norm);
            rs[j] = row_coef(d, p.norm);
        }
        if (dup)
            for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr[base + j] - e0;
        __syncthreads();

        // ---- layer 1: F -> 32 --------------------------------------------------------
        if (f <= kSmallF) {
            // aggregate first: stage c_j * x_j channel-major in bufB, scalar sums per channel
            float* xs = bufB;
            for (int idx = tid; idx < n * f; idx += nthreads) {
                int j = idx / f, k = idx - j * f;
                xs[k * nmax + j] = cs[j] * p.x[(int64_t)(base + j) * p.ldx + k];
            }
            __syncthreads();
            for (int i = warp; i < n; i += nwarps) {
                const float r = rs[i];
                float acc = b1s[lane];
                for (int k = 0; k < f; ++k) {
                    const float a = r * scalar_row_sum(xs + k * nmax, bm + i * wpr, wpr, dup, rp,
                                                       col_g, base, i);
                    acc = fmaf(a, w1t[k * kHid + lane], acc);
                }
                bufA[i * kHid + lane] = tanhf(acc);
            }
        } else {
            // project first: H = x W1^T into bufB, then the 32-wide aggregation
            for (int j = warp; j < n; j += nwarps) {
                const float* xr = p.x + (int64_t)(base + j) * p.ldx;
                float acc = 0.f;
                for (int k = 0; k < f; ++k) acc = fmaf(xr[k], w1t[k * kHid + lane], acc);
                bufB[j * kHid + lane] = cs[j] * acc;
            }
            __syncthreads();
            {   // column sums of the scaled H for the complement trick
                float part = 0.f;
                for (int j = warp; j < n; j += nwarps) part += bufB[j * kHid + lane];
                red[warp * kHid + lane] = part;
                __syncthreads();
                if (tid < kHid) {
                    float s = 0.f;
                    for (int w = 0; w < nwarps; ++w) s += red[w * kHid + tid];
                    colsum[tid] = s;
                }
                __syncthreads();
            }
            aggregate32<false, false>(bufB, bufA, bm, wpr, n, dup, rp, col_g, base, rs, colsum,
                                      nullptr, b1s, nullptr, nullptr, nullptr);
        }
        __syncthreads();

        // ---- layers 2 and 3: 32 -> 32, ping-pong A -> B -> A ---------------------------
#pragma unroll 1
        for (int layer = 1; layer <= 2; ++layer) {
            float* bin = (layer == 1) ? bufA : bufB;
            float* bout = (layer == 1) ? bufB : bufA;
            // prepare: copy x_layer out to HBM (coalesced rows), scale by c_j in place, column sums
            {
                float part = 0.f;
                float* xo = xc + (layer - 1) * kHid;
                for (int j = warp; j < n; j += nwarps) {
                    float val = bin[j * kHid + lane];
                    xo[(int64_t)j * p.ldc + lane] = val;
                    val *= cs[j];
                    bin[j * kHid + lane] = val;
                    part += val;
                }
                red[warp * kHid + lane] = part;
                __syncthreads();
                if (tid < kHid) {
                    float s = 0.f;
                    for (int w = 0; w < nwarps; ++w) s += red[w * kHid + tid];
                    colsum[tid] = s;
                }
                __syncthreads();
            }
            if (layer == 1)
                aggregate32<true, false>(bin, bout, bm, wpr, n, dup, rp, col_g, base, rs, colsum,
                                         w2t, b2s, nullptr, nullptr, nullptr);
            else
                aggregate32<true, true>(bin, bout, bm, wpr, n, dup, rp, col_g, base, rs, colsum,
                                        w3t, b3s, w4s, cs, v);
            __syncthreads();
        }

        // ---- layer 4: 32 -> 1 (already projected into v), plus copy-out of x_3 -----------
        for (int j = warp; j < n; j += nwarps)
            xc[(int64_t)j * p.ldc + 2 * kHid + lane] = bufA[j * kHid + lane];
        for (int i = warp; i < n; i += nwarps) {
            const float s = scalar_row_sum(v, bm + i * wpr, wpr, dup, rp, col_g, base, i);
            if (lane == 0) {
                const float x4 = tanhf(fmaf(rs[i], s, b4));
                key[i] = x4;
                xc[(int64_t)i * p.ldc + 3 * kHid] = x4;
            }
        }
        __syncthreads();

        // ---- SortPool: order by x_4 descending, ties by node index -----------------------
        uint64_t* comp = reinterpret_cast<uint64_t*>(bufB);      // x_2 is dead by now
        if (n <= 256) {
            for (int j = tid; j < n; j += nthreads)
                comp[j] = ((uint64_t)descending_key_bits(key[j]) << 32) | (uint32_t)j;
            __syncthreads();
            for (int i = tid; i < n; i += nthreads) {
                const uint64_t mine = comp[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += comp[j] < mine;
                if (rank < keep) order[rank] = i;
            }
        } else {
            const uint32_t pw = next_pow2((uint32_t)n);
            for (uint32_t j = tid; j < pw; j += nthreads)
                comp[j] = (j < (uint32_t)n)
                              ? (((uint64_t)descending_key_bits(key[j]) << 32) | j) : ~0ull;
            bitonic_sort_block(comp, pw);
            for (int r = tid; r < keep; r += nthreads) order[r] = (int)(uint32_t)(comp[r] & 0xffffffffu);
        }
        __syncthreads();

        // ---- gather the k winners (rows of x_cat this CTA just wrote: L2 hits) ------------
        // flat index over keep*97 elements: contiguous writes, row-contiguous reads, and
        // several independent loads in flight per thread
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
        __syncthreads();   // shared buffers are reused by the next graph
    }
}

}  // namespace dgcnn

using namespace dgcnn;

int dgcnn_stack_fwd_fma_supported(int32_t num_features, int64_t max_nodes) {
    if (num_features < 1 || num_features > kMaxF || max_nodes < 1 || max_nodes > 4096) return 0;
    const int nmax = stack_nmax_for(max_nodes);
    const StackLayout L = stack_layout(num_features, nmax, stack_threads_for(nmax) / 32);
    return (size_t)L.total * 4 + 64 <= (size_t)kSmemBudget ? 1 : 0;
}

// FMA-gather variant of dgcnn_stack_fwd (argument checks and the work-queue reset are
// done by the dispatcher in graph_stack_mma.cu)
int dgcnn_stack_fwd_fma(const float* x, int64_t ldx, int32_t num_features, const int32_t* rowptr,
                        const int32_t* col, const float* dis, const int32_t* gptr, const int32_t* gorder,
                        const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                        int64_t num_nodes, int64_t num_graphs, int64_t max_nodes, const float* w1,
                        const float* b1,
                        const float* w2, const float* b2, const float* w3, const float* b3,
                        const float* w4, const float* b4, float* xcat, int64_t ldc, float* pooled,
                        int32_t* perm, int32_t k, int32_t norm, int32_t* status, int32_t* counter,
                        cudaStream_t st) {
    (void)num_nodes;
    StackFwdParams p{};
    p.x = x; p.ldx = ldx; p.f = num_features;
    p.rowptr = rowptr; p.col = col; p.dis = dis; p.gptr = gptr; p.num_graphs = (int)num_graphs;
    p.bitmap = bitmap; p.bmoff = bmoff; p.gflags = gflags;
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.w4 = w4; p.b4 = b4;
    p.xcat = xcat; p.ldc = ldc; p.pooled = pooled; p.perm = perm; p.k = k;
    p.norm = norm; p.nmax = stack_nmax_for(max_nodes);
    p.gorder = gorder; p.counter = counter; p.status = status;

    const int threads = stack_threads_for(p.nmax);
    const size_t smem = (size_t)stack_layout(p.f, p.nmax, threads / 32).total * 4;
    if (cudaFuncSetAttribute(stack_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_fwd_kernel, threads, smem) !=
            cudaSuccess || per_sm < 1)
        per_sm = 1;
    int64_t grid = (int64_t)per_sm * DGCNN_NUM_SMS;
    if (grid > num_graphs) grid = num_graphs;
    DGCNN_LAUNCH(stack_fwd_kernel, (unsigned)grid, threads, smem, st, p);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
