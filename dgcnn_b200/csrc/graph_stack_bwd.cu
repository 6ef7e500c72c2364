// Fused backward of the hot path (autograd of model.py:28-35, train.py:40) for the
// model's fixed widths, one CTA per graph -- the mirror image of graph_stack.cu:
//
//   dpooled [B,k*97] --(SortPool bwd: rows scattered by perm)--> d x_cat
//   layer 4..1:  dpre = dy * (1 - y^2);  db += sum dpre;  dh = A_hat^T dpre;
//                dW  += dh^T x_in;       dx_in = dh W  (+ the pooled gradient of x_in's slice)
//
// Per graph everything lives in shared memory: the bitmap of the TRANSPOSED adjacency
// (from the CSR by source), one [n,32] gradient buffer that is rewritten in place layer
// after layer, and one [n,32] buffer for dh.  The saved activations x_cat and the pooled
// gradient are streamed from HBM/L2 exactly once per use.  The parameter gradients are
// accumulated per CTA in shared memory (each output owned by one thread, graphs visited
// in a fixed order), written as one partial vector per CTA and summed by a second tiny
// kernel in CTA order: deterministic, no float atomics.
//
// Flat gradient layout (= PyG parameter order, model.py:13-16):
//   [ conv1.lin.weight 32xF | conv1.bias 32 | conv2.lin.weight 32x32 | conv2.bias 32 |
//     conv3.lin.weight 32x32 | conv3.bias 32 | conv4.lin.weight 1x32 | conv4.bias 1 ]
// The gradient w.r.t. the input features x is not produced (train.py never needs it);
// callers that do need it use the per-layer kernels.
#include "graph_stack.cuh"

namespace dgcnn {

struct StackBwdParams {
    const float* dpooled; const int32_t* perm; int k;
    const float* xcat; int64_t ldc;
    const float* x; int64_t ldx; int f;
    const int32_t* rowptr_t; const int32_t* col_t; const float* dis; const int32_t* gptr;
    const int32_t* gorder; int num_graphs;
    // K0b bitmaps: of A_hat (used when the batch was proven symmetric) and of A_hat^T
    const uint32_t* bitmap; const int32_t* bmoff; const int32_t* gflags;
    const uint32_t* bitmap_t; const int32_t* bmoff_t; const int32_t* gflags_t;
    const float* w2; const float* w3; const float* w4;
    int norm; int nmax;
    float* partials;     // [gridDim.x][P]
    int32_t* status;
};

struct GradOffsets { int w1, b1, w2, b2, w3, b3, w4, b4, total; };

__host__ __device__ inline GradOffsets grad_offsets(int f) {
    GradOffsets g;
    int o = 0;
    g.w1 = o; o += kHid * f;
    g.b1 = o; o += kHid;
    g.w2 = o; o += kHid * kHid;
    g.b2 = o; o += kHid;
    g.w3 = o; o += kHid * kHid;
    g.b3 = o; o += kHid;
    g.w4 = o; o += kHid;
    g.b4 = o; o += 1;
    g.total = o;
    return g;
}

struct StackBwdLayout {
    int w2, w3, w4, sacc, colsum, red, bufA, bufB, bm, cs, rs, p4, hv, rank, rp, total;
};

__host__ __device__ inline StackBwdLayout stack_bwd_layout(int f, int nmax, int nwarps) {
    StackBwdLayout L;
    int o = 0;
    const int wpr = (nmax + 31) >> 5;
    L.w2 = o; o += kHid * kHid;
    L.w3 = o; o += kHid * kHid;
    L.w4 = o; o += kHid;
    L.sacc = o; o += al4(grad_offsets(f).total);
    L.colsum = o; o += kHid;
    L.red = o; o += 2 * nwarps * kHid;
    L.bufA = o; o += nmax * kHid;
    L.bufB = o; o += nmax * kHid;
    L.bm = o; o += al4(nmax * wpr);
    L.cs = o; o += al4(nmax);
    L.rs = o; o += al4(nmax);
    L.p4 = o; o += al4(nmax);
    L.hv = o; o += al4(nmax);
    L.rank = o; o += al4(nmax);
    L.rp = o; o += al4(nmax + 1);
    L.total = o;
    return L;
}

// sum the per-warp partial rows red[w][0..31] into one value per channel (tid < 32)
__device__ __forceinline__ float reduce_rows(const float* red, int nwarps, int c) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w * kHid + c];
    return s;
}

__global__ void __launch_bounds__(kStackMaxThreads, 2) stack_bwd_kernel(StackBwdParams p) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int grp = lane >> 3, q = lane & 7;
    const int f = p.f, nmax = p.nmax;
    const StackBwdLayout L = stack_bwd_layout(f, nmax, nwarps);
    const GradOffsets G = grad_offsets(f);
    float* w2s = sm + L.w2;  float* w3s = sm + L.w3;  float* w4s = sm + L.w4;
    float* sacc = sm + L.sacc;  float* colsum = sm + L.colsum;
    float* red0 = sm + L.red;  float* red1 = red0 + nwarps * kHid;
    float* bufA = sm + L.bufA;  float* bufB = sm + L.bufB;
    uint32_t* bm = reinterpret_cast<uint32_t*>(sm + L.bm);
    float* cs = sm + L.cs;  float* rs = sm + L.rs;  float* p4 = sm + L.p4;  float* hv = sm + L.hv;
    int* rank = reinterpret_cast<int*>(sm + L.rank);
    int* rp = reinterpret_cast<int*>(sm + L.rp);

    // W2, W3 row-major [c][k] as stored: dx[k] = sum_c dh[c] W[c][k]
    for (int idx = tid; idx < kHid * kHid; idx += nthreads) { w2s[idx] = p.w2[idx]; w3s[idx] = p.w3[idx]; }
    if (tid < kHid) w4s[tid] = p.w4[tid];
    for (int idx = tid; idx < G.total; idx += nthreads) sacc[idx] = 0.f;
    __syncthreads();

    // A_hat^T: the forward bitmap when K0 proved the batch symmetric, else the transposed one
    const bool use_t = p.status && (*p.status & DGCNN_GRAPH_GENERIC) && p.bitmap_t;
    const uint32_t* gbm = use_t ? p.bitmap_t : p.bitmap;
    const int32_t* gbo = use_t ? p.bmoff_t : p.bmoff;
    const int32_t* gfl = use_t ? p.gflags_t : p.gflags;

    // dW tile workers: the first DWW warps own the 32x32 outputs as 4(c) x 2(k) tiles while
    // the other warps run the row-local projection; with 4 warps everybody does both in turn
    const int dww = nwarps >= 8 ? 4 : nwarps;
    const bool split = nwarps >= 8;

    for (int gq = blockIdx.x; gq < p.num_graphs; gq += gridDim.x) {  // fixed order: deterministic
        const int g = p.gorder ? p.gorder[gq] : gq;
        const int base = p.gptr[g];
        int n = p.gptr[g + 1] - base;
        if (n > nmax) {
            if (tid == 0 && p.status) atomicOr(p.status, DGCNN_GRAPH_BAD_BATCH);
            n = 0;
        }
        if (n == 0) continue;
        const int keep = min(n, p.k);
        const int wpr = (n + 31) >> 5;
        const bool dup = (gfl[g] & 1) != 0;
        const int e0 = dup ? p.rowptr_t[base] : 0;
        const int32_t* col_g = p.col_t + e0;
        const float* xc = p.xcat + (int64_t)base * p.ldc;
        const float* dp = p.dpooled + (int64_t)g * p.k * kCat;
        const int32_t* perm_g = p.perm + (int64_t)g * p.k;
        const uint32_t tailmask = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;

        // ---- phase 0 ------------------------------------------------------------------
        load_bitmap(gbm + gbo[g], bm, n * wpr, tid, nthreads);
        for (int j = tid; j < n; j += nthreads) {
            const float d = p.dis[base + j];
            cs[j] = col_coef(d, p.norm);
            rs[j] = row_coef(d, p.norm);
            rank[j] = -1;
        }
        if (dup)
            for (int j = tid; j <= n; j += nthreads) rp[j] = p.rowptr_t[base + j] - e0;
        __syncthreads();
        // rank[node] = pooled row that took this node (the inverse of perm), -1 if truncated
        for (int r = tid; r < keep; r += nthreads) {
            const int node = perm_g[r] - base;
            if ((unsigned)node < (unsigned)n) rank[node] = r;
        }
        __syncthreads();

        // ---- layer 4 (32 -> 1) -----------------------------------------------------------
        {
            float dbp = 0.f;
            for (int i = tid; i < n; i += nthreads) {
                const float y = xc[(int64_t)i * p.ldc + 3 * kHid];
                const float gy = rank[i] >= 0 ? dp[rank[i] * kCat + 3 * kHid] : 0.f;
                const float d = gy * (1.f - y * y);
                dbp += d;
                p4[i] = rs[i] * d;                                   // A_hat^T: row coef rides along
            }
            dbp = warp_sum(dbp);
            if (lane == 0) red0[warp] = dbp;
        }
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w) s += red0[w];
            sacc[G.b4] += s;
        }
        for (int i = warp; i < n; i += nwarps) {
            const float s = scalar_row_sum(p4, bm + i * wpr, wpr, dup, rp, col_g, base, i);
            if (lane == 0) hv[i] = cs[i] * s;                        // dh4[i]
        }
        __syncthreads();
        {   // dW4[k] += sum_i dh4[i] x3[i][k];  G3[i][k] = dh4[i] w4[k] + pooled grad of x3
            float dwp = 0.f;
            const float w4k = w4s[lane];
            for (int i = warp; i < n; i += nwarps) {
                const float h = hv[i];
                dwp = fmaf(h, xc[(int64_t)i * p.ldc + 2 * kHid + lane], dwp);
                const int r = rank[i];
                const float gp = r >= 0 ? dp[r * kCat + 2 * kHid + lane] : 0.f;
                bufA[i * kHid + lane] = fmaf(h, w4k, gp);
            }
            red0[warp * kHid + lane] = dwp;
        }
        __syncthreads();
        if (tid < kHid) sacc[G.w4 + tid] += reduce_rows(red0, nwarps, tid);
        __syncthreads();

        // ---- layers 3, 2, 1 (32 wide outputs) ----------------------------------------------
#pragma unroll 1
        for (int layer = 3; layer >= 1; --layer) {
            const int offy = (layer - 1) * kHid;
            // a: bufA <- r_i * dpre (in place), db partials, column sums for the complement
            {
                float dbp = 0.f, sp = 0.f;
                for (int i = warp; i < n; i += nwarps) {
                    const float y = xc[(int64_t)i * p.ldc + offy + lane];
                    const float d = bufA[i * kHid + lane] * (1.f - y * y);
                    dbp += d;
                    const float sc = rs[i] * d;
                    bufA[i * kHid + lane] = sc;
                    sp += sc;
                }
                red0[warp * kHid + lane] = dbp;
                red1[warp * kHid + lane] = sp;
            }
            __syncthreads();
            if (tid < kHid) {
                const int ob = layer == 3 ? G.b3 : (layer == 2 ? G.b2 : G.b1);
                sacc[ob + tid] += reduce_rows(red0, nwarps, tid);
                colsum[tid] = reduce_rows(red1, nwarps, tid);
            }
            __syncthreads();
            // b: bufB <- dh = c_i * sum_{d in out(i) U {i}} bufA[d]
            {
                const float4* in4 = reinterpret_cast<const float4*>(bufA);
                float4* out4 = reinterpret_cast<float4*>(bufB);
                for (int i0 = warp * 4; i0 < n; i0 += nwarps * 4) {
                    const int i = i0 + grp;
                    if (i < n) {
                        float4 acc = gather_row32(in4, bm, wpr, n, dup, rp, col_g, base, colsum, i, q,
                                                  tailmask);
                        const float c = cs[i];
                        out4[i * 8 + q] = make_float4(acc.x * c, acc.y * c, acc.z * c, acc.w * c);
                    }
                }
            }
            __syncthreads();
            if (layer == 1) {
                // c1: dW1[c][k] += sum_i dh[i][c] x0[i][k]   (no input gradient needed)
                for (int o = tid; o < kHid * f; o += nthreads) {
                    const int c = o & 31, k = o >> 5;
                    const float* xr = p.x + (int64_t)base * p.ldx + k;
                    float a0 = 0.f, a1 = 0.f;
                    int i = 0;
                    for (; i + 1 < n; i += 2) {
                        a0 = fmaf(bufB[i * kHid + c], xr[(int64_t)i * p.ldx], a0);
                        a1 = fmaf(bufB[(i + 1) * kHid + c], xr[(int64_t)(i + 1) * p.ldx], a1);
                    }
                    if (i < n) a0 = fmaf(bufB[i * kHid + c], xr[(int64_t)i * p.ldx], a0);
                    sacc[G.w1 + c * f + k] += a0 + a1;
                }
            } else {
                const int offx = (layer - 2) * kHid;
                const float* ws = layer == 3 ? w3s : w2s;
                const int ow = layer == 3 ? G.w3 : G.w2;
                // c(i): dW[c][k] += sum_i dh[i][c] x_in[i][k], 4x2 register tiles on 128 threads
                if (warp < dww) {
                    for (int u = tid; u < 128; u += dww * 32) {
                        const int cq = u >> 4, k0 = (u & 15) * 2;
                        float a[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
                        const float4* hb4 = reinterpret_cast<const float4*>(bufB) + cq;
                        const float* xin = xc + offx + k0;
#pragma unroll 4
                        for (int i = 0; i < n; ++i) {
                            const float4 h = hb4[i * 8];
                            const float x0 = xin[(int64_t)i * p.ldc], x1 = xin[(int64_t)i * p.ldc + 1];
                            a[0][0] = fmaf(h.x, x0, a[0][0]); a[0][1] = fmaf(h.x, x1, a[0][1]);
                            a[1][0] = fmaf(h.y, x0, a[1][0]); a[1][1] = fmaf(h.y, x1, a[1][1]);
                            a[2][0] = fmaf(h.z, x0, a[2][0]); a[2][1] = fmaf(h.z, x1, a[2][1]);
                            a[3][0] = fmaf(h.w, x0, a[3][0]); a[3][1] = fmaf(h.w, x1, a[3][1]);
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            sacc[ow + (4 * cq + r) * kHid + k0] += a[r][0];
                            sacc[ow + (4 * cq + r) * kHid + k0 + 1] += a[r][1];
                        }
                    }
                }
                // c(ii): bufA <- dx_in = dh W + pooled gradient of x_in's slice
                if (!split || warp >= dww) {
                    const int pw = split ? warp - dww : warp;
                    const int pn = split ? nwarps - dww : nwarps;
                    const float4* in4 = reinterpret_cast<const float4*>(bufB);
                    const float4* w4m = reinterpret_cast<const float4*>(ws);
                    for (int i0 = pw * 4; i0 < n; i0 += pn * 4) {
                        const int i = i0 + grp;
                        const bool active = i < n;
                        const float4 acc = active ? in4[i * 8 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 init = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (active && rank[i] >= 0) {
                            const float* gp = dp + rank[i] * kCat + offx + 4 * q;
                            init = make_float4(gp[0], gp[1], gp[2], gp[3]);
                        }
                        const float4 y = project32(acc, init, w4m, lane, q);
                        if (active) reinterpret_cast<float4*>(bufA)[i * 8 + q] = y;
                    }
                }
            }
            __syncthreads();
        }
    }

    // one partial vector per CTA
    __syncthreads();
    float* out = p.partials + (int64_t)blockIdx.x * G.total;
    for (int idx = tid; idx < G.total; idx += nthreads) out[idx] = sacc[idx];
}

// grads[o] = sum over CTAs, in CTA order
__global__ void __launch_bounds__(256)
stack_bwd_reduce(const float* __restrict__ partials, int parts, int total, float* __restrict__ grads) {
    DGCNN_PDL_WAIT();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    float s = 0.f;
    for (int b = 0; b < parts; ++b) s += partials[(int64_t)b * total + o];
    grads[o] = s;
}

static int stack_bwd_grid(int f, int nmax, int threads, size_t smem, int64_t num_graphs) {
    (void)f; (void)nmax;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_bwd_kernel, threads, smem) !=
            cudaSuccess || per_sm < 1)
        per_sm = 1;
    int64_t grid = (int64_t)per_sm * DGCNN_NUM_SMS;
    if (grid > num_graphs) grid = num_graphs;
    return (int)(grid < 1 ? 1 : grid);
}

}  // namespace dgcnn

using namespace dgcnn;

// tensor-core variant, graph_stack_bwd_mma.cu
int dgcnn_stack_bwd_mma_supported(int32_t f, int64_t max_nodes, bool conv5 = false);
size_t dgcnn_stack_bwd_mma_workspace_bytes(int32_t f, int64_t num_graphs, int64_t num_nodes);
int dgcnn_stack_bwd_mma(const float* dpooled, const int32_t* perm, int32_t k, const float* xcat,
                        int64_t ldc, const float* x, int64_t ldx, int32_t f, const int32_t* rowptr_t,
                        const int32_t* col_t, const float* dis, const int32_t* gptr,
                        const int32_t* gorder, const int32_t* gdesc, const uint32_t* fragmap,
                        const uint32_t* bitmap, const int32_t* bmoff,
                        const int32_t* gflags, const uint32_t* bitmap_t, const int32_t* bmoff_t,
                        const int32_t* gflags_t, int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                        const float* w2, const float* w3, const float* w4, int32_t norm, float* grads, int32_t* status,
                        void* workspace, cudaStream_t st, const float* dh1 = nullptr,
                        const uint8_t* arg = nullptr, const float* w5 = nullptr);

static int fma_bwd_supported(int32_t num_features, int64_t max_nodes) {
    if (num_features < 1 || num_features > kMaxF || max_nodes < 1 || max_nodes > 1024) return 0;
    const int nmax = stack_nmax_for(max_nodes);
    const StackBwdLayout L = stack_bwd_layout(num_features, nmax, stack_threads_for(nmax) / 32);
    return (size_t)L.total * 4 + 64 <= (size_t)kSmemBudget ? 1 : 0;
}

extern "C" int dgcnn_stack_bwd_supported(int32_t num_features, int64_t max_nodes) {
    // 1: the tensor-core variant fits; 2: only the FMA variant fits (it needs less shared
    // memory per node); 0: neither
    if (dgcnn_stack_bwd_mma_supported(num_features, max_nodes)) return 1;
    return fma_bwd_supported(num_features, max_nodes) ? 2 : 0;
}

extern "C" int64_t dgcnn_stack_num_params(int32_t num_features) {
    return num_features < 1 ? 0 : grad_offsets(num_features).total;
}

extern "C" size_t dgcnn_stack_bwd_workspace_bytes(int32_t num_features, int64_t num_graphs,
                                                  int64_t num_nodes) {
    if (num_features < 1) return 0;
    // FMA variant: one partial gradient vector per CTA (at most 8 CTAs per SM);
    // tensor-core variant: one per graph
    const size_t fma = sizeof(float) * (size_t)grad_offsets(num_features).total * 8 * DGCNN_NUM_SMS + 256;
    const size_t mma = dgcnn_stack_bwd_mma_workspace_bytes(num_features, num_graphs, num_nodes);
    return fma > mma ? fma : mma;
}

extern "C" int dgcnn_stack_bwd(const float* dpooled, const int32_t* perm, int32_t k,
                               const float* xcat, int64_t ldc, const float* x, int64_t ldx,
                               int32_t num_features, const int32_t* rowptr_t, const int32_t* col_t,
                               const float* dis, const int32_t* gptr, const int32_t* gorder,
                               const int32_t* gdesc, const uint32_t* fragmap,
                               const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                               const uint32_t* bitmap_t, const int32_t* bmoff_t,
                               const int32_t* gflags_t,
                               int64_t num_nodes, int64_t num_graphs, int64_t max_nodes, const float* w2,
                               const float* w3, const float* w4, int32_t norm, int32_t variant,
                               float* grads, int32_t* status, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (num_nodes < 0 || num_graphs < 0 || k < 1 || num_features < 1 || ldx < num_features ||
        ldc < kCat || !grads)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const GradOffsets G = grad_offsets(num_features);
    if (num_graphs == 0 || num_nodes == 0) {
        if (cudaMemsetAsync(grads, 0, sizeof(float) * G.total, st) != cudaSuccess) return DGCNN_ERR_CUDA;
        return DGCNN_OK;
    }
    if (variant != DGCNN_STACK_MMA && variant != DGCNN_STACK_FMA) return DGCNN_ERR_INVALID_ARGUMENT;
    if (variant == DGCNN_STACK_MMA ? !dgcnn_stack_bwd_mma_supported(num_features, max_nodes)
                                   : !fma_bwd_supported(num_features, max_nodes))
        return DGCNN_ERR_UNSUPPORTED;
    if (num_graphs >= INT32_MAX || num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (!bitmap || !bmoff || !gflags || !status) return DGCNN_ERR_INVALID_ARGUMENT;
    if (max_nodes > 1024) return DGCNN_ERR_UNSUPPORTED;
    if (!dpooled || !perm || !xcat || !x || !rowptr_t || !dis || !gptr || !w2 || !w3 || !w4)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_stack_bwd_workspace_bytes(num_features, num_graphs, num_nodes))
        return DGCNN_ERR_WORKSPACE;
    if (variant == DGCNN_STACK_MMA)
        return dgcnn_stack_bwd_mma(dpooled, perm, k, xcat, ldc, x, ldx, num_features, rowptr_t, col_t, dis,
                                   gptr, gorder, gdesc, fragmap, bitmap, bmoff, gflags, bitmap_t, bmoff_t, gflags_t,
                                   num_nodes, num_graphs, max_nodes, w2, w3, w4, norm, grads, status, workspace, st);

    StackBwdParams p{};
    p.dpooled = dpooled; p.perm = perm; p.k = k; p.xcat = xcat; p.ldc = ldc;
    p.x = x; p.ldx = ldx; p.f = num_features;
    p.rowptr_t = rowptr_t; p.col_t = col_t; p.dis = dis; p.gptr = gptr; p.gorder = gorder; p.num_graphs = (int)num_graphs;
    p.bitmap = bitmap; p.bmoff = bmoff; p.gflags = gflags;
    p.bitmap_t = bitmap_t; p.bmoff_t = bmoff_t; p.gflags_t = gflags_t;
    p.w2 = w2; p.w3 = w3; p.w4 = w4; p.norm = norm; p.nmax = stack_nmax_for(max_nodes);
    p.partials = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    p.status = status;

    const int threads = stack_threads_for(p.nmax);
    const size_t smem = (size_t)stack_bwd_layout(p.f, p.nmax, threads / 32).total * 4;
    if (cudaFuncSetAttribute(stack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    int grid = stack_bwd_grid(p.f, p.nmax, threads, smem, num_graphs);
    if (grid > 8 * DGCNN_NUM_SMS) grid = 8 * DGCNN_NUM_SMS;
    DGCNN_LAUNCH(stack_bwd_kernel, grid, threads, smem, st, p);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(stack_bwd_reduce, (G.total + 255) / 256, 256, 0, st, p.partials, grid, G.total, grads);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

// ---- the fused backward fed with d(h1) instead of d(pooled) -------------------------
extern "C" int dgcnn_stack_bwd_conv5_supported(int32_t num_features, int64_t max_nodes) {
    return dgcnn_stack_bwd_mma_supported(num_features, max_nodes, true);
}

extern "C" int64_t dgcnn_stack_conv5_num_params(int32_t num_features) {
    return num_features < 1 ? 0 : grad_offsets(num_features).total + 16 * kCat + 16;
}

extern "C" int dgcnn_stack_bwd_conv5(const float* dh1, const uint8_t* arg, const int32_t* perm, int32_t k,
                                     const float* xcat, int64_t ldc, const float* x, int64_t ldx,
                                     int32_t num_features, const int32_t* rowptr_t, const int32_t* col_t,
                                     const float* dis, const int32_t* gptr, const int32_t* gorder,
                                     const int32_t* gdesc, const uint32_t* fragmap,
                                     const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                                     const uint32_t* bitmap_t, const int32_t* bmoff_t,
                                     const int32_t* gflags_t,
                                     int64_t num_nodes, int64_t num_graphs, int64_t max_nodes, const float* w2,
                                     const float* w3, const float* w4, const float* w5, int32_t norm,
                                     float* grads, int32_t* status, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (num_nodes < 0 || num_graphs < 0 || k < 2 || num_features < 1 || ldx < num_features ||
        ldc < kCat || !grads)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (norm != DGCNN_NORM_SYM && norm != DGCNN_NORM_RW) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = dgcnn_stack_conv5_num_params(num_features);
    if (num_graphs == 0) {
        if (cudaMemsetAsync(grads, 0, sizeof(float) * total, st) != cudaSuccess) return DGCNN_ERR_CUDA;
        return DGCNN_OK;
    }
    if (!dgcnn_stack_bwd_mma_supported(num_features, max_nodes, true)) return DGCNN_ERR_UNSUPPORTED;
    if (num_graphs >= INT32_MAX || num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (!bitmap || !bmoff || !gflags || !status || !gdesc) return DGCNN_ERR_INVALID_ARGUMENT;
    if (!dh1 || !arg || !perm || !xcat || (num_nodes > 0 && !x) || !rowptr_t || !dis || !gptr || !w2 || !w3 ||
        !w4 || !w5)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_stack_bwd_workspace_bytes(num_features, num_graphs, num_nodes))
        return DGCNN_ERR_WORKSPACE;
    return dgcnn_stack_bwd_mma(nullptr, perm, k, xcat, ldc, x, ldx, num_features, rowptr_t, col_t, dis, gptr,
                               gorder, gdesc, fragmap, bitmap, bmoff, gflags, bitmap_t, bmoff_t, gflags_t,
                               num_nodes, num_graphs, max_nodes, w2, w3, w4, norm, grads, status, workspace, st,
                               dh1, arg, w5);
}
