// K2 / K4 -- SortPooling forward and backward.
//
// Reference call site: model.py:17,35  self.sort_pool = SortAggregation(k=30);
// x = self.sort_pool(x, batch).  PyG's SortAggregation.forward scatters x into a
// dense [B, Nmax, D] tensor padded with (x.min()-1), sorts every padded row of
// length Nmax on the last channel, gathers ALL B*Nmax rows, slices/pads to k and
// finally overwrites the fill value with 0.  Here one CTA owns one graph: it
// sorts that graph's n_g keys only (64-bit key|index composites, bitonic network
// in shared memory), then copies just the min(n_g,k) winning rows and zero-fills
// the rest.  No dense detour, no host sync, and the permutation is emitted for
// the backward scatter.
//
// Order contract (pinned by the oracle with stable=True):
// descending key, ties by ascending node index, -0.0 == +0.0, NaN first.
#include "common.cuh"
#include "sort_key.cuh"

namespace dgcnn {

// Top-k SELECTION for graphs much larger than k (D&D: 5748 nodes, k = 291): sorting all n keys is
// wasted work -- only the k winners are needed, in order.  Radix select on the 32-bit order keys
// (four 8-bit passes, shared-memory histograms) finds the k-th key T; an ORDERED count settles
// the ties at T by ascending node index (the contract); the k winners are compacted and only THEY
// are sorted (bitonic on <= next_pow2(k) composites).  Returns with the winners' composites sorted
// in buf[0..keep).  Keys are re-read from x (L2 hits) instead of kept: any n fits.
__device__ void select_top_k(const float* __restrict__ x, int64_t ldx, int d, int base, int n, int keep,
                             uint64_t* buf /* >= next_pow2(keep) */, uint32_t* hist /* [256 + 8] */) {
    const int tid = threadIdx.x, nt = blockDim.x;
    auto key_of = [&](int t) { return descending_key_bits(x[(int64_t)(base + t) * ldx + (d - 1)]); };
    uint32_t prefix = 0, want = (uint32_t)keep;            // the `want`-th smallest among keys matching prefix
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < 256; i += nt) hist[i] = 0;
        __syncthreads();
        for (int t = tid; t < n; t += nt) {
            const uint32_t kb = key_of(t);
            if ((kb & himask) == prefix) atomicAdd(&hist[(kb >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {                                    // 256 buckets: a serial scan is ~256 cycles
            uint32_t cum = 0, b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] >= want) break;
                cum += hist[b];
            }
            hist[256] = b;
            hist[257] = want - cum;
        }
        __syncthreads();
        prefix |= hist[256] << shift;
        want = hist[257];
        __syncthreads();
    }
    const uint32_t T = prefix, r_eq = want;                // take every key < T and the first r_eq keys == T
    // ordered count of the keys == T: contiguous chunks per thread, block scan of the chunk counts
    const int chunk = (n + nt - 1) / nt;
    const int t0 = min(n, tid * chunk), t1 = min(n, t0 + chunk);
    uint32_t mine = 0;
    for (int t = t0; t < t1; ++t) mine += key_of(t) == T;
    uint32_t* scan = reinterpret_cast<uint32_t*>(buf);      // [nt] (buf is free until the compaction)
    __syncthreads();
    scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < nt; ++i) { const uint32_t v = scan[i]; scan[i] = run; run += v; }
        hist[258] = 0;                                      // compaction cursor
    }
    __syncthreads();
    uint32_t eq_before = scan[tid];
    __syncthreads();                                        // scan[] (= buf) is rewritten below
    const uint32_t p2 = next_pow2((uint32_t)keep);
    for (uint32_t i = tid + keep; i < p2; i += nt) buf[i] = kPadComposite;
    for (int t = t0; t < t1; ++t) {
        const uint32_t kb = key_of(t);
        bool take = kb < T;
        if (kb == T) { take = eq_before < r_eq; ++eq_before; }
        if (take) buf[atomicAdd(&hist[258], 1u)] = ((uint64_t)kb << 32) | (uint32_t)t;
    }
    bitonic_sort_block(buf, p2);                            // composites are unique: the order is fixed
}

// one CTA per graph (grid-stride). smem_cap = composites that fit the dynamic
// shared buffer; larger graphs sort in the global workspace slice [2*base, 2*base+P)
__global__ void __launch_bounds__(1024)
sp_fwd_kernel(const float* __restrict__ x, int64_t ldx, int d,
                              const int32_t* __restrict__ gptr, int64_t num_graphs, int k,
                              float* __restrict__ out, int32_t* __restrict__ perm,
                              uint64_t* __restrict__ workspace, uint32_t smem_cap) {
    DGCNN_PDL_WAIT();
    extern __shared__ uint64_t sbuf[];
    __shared__ uint32_t hist[264];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int64_t g = blockIdx.x; g < num_graphs; g += gridDim.x) {
        const int base = gptr[g];
        const int n = gptr[g + 1] - base;
        const int keep = min(n, k);
        if (n > 1) {
            const uint32_t p = next_pow2((uint32_t)n);
            // many more nodes than winners: select, then sort the winners only
            const bool select = n >= 4 * keep && n > 1024 && next_pow2((uint32_t)keep) <= smem_cap &&
                                (uint32_t)blockDim.x * 4u <= smem_cap * 8u;
            uint64_t* buf = (select || p <= smem_cap) ? sbuf : workspace + 2 * (int64_t)base;
            __syncthreads();  // previous graph's readers are done with sbuf
            if (select) {
                select_top_k(x, ldx, d, base, n, keep, buf, hist);
            } else {
                for (uint32_t t = threadIdx.x; t < p; t += blockDim.x) {
                    uint64_t c = kPadComposite;
                    if (t < (uint32_t)n) {
                        float key = x[(int64_t)(base + t) * ldx + (d - 1)];
                        c = ((uint64_t)descending_key_bits(key) << 32) | t;
                    }
                    buf[t] = c;
                }
                bitonic_sort_block(buf, p);
            }
            // the winners, one warp per row, eight rows in flight per warp (the copy is L2 / HBM latency)
            constexpr int R = 8;
            for (int r0 = warp * R; r0 < keep; r0 += nwarps * R) {
                int src[R];
#pragma unroll
                for (int u = 0; u < R; ++u) src[u] = base + (int)(uint32_t)(buf[min(r0 + u, keep - 1)] & 0xffffffffu);
                for (int c0 = 0; c0 < d; c0 += 32) {
                    const int c = c0 + lane;
                    float v[R];
#pragma unroll
                    for (int u = 0; u < R; ++u) v[u] = c < d ? x[(int64_t)src[u] * ldx + c] : 0.f;
#pragma unroll
                    for (int u = 0; u < R; ++u)
                        if (c < d && r0 + u < keep) out[((int64_t)g * k + r0 + u) * d + c] = v[u];
                }
                if (lane < R && r0 + lane < keep) perm[g * k + r0 + lane] = src[lane];
            }
        } else if (n == 1) {
            const float* xr = x + (int64_t)base * ldx;
            float* orow = out + (int64_t)g * k * d;
            if (keep == 1) {
                for (int c = threadIdx.x; c < d; c += blockDim.x) orow[c] = xr[c];
                if (threadIdx.x == 0) perm[g * k] = base;
            }
        }
        // zero padding for rows keep..k-1 (contiguous)
        float* pad = out + ((int64_t)g * k + keep) * d;
        const int64_t pad_elems = (int64_t)(k - keep) * d;
        for (int64_t i = threadIdx.x; i < pad_elems; i += blockDim.x) pad[i] = 0.0f;
        for (int r = keep + threadIdx.x; r < k; r += blockDim.x) perm[g * k + r] = -1;
    }
}

__global__ void __launch_bounds__(256)
sp_bwd_zero(float* __restrict__ dx, int64_t lddx, int d, int64_t n) {
    DGCNN_PDL_WAIT();
    int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / d;
        dx[r * lddx + (i - r * d)] = 0.0f;
    }
}

// warp per pooled row: dx[perm] = dout
__global__ void __launch_bounds__(256)
sp_bwd_scatter(const float* __restrict__ dout, const int32_t* __restrict__ perm, int64_t rows, int d,
               float* __restrict__ dx, int64_t lddx, int64_t n) {
    DGCNN_PDL_WAIT();
    const int lane = threadIdx.x & 31;
    int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
         r += warps) {
        int p = perm[r];
        if (p < 0 || p >= n) continue;
        const float* g = dout + r * d;
        float* o = dx + (int64_t)p * lddx;
        for (int c = lane; c < d; c += 32) o[c] = g[c];
    }
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_sort_pool_workspace_bytes(int64_t num_nodes, int64_t num_graphs) {
    (void)num_graphs;
    if (num_nodes < 0) return 0;
    // fallback sort buffer for graphs too large for shared memory: next_pow2(n_g) <= 2 n_g
    return sizeof(uint64_t) * 2 * (size_t)num_nodes + 256;
}

extern "C" int dgcnn_sort_pool_fwd(const float* x, int64_t ldx, int32_t d, const int32_t* gptr,
                                   int64_t num_nodes, int64_t num_graphs, int32_t k,
                                   int64_t max_nodes_hint, float* out,
                                   int32_t* perm, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (num_graphs < 0 || num_nodes < 0 || d < 1 || k < 1 || ldx < d)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs == 0) return DGCNN_OK;
    if (!gptr || !out || !perm || (num_nodes > 0 && !x)) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs * (int64_t)k >= INT32_MAX || num_nodes >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < dgcnn_sort_pool_workspace_bytes(num_nodes, num_graphs))
        return DGCNN_ERR_WORKSPACE;
    uintptr_t aligned = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;

    // shared sort buffer: sized from the hint, 64 KB (8192 nodes) when unknown,
    // never above 128 KB (16384 nodes) so that at least one CTA/SM stays resident
    uint32_t cap = 8192;
    if (max_nodes_hint > 0) {
        cap = next_pow2((uint32_t)(max_nodes_hint > 16384 ? 16384 : max_nodes_hint));
        if (cap < 64) cap = 64;
    }
    size_t smem = sizeof(uint64_t) * cap;
    int threads = cap >= 2048 ? 1024 : (cap >= 512 ? 256 : 128);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(sp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    int grid = (int)(num_graphs < 8 * DGCNN_NUM_SMS ? num_graphs : 8 * DGCNN_NUM_SMS);
    DGCNN_LAUNCH(sp_fwd_kernel, grid, threads, smem, static_cast<cudaStream_t>(stream), 
        x, ldx, d, gptr, num_graphs, k, out, perm, reinterpret_cast<uint64_t*>(aligned), cap);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_sort_pool_bwd(const float* dout, const int32_t* perm, int64_t num_graphs,
                                   int32_t k, int32_t d, float* dx, int64_t lddx, int64_t num_nodes,
                                   void* stream) {
    if (num_graphs < 0 || num_nodes < 0 || d < 1 || k < 1 || lddx < d)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_nodes == 0) return DGCNN_OK;
    if (!dx) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (lddx == d) {
        if (cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)num_nodes * d, st) != cudaSuccess)
            return DGCNN_ERR_CUDA;
    } else {
        DGCNN_LAUNCH(sp_bwd_zero, grid_for(num_nodes * d, 256, 8), 256, 0, st, dx, lddx, d, num_nodes);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    if (num_graphs == 0) return DGCNN_OK;
    if (!dout || !perm) return DGCNN_ERR_INVALID_ARGUMENT;
    int64_t rows = num_graphs * k;
    DGCNN_LAUNCH(sp_bwd_scatter, grid_for(rows, 8, 8), 256, 0, st, dout, perm, rows, d, dx, lddx, num_nodes);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
