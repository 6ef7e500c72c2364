// KT -- the dense tail of the model, model.py:36-43, forward and backward, plus a flat Adam
// (train.py:41).  SURVEY.md 8f rows N2/N3: the callers either side of the graph hot path.
//
//   pooled [B, k*97] --view--> [B,1,k*97]
//   conv5  Conv1d(1,16,97,97) + ReLU      model.py:19,37   (kernel == stride: a per-row 97->16 linear)
//   pool   MaxPool1d(2,2)                 model.py:20,38
//   conv6  Conv1d(16,32,5,1) + ReLU       model.py:19,39
//   fc1    Linear(32*(k/2-4),128) + ReLU  model.py:21,41
//   drop   Dropout(0.5)                   model.py:22,42
//   fc2    Linear(128,C) + log_softmax    model.py:23,43
//
// Stock torch runs this as ~60 small launches (cuDNN implicit-GEMM convolutions, layout
// shuffles, separate bias/ReLU/pool kernels, cuBLAS SIMT GEMMs picked for the wrong shape)
// that cost 5x the graph kernels once those are fused.  Here: 5 forward and 9 backward
// launches of plain fp32 FMA kernels, every reduction in a fixed order (no float atomics).
#include <mutex>

#include "common.cuh"

namespace dgcnn {

constexpr int kC5 = 16, kKW = 97, kC6 = 32, kK6 = 5, kFc = 128;
constexpr int kFc1Splits = 16;     // split-K slabs of the fc1 forward GEMM
constexpr int kDwSplits = 4;       // batch splits of the fc1 weight-gradient GEMM

// ------------------------------------------------------------------------------------------
// conv5 + ReLU + MaxPool(2,2).  A pooled row pair (2j, 2j+1) is 194 contiguous floats; a CTA
// stages 64 pairs in shared memory and every thread owns one pair x 4 channels (both rows,
// so the pooling max stays in registers): 3 shared loads per 8 FMAs.  h1[b][c][j];
// arg[b][c][j] = 0/1 the winning row, 2 when the max is not positive (ReLU dead).
// ------------------------------------------------------------------------------------------
constexpr int kC5Pairs = 64;                       // row pairs staged per CTA iteration
constexpr int kC5Row = 2 * kKW;                    // one pair = 194 contiguous floats
constexpr size_t kC5StageBytes = sizeof(float) * (kC5Pairs * kC5Row + kKW * kC5 + kC5);

// stage `count` row pairs starting at pair index pr0 into xs[pp][194]; warp w copies pairs
// w, w+8, ...; all 7 loads of a lane are issued before the first store
__device__ __forceinline__ void c5_stage_pairs(const float* __restrict__ pooled, int64_t pr0, int count,
                                               int k, int L1, float* __restrict__ xs,
                                               unsigned char* __restrict__ nonzero = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // two pairs (14 loads per lane) in flight per warp: the copy is L2 latency
    for (int pp0 = warp; pp0 < kC5Pairs; pp0 += 16) {
        float v[2][7];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pp = pp0 + 8 * h;
            const bool live = pp < count;
            const int64_t pr = pr0 + pp;
            const int64_t b = live ? pr / L1 : 0;
            const int j = live ? (int)(pr - b * L1) : 0;
            const float* src = pooled + (b * k + 2 * j) * kKW;
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int i = lane + 32 * u;
                v[h][u] = (live && i < kC5Row) ? src[i] : 0.f;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pp = pp0 + 8 * h;
            bool any = false;
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int i = lane + 32 * u;
                if (i < kC5Row) xs[pp * kC5Row + i] = v[h][u];
                any |= v[h][u] != 0.f;                    // NaN != 0: a NaN row is not skipped
            }
            if (nonzero) {
                any = __any_sync(DGCNN_FULL_MASK, any);
                if (lane == 0) nonzero[pp] = any ? 1 : 0;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
tail_c5_fwd(const float* __restrict__ pooled, int64_t B, int k, int L1, const float* __restrict__ w5,
            const float* __restrict__ b5, float* __restrict__ h1, uint8_t* __restrict__ arg) {
    DGCNN_PDL_WAIT();
    extern __shared__ __align__(16) float c5sm[];
    float* xs = c5sm;                                   // [64][194]
    float* w5t = xs + kC5Pairs * kC5Row;                // [97][16]
    float* sb = w5t + kKW * kC5;                        // [16]
    __shared__ unsigned char nonzero[kC5Pairs];         // SortPooling pads with all-zero rows: skip them
    for (int idx = threadIdx.x; idx < kKW * kC5; idx += 256) {
        int c = idx / kKW, i = idx - c * kKW;
        w5t[i * kC5 + c] = w5[idx];
    }
    if (threadIdx.x < kC5) sb[threadIdx.x] = b5[threadIdx.x];
    const int pp = threadIdx.x >> 2, cg = threadIdx.x & 3;      // pair in the stage, channels 4cg..4cg+3
    const int64_t pairs = B * L1;
    for (int64_t pr0 = (int64_t)blockIdx.x * kC5Pairs; pr0 < pairs; pr0 += (int64_t)gridDim.x * kC5Pairs) {
        const int count = (int)min((int64_t)kC5Pairs, pairs - pr0);
        __syncthreads();
        c5_stage_pairs(pooled, pr0, count, k, L1, xs, nonzero);
        __syncthreads();
        const float* x0 = xs + pp * kC5Row;
        const float* x1 = x0 + kKW;
        float a0[4], a1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a0[q] = a1[q] = sb[4 * cg + q];
        // a warp covers 8 consecutive pairs: padding pairs are contiguous, so whole warps skip
        const int iters = __any_sync(DGCNN_FULL_MASK, nonzero[pp] != 0) ? kKW : 0;
#pragma unroll 4
        for (int i = 0; i < iters; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(w5t + i * kC5 + 4 * cg);
            const float u0 = x0[i], u1 = x1[i];
            a0[0] = fmaf(w.x, u0, a0[0]); a0[1] = fmaf(w.y, u0, a0[1]);
            a0[2] = fmaf(w.z, u0, a0[2]); a0[3] = fmaf(w.w, u0, a0[3]);
            a1[0] = fmaf(w.x, u1, a1[0]); a1[1] = fmaf(w.y, u1, a1[1]);
            a1[2] = fmaf(w.z, u1, a1[2]); a1[3] = fmaf(w.w, u1, a1[3]);
        }
        // results go through shared memory (the staged inputs are dead): h1/arg are [b][c][j], so a
        // thread's 4 channels are L1 floats apart -- written directly they are 2 M scattered 4-byte
        // and 1-byte stores; re-ordered, each channel's 64 consecutive j are one coalesced run
        __syncthreads();
        float* so = xs;                                   // [16][64] maxima
        unsigned char* sa = reinterpret_cast<unsigned char*>(xs + kC5 * kC5Pairs);   // [16][64] argmax
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float z0 = fmaxf(a0[q], 0.f), z1 = fmaxf(a1[q], 0.f);
            const float m = fmaxf(z0, z1);
            so[(4 * cg + q) * kC5Pairs + pp] = m;
            sa[(4 * cg + q) * kC5Pairs + pp] = (unsigned char)(m <= 0.f ? 2 : (z0 >= z1 ? 0 : 1));
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC5 * kC5Pairs; idx += 256) {
            const int c = idx / kC5Pairs, q = idx - c * kC5Pairs;
            if (q < count) {
                const int64_t pr = pr0 + q;
                const int64_t b = pr / L1;
                const int j = (int)(pr - b * L1);
                const int64_t o = (b * kC5 + c) * L1 + j;
                h1[o] = so[idx];
                arg[o] = sa[idx];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// conv6 + ReLU: one CTA per graph; lane = output channel, a warp walks output positions.
// h2 is written flattened as torch's x.view(B,-1) sees it: [b][o * L2 + t].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tail_c6_fwd(const float* __restrict__ h1, int64_t B, int L1, const float* __restrict__ w6,
            const float* __restrict__ b6, float* __restrict__ h2) {
    DGCNN_PDL_WAIT();
    extern __shared__ float sm[];
    float* w6t = sm;                         // [(c*5+d)][o]
    float* sb = sm + kC5 * kK6 * kC6;        // [32]
    float* h1s = sb + kC6;                   // [16][L1 + 8] (zero tail: windows may run past L1)
    const int L2 = L1 - (kK6 - 1);
    const int L1P = L1 + 8;
    float* outs = h1s + kC5 * L1P;           // [32][L2] results of this graph
    for (int idx = threadIdx.x; idx < kC6 * kC5 * kK6; idx += 256) {
        int o = idx / (kC5 * kK6), r = idx - o * (kC5 * kK6);
        w6t[r * kC6 + o] = w6[idx];
    }
    if (threadIdx.x < kC6) sb[threadIdx.x] = b6[threadIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC5 * L1P; idx += 256) {
            const int c = idx / L1P, t = idx - c * L1P;
            h1s[idx] = t < L1 ? h1[(b * kC5 + c) * L1 + t] : 0.f;
        }
        __syncthreads();
        // lane = output channel; a warp owns runs of 8 consecutive positions and slides a
        // 5-wide window over each input channel: 17 shared loads per 40 FMAs
        for (int t0 = warp * 8; t0 < L2; t0 += 64) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = sb[lane];
#pragma unroll 2
            for (int c = 0; c < kC5; ++c) {
                float w[kK6];
#pragma unroll
                for (int d = 0; d < kK6; ++d) w[d] = w6t[(c * kK6 + d) * kC6 + lane];
                const float* hr = h1s + c * L1P + t0;
                float win[kK6 + 7];
#pragma unroll
                for (int u = 0; u < kK6 + 7; ++u) win[u] = hr[u];
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int d = 0; d < kK6; ++d) acc[u] = fmaf(w[d], win[u + d], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < L2) outs[lane * L2 + t0 + u] = fmaxf(acc[u], 0.f);
        }
        // the graph's [32][L2] block is contiguous in h2: one coalesced copy instead of stores
        // that are L2 floats apart across the lanes of a warp
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC6 * L2; idx += 256) h2[b * kC6 * L2 + idx] = outs[idx];
    }
}

// ------------------------------------------------------------------------------------------
// fp32-accurate GEMM  C[m][n] = sum_k A(m,k) * Bm(k,n)  on the tensor cores: 3xTF32
// (mma.sync.m16n8k8.tf32; every operand split x = big + small with big = tf32(x), small =
// tf32(x - big); big*big + big*small + small*big accumulated in fp32 -- the dropped
// small*small term is < 2^-22 relative).  64 x 128 tiles, 256 threads = 2 x 4 warps,
// shared-memory tiles As[k][m] / Bs[k][n] with strides = 8 (mod 32): conflict-free fragment
// loads.  The next k-tile's global loads are issued before the current tile's MMAs.
// A_KM: A is stored [k][m] (else [m][k]); B_KN: Bm is stored [k][n] (else [n][k]).  blockIdx.z
// splits K; each split writes its own [M][N] slab of C (the caller adds the slabs in order).
// RELU_MASK: C *= (mask > 0), the ReLU backward.  The FMA version of this kernel spent
// 7-8 M warp instructions per GEMM at 50-60 % issue utilisation on 128 CTAs (profiles/).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tf32_split(float x, uint32_t& big, uint32_t& small) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(big) : "f"(x));
    const float rem = x - __uint_as_float(big);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(small) : "f"(rem));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CTA tile of the GEMM: 64 x 64 (2 x 4 warps of 32 x 16).  Narrow tiles on purpose: these GEMMs
// are small (M = batch, K or N = 128) and latency-bound, so several CTAs per SM matter more than
// operand reuse.
constexpr int kGemmBM = 64, kGemmBN = 64;

template <bool A_KM, bool B_KN, bool RELU_MASK>
__global__ void __launch_bounds__(256)
gemm_f32(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm, int64_t ldb,
         float* __restrict__ C, int M, int N, int K, int kchunk, const float* __restrict__ mask) {
    DGCNN_PDL_WAIT();
    constexpr int BM = kGemmBM, BN = kGemmBN, BK = 16, BMP = BM + 8, BNP = BN + 8;
    constexpr int NI = BN / 32;                            // 8-column MMA tiles per warp (4 warps along N)
    __shared__ float As[BK * BMP];
    __shared__ float Bs[BK * BNP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * (BN / 4);   // warp tile origin inside the CTA tile
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    float acc[2][NI][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.f;
    constexpr int RA = (BM * BK) / 256, RB = (BN * BK) / 256;
    float ra[RA], rb[RB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < RA; ++r) {
            const int idx = threadIdx.x + 256 * r;
            int kk, m;
            if (A_KM) { kk = idx / BM; m = idx - kk * BM; } else { m = idx / BK; kk = idx - m * BK; }
            const int gm = m0 + m, gk = k0 + kk;
            ra[r] = 0.f;
            if (gm < M && gk < kend) ra[r] = A_KM ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int idx = threadIdx.x + 256 * r;
            int kk, n;
            if (B_KN) { kk = idx / BN; n = idx - kk * BN; } else { n = idx / BK; kk = idx - n * BK; }
            const int gn = n0 + n, gk = k0 + kk;
            rb[r] = 0.f;
            if (gn < N && gk < kend) rb[r] = B_KN ? Bm[(int64_t)gk * ldb + gn] : Bm[(int64_t)gn * ldb + gk];
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int r = 0; r < RA; ++r) {
            const int idx = threadIdx.x + 256 * r;
            int kk, m;
            if (A_KM) { kk = idx / BM; m = idx - kk * BM; } else { m = idx / BK; kk = idx - m * BK; }
            As[kk * BMP + m] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int idx = threadIdx.x + 256 * r;
            int kk, n;
            if (B_KN) { kk = idx / BN; n = idx - kk * BN; } else { n = idx / BK; kk = idx - n * BK; }
            Bs[kk * BNP + n] = rb[r];
        }
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            uint32_t abig[2][4], asml[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const float* ap = As + (ks * 8 + t) * BMP + wm + mi * 16 + g;
                tf32_split(ap[0], abig[mi][0], asml[mi][0]);
                tf32_split(ap[8], abig[mi][1], asml[mi][1]);
                tf32_split(ap[4 * BMP], abig[mi][2], asml[mi][2]);
                tf32_split(ap[4 * BMP + 8], abig[mi][3], asml[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const float* bp = Bs + (ks * 8 + t) * BNP + wn + ni * 8 + g;
                uint32_t bb0, bs0, bb1, bs1;
                tf32_split(bp[0], bb0, bs0);
                tf32_split(bp[4 * BNP], bb1, bs1);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_tf32(acc[mi][ni], asml[mi], bb0, bb1);
                    mma_tf32(acc[mi][ni], abig[mi], bs0, bs1);
                    mma_tf32(acc[mi][ni], abig[mi], bb0, bb1);
                }
            }
        }
        __syncthreads();
    }
    float* out = C + (int64_t)blockIdx.z * M * N;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int gm = m0 + wm + mi * 16 + g + 8 * half;
            if (gm >= M) continue;
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int gn = n0 + wn + ni * 8 + 2 * t + q;
                    if (gn >= N) continue;
                    float v = acc[mi][ni][2 * half + q];
                    if (RELU_MASK) v = mask[(int64_t)gm * N + gn] > 0.f ? v : 0.f;
                    out[(int64_t)gm * N + gn] = v;
                }
        }
}

// 32-bit mix (murmur3 finaliser) of (seed, offset, index): one dropout decision per element
__device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t offset, uint32_t idx) {
    uint32_t h = (uint32_t)seed ^ (uint32_t)(seed >> 32) ^ ((uint32_t)offset * 0x9E3779B9u) ^
                 (uint32_t)(offset >> 32);
    h ^= idx * 0x85EBCA6Bu + 0x7F4A7C15u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// fc1 epilogue: sum the split-K slabs in order, + bias, ReLU, Dropout(0.5) when training.
// keep[b][j] = 0 dropped (or ReLU dead), else the multiplier applied (1 or 2) as uint8.
__global__ void __launch_bounds__(256)
tail_fc1_epilogue(const float* __restrict__ slabs, int splits, int64_t total, const float* __restrict__ bias,
                  int training, uint64_t seed, const int64_t* __restrict__ rng_offset,
                  float* __restrict__ h3, uint8_t* __restrict__ keep) {
    DGCNN_PDL_WAIT();
    const uint64_t off = (training && rng_offset) ? (uint64_t)*rng_offset : 0ull;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        float s = bias[i % kFc];
        for (int sp = 0; sp < splits; ++sp) s += slabs[(int64_t)sp * total + i];
        s = fmaxf(s, 0.f);
        uint8_t kp = s > 0.f ? 1 : 0;
        if (training) {
            if (mix32(seed, off, (uint32_t)i) & 0x80000000u) { s *= 2.f; kp = kp ? 2 : 0; }
            else { s = 0.f; kp = 0; }
        }
        h3[i] = s;
        keep[i] = kp;
    }
}

// fc2 + log_softmax: one warp per graph (C <= 32); the last block bumps the dropout offset
__global__ void __launch_bounds__(256)
tail_fc2_lsm_fwd(const float* __restrict__ h3, int64_t B, int C, const float* __restrict__ w2,
                 const float* __restrict__ b2, float* __restrict__ logp, int64_t* rng_offset) {
    DGCNN_PDL_WAIT();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        const float* hr = h3 + b * kFc;
        float mylogit = -INFINITY;
        for (int c = 0; c < C; ++c) {
            float s = 0.f;
            for (int j = lane; j < kFc; j += 32) s = fmaf(hr[j], w2[c * kFc + j], s);
            s = warp_sum(s) + b2[c];
            if (lane == c) mylogit = s;
        }
        float mx = mylogit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(DGCNN_FULL_MASK, mx, o));
        const float e = lane < C ? expf(mylogit - mx) : 0.f;
        const float lse = mx + logf(warp_sum(e));
        if (lane < C) logp[b * C + lane] = mylogit - lse;
    }
    if (rng_offset && blockIdx.x == 0 && threadIdx.x == 0) *rng_offset += 1;
}


// ---- fc1 epilogue + fc2 + log_softmax + NLL + d(logits) + fc2's row backward in ONE launch ---------
// Everything between the fc1 GEMM of the forward and the fc1 GEMMs of the backward is row-local on
// [B,128] / [B,C]: four launches of 4-8 us each (tail_fc1_epilogue, tail_fc2_lsm_fwd, nll_sum_kernel,
// tail_fc2_bwd_rows) for work that fits one warp per graph.  The arithmetic of every value is the
// same as in those kernels (bit-identical h3, keep, logp, dlogit, dz3); the per-graph loss / hit
// scalars are summed in a fixed order by one extra block of tail_fc2_bwd_params (the next kernel of
// the parameter-gradient chain), which also bumps the dropout offset.  grad of the SUM of the NLL.
__global__ void __launch_bounds__(128)
tail_head_kernel(const float* __restrict__ slabs, int splits, int64_t B, int C, const float* __restrict__ bf1,
                 const float* __restrict__ w2, const float* __restrict__ b2, const int64_t* __restrict__ y,
                 int training, uint64_t seed, int64_t* rng_offset, float* __restrict__ h3,
                 uint8_t* __restrict__ keep, float* __restrict__ logp, float* __restrict__ dlogit,
                 float* __restrict__ dz3, float* __restrict__ per_graph /* [B][2] */) {
    DGCNN_PDL_WAIT();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t off = (training && rng_offset) ? (uint64_t)*rng_offset : 0ull;
    const int64_t total = B * kFc;
    for (int64_t b = (int64_t)blockIdx.x * 4 + warp; b < B; b += (int64_t)gridDim.x * 4) {
        // fc1 epilogue: slabs in order + bias, ReLU, Dropout(0.5)
        float hv[kFc / 32];
        uint8_t kp[kFc / 32];
#pragma unroll
        // all slab loads of the row are issued before the first add (same sums, in slab order)
        float part[kFc / 32][kFc1Splits];
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q)
#pragma unroll
            for (int sp = 0; sp < kFc1Splits; ++sp)
                part[q][sp] = sp < splits ? slabs[(int64_t)sp * total + b * kFc + lane + 32 * q] : 0.f;
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q) {
            const int64_t i = b * kFc + lane + 32 * q;
            float sv = bf1[lane + 32 * q];
#pragma unroll
            for (int sp = 0; sp < kFc1Splits; ++sp)
                if (sp < splits) sv += part[q][sp];
            sv = fmaxf(sv, 0.f);
            uint8_t k8 = sv > 0.f ? 1 : 0;
            if (training) {
                if (mix32(seed, off, (uint32_t)i) & 0x80000000u) { sv *= 2.f; k8 = k8 ? 2 : 0; }
                else { sv = 0.f; k8 = 0; }
            }
            hv[q] = sv; kp[q] = k8;
            h3[i] = sv; keep[i] = k8;
        }
        // fc2 + log_softmax (lane c holds class c)
        float mylogit = -INFINITY;
        for (int c = 0; c < C; ++c) {
            float sv = 0.f;
#pragma unroll
            for (int q = 0; q < kFc / 32; ++q) sv = fmaf(hv[q], w2[c * kFc + lane + 32 * q], sv);
            sv = warp_sum(sv) + b2[c];
            if (lane == c) mylogit = sv;
        }
        float mx = mylogit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(DGCNN_FULL_MASK, mx, o));
        const float e = lane < C ? expf(mylogit - mx) : 0.f;
        const float lse = mx + logf(warp_sum(e));
        const float lp = mylogit - lse;
        if (lane < C) logp[b * C + lane] = lp;
        // NLL of this graph, prediction, d(sum of NLL)/d(logp) = -[c == y]
        const int yb = (int)y[b];
        float bv = __shfl_sync(DGCNN_FULL_MASK, lp, 0);
        int best = 0;
        for (int c = 1; c < C; ++c) {
            const float v = __shfl_sync(DGCNN_FULL_MASK, lp, c);
            if (v > bv) { bv = v; best = c; }
        }
        const float ly = (yb >= 0 && yb < C) ? __shfl_sync(DGCNN_FULL_MASK, lp, yb & 31) : 0.f;
        if (lane == 0) {
            per_graph[2 * b] = (yb >= 0 && yb < C) ? -ly : 0.f;
            per_graph[2 * b + 1] = best == yb ? 1.f : 0.f;
        }
        // log_softmax / fc2 backward rows: dlogit = dlogp - softmax * sum(dlogp); dz3 = (dlogit W2) * keep
        const float gd = (lane < C && lane == yb) ? -1.f : 0.f;
        const float sum = warp_sum(gd);
        const float dl = lane < C ? gd - expf(lp) * sum : 0.f;
        if (lane < C) dlogit[b * C + lane] = dl;
        float acc[kFc / 32];
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q) acc[q] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float dc = __shfl_sync(DGCNN_FULL_MASK, dl, c);
#pragma unroll
            for (int q = 0; q < kFc / 32; ++q) acc[q] = fmaf(dc, w2[c * kFc + lane + 32 * q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q) dz3[b * kFc + lane + 32 * q] = acc[q] * (float)kp[q];
    }
}

// ---- backward -----------------------------------------------------------------------------
// fc2 / log_softmax backward, part A (one warp per graph):
//   dlogit = dlogp - softmax * sum(dlogp);   dz3 = (dlogit W2) * keep
__global__ void __launch_bounds__(256)
tail_fc2_bwd_rows(const float* __restrict__ dlogp, const float* __restrict__ logp,
                  const uint8_t* __restrict__ keep, int64_t B, int C, const float* __restrict__ w2,
                  float* __restrict__ dlogit, float* __restrict__ dz3) {
    DGCNN_PDL_WAIT();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        const float g = lane < C ? dlogp[b * C + lane] : 0.f;
        const float sum = warp_sum(g);
        const float dl = lane < C ? g - expf(logp[b * C + lane]) * sum : 0.f;
        if (lane < C) dlogit[b * C + lane] = dl;
        float acc[kFc / 32];
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q) acc[q] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float dc = __shfl_sync(DGCNN_FULL_MASK, dl, c);
#pragma unroll
            for (int q = 0; q < kFc / 32; ++q) acc[q] = fmaf(dc, w2[c * kFc + lane + 32 * q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < kFc / 32; ++q) {
            const int64_t i = b * kFc + lane + 32 * q;
            dz3[i] = acc[q] * (float)keep[i];
        }
    }
}

// part B: dW2[c][j] = sum_b dlogit[b][c] h3[b][j], db2[c] = sum_b dlogit, dbf1[j] = sum_b dz3.
// Block = 32 outputs x 8 batch lanes, batch lanes combined in order (deterministic).
__global__ void __launch_bounds__(256)
tail_fc2_bwd_params(const float* __restrict__ dlogit, const float* __restrict__ h3,
                    const float* __restrict__ dz3, int64_t B, int C, float* __restrict__ dw2,
                    float* __restrict__ db2, float* __restrict__ dbf1,
                    const float* __restrict__ per_graph = nullptr, float* __restrict__ stats = nullptr,
                    int64_t* rng_offset = nullptr) {
    DGCNN_PDL_WAIT();
    __shared__ float red[8][33];
    const int ox = threadIdx.x & 31, gy = threadIdx.x >> 5;
    if (per_graph && blockIdx.x == gridDim.x - 1) {
        // (launched with one block more than the gradients need) loss sum / #correct of the step from
        // tail_head_kernel's per-graph scalars, in a fixed order; the dropout stream moves on
        float l = 0.f, cnt = 0.f;
        for (int64_t b = threadIdx.x; b < B; b += 256) { l += per_graph[2 * b]; cnt += per_graph[2 * b + 1]; }
        l = warp_sum(l); cnt = warp_sum(cnt);
        if (ox == 0) { red[0][gy] = l; red[1][gy] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, c2 = 0.f;
            for (int w = 0; w < 8; ++w) { a += red[0][w]; c2 += red[1][w]; }
            stats[0] = a;
            stats[1] = c2;
            if (rng_offset) *rng_offset += 1;
        }
        return;
    }
    const int o = blockIdx.x * 32 + ox;
    const int total = C * kFc + C + kFc;
    float s = 0.f;
    if (o < C * kFc) {
        const int c = o / kFc, j = o - c * kFc;
        int64_t b = gy;
        for (; b + 56 < B; b += 64) {                       // eight loads in flight, added in batch order
            float dv[8], hv8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { dv[u] = dlogit[(b + 8 * u) * C + c]; hv8[u] = h3[(b + 8 * u) * kFc + j]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) s = fmaf(dv[u], hv8[u], s);
        }
        for (; b < B; b += 8) s = fmaf(dlogit[b * C + c], h3[b * kFc + j], s);
    } else if (o < C * kFc + C) {
        const int c = o - C * kFc;
        for (int64_t b = gy; b < B; b += 8) s += dlogit[b * C + c];
    } else if (o < total) {
        const int j = o - C * kFc - C;
        for (int64_t b = gy; b < B; b += 8) s += dz3[b * kFc + j];
    }
    red[gy][ox] = s;
    __syncthreads();
    if (gy == 0 && o < total) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) t += red[y][ox];
        if (o < C * kFc) dw2[o] = t;
        else if (o < C * kFc + C) db2[o - C * kFc] = t;
        else dbf1[o - C * kFc - C] = t;
    }
}

// conv6 backward w.r.t. its input: dh1[b][c][s] = sum_{o,d} dz2[b][o][s-d] W6[o][c][d].
// One CTA per graph; thread = (input channel c, run of 5 positions), sliding window over dz2.
__global__ void __launch_bounds__(256)
tail_c6_bwd_input(const float* __restrict__ dz2, int64_t B, int L1, const float* __restrict__ w6,
                  float* __restrict__ dh1) {
    DGCNN_PDL_WAIT_ONLY();
    extern __shared__ float sm[];
    float* w6c = sm;                         // [(o*5+d)][c]
    const int L2 = L1 - (kK6 - 1);
    const int LZ = L2 + 2 * (kK6 - 1) + 8;   // dz row with 4 zeros in front, zeros behind
    float* dzs = sm + kC6 * kK6 * kC5;       // [32][LZ]
    float* outs = dzs + kC6 * LZ;            // [16][L1] results of this graph
    for (int idx = threadIdx.x; idx < kC6 * kC5 * kK6; idx += 256) {
        int o = idx / (kC5 * kK6), r = idx - o * (kC5 * kK6);
        int c = r / kK6, d = r - c * kK6;
        w6c[(o * kK6 + d) * kC5 + c] = w6[idx];
    }
    const int c = threadIdx.x & 15, run = threadIdx.x >> 4;     // 16 runs
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC6 * LZ; idx += 256) {
            const int o = idx / LZ, t = idx - o * LZ - (kK6 - 1);
            dzs[idx] = (t >= 0 && t < L2) ? dz2[(b * kC6 + o) * L2 + t] : 0.f;
        }
        __syncthreads();
        for (int s0 = run * 5; s0 < L1; s0 += 80) {
            float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int o = 0; o < kC6; ++o) {
                float w[kK6];
#pragma unroll
                for (int d = 0; d < kK6; ++d) w[d] = w6c[(o * kK6 + d) * kC5 + c];
                // dz[o][s - d] for s = s0..s0+4, d = 0..4  ->  padded index s - d + 4
                const float* dr = dzs + o * LZ + s0;
                float win[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) win[u] = dr[u];
#pragma unroll
                for (int u = 0; u < 5; ++u)
#pragma unroll
                    for (int d = 0; d < kK6; ++d) acc[u] = fmaf(win[u + 4 - d], w[d], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 5; ++u)
                if (s0 + u < L1) outs[c * L1 + s0 + u] = acc[u];
        }
        // the graph's [16][L1] block is contiguous in dh1: one coalesced copy
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC5 * L1; idx += 256) dh1[b * kC5 * L1 + idx] = outs[idx];
    }
}

// conv6 weight/bias gradient: persistent CTAs; thread = (pair of output channels, input
// channel) owns all 5 taps (10 outputs) and slides a 5-wide window over h1 while it walks
// dz2: 3 shared loads per 10 FMAs.  One partial vector [2560 + 32] per CTA.
__global__ void __launch_bounds__(256)
tail_c6_bwd_weight(const float* __restrict__ dz2, const float* __restrict__ h1, int64_t B, int L1,
                   float* __restrict__ partials) {
    DGCNN_PDL_WAIT();
    extern __shared__ float sm[];
    const int L2 = L1 - (kK6 - 1);
    const int L1P = L1 | 1;          // odd row stride: the 16 channel rows hit distinct banks
    float* dzs = sm;                 // [32][L2]
    float* h1s = sm + kC6 * L2;      // [16][L1P]
    constexpr int NW = kC6 * kC5 * kK6;   // 2560
    const int c = threadIdx.x & 15, op = threadIdx.x >> 4;     // output channels 2op, 2op+1
    float a0[kK6] = {0.f, 0.f, 0.f, 0.f, 0.f}, a1[kK6] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float sb0 = 0.f, sb1 = 0.f;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kC6 * L2; idx += 256) dzs[idx] = dz2[b * kC6 * L2 + idx];
        for (int idx = threadIdx.x; idx < kC5 * L1; idx += 256) {
            const int cc = idx / L1, t = idx - cc * L1;
            h1s[cc * L1P + t] = h1[b * kC5 * L1 + idx];
        }
        __syncthreads();
        const float* d0 = dzs + (2 * op) * L2;
        const float* d1 = d0 + L2;
        const float* hr = h1s + c * L1P;
        float w0 = hr[0], w1 = hr[1], w2 = hr[2], w3 = hr[3];
        for (int t = 0; t < L2; ++t) {
            const float w4 = hr[t + 4];
            const float z0 = d0[t], z1 = d1[t];
            a0[0] = fmaf(z0, w0, a0[0]); a0[1] = fmaf(z0, w1, a0[1]); a0[2] = fmaf(z0, w2, a0[2]);
            a0[3] = fmaf(z0, w3, a0[3]); a0[4] = fmaf(z0, w4, a0[4]);
            a1[0] = fmaf(z1, w0, a1[0]); a1[1] = fmaf(z1, w1, a1[1]); a1[2] = fmaf(z1, w2, a1[2]);
            a1[3] = fmaf(z1, w3, a1[3]); a1[4] = fmaf(z1, w4, a1[4]);
            if (c == 0) { sb0 += z0; sb1 += z1; }
            w0 = w1; w1 = w2; w2 = w3; w3 = w4;
        }
    }
    float* out = partials + (int64_t)blockIdx.x * (NW + kC6);
#pragma unroll
    for (int d = 0; d < kK6; ++d) {
        out[((2 * op) * kC5 + c) * kK6 + d] = a0[d];
        out[((2 * op + 1) * kC5 + c) * kK6 + d] = a1[d];
    }
    if (c == 0) { out[NW + 2 * op] = sb0; out[NW + 2 * op + 1] = sb1; }
}

// conv5 / pool / ReLU backward w.r.t. pooled: one warp per row pair
__global__ void __launch_bounds__(256)
tail_c5_bwd_input(const float* __restrict__ dh1, const uint8_t* __restrict__ arg, int64_t B, int k, int L1,
                  const float* __restrict__ w5, float* __restrict__ dpooled) {
    DGCNN_PDL_WAIT_ONLY();
    __shared__ float w5s[kC5 * kKW];
    constexpr int TP = 32;                              // pairs per CTA iteration (4 per warp)
    __shared__ float zv[TP][kC5 + 1];
    __shared__ unsigned char za[TP][kC5];
    for (int idx = threadIdx.x; idx < kC5 * kKW; idx += 256) w5s[idx] = w5[idx];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pairs = B * L1;
    for (int64_t pr0 = (int64_t)blockIdx.x * TP; pr0 < pairs; pr0 += (int64_t)gridDim.x * TP) {
        __syncthreads();
        // tile [16 channels][32 consecutive pairs] of dh1 / arg: 32 consecutive j of one channel
        // are contiguous (within a graph), so these loads are coalesced -- one pair at a time they
        // were 16 scattered floats + 16 scattered bytes per pair
        for (int idx = threadIdx.x; idx < kC5 * TP; idx += 256) {
            const int c = idx / TP, q = idx - c * TP;
            const int64_t pr = pr0 + q;
            float v = 0.f;
            unsigned char a = 2;
            if (pr < pairs) {
                const int64_t b = pr / L1;
                const int j = (int)(pr - b * L1);
                const int64_t o = (b * kC5 + c) * L1 + j;
                v = dh1[o];
                a = arg[o];
            }
            zv[q][c] = v;
            za[q][c] = a;
        }
        __syncthreads();
        for (int q = warp; q < TP; q += 8) {
            const int64_t pr = pr0 + q;
            if (pr >= pairs) break;
            const int64_t b = pr / L1;
            const int j = (int)(pr - b * L1);
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < kC5; ++c) {
                const int r = za[q][c];
                if (r > 1) continue;                       // ReLU dead
                const float v = zv[q][c];
                const float* wr = w5s + c * kKW;
                if (r == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) { int i = lane + 32 * u; if (i < kKW) a0[u] = fmaf(v, wr[i], a0[u]); }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) { int i = lane + 32 * u; if (i < kKW) a1[u] = fmaf(v, wr[i], a1[u]); }
                }
            }
            float* dst = dpooled + (b * k + 2 * j) * kKW;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = lane + 32 * u;
                if (i < kKW) { dst[i] = a0[u]; dst[kKW + i] = a1[u]; }
            }
        }
    }
    // an odd k leaves the last row outside every pooling window: zero gradient
    if ((k & 1) && blockIdx.x == 0)
        for (int64_t idx = threadIdx.x; idx < B * kKW; idx += 256) {
            const int64_t b = idx / kKW;
            dpooled[(b * k + (k - 1)) * kKW + (idx - b * kKW)] = 0.f;
        }
}

// conv5 weight/bias gradient: persistent CTAs stage 64 row pairs at a time (coalesced, all
// loads in flight) and every thread owns 7 of the 16 x 97 outputs across the whole sweep:
// thread = (channel c, column group ig), columns ig + 16 q.  Only the row that won the
// pooling (arg) contributes.  One partial vector [1552 + 16] per CTA.
__global__ void __launch_bounds__(256)
tail_c5_bwd_weight(const float* __restrict__ dh1, const uint8_t* __restrict__ arg,
                   const float* __restrict__ pooled, int64_t B, int k, int L1,
                   float* __restrict__ partials) {
    DGCNN_PDL_WAIT();
    constexpr int NW = kC5 * kKW;                           // 1552
    extern __shared__ __align__(16) float c5sm[];
    float* xs = c5sm;                                       // [64][194]
    float* zv = xs + kC5Pairs * kC5Row;                     // [64][16]
    int* win = reinterpret_cast<int*>(zv + kC5Pairs * kC5); // [64][16]
    const int c = threadIdx.x & 15, ig = threadIdx.x >> 4;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float accb = 0.f;
    const int64_t pairs = B * L1;
    for (int64_t pr0 = (int64_t)blockIdx.x * kC5Pairs; pr0 < pairs; pr0 += (int64_t)gridDim.x * kC5Pairs) {
        const int count = (int)min((int64_t)kC5Pairs, pairs - pr0);
        __syncthreads();
        c5_stage_pairs(pooled, pr0, count, k, L1, xs);
        for (int idx = threadIdx.x; idx < kC5Pairs * kC5; idx += 256) {
            const int pp = idx >> 4, cc = idx & 15;
            float v = 0.f;
            int a = 0;
            if (pp < count) {
                const int64_t pr = pr0 + pp;
                const int64_t b = pr / L1;
                const int j = (int)(pr - b * L1);
                const int64_t o = (b * kC5 + cc) * L1 + j;
                a = arg[o];
                v = a < 2 ? dh1[o] : 0.f;
                a = a < 2 ? a : 0;
            }
            zv[idx] = v;
            win[idx] = a;
        }
        __syncthreads();
        for (int pp = 0; pp < count; ++pp) {
            const float v = zv[pp * kC5 + c];
            const float* xr = xs + pp * kC5Row + win[pp * kC5 + c] * kKW + ig;
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[q] = fmaf(v, xr[16 * q], acc[q]);
            if (ig == 0) { acc[6] = fmaf(v, xr[96], acc[6]); accb += v; }
        }
    }
    float* out = partials + (int64_t)blockIdx.x * (NW + kC5);
#pragma unroll
    for (int q = 0; q < 6; ++q) out[c * kKW + ig + 16 * q] = acc[q];
    if (ig == 0) { out[c * kKW + 96] = acc[6]; out[NW + c] = accb; }
}

// out[o] = sum_p partials[p][o] (fixed partition and order: deterministic); block = 32 outputs
// x 8 partial lanes; optionally split over two destinations
__global__ void __launch_bounds__(256)
tail_reduce_partials(const float* __restrict__ partials, int parts, int total, int split_at,
                     float* __restrict__ out_a, float* __restrict__ out_b) {
    DGCNN_PDL_WAIT();
    __shared__ float red[8][33];
    const int ox = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + ox;
    float s = 0.f;
    if (o < total)
        for (int p = gy; p < parts; p += 8) s += partials[(int64_t)p * total + o];
    red[gy][ox] = s;
    __syncthreads();
    if (gy == 0 && o < total) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) t += red[y][ox];
        if (o < split_at) out_a[o] = t; else out_b[o - split_at] = t;
    }
}

// Adam on flat buffers (torch.optim.Adam defaults semantics, train.py:99): the step counter
// lives on the device so that the whole step can be replayed from a CUDA graph.
__global__ void __launch_bounds__(256)
adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
          int64_t n, int64_t* step, float lr, float beta1, float beta2, float eps, float grad_scale) {
    DGCNN_PDL_WAIT();
    const int64_t t = *step + 1;
    const float bc1 = 1.f - powf(beta1, (float)t), bc2 = 1.f - powf(beta2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}
__global__ void adam_bump(int64_t* step) {
    DGCNN_PDL_WAIT(); *step += 1; }

// NLL (train.py:39,44-45) on log-probabilities, single CTA: stats[0] = -sum_b logp[b][y_b],
// stats[1] = #{argmax == y};  dlogp[b][c] = -grad_scale at c == y_b, else 0.
__global__ void __launch_bounds__(1024)
nll_sum_kernel(const float* __restrict__ logp, const int64_t* __restrict__ y, int64_t B, int C,
               float grad_scale, float* __restrict__ stats, float* __restrict__ dlogp) {
    DGCNN_PDL_WAIT();
    __shared__ float sl[32], sc[32];
    float loss = 0.f, correct = 0.f;
    for (int64_t b = threadIdx.x; b < B; b += 1024) {
        const int yb = (int)y[b];
        int best = 0;
        float bv = logp[b * C];
        for (int c = 0; c < C; ++c) {
            const float v = logp[b * C + c];
            if (v > bv) { bv = v; best = c; }
            if (dlogp) dlogp[b * C + c] = c == yb ? -grad_scale : 0.f;
        }
        if (yb >= 0 && yb < C) loss -= logp[b * C + yb];
        correct += best == yb ? 1.f : 0.f;
    }
    loss = warp_sum(loss);
    correct = warp_sum(correct);
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = loss; sc[threadIdx.x >> 5] = correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c2 = 0.f;
        for (int w = 0; w < 32; ++w) { a += sl[w]; c2 += sc[w]; }
        stats[0] = a;
        stats[1] = c2;
    }
}

struct TailDims { int L1, L2, D1; };
__host__ inline TailDims tail_dims(int k) {
    TailDims d;
    d.L1 = k / 2;
    d.L2 = d.L1 - (kK6 - 1);
    d.D1 = kC6 * d.L2;
    return d;
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_tail_workspace_bytes(int64_t num_graphs, int32_t k, int32_t num_classes) {
    if (num_graphs < 0 || k < 10 || num_classes < 1) return 0;
    const TailDims d = tail_dims(k);
    size_t fwd = sizeof(float) * (size_t)kFc1Splits * num_graphs * kFc;
    size_t bwd = sizeof(float) * ((size_t)64 + 2 * (size_t)num_graphs          // ticket (256 B) + per-graph loss / hit
                                  + (size_t)num_graphs * num_classes            // dlogit
                                  + (size_t)num_graphs * kFc                   // dz3
                                  + (size_t)num_graphs * d.D1                  // dz2
                                  + (size_t)num_graphs * kC5 * d.L1            // dh1
                                  + (size_t)num_graphs * (kC6 * kC5 * kK6 + kC6)   // conv6 partials
                                  + (size_t)4 * DGCNN_NUM_SMS * (kC5 * kKW + kC5)   // conv5 partials
                                  + (size_t)kDwSplits * kFc * d.D1);                // dWf1 slabs
    return (fwd > bwd ? fwd : bwd) + 1024;
}

extern "C" int dgcnn_tail_fwd(const float* pooled, int64_t num_graphs, int32_t k, const float* w5,
                              const float* b5, const float* w6, const float* b6, const float* wf1,
                              const float* bf1, const float* wf2, const float* bf2, int32_t num_classes,
                              int32_t training, uint64_t seed, int64_t* rng_offset, float* h1,
                              uint8_t* arg, float* h2, float* h3, uint8_t* keep, float* logp,
                              void* workspace, size_t workspace_bytes, void* stream) {
    const int64_t B = num_graphs;
    if (B < 0 || k < 10 || num_classes < 1) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_classes > 32) return DGCNN_ERR_UNSUPPORTED;
    if (B == 0) return DGCNN_OK;
    // pooled == NULL: h1 / arg were already produced by dgcnn_stack_fwd_conv5
    if ((pooled && (!w5 || !b5)) || !w6 || !b6 || !wf1 || !bf1 || !wf2 || !bf2 || !h1 || !arg || !h2 ||
        !h3 || !keep || !logp)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_tail_workspace_bytes(B, k, num_classes))
        return DGCNN_ERR_WORKSPACE;
    const TailDims d = tail_dims(k);
    if (B * (int64_t)d.D1 >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* slabs = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);

    if (pooled) {
        if (cudaFuncSetAttribute(tail_c5_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kC5StageBytes) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        DGCNN_LAUNCH(tail_c5_fwd, grid_for(B * d.L1, kC5Pairs, 4), 256, kC5StageBytes, st, pooled, B, k, d.L1, w5, b5, h1,
                                                                                 arg);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    const size_t smem6 = sizeof(float) * (kC5 * kK6 * kC6 + kC6 + kC5 * (d.L1 + 8) + kC6 * d.L2);
    if (smem6 > 96 * 1024) return DGCNN_ERR_UNSUPPORTED;
    if (smem6 > 48 * 1024 &&
        cudaFuncSetAttribute(tail_c6_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    DGCNN_LAUNCH(tail_c6_fwd, grid_for(B, 1, 4), 256, smem6, st, h1, B, d.L1, w6, b6, h2);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    // fc1: [B, D1] x Wf1^T [D1, 128], split-K slabs
    const int kchunk = (int)ceil_div(ceil_div(d.D1, kFc1Splits), 16) * 16;
    const int splits = (int)ceil_div(d.D1, kchunk);
    dim3 g1((unsigned)ceil_div(kFc, kGemmBN), (unsigned)ceil_div(B, kGemmBM), (unsigned)splits);
    DGCNN_LAUNCH((gemm_f32<false, false, false>), g1, 256, 0, st, h2, d.D1, wf1, d.D1, slabs, (int)B, kFc, d.D1, kchunk,
                                                      nullptr);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(tail_fc1_epilogue, grid_for(B * kFc, 256, 4), 256, 0, st, slabs, splits, B * kFc, bf1, training, seed,
                                                                 rng_offset, h3, keep);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(tail_fc2_lsm_fwd, grid_for(B, 8, 2), 256, 0, st, h3, B, num_classes, wf2, bf2, logp,
                                                        training ? rng_offset : nullptr);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

// dgcnn_tail_fwd + dgcnn_nll_sum + the first kernel of the backward in one call for a TRAINING step:
// conv5 (unless pooled == NULL: h1 / arg are inputs) -> conv6 -> fc1 GEMM -> tail_head_kernel.
// Leaves logp, h3, keep as dgcnn_tail_fwd does, and dlogit / dz3 / the per-graph loss and hit scalars in
// `workspace_bwd` where dgcnn_tail_bwd_after_loss picks them up (it also writes stats = [sum of NLL,
// #correct] and advances the dropout stream).  The gradient is that of the SUM of the NLL.
extern "C" int dgcnn_tail_fwd_loss(const float* pooled, int64_t num_graphs, int32_t k, const float* w5,
                                   const float* b5, const float* w6, const float* b6, const float* wf1,
                                   const float* bf1, const float* wf2, const float* bf2, int32_t num_classes,
                                   const int64_t* y, int32_t training, uint64_t seed, int64_t* rng_offset,
                                   float* h1, uint8_t* arg, float* h2, float* h3, uint8_t* keep, float* logp,
                                   void* workspace_fwd, size_t workspace_fwd_bytes, void* workspace_bwd,
                                   size_t workspace_bwd_bytes, void* stream) {
    const int64_t B = num_graphs;
    if (B < 0 || k < 10 || num_classes < 1) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_classes > 32) return DGCNN_ERR_UNSUPPORTED;
    if (B == 0) return DGCNN_OK;
    if ((pooled && (!w5 || !b5)) || !w6 || !b6 || !wf1 || !bf1 || !wf2 || !bf2 || !h1 || !arg || !h2 ||
        !h3 || !keep || !logp || !y)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (training && !rng_offset) return DGCNN_ERR_INVALID_ARGUMENT;
    const size_t need = dgcnn_tail_workspace_bytes(B, k, num_classes);
    if (!workspace_fwd || workspace_fwd_bytes < need || !workspace_bwd || workspace_bwd_bytes < need)
        return DGCNN_ERR_WORKSPACE;
    const TailDims d = tail_dims(k);
    if (B * (int64_t)d.D1 >= INT32_MAX) return DGCNN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* slabs = reinterpret_cast<float*>(((uintptr_t)workspace_fwd + 255) & ~(uintptr_t)255);
    float* wb = reinterpret_cast<float*>(((uintptr_t)workspace_bwd + 255) & ~(uintptr_t)255);
    float* per_graph = wb + 64;
    float* dlogit = wb + 64 + 2 * B;
    float* dz3 = dlogit + B * num_classes;
    if (pooled) {
        if (cudaFuncSetAttribute(tail_c5_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kC5StageBytes) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        DGCNN_LAUNCH(tail_c5_fwd, grid_for(B * d.L1, kC5Pairs, 4), 256, kC5StageBytes, st, pooled, B, k, d.L1, w5, b5, h1,
                                                                                 arg);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    const size_t smem6 = sizeof(float) * (kC5 * kK6 * kC6 + kC6 + kC5 * (d.L1 + 8) + kC6 * d.L2);
    if (smem6 > 96 * 1024) return DGCNN_ERR_UNSUPPORTED;
    if (smem6 > 48 * 1024 &&
        cudaFuncSetAttribute(tail_c6_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    DGCNN_LAUNCH(tail_c6_fwd, grid_for(B, 1, 4), 256, smem6, st, h1, B, d.L1, w6, b6, h2);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    const int kchunk = (int)ceil_div(ceil_div(d.D1, kFc1Splits), 16) * 16;
    const int splits = (int)ceil_div(d.D1, kchunk);
    dim3 g1((unsigned)ceil_div(kFc, kGemmBN), (unsigned)ceil_div(B, kGemmBM), (unsigned)splits);
    DGCNN_LAUNCH((gemm_f32<false, false, false>), g1, 256, 0, st, h2, d.D1, wf1, d.D1, slabs, (int)B, kFc, d.D1, kchunk,
                                                      nullptr);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(tail_head_kernel, grid_for(B, 4, 4), 128, 0, st, slabs, splits, B, num_classes, bf1, wf2, bf2, y, training,
                                                        seed, rng_offset, h3, keep, logp, dlogit, dz3, per_graph);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

// Side stream of the parameter-gradient chain of dgcnn_tail_bwd (overlap != 0): one per
// device, created on first use and kept for the life of the process.
struct SideStream { cudaStream_t stream; cudaEvent_t ev[3]; cudaEvent_t done; bool ready; };
static SideStream g_side[64];
static std::mutex g_side_mutex;

static SideStream* side_stream() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream& s = g_side[dev];
    if (!s.ready) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 3; ++i)
            if (cudaEventCreateWithFlags(&s.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ready = true;
    }
    return &s;
}

// dh1_ext != NULL: stop at d(h1) (written there); conv5's backward -- dpooled, dw5, db5 -- is then
// done by the caller (dgcnn_stack_bwd_conv5) and pooled / dpooled / dw5 / db5 are unused.
static int tail_bwd_impl(const float* dlogp, const float* pooled, int64_t num_graphs, int32_t k,
                         const float* w5, const float* w6, const float* wf1, const float* wf2,
                         int32_t num_classes, const float* h1, const uint8_t* arg, const float* h2,
                         const float* h3, const uint8_t* keep, const float* logp, float* dpooled,
                         float* dw5, float* db5, float* dw6, float* db6, float* dwf1, float* dbf1,
                         float* dwf2, float* dbf2, int32_t overlap, void* workspace,
                         size_t workspace_bytes, void* stream, bool to_h1, float* dh1_ext,
                         float* stats_after_loss = nullptr, int64_t* rng_offset = nullptr) {
    // stats_after_loss != NULL: dgcnn_tail_fwd_loss already left dlogit / dz3 / the per-graph scalars in
    // THIS workspace; fc2's row backward is skipped and the scalars are summed into stats here
    const int64_t B = num_graphs;
    if (B < 0 || k < 10 || num_classes < 1 || overlap < 0 || overlap > 2) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_classes > 32) return DGCNN_ERR_UNSUPPORTED;
    if ((!to_h1 && (!dw5 || !db5)) || !dw6 || !db6 || !dwf1 || !dbf1 || !dwf2 || !dbf2)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < dgcnn_tail_workspace_bytes(B, k, num_classes))
        return DGCNN_ERR_WORKSPACE;
    const TailDims d = tail_dims(k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0) {
        if (stats_after_loss) cudaMemsetAsync(stats_after_loss, 0, 2 * sizeof(float), st);
        if (!to_h1) { cudaMemsetAsync(dw5, 0, sizeof(float) * kC5 * kKW, st);  cudaMemsetAsync(db5, 0, sizeof(float) * kC5, st); }
        cudaMemsetAsync(dw6, 0, sizeof(float) * kC6 * kC5 * kK6, st);  cudaMemsetAsync(db6, 0, sizeof(float) * kC6, st);
        cudaMemsetAsync(dwf1, 0, sizeof(float) * kFc * d.D1, st);  cudaMemsetAsync(dbf1, 0, sizeof(float) * kFc, st);
        cudaMemsetAsync(dwf2, 0, sizeof(float) * num_classes * kFc, st);
        cudaMemsetAsync(dbf2, 0, sizeof(float) * num_classes, st);
        return DGCNN_OK;
    }
    const bool after_loss = stats_after_loss != nullptr;
    if ((!dlogp && !after_loss) || (!to_h1 && (!pooled || !w5 || !dpooled || !arg)) || !w6 || !wf1 || !wf2 || !h1 ||
        !h2 || !h3 || !keep || !logp)
        return DGCNN_ERR_INVALID_ARGUMENT;
    float* ws = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* per_graph = ws + 64;
    ws += 64 + 2 * B;                           // (per-graph loss / hit scalars of dgcnn_tail_fwd_loss)
    float* dlogit = ws;                         ws += B * num_classes;
    float* dz3 = ws;                            ws += B * kFc;
    float* dz2 = ws;                            ws += B * (int64_t)d.D1;
    float* dh1 = to_h1 ? dh1_ext : ws;          ws += B * (int64_t)kC5 * d.L1;
    float* part6 = ws;                          ws += B * (kC6 * kC5 * kK6 + kC6);
    float* part5 = ws;                          ws += 4 * DGCNN_NUM_SMS * (kC5 * kKW + kC5);
    float* slabw = ws;

    // Two dependency chains: the INPUT gradients (fc2 rows -> fc1 dx -> conv6 dx -> conv5 dx ->
    // dpooled), which the graph backward is waiting for, and the PARAMETER gradients, which only
    // the optimizer needs.  With overlap != 0 the second chain runs on the library's side
    // stream, forked/joined with events (capturable in a CUDA graph).
    SideStream* side = overlap ? side_stream() : nullptr;
    if (overlap && !side) return DGCNN_ERR_CUDA;
    cudaStream_t sw = side ? side->stream : st;      // stream of the parameter-gradient chain
    auto fork = [&](int i) -> bool {                 // sw waits for everything issued on st so far
        if (!side) return true;
        return cudaEventRecord(side->ev[i], st) == cudaSuccess &&
               cudaStreamWaitEvent(sw, side->ev[i], 0) == cudaSuccess;
    };

    if (!after_loss) {
        DGCNN_LAUNCH(tail_fc2_bwd_rows, grid_for(B, 8, 4), 256, 0, st, dlogp, logp, keep, B, num_classes, wf2, dlogit, dz3);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    if (!fork(0)) return DGCNN_ERR_CUDA;
    DGCNN_LAUNCH(tail_fc2_bwd_params, (num_classes * kFc + num_classes + kFc + 31) / 32 + (after_loss ? 1 : 0), 256, 0, sw, 
        dlogit, h3, dz3, B, num_classes, dwf2, dbf2, dbf1, after_loss ? per_graph : nullptr, stats_after_loss,
        rng_offset);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    // dz2 = (dz3 Wf1) * (h2 > 0):  [B,128] x [128,D1]
    dim3 ga((unsigned)ceil_div(d.D1, kGemmBN), (unsigned)ceil_div(B, kGemmBM), 1);
    DGCNN_LAUNCH((gemm_f32<false, true, true>), ga, 256, 0, st, dz3, kFc, wf1, d.D1, dz2, (int)B, d.D1, kFc, kFc, h2);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    // dWf1 = dz3^T h2:  [128,B] x [B,D1]
    {   // split over the batch, slabs summed in order
        const int kchunk = (int)ceil_div(ceil_div(B, kDwSplits), 16) * 16;
        const int splits = (int)ceil_div(B, kchunk);
        dim3 gb((unsigned)ceil_div(d.D1, kGemmBN), (unsigned)ceil_div(kFc, kGemmBM), (unsigned)splits);
        DGCNN_LAUNCH((gemm_f32<true, true, false>), gb, 256, 0, sw, dz3, kFc, h2, d.D1, slabw, kFc, d.D1, (int)B, kchunk,
                                                        nullptr);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        const int total = kFc * d.D1;
        DGCNN_LAUNCH(tail_reduce_partials, (total + 31) / 32, 256, 0, sw, slabw, splits, total, total, dwf1, dwf1);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    const size_t smem_in = sizeof(float) * (kC6 * kK6 * kC5 + kC6 * (d.L2 + 2 * (kK6 - 1) + 8) + kC5 * d.L1);
    const size_t smem_w = sizeof(float) * (kC6 * d.L2 + kC5 * (d.L1 | 1));
    if (smem_in > 200 * 1024 || smem_w > 48 * 1024) return DGCNN_ERR_UNSUPPORTED;
    if (smem_in > 48 * 1024 &&
        cudaFuncSetAttribute(tail_c6_bwd_input, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem_in) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    if (!fork(1)) return DGCNN_ERR_CUDA;             // dz2 is ready
    DGCNN_LAUNCH(tail_c6_bwd_input, grid_for(B, 1, 4), 256, smem_in, st, dz2, B, d.L1, w6, dh1);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    const int parts6 = grid_for(B, 1, 2);
    DGCNN_LAUNCH(tail_c6_bwd_weight, parts6, 256, smem_w, sw, dz2, h1, B, d.L1, part6);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    const int n6 = kC6 * kC5 * kK6;
    DGCNN_LAUNCH(tail_reduce_partials, (n6 + kC6 + 31) / 32, 256, 0, sw, part6, parts6, n6 + kC6, n6, dw6, db6);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (!to_h1) {
        if (!fork(2)) return DGCNN_ERR_CUDA;             // dh1 is ready
        DGCNN_LAUNCH(tail_c5_bwd_input, grid_for(B * d.L1, 32, 8), 256, 0, st, dh1, arg, B, k, d.L1, w5, dpooled);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        const size_t smem5 = sizeof(float) * (kC5Pairs * kC5Row + 2 * kC5Pairs * kC5);
        if (cudaFuncSetAttribute(tail_c5_bwd_weight, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem5) != cudaSuccess)
            return DGCNN_ERR_CUDA;
        const int parts5 = grid_for(B * d.L1, kC5Pairs, 2);
        DGCNN_LAUNCH(tail_c5_bwd_weight, parts5, 256, smem5, sw, dh1, arg, pooled, B, k, d.L1, part5);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
        const int n5 = kC5 * kKW;
        DGCNN_LAUNCH(tail_reduce_partials, (n5 + kC5 + 31) / 32, 256, 0, sw, part5, parts5, n5 + kC5, n5, dw5, db5);
        DGCNN_RETURN_IF_LAUNCH_FAILED();
    }
    if (side) {
        if (cudaEventRecord(side->done, sw) != cudaSuccess) return DGCNN_ERR_CUDA;
        if (overlap == 1 && cudaStreamWaitEvent(st, side->done, 0) != cudaSuccess) return DGCNN_ERR_CUDA;
    }
    return DGCNN_OK;
}

extern "C" int dgcnn_tail_bwd(const float* dlogp, const float* pooled, int64_t num_graphs, int32_t k,
                              const float* w5, const float* w6, const float* wf1, const float* wf2,
                              int32_t num_classes, const float* h1, const uint8_t* arg, const float* h2,
                              const float* h3, const uint8_t* keep, const float* logp, float* dpooled,
                              float* dw5, float* db5, float* dw6, float* db6, float* dwf1, float* dbf1,
                              float* dwf2, float* dbf2, int32_t overlap, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return tail_bwd_impl(dlogp, pooled, num_graphs, k, w5, w6, wf1, wf2, num_classes, h1, arg, h2, h3, keep, logp,
                         dpooled, dw5, db5, dw6, db6, dwf1, dbf1, dwf2, dbf2, overlap, workspace, workspace_bytes,
                         stream, false, nullptr);
}

// The tail's backward down to d(h1) [B,16,k/2] only: conv5's own backward runs
// inside dgcnn_stack_bwd_conv5.  dw6 .. dbf2 as in dgcnn_tail_bwd.
extern "C" int dgcnn_tail_bwd_h1(const float* dlogp, int64_t num_graphs, int32_t k, const float* w6,
                                 const float* wf1, const float* wf2, int32_t num_classes, const float* h1,
                                 const float* h2, const float* h3, const uint8_t* keep, const float* logp,
                                 float* dh1, float* dw6, float* db6, float* dwf1, float* dbf1, float* dwf2,
                                 float* dbf2, int32_t overlap, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!dh1 && num_graphs > 0) return DGCNN_ERR_INVALID_ARGUMENT;
    return tail_bwd_impl(dlogp, nullptr, num_graphs, k, nullptr, w6, wf1, wf2, num_classes, h1, nullptr, h2, h3,
                         keep, logp, nullptr, nullptr, nullptr, dw6, db6, dwf1, dbf1, dwf2, dbf2, overlap,
                         workspace, workspace_bytes, stream, true, dh1);
}

// The backward after dgcnn_tail_fwd_loss: `workspace` is the workspace_bwd of that call (dlogit, dz3 and
// the per-graph scalars are in it), `stats` receives [sum of NLL, #correct], `rng_offset` advances.
// dh1 != NULL: stop at d(h1) (conv5's backward runs in dgcnn_stack_bwd_conv5; pooled, arg,
// w5, dpooled, dw5, db5 unused); dh1 == NULL: the whole tail down to dpooled.
extern "C" int dgcnn_tail_bwd_after_loss(const float* pooled, int64_t num_graphs, int32_t k, const float* w5,
                                         const float* w6, const float* wf1, const float* wf2,
                                         int32_t num_classes, const float* h1, const uint8_t* arg,
                                         const float* h2, const float* h3, const uint8_t* keep,
                                         const float* logp, float* dpooled, float* dh1, float* dw5, float* db5,
                                         float* dw6, float* db6, float* dwf1, float* dbf1, float* dwf2,
                                         float* dbf2, float* stats, int64_t* rng_offset, int32_t overlap,
                                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!stats) return DGCNN_ERR_INVALID_ARGUMENT;
    return tail_bwd_impl(nullptr, pooled, num_graphs, k, w5, w6, wf1, wf2, num_classes, h1, arg, h2, h3, keep, logp,
                         dpooled, dw5, db5, dw6, db6, dwf1, dbf1, dwf2, dbf2, overlap, workspace, workspace_bytes,
                         stream, dh1 != nullptr, dh1, stats, rng_offset);
}

extern "C" int dgcnn_tail_bwd_join(void* stream) {
    SideStream* side = side_stream();
    if (!side) return DGCNN_ERR_CUDA;
    if (cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), side->done, 0) != cudaSuccess)
        return DGCNN_ERR_CUDA;
    return DGCNN_OK;
}


// h[n][32] = x[n][cin] W^T (W [32][cin], PyG's `lin` of GCNConv.forward) on the tensor cores
// (3xTF32, fp32-accurate): project FIRST when the input is wider than the 32 output channels
// (D&D F = 90, power-law F = 64), then aggregate 32-wide rows with dgcnn_graph_conv_fwd(weight =
// NULL) -- the order PyG itself uses.
extern "C" int dgcnn_project_rows(const float* x, int64_t ldx, int32_t cin, const float* weight, float* h,
                                  int64_t num_nodes, void* stream) {
    if (num_nodes < 0 || cin < 1 || ldx < cin) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_nodes >= (int64_t)kGemmBM * 65535 || num_nodes >= INT32_MAX / 64) return DGCNN_ERR_UNSUPPORTED;
    if (num_nodes == 0) return DGCNN_OK;
    if (!x || !weight || !h) return DGCNN_ERR_INVALID_ARGUMENT;
    dim3 grid(1, (unsigned)ceil_div(num_nodes, kGemmBM), 1);
    const int kchunk = (int)ceil_div(cin, 16) * 16;
    DGCNN_LAUNCH((gemm_f32<false, false, false>), grid, 256, 0, static_cast<cudaStream_t>(stream), 
        x, ldx, weight, cin, h, (int)num_nodes, 32, cin, kchunk, nullptr);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                               int64_t n, int64_t* step, float lr, float beta1, float beta2, float eps,
                               float grad_scale, void* stream) {
    if (n < 0 || !step) return DGCNN_ERR_INVALID_ARGUMENT;
    if (n == 0) return DGCNN_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DGCNN_LAUNCH(adam_flat, grid_for(n, 256, 4), 256, 0, st, params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1,
                                                   beta2, eps, grad_scale);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(adam_bump, 1, 1, 0, st, step);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}

extern "C" int dgcnn_nll_sum(const float* logp, const int64_t* y, int64_t num_graphs, int32_t num_classes,
                             float grad_scale, float* stats, float* dlogp, void* stream) {
    if (num_graphs < 0 || num_classes < 1 || !stats) return DGCNN_ERR_INVALID_ARGUMENT;
    if (num_graphs > 0 && (!logp || !y)) return DGCNN_ERR_INVALID_ARGUMENT;
    DGCNN_LAUNCH(nll_sum_kernel, 1, 1024, 0, static_cast<cudaStream_t>(stream), logp, y, num_graphs, num_classes,
                                                                     grad_scale, stats, dlogp);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
