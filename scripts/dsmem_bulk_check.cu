// Stand-alone check of the distributed-shared-memory bulk exchange used by the split graphs of KS
// (graph_mma.cuh: push_to_peer / mbar_*): each CTA of a cluster pair fills its half of a buffer,
// pushes it into the peer with cp.async.bulk shared::cta -> shared::cluster and waits on its own
// mbarrier; both CTAs must end up with the whole buffer.   nvcc -arch=sm_100a dsmem_bulk_check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dgcnn_b200/csrc/graph_mma.cuh"
using namespace dgcnn;

__global__ void k(float* out, int rounds) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ __align__(8) uint64_t mbar;
    uint32_t rank; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t mb = smem_addr_u32(&mbar);
    if (threadIdx.x == 0) mbar_init(mb, 1);
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    float* f = reinterpret_cast<float*>(sm);
    const uint32_t peer_mbar = map_to_peer(mb, rank ^ 1);
    uint32_t parity = 0;
    for (int r = 0; r < rounds; ++r) {
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) f[rank * 1024 + i] = r * 10000.f + rank * 1000.f + i;
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(mb, 4096);
            push_to_peer(smem_addr_u32(f + rank * 1024), 4096, rank ^ 1, peer_mbar);
        }
        mbar_wait(mb, parity);
        parity ^= 1;
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[(r * gridDim.x + blockIdx.x) * 2048 + i] = f[i];
        // my half is rewritten next round: the peer has it once it passed its wait (cluster barrier)
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

int main() {
    const int rounds = 50, ctas = 148;
    float* d; cudaMalloc(&d, sizeof(float) * rounds * ctas * 2048);
    cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 8192; cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, d, rounds);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    float* h = new float[(size_t)rounds * ctas * 2048];
    cudaMemcpy(h, d, sizeof(float) * rounds * ctas * 2048, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int r = 0; r < rounds; ++r) for (int b = 0; b < ctas; ++b) for (int i = 0; i < 2048; ++i) {
        float want = r * 10000.f + (i / 1024) * 1000.f + (i % 1024);
        if (h[((size_t)r * ctas + b) * 2048 + i] != want) ++bad;
    }
    printf("dsmem bulk exchange: %ld mismatches over %d rounds x %d CTAs\n", bad, rounds, ctas);
    return bad != 0;
}
