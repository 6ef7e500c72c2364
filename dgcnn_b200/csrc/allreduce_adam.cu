// X1 + N3 -- the step's one collective fused with the optimizer: a one-shot all-reduce of the
// flat gradient buffer over NVLink / NVSwitch PEER MEMORY followed, in the same kernel, by the
// flat Adam update (train.py:40-42 across ranks; SURVEY.md 8e: one sum of 0.2-4 MB per step,
// latency-bound).
//
// Every rank owns an exchange buffer that all the others map through CUDA IPC:
//
//     [ header: arrival counters of parity 0 / 1, 128 B apart ]
//     [ parity 0: W slots of n floats, slot r written by rank r ][ parity 1: the same ]
//
// PUSH model (remote stores are fire-and-forget, remote loads cost a NVLink round trip each):
// one launch (G CTAs, all resident) per step e, parity = e & 1:
//   1. every rank stores its local gradient sums into slot [rank] of EVERY rank's buffer
//      (its own included), fences to system scope, and adds one arrival per CTA to every
//      rank's counter (remote atomics);
//   2. thread 0 of every CTA polls the LOCAL counter until it shows G * W * (e/2 + 1);
//   3. the W local slots are summed IN RANK ORDER (every rank computes bit-identical sums), the
//      global sums go back to `grads` (the loss / accuracy scalars ride along), Adam is applied.
// Two parity regions make one barrier per step enough: a rank pushes step e+2 only after it
// passed the barrier of step e+1, which every rank joins only after its step-e kernel is over,
// so nobody still reads the region that step e+2 overwrites.  A poll that does not complete
// within ~20 s raises DGCNN_COMM_TIMEOUT in `status` instead of hanging the GPU.
//
// NCCL costs ~90 us per step here (launch + ring + rank skew); see DESIGN.md for this kernel.
#include <cstring>

#include "common.cuh"

namespace dgcnn {

constexpr int kArMaxWorld = 16;
constexpr int kArCtas = DGCNN_NUM_SMS;     // one CTA per SM, all co-resident (the kernel is its own barrier)
constexpr int kArThreads = 512;            // ~1 MB per rank: about one 16-byte element per thread
constexpr int kArHeaderBytes = 256;

struct AllreduceAdamParams {
    float* p; float* g; float* m; float* v;
    int64_t n_params, n_total;
    const int64_t* step; const int64_t* epoch;
    float lr, beta1, beta2, eps, grad_scale;
    unsigned char* exch[kArMaxWorld];
    int world, rank;
    int32_t* status;
    int64_t* trace;      // optional debug timeline (dgcnn_allreduce_set_trace): [steps][4] globaltimer ns
};

__device__ __forceinline__ int64_t ar_global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (int64_t)t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kArThreads)
allreduce_adam_kernel(AllreduceAdamParams a) {
    DGCNN_PDL_WAIT();
    const int64_t e = *a.epoch;
    const int parity = (int)(e & 1);
    const bool tracer = a.trace && blockIdx.x == 0 && threadIdx.x == 0;
    int64_t* tr = tracer ? a.trace + (e & 1023) * 4 : nullptr;
    if (tracer) tr[0] = ar_global_ns();                   // kernel entered: this rank's step is over
    const uint32_t target = (uint32_t)(gridDim.x * (uint64_t)a.world * (uint64_t)(e / 2 + 1));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = (a.n_total + 31) / 32 * 32;             // floats per slot
    const int64_t region = (int64_t)parity * a.world * slot;     // floats before this parity's slots

    // 1. push the local sums into slot [rank] of every rank's buffer: 16-byte remote stores (the
    //    slots are 128-byte aligned; `g` is when n_total is a multiple of 4 and the buffer aligned)
    const bool vec = ((reinterpret_cast<uintptr_t>(a.g) & 15) == 0);
    const int64_t n4 = vec ? a.n_total / 4 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        const float4 gi = reinterpret_cast<const float4*>(a.g)[i];
        for (int r = 0; r < a.world; ++r)
            reinterpret_cast<float4*>(reinterpret_cast<float*>(a.exch[r] + kArHeaderBytes) + region +
                                      (int64_t)a.rank * slot)[i] = gi;
    }
    for (int64_t i = 4 * n4 + tid; i < a.n_total; i += stride) {
        const float gi = a.g[i];
        for (int r = 0; r < a.world; ++r)
            reinterpret_cast<float*>(a.exch[r] + kArHeaderBytes)[region + (int64_t)a.rank * slot + i] = gi;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        for (int r = 0; r < a.world; ++r)
            atomicAdd_system(reinterpret_cast<uint32_t*>(a.exch[r] + parity * 128), 1u);
    if (tracer) tr[1] = ar_global_ns();                   // sums pushed, arrival signalled

    // 2. wait until every CTA of every rank has arrived at the LOCAL counter.  A peer that never
    //    shows up (~20 s) raises DGCNN_COMM_TIMEOUT and the step is ABANDONED: no sum, no Adam, no
    //    counter bump -- stale or missing slots must never reach the parameters.
    __shared__ int timed_out;
    if (threadIdx.x == 0) {
        timed_out = 0;
        const uint32_t* cnt = reinterpret_cast<const uint32_t*>(a.exch[a.rank] + parity * 128);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(cnt) - target) < 0) {
            if ((a.status && (*reinterpret_cast<volatile int32_t*>(a.status) & DGCNN_COMM_TIMEOUT)) ||
                clock64() - t0 > 40000000000ll) {                 // ~20 s: a peer is gone
                if (a.status) atomicOr(a.status, DGCNN_COMM_TIMEOUT);
                timed_out = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (timed_out) return;
    if (tracer) tr[2] = ar_global_ns();                   // every rank has arrived (wait = [2] - [1])

    // 3. sum the local slots in rank order, Adam on the parameters (16 bytes per thread and step
    //    where the layout allows: slots are 128-byte aligned, the state arrays are torch allocations)
    const int64_t t = *a.step + 1;
    const float bc1 = 1.f - powf(a.beta1, (float)t), bc2 = 1.f - powf(a.beta2, (float)t);
    const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float* mine = reinterpret_cast<const float*>(a.exch[a.rank] + kArHeaderBytes) + region;
    auto adam1 = [&](float s, float& pi, float& mi, float& vi) {
        const float gi = s * a.grad_scale;
        mi = a.beta1 * mi + (1.f - a.beta1) * gi;
        vi = a.beta2 * vi + (1.f - a.beta2) * gi * gi;
        pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + a.eps);
    };
    const bool vec3 = vec && (((reinterpret_cast<uintptr_t>(a.p) | reinterpret_cast<uintptr_t>(a.m) |
                                reinterpret_cast<uintptr_t>(a.v)) & 15) == 0);
    const int64_t p4 = vec3 ? a.n_params / 4 : 0;              // float4 groups that are all parameters
    for (int64_t i = tid; i < p4; i += stride) {
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < a.world; ++r) {
            const float4 x4 = __ldcv(reinterpret_cast<const float4*>(mine + (int64_t)r * slot) + i);
            s4.x += x4.x; s4.y += x4.y; s4.z += x4.z; s4.w += x4.w;
        }
        reinterpret_cast<float4*>(a.g)[i] = s4;
        float4 pv = reinterpret_cast<float4*>(a.p)[i], mv = reinterpret_cast<float4*>(a.m)[i],
               vv = reinterpret_cast<float4*>(a.v)[i];
        adam1(s4.x, pv.x, mv.x, vv.x); adam1(s4.y, pv.y, mv.y, vv.y);
        adam1(s4.z, pv.z, mv.z, vv.z); adam1(s4.w, pv.w, mv.w, vv.w);
        reinterpret_cast<float4*>(a.p)[i] = pv;
        reinterpret_cast<float4*>(a.m)[i] = mv;
        reinterpret_cast<float4*>(a.v)[i] = vv;
    }
    for (int64_t i = 4 * p4 + tid; i < a.n_total; i += stride) {
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) s += __ldcv(mine + (int64_t)r * slot + i);
        a.g[i] = s;
        if (i < a.n_params) {
            float pi = a.p[i], mi = a.m[i], vi = a.v[i];
            adam1(s, pi, mi, vi);
            a.p[i] = pi; a.m[i] = mi; a.v[i] = vi;
        }
    }
    if (tracer) tr[3] = ar_global_ns();                   // CTA 0 done with its share of sum + Adam
}

__global__ void allreduce_adam_bump(int64_t* step, int64_t* epoch, const int32_t* status) {
    DGCNN_PDL_WAIT();
    if (status && (*status & DGCNN_COMM_TIMEOUT)) return;         // the step was abandoned
    *step += 1;
    *epoch += 1;
}

}  // namespace dgcnn

using namespace dgcnn;

// ---- exchange buffers: set-up calls (these allocate; the per-step call never does) -----------
// The buffer is a cudaMalloc allocation of its own, so that the IPC handle maps exactly it.
// A peer opens the handle WITH ITS OWN DEVICE CURRENT: cudaIpcOpenMemHandle then maps the
// exporter's memory into the importer's context and enables peer access between the two
// devices (importing it into the exporter's device context of the importing process does not
// make it visible to kernels of another device).
static size_t exchange_bytes(int64_t n_total, int world) {
    return (size_t)kArHeaderBytes + 2 * (size_t)world * sizeof(float) * (size_t)((n_total + 31) / 32 * 32);
}

extern "C" int dgcnn_exchange_create(int64_t n_total, int32_t world, void** local_ptr, unsigned char* handle64) {
    if (n_total < 0 || world < 1 || world > kArMaxWorld || !local_ptr || !handle64)
        return DGCNN_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const size_t bytes = exchange_bytes(n_total, world);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return DGCNN_ERR_CUDA;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaFree(p);
        return DGCNN_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return DGCNN_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    *local_ptr = p;
    return DGCNN_OK;
}

extern "C" int dgcnn_exchange_open(const unsigned char* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) return DGCNN_ERR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        return DGCNN_ERR_UNSUPPORTED;
    }
    *peer_ptr = p;
    return DGCNN_OK;
}

extern "C" int dgcnn_exchange_close(void* peer_ptr) {
    return (!peer_ptr || cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess) ? DGCNN_OK : DGCNN_ERR_CUDA;
}

extern "C" int dgcnn_exchange_destroy(void* local_ptr) {
    return (!local_ptr || cudaFree(local_ptr) == cudaSuccess) ? DGCNN_OK : DGCNN_ERR_CUDA;
}

static int64_t* g_ar_trace = nullptr;
extern "C" void dgcnn_allreduce_set_trace(int64_t* device_buffer) { g_ar_trace = device_buffer; }

extern "C" size_t dgcnn_allreduce_adam_exchange_bytes(int64_t n_total, int32_t world) {
    if (n_total < 0 || world < 1) return 0;
    return exchange_bytes(n_total, world);
}

extern "C" int dgcnn_allreduce_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                    int64_t n_params, int64_t n_total, int64_t* step, int64_t* epoch,
                                    float lr, float beta1, float beta2, float eps, float grad_scale,
                                    void* const* exchange, int32_t world, int32_t rank, int32_t* status,
                                    void* stream) {
    if (n_params < 0 || n_total < n_params || !step || !epoch || !exchange) return DGCNN_ERR_INVALID_ARGUMENT;
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world) return DGCNN_ERR_INVALID_ARGUMENT;
    if (n_total == 0) return DGCNN_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return DGCNN_ERR_INVALID_ARGUMENT;
    AllreduceAdamParams a{};
    a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq;
    a.n_params = n_params; a.n_total = n_total; a.step = step; a.epoch = epoch;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
    for (int r = 0; r < world; ++r) {
        if (!exchange[r]) return DGCNN_ERR_INVALID_ARGUMENT;
        a.exch[r] = static_cast<unsigned char*>(exchange[r]);
    }
    a.world = world; a.rank = rank; a.status = status;
    a.trace = g_ar_trace;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DGCNN_LAUNCH(allreduce_adam_kernel, kArCtas, kArThreads, 0, st, a);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    DGCNN_LAUNCH(allreduce_adam_bump, 1, 1, 0, st, step, epoch, status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    return DGCNN_OK;
}
