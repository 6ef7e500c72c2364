// One training step of the reference loop (train.py:35-45: model(data) -> NLL -> backward ->
// Adam) as ONE host call: K0 -> K0b -> KS -> KT forward -> NLL -> KT backward -> KSB ->
// [peer all-reduce +] Adam, every intermediate buffer carved from a caller-owned arena.
// The Python trainer makes the same sequence of C-ABI calls one by one (~30 launches, ~45
// allocations: ~0.5 ms of interpreter time per step, more than the GPU needs for the
// reference's small batches); this entry point is the same sequence without the interpreter.
// No allocation, no synchronisation: capturable in a CUDA graph.
#include <cstdlib>

#include "common.cuh"

// conv5 + ReLU + max-pool run inside KS / KSB whenever the batch fits (no pooled /
// dpooled round trip).  dgcnn_train_step_configure(0) or DGCNN_FUSE_CONV5=0 keeps the unfused
// sequence (A/B timing, tests).
static int g_fuse_conv5 = -1;
extern "C" void dgcnn_train_step_configure(int32_t fuse_conv5) { g_fuse_conv5 = fuse_conv5 < 0 ? -1 : (fuse_conv5 != 0); }
// Lazy adjacency maps: K0b shrinks to the offsets / descriptors (and the gated A_hat^T bitmap of an
// asymmetric batch); the forward kernel expands each graph's CSR rows into its fragment map itself
// and exports it for the backward kernel -- two launches and one pass over col[] less per step.
// dgcnn_train_step_configure_maps(0) or DGCNN_LAZY_MAPS=0 keeps the full K0b.
static int g_lazy_maps = -1;
extern "C" void dgcnn_train_step_configure_maps(int32_t lazy) { g_lazy_maps = lazy < 0 ? -1 : (lazy != 0); }
static bool lazy_maps_enabled() {
    if (g_lazy_maps < 0) {
        const char* env = getenv("DGCNN_LAZY_MAPS");
        g_lazy_maps = !(env && env[0] == '0');
    }
    return g_lazy_maps != 0;
}
void dgcnn_stack_fwd_next_lazy(int on);                  // graph_stack_mma.cu
int dgcnn_build_bitmaps_impl(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* col_t,
                             const int32_t* gptr, const int64_t* batch, const int32_t* batch32, int64_t num_nodes,
                             int64_t num_graphs, int64_t max_nodes, uint32_t* bitmap, uint32_t* bitmap_t,
                             int64_t bitmap_words, int32_t* bmoff, int32_t* gflags, int32_t* gflags_t,
                             uint32_t* fragmap, int64_t fragmap_words, int32_t* fgoff, const int32_t* gorder,
                             int32_t* gdesc, const int32_t* gate_word, int32_t gate_mask, void* stream, int lazy);

static bool fuse_conv5_enabled() {
    if (g_fuse_conv5 < 0) {
        const char* env = getenv("DGCNN_FUSE_CONV5");
        g_fuse_conv5 = !(env && env[0] == '0');
    }
    return g_fuse_conv5 != 0;
}

// graph_status is int32[2]: [0] the flags of the CURRENT step (DGCNN_GRAPH_GENERIC is per-batch
// state the kernels read back), [1] the OR of the error flags of every earlier step since the
// caller last cleared it -- a bad batch in the middle of an epoch is not lost (one tiny launch
// in place of the memset that used to clear the word).
__global__ void step_status_begin(int32_t* status) {
    DGCNN_PDL_WAIT();
    status[1] |= status[0] & ~DGCNN_GRAPH_GENERIC;
    status[0] = 0;
}

namespace {

struct Arena {
    char* base;
    size_t off;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += sizeof(T) * (count ? count : 1);
        return p;
    }
};

struct StepBuffers {
    int32_t *rowptr, *col, *rowptr_t, *col_t, *gptr, *gorder, *bmoff, *gflags, *gflags_t, *fgoff, *gdesc, *perm;
    uint32_t *bitmap, *bitmap_t, *fragmap;
    float *dis, *xcat, *pooled, *h1, *h2, *h3, *logp, *dlogp, *dpooled, *dh1;
    uint8_t *arg, *keep;
    void *ws_build, *ws_fwd, *ws_tail_f, *ws_tail_b, *ws_bwd;
    size_t n_build, n_fwd, n_tail, n_bwd;
    int64_t bm_words, fm_words;
    // resident data set (dgcnn_train_step_resident): the collated batch lives in the arena too
    float* x_batch;
    int32_t* batch32;
    int64_t* y_batch;
};

constexpr int kXcatLd = 100;

// does this step run conv5 + ReLU + max-pool inside KS / KSB (no pooled / dpooled at all)?
bool step_fuses_conv5(int32_t F, int64_t max_nodes) {
    return fuse_conv5_enabled() && dgcnn_stack_bwd_supported(F, max_nodes) == 1 &&
           dgcnn_stack_fwd_conv5_supported(F, max_nodes) && dgcnn_stack_bwd_conv5_supported(F, max_nodes);
}

StepBuffers carve(Arena& a, int64_t N, int64_t E, int64_t B, int32_t F, int32_t k, int32_t C,
                  int64_t max_nodes, bool resident) {
    StepBuffers s{};
    const bool fused5 = step_fuses_conv5(F, max_nodes);
    const int64_t l1 = k / 2, d1 = 32 * (l1 - 4);
    s.bm_words = dgcnn_graph_bitmap_words(N, B, max_nodes);
    s.fm_words = dgcnn_graph_fragmap_words(N, B, max_nodes);
    s.rowptr = a.take<int32_t>(N + 1);
    s.col = a.take<int32_t>(E);
    s.rowptr_t = a.take<int32_t>(N + 1);
    s.col_t = a.take<int32_t>(E);
    s.dis = a.take<float>(N);
    s.gptr = a.take<int32_t>(B + 1);
    s.gorder = a.take<int32_t>(B);
    s.bitmap = a.take<uint32_t>(2 * s.bm_words);          // A_hat | A_hat^T, adjacent: one memset
    s.bitmap_t = s.bitmap ? s.bitmap + s.bm_words : nullptr;
    s.bmoff = a.take<int32_t>(B + 1);
    s.gflags = a.take<int32_t>(B);
    s.gflags_t = a.take<int32_t>(B);
    s.fragmap = a.take<uint32_t>(s.fm_words);
    s.fgoff = a.take<int32_t>(B + 1);
    s.gdesc = a.take<int32_t>(4 * B);
    s.xcat = a.take<float>(N * kXcatLd);
    s.pooled = fused5 ? nullptr : a.take<float>(B * k * 97);   // N2: SortPooling's output is never materialised
    s.perm = a.take<int32_t>(B * k);
    s.h1 = a.take<float>(B * 16 * l1);
    s.arg = a.take<uint8_t>(B * 16 * l1);
    s.h2 = a.take<float>(B * d1);
    s.h3 = a.take<float>(B * 128);
    s.keep = a.take<uint8_t>(B * 128);
    s.logp = a.take<float>(B * C);
    s.dlogp = a.take<float>(B * C);
    s.dpooled = fused5 ? nullptr : a.take<float>(B * k * 97);
    s.dh1 = a.take<float>(B * 16 * l1);
    s.n_build = resident ? dgcnn_collate_workspace_bytes(B) : dgcnn_build_graph_workspace_bytes(N, E);
    s.n_fwd = dgcnn_stack_fwd_workspace_bytes();
    s.n_tail = dgcnn_tail_workspace_bytes(B, k, C);
    s.n_bwd = dgcnn_stack_bwd_workspace_bytes(F, B, N);
    s.ws_build = a.take<char>(s.n_build);
    s.ws_fwd = a.take<char>(s.n_fwd);
    s.ws_tail_f = a.take<char>(s.n_tail);
    s.ws_tail_b = a.take<char>(s.n_tail);
    s.ws_bwd = a.take<char>(s.n_bwd);
    if (resident) {
        s.x_batch = a.take<float>(N * F);
        s.batch32 = a.take<int32_t>(N);
        s.y_batch = a.take<int64_t>(B);
    }
    return s;
}

}  // namespace

extern "C" size_t dgcnn_train_step_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t num_graphs,
                                                   int32_t num_features, int32_t k, int32_t num_classes,
                                                   int64_t max_nodes) {
    if (num_nodes < 0 || num_edges < 0 || num_graphs < 1 || num_features < 1 || k < 10 || num_classes < 1 ||
        max_nodes < 1)
        return 0;
    Arena a{nullptr, 0};
    carve(a, num_nodes, num_edges, num_graphs, num_features, k, num_classes, max_nodes, false);
    return a.off + 512;
}

extern "C" size_t dgcnn_train_step_resident_workspace_bytes(int64_t num_nodes, int64_t num_edges,
                                                            int64_t num_graphs, int32_t num_features,
                                                            int32_t k, int32_t num_classes,
                                                            int64_t max_nodes) {
    if (num_nodes < 0 || num_edges < 0 || num_graphs < 1 || num_features < 1 || k < 10 || num_classes < 1 ||
        max_nodes < 1)
        return 0;
    Arena a{nullptr, 0};
    carve(a, num_nodes, num_edges, num_graphs, num_features, k, num_classes, max_nodes, true);
    return a.off + 512;
}

extern "C" int64_t dgcnn_train_step_num_params(int32_t num_features, int32_t k, int32_t num_classes) {
    const int64_t d1 = 32 * ((int64_t)k / 2 - 4);
    return dgcnn_stack_num_params(num_features) + 16 * 97 + 16 + 32 * 16 * 5 + 32 + 128 * d1 + 128 +
           (int64_t)num_classes * 128 + num_classes;
}

namespace {

#define DGCNN_TRY(call) do { const int rc_ = (call); if (rc_ != DGCNN_OK) return rc_; } while (0)

struct StepArgs {
    int64_t N, E, B;
    int32_t F, k, C;
    int64_t max_nodes;
    int32_t norm;
    float *params, *grads, *exp_avg, *exp_avg_sq;
    int64_t* step;
    float lr, beta1, beta2, eps;
    int64_t global_batch;
    int32_t training;
    uint64_t seed;
    int64_t* rng_offset;
    void* const* exchange;
    int32_t world, rank;
    int64_t* epoch;
    int32_t *comm_status, *graph_status;
    void* workspace;
    size_t workspace_bytes;
    void* stream;
};

int check_step_args(const StepArgs& t, bool resident) {
    if (t.N < 1 || t.E < 0 || t.B < 1 || t.F < 1 || t.k < 10 || t.C < 1 || t.max_nodes < 1 || t.global_batch < 1)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (!t.params || !t.grads || !t.exp_avg || !t.exp_avg_sq || !t.step || !t.graph_status || !t.workspace)
        return DGCNN_ERR_INVALID_ARGUMENT;
    if (t.training && !t.rng_offset) return DGCNN_ERR_INVALID_ARGUMENT;
    if (t.C > 32) return DGCNN_ERR_UNSUPPORTED;
    if (!dgcnn_stack_fwd_supported(t.F, t.max_nodes) || dgcnn_stack_bwd_supported(t.F, t.max_nodes) == 0)
        return DGCNN_ERR_UNSUPPORTED;
    const size_t need = resident
        ? dgcnn_train_step_resident_workspace_bytes(t.N, t.E, t.B, t.F, t.k, t.C, t.max_nodes)
        : dgcnn_train_step_workspace_bytes(t.N, t.E, t.B, t.F, t.k, t.C, t.max_nodes);
    if (t.workspace_bytes < need) return DGCNN_ERR_WORKSPACE;
    return DGCNN_OK;
}

// Everything after the graph build: K0b -> KS -> KT forward -> NLL -> KT backward -> KSB ->
// [peer all-reduce +] Adam.  batch64 / batch32: the node -> graph vector in whichever form
// the batch has it (K0b saves a search per row with it).
int step_after_build(const StepArgs& t, const StepBuffers& s, const float* x, int64_t ldx,
                     const int64_t* batch64, const int32_t* batch32, const int64_t* y, bool maps_ready) {
    const int64_t N = t.N, B = t.B;
    const int32_t F = t.F, C = t.C, k = t.k;
    void* stream = t.stream;
    // parameters and gradients: the flat layout of FusedTrainer (PyG order for the graph
    // convolutions, then conv5, conv6, classifier_1, classifier_2; weight before bias)
    const int64_t d1 = 32 * ((int64_t)k / 2 - 4);
    const int64_t sizes[16] = {32LL * F, 32, 32 * 32, 32, 32 * 32, 32, 32, 1,
                               16 * 97, 16, 32 * 16 * 5, 32, 128 * d1, 128, (int64_t)C * 128, C};
    float* p[16];
    float* g[16];
    int64_t off = 0;
    for (int i = 0; i < 16; ++i) { p[i] = t.params + off; g[i] = t.grads + off; off += sizes[i]; }
    const int64_t n_params = off;
    float* stats = t.grads + n_params;                // [sum of NLL, #correct] ride the all-reduce
    const int bwd_kind = dgcnn_stack_bwd_supported(F, t.max_nodes);   // 1 MMA, 2 FMA only

    // (the FMA backward fallback reads the row bitmaps: full K0b then)
    const bool lazy = !maps_ready && bwd_kind == 1 && lazy_maps_enabled();
    if (!maps_ready)                                  // (a resident data set hands K0b's outputs over)
        DGCNN_TRY(dgcnn_build_bitmaps_impl(s.rowptr, s.col, s.rowptr_t, s.col_t, s.gptr, batch64, batch32, N, B,
                                           t.max_nodes, s.bitmap, s.bitmap_t, s.bm_words, s.bmoff, s.gflags,
                                           s.gflags_t, s.fragmap, s.fm_words, s.fgoff, s.gorder, s.gdesc,
                                           t.graph_status, DGCNN_GRAPH_GENERIC, stream, lazy ? 1 : 0));
    dgcnn_stack_fwd_next_lazy(lazy ? 1 : 0);          // the forward call below fills fragmap / gflags itself
    if (step_fuses_conv5(F, t.max_nodes)) {
        // N2: KS emits h1 / arg, the tail starts at conv6, its backward stops at d(h1), KSB does the
        // rest and writes the ten gradients g[0..9] (GraphConv + conv5: one contiguous slice)
        DGCNN_TRY(dgcnn_stack_fwd_conv5(x, ldx, F, s.rowptr, s.col, s.dis, s.gptr, s.gorder, s.bitmap, s.bmoff,
                                        s.gflags, s.fragmap, s.fgoff, s.gdesc, N, B, t.max_nodes, p[0], p[1], p[2],
                                        p[3], p[4], p[5], p[6], p[7], p[8], p[9], s.xcat, kXcatLd, nullptr, s.perm,
                                        k, s.h1, s.arg, t.norm, t.graph_status, s.ws_fwd, s.n_fwd, stream));
        // conv6 -> fc1 -> [fc1 epilogue + fc2 + log_softmax + NLL + d(logits) + fc2 rows: one kernel]
        DGCNN_TRY(dgcnn_tail_fwd_loss(nullptr, B, k, p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15], C, y,
                                      t.training, t.seed, t.rng_offset, s.h1, s.arg, s.h2, s.h3, s.keep, s.logp,
                                      s.ws_tail_f, s.n_tail, s.ws_tail_b, s.n_tail, stream));
        DGCNN_TRY(dgcnn_tail_bwd_after_loss(nullptr, B, k, nullptr, p[10], p[12], p[14], C, s.h1, nullptr, s.h2,
                                            s.h3, s.keep, s.logp, nullptr, s.dh1, nullptr, nullptr, g[10], g[11],
                                            g[12], g[13], g[14], g[15], stats, t.training ? t.rng_offset : nullptr,
                                            2, s.ws_tail_b, s.n_tail, stream));
        DGCNN_TRY(dgcnn_stack_bwd_conv5(s.dh1, s.arg, s.perm, k, s.xcat, kXcatLd, x, ldx, F, s.rowptr_t, s.col_t,
                                        s.dis, s.gptr, s.gorder, s.gdesc, s.fragmap, s.bitmap, s.bmoff, s.gflags,
                                        s.bitmap_t, s.bmoff, s.gflags_t, N, B, t.max_nodes, p[2], p[4], p[6], p[8],
                                        t.norm, t.grads, t.graph_status, s.ws_bwd, s.n_bwd, stream));
    } else {
    DGCNN_TRY(dgcnn_stack_fwd(x, ldx, F, s.rowptr, s.col, s.dis, s.gptr, s.gorder, s.bitmap, s.bmoff,
                              s.gflags, s.fragmap, s.fgoff, s.gdesc, N, B, t.max_nodes, p[0], p[1], p[2], p[3],
                              p[4], p[5], p[6], p[7], s.xcat, kXcatLd, s.pooled, s.perm, k, t.norm,
                              DGCNN_STACK_MMA, t.graph_status, s.ws_fwd, s.n_fwd, stream));
    DGCNN_TRY(dgcnn_tail_fwd_loss(s.pooled, B, k, p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15], C, y,
                                  t.training, t.seed, t.rng_offset, s.h1, s.arg, s.h2, s.h3, s.keep, s.logp,
                                  s.ws_tail_f, s.n_tail, s.ws_tail_b, s.n_tail, stream));
    // the tail's parameter gradients run on the library's side stream underneath KSB
    DGCNN_TRY(dgcnn_tail_bwd_after_loss(s.pooled, B, k, p[8], p[10], p[12], p[14], C, s.h1, s.arg, s.h2, s.h3,
                                        s.keep, s.logp, s.dpooled, nullptr, g[8], g[9], g[10], g[11], g[12], g[13],
                                        g[14], g[15], stats, t.training ? t.rng_offset : nullptr, 2, s.ws_tail_b,
                                        s.n_tail, stream));
    DGCNN_TRY(dgcnn_stack_bwd(s.dpooled, s.perm, k, s.xcat, kXcatLd, x, ldx, F, s.rowptr_t, s.col_t, s.dis,
                              s.gptr, s.gorder, s.gdesc, s.fragmap, s.bitmap, s.bmoff, s.gflags, s.bitmap_t,
                              s.bmoff, s.gflags_t, N, B, t.max_nodes, p[2], p[4], p[6], t.norm,
                              bwd_kind == 1 ? DGCNN_STACK_MMA : DGCNN_STACK_FMA, t.grads, t.graph_status,
                              s.ws_bwd, s.n_bwd, stream));
    }
    DGCNN_TRY(dgcnn_tail_bwd_join(stream));
    const float scale = 1.0f / (float)t.global_batch;
    if (t.world > 1 && t.exchange) {
        if (!t.epoch) return DGCNN_ERR_INVALID_ARGUMENT;
        DGCNN_TRY(dgcnn_allreduce_adam(t.params, t.grads, t.exp_avg, t.exp_avg_sq, n_params, n_params + 2,
                                       t.step, t.epoch, t.lr, t.beta1, t.beta2, t.eps, scale, t.exchange,
                                       t.world, t.rank, t.comm_status, stream));
    } else {
        DGCNN_TRY(dgcnn_adam_step(t.params, t.grads, t.exp_avg, t.exp_avg_sq, n_params, t.step, t.lr, t.beta1,
                                  t.beta2, t.eps, scale, stream));
    }
    return DGCNN_OK;
}

}  // namespace

extern "C" int dgcnn_train_step(const float* x, int64_t ldx, const void* edge_index, int32_t index_is_i32,
                                const void* batch, const int64_t* y, int64_t num_nodes, int64_t num_edges,
                                int64_t num_graphs, int32_t num_features, int32_t k, int32_t num_classes,
                                int64_t max_nodes, int32_t norm, float* params, float* grads,
                                float* exp_avg, float* exp_avg_sq, int64_t* step, float lr, float beta1,
                                float beta2, float eps, int64_t global_batch, int32_t training,
                                uint64_t seed, int64_t* rng_offset, void* const* exchange, int32_t world,
                                int32_t rank, int64_t* epoch, int32_t* comm_status, int32_t* graph_status,
                                void* workspace, size_t workspace_bytes, void* stream) {
    const StepArgs t{num_nodes, num_edges, num_graphs, num_features, k, num_classes, max_nodes, norm,
                     params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, global_batch, training,
                     seed, rng_offset, exchange, world, rank, epoch, comm_status, graph_status, workspace,
                     workspace_bytes, stream};
    if (!x || !edge_index || !batch || !y) return DGCNN_ERR_INVALID_ARGUMENT;
    DGCNN_TRY(check_step_args(t, false));
    const int64_t N = num_nodes, E = num_edges, B = num_graphs;
    Arena a{reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255), 0};
    const StepBuffers s = carve(a, N, E, B, num_features, k, num_classes, max_nodes, false);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DGCNN_LAUNCH(step_status_begin, 1, 1, 0, st, graph_status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    if (index_is_i32)
        DGCNN_TRY(dgcnn_build_graph_i32(static_cast<const int32_t*>(edge_index), E,
                                        static_cast<const int32_t*>(batch), N, B, s.rowptr, s.col, s.rowptr_t,
                                        s.col_t, s.dis, s.gptr, s.gorder, graph_status, 0, s.ws_build,
                                        s.n_build, stream));
    else
        DGCNN_TRY(dgcnn_build_graph(static_cast<const int64_t*>(edge_index), E,
                                    static_cast<const int64_t*>(batch), N, B, s.rowptr, s.col, s.rowptr_t,
                                    s.col_t, s.dis, s.gptr, s.gorder, graph_status, 0, s.ws_build, s.n_build,
                                    stream));
    return step_after_build(t, s, x, ldx, index_is_i32 ? nullptr : static_cast<const int64_t*>(batch),
                            index_is_i32 ? static_cast<const int32_t*>(batch) : nullptr, y, false);
}

// The same step fed from a data set that is resident in HBM: dgcnn_collate
// gathers the batch's CSR, dis, x, batch and y from the data-set arrays -- no host-to-device
// copy of the batch (train.py:36) and no K0.
extern "C" int dgcnn_train_step_resident(const dgcnn_dataset* dataset, const int32_t* ids, int64_t num_nodes,
                                         int64_t num_edges, int64_t num_graphs, int32_t k,
                                         int32_t num_classes, int64_t max_nodes, int32_t norm, float* params,
                                         float* grads, float* exp_avg, float* exp_avg_sq, int64_t* step,
                                         float lr, float beta1, float beta2, float eps, int64_t global_batch,
                                         int32_t training, uint64_t seed, int64_t* rng_offset,
                                         void* const* exchange, int32_t world, int32_t rank, int64_t* epoch,
                                         int32_t* comm_status, int32_t* graph_status, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    if (!dataset || !ids || !dataset->x || !dataset->y) return DGCNN_ERR_INVALID_ARGUMENT;
    const StepArgs t{num_nodes, num_edges, num_graphs, dataset->num_features, k, num_classes, max_nodes, norm,
                     params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, global_batch, training,
                     seed, rng_offset, exchange, world, rank, epoch, comm_status, graph_status, workspace,
                     workspace_bytes, stream};
    DGCNN_TRY(check_step_args(t, true));
    const int64_t N = num_nodes, E = num_edges, B = num_graphs;
    const int32_t F = dataset->num_features;
    Arena a{reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255), 0};
    StepBuffers s = carve(a, N, E, B, F, k, num_classes, max_nodes, true);
    if (dataset->symmetric) {                          // one CSR serves both directions
        s.rowptr_t = s.rowptr;
        s.col_t = s.col;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DGCNN_LAUNCH(step_status_begin, 1, 1, 0, st, graph_status);
    DGCNN_RETURN_IF_LAUNCH_FAILED();
    // K0b's outputs are gathered too when the data set carries them and every graph of the batch
    // owns a bitmap (max_nodes <= 1024); otherwise K0b runs on the gathered CSR
    const bool maps = dataset->bitmap && dataset->fragmap && max_nodes <= 1024;
    dgcnn_batch_graph out{};
    out.x = s.x_batch; out.ldx = F; out.batch32 = s.batch32; out.y = s.y_batch;
    out.rowptr = s.rowptr; out.col = s.col; out.rowptr_t = s.rowptr_t; out.col_t = s.col_t;
    out.dis = s.dis; out.gptr = s.gptr; out.gorder = s.gorder;
    if (maps) {
        out.bitmap = s.bitmap; out.bitmap_t = s.bitmap_t; out.bmoff = s.bmoff; out.gflags = s.gflags;
        out.gflags_t = s.gflags_t; out.fragmap = s.fragmap; out.fgoff = s.fgoff; out.gdesc = s.gdesc;
    }
    DGCNN_TRY(dgcnn_collate(dataset, ids, B, N, E, &out, graph_status, s.ws_build, s.n_build, stream));
    return step_after_build(t, s, s.x_batch, F, nullptr, s.batch32, s.y_batch, maps);
}
#undef DGCNN_TRY

// The resident step replayed as a CUDA graph (the product loop, driver.train_epoch): every call
// stream-captures the step -- host work only -- and UPDATES the executable graph of the previous
// call in place (cudaGraphExecUpdate: same topology, new kernel arguments and grids), then launches
// it.  The device runs the step without the gaps of 16 eagerly launched kernels and their
// side-stream events (233 us -> the captured step's time); a change of topology (conv5 fusion on /
// off, more than 1024 graphs) re-instantiates.  The first call on a device runs eagerly: it
// creates what must not be created under capture (side stream, function attributes, probes).
namespace {
struct StepGraph { cudaGraphExec_t exec = nullptr; bool warmed = false; };
StepGraph g_step_graph[64];
int64_t g_step_graph_counts[4] = {0, 0, 0, 0};        // updated in place, instantiated, eager, capture failed
}  // namespace

// debug: how the graph-replayed steps of this process were launched
// {updated in place, newly instantiated, run eagerly, capture failed}
extern "C" void dgcnn_train_step_graph_counts(int64_t* out4) {
    for (int i = 0; i < 4; ++i) out4[i] = g_step_graph_counts[i];
}

extern "C" int dgcnn_train_step_resident_graphed(const dgcnn_dataset* dataset, const int32_t* ids,
                                                 int64_t num_nodes, int64_t num_edges, int64_t num_graphs,
                                                 int32_t k, int32_t num_classes, int64_t max_nodes, int32_t norm,
                                                 float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                                 int64_t* step, float lr, float beta1, float beta2, float eps,
                                                 int64_t global_batch, int32_t training, uint64_t seed,
                                                 int64_t* rng_offset, void* const* exchange, int32_t world,
                                                 int32_t rank, int64_t* epoch, int32_t* comm_status,
                                                 int32_t* graph_status, void* workspace, size_t workspace_bytes,
                                                 void* stream) {
    auto plain = [&]() {
        return dgcnn_train_step_resident(dataset, ids, num_nodes, num_edges, num_graphs, k, num_classes, max_nodes,
                                         norm, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps,
                                         global_batch, training, seed, rng_offset, exchange, world, rank, epoch,
                                         comm_status, graph_status, workspace, workspace_bytes, stream);
    };
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0;
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || st == nullptr ||
        cudaStreamIsCapturing(st, &capturing) != cudaSuccess || capturing != cudaStreamCaptureStatusNone) {
        ++g_step_graph_counts[2];
        return plain();                                   // (legacy stream / already inside a capture)
    }
    StepGraph& sg = g_step_graph[dev];
    if (!sg.warmed) {
        sg.warmed = true;
        ++g_step_graph_counts[2];
        return plain();
    }
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        ++g_step_graph_counts[3];
        return plain();
    }
    const int rc = plain();
    cudaGraph_t graph = nullptr;
    const cudaError_t ended = cudaStreamEndCapture(st, &graph);
    if (rc != DGCNN_OK || ended != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        ++g_step_graph_counts[3];
        return rc != DGCNN_OK ? rc : DGCNN_ERR_CUDA;
    }
    bool updated = sg.exec != nullptr;
    if (sg.exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(sg.exec, graph, &info) != cudaSuccess) {
            cudaGetLastError();                           // topology changed: build a new executable
            cudaGraphExecDestroy(sg.exec);
            sg.exec = nullptr;
            updated = false;
        }
    }
    if (!sg.exec && cudaGraphInstantiate(&sg.exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        sg.exec = nullptr;
        cudaGraphDestroy(graph);
        return plain();
    }
    cudaGraphDestroy(graph);
    ++g_step_graph_counts[updated ? 0 : 1];
    if (cudaGraphLaunch(sg.exec, st) != cudaSuccess) return DGCNN_ERR_CUDA;
    return DGCNN_OK;
}

