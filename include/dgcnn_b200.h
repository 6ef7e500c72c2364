/*
 * dgcnn_b200 -- C ABI of the B200-native DGCNN hot path.
 *
 * The reference (leftthomas/DGCNN) has no FFI layer: its hot path is the PyG
 * Module API that model.py consumes.  Each entry point below replaces the
 * arithmetic behind one of those call sites (cited per function); the Python
 * mirror of the Module API (dgcnn_b200/nn.py) binds them with ctypes, exactly as
 * INTEGRATION.md shows for a maintainer of the reference.
 *
 * Rules that hold for EVERY function:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the
 *     current CUDA device unless the name says `host`;
 *   - returns DGCNN_OK (0) or a negative dgcnn_status; never throws, never
 *     allocates, never synchronises, launches only on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - the caller owns every buffer; scratch space is sized by the matching
 *     *_workspace_bytes() query (pure host arithmetic, no CUDA call);
 *   - safe under CUDA-graph stream capture; re-entrant and thread-safe;
 *   - features are fp32 row-major with an explicit leading dimension (`ld*`,
 *     in floats) so a layer can read/write its column slice of the
 *     concatenated [N,97] buffer in place (model.py:34 needs no copy);
 *   - indices are int64 at the boundary the reference defines
 *     (edge_index, batch) and int32 inside (CSR, gptr, perm).
 */
#ifndef DGCNN_B200_H_
#define DGCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGCNN_B200_ABI_VERSION 2

typedef enum dgcnn_status {
    DGCNN_OK = 0,
    DGCNN_ERR_INVALID_ARGUMENT = -1, /* null pointer, negative size, bad enum            */
    DGCNN_ERR_UNSUPPORTED = -2,      /* size outside what the kernels cover (see each fn) */
    DGCNN_ERR_WORKSPACE = -3,        /* workspace smaller than *_workspace_bytes()        */
    DGCNN_ERR_CUDA = -4              /* a launch failed; cudaGetLastError was consumed    */
} dgcnn_status;

/* GCNConv normalisation.  SYM is what the reference computes (PyG GCNConv
 * default, model.py:13-16); RW is the AAAI-18 paper's D^-1 (A+I). */
typedef enum dgcnn_norm { DGCNN_NORM_SYM = 0, DGCNN_NORM_RW = 1 } dgcnn_norm;

/* Activation fused behind the convolution: NONE for a bare GCNConv call,
 * TANH for model.py:30-33's torch.tanh(convN(...)). */
typedef enum dgcnn_act { DGCNN_ACT_NONE = 0, DGCNN_ACT_TANH = 1 } dgcnn_act;

/* Bits OR-ed by the kernels into the optional device word `status`
 * (graph validation happens on the device, without a host sync). */
#define DGCNN_GRAPH_BAD_EDGE 1    /* an edge_index entry outside [0, N)      */
#define DGCNN_GRAPH_BAD_BATCH 2   /* batch not non-decreasing / outside [0,B) */
#define DGCNN_GRAPH_RANGE 4       /* a projected feature exceeded the fp16 split range */
#define DGCNN_COMM_TIMEOUT 16     /* dgcnn_allreduce_adam: a peer did not arrive within ~20 s */
#define DGCNN_GRAPH_GENERIC 8     /* (informational) input was not a sorted symmetric edge list:
                                     K0 took the generic path, A_hat^T != A_hat may hold */

/* Implementations of the fused forward (dgcnn_stack_fwd `variant`). */
#define DGCNN_STACK_MMA 0         /* tensor-core aggregation + projection (default)    */
#define DGCNN_STACK_FMA 1         /* fp32 FMA gather through shared memory            */

int dgcnn_abi_version(void);
const char* dgcnn_status_string(int status);

/* ------------------------------------------------------------------------
 * K0  graph build.  Replaces, once per batch instead of once per layer:
 *   model.py:28       remove_self_loops(edge_index)
 *   model.py:30-33 -> GCNConv.forward -> gcn_norm (PyG nn/conv/gcn_conv.py):
 *                     add_remaining_self_loops, in-degree scatter, deg^-1/2
 *   model.py:35    -> SortAggregation -> to_dense_batch's per-graph offsets
 *
 * edge_index [2,E] int64 (row 0 = source j, row 1 = target i), any order,
 * self loops and duplicates allowed (loops dropped, duplicates kept: PyG counts
 * multi-edges in both the degree and the sum).  batch [N] int64 non-decreasing.
 *
 * Outputs:
 *   rowptr  [N+1], col  [E]  CSR by TARGET: col = sources of row i, ascending
 *   rowptr_t[N+1], col_t[E]  CSR by SOURCE (the transpose, for backward);
 *                            pass both NULL to skip
 *   dis     [N]   (1 + in_degree)^-1/2     (the implicit self loop is the +1)
 *   gptr    [B+1] node offset of every graph (empty graphs allowed)
 *   gorder  [B]   optional: graph ids by descending size, the order in which the
 *                 per-graph kernels (KS, KSB) drain their work queue
 *   status  optional device int32, OR-ed with DGCNN_GRAPH_* on bad input
 * Only the first rowptr[N] entries of col are meaningful.
 * Fast path: a strictly (src,dst)-sorted, loop-free, SYMMETRIC list (what TUDataset/PyG
 * batches are) is converted in one streaming pass; symmetry is established by two 64-bit
 * multiset fingerprints of {(s,d)} vs {(d,s)} (false accept ~2^-64), or, with
 * exact_verify != 0, by one binary search per edge (slower, exact).  Anything else takes
 * the generic count/scan/fill/sort pipeline; the choice is made on the device.
 * Limits: N, E < 2^31.
 * ------------------------------------------------------------------------ */
size_t dgcnn_build_graph_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int dgcnn_build_graph(const int64_t* edge_index, int64_t num_edges,
                      const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                      int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t,
                      float* dis, int32_t* gptr, int32_t* gorder, int32_t* status,
                      int32_t exact_verify, void* workspace, size_t workspace_bytes, void* stream);
/* The same for COMPACT host batches: edge_index [2,E] and batch [N] as int32 (a loader that
 * collates int32 indices halves the host-to-device copy of train.py:36, which is what bounds
 * the end-to-end rate: 16 of the 16.2 bytes per edge are indices). */
int dgcnn_build_graph_i32(const int32_t* edge_index, int64_t num_edges,
                          const int32_t* batch, int64_t num_nodes, int64_t num_graphs,
                          int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t,
                          float* dis, int32_t* gptr, int32_t* gorder, int32_t* status,
                          int32_t exact_verify, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * K0b  per-graph adjacency bitmaps (with GCNConv's self loop) for the fused per-graph
 * kernels KS / KSB, built once per batch from the CSRs of K0: `bitmap` from rowptr/col and,
 * when bitmap_t is given, `bitmap_t` (A_hat^T) from rowptr_t/col_t -- that half is skipped ON
 * THE DEVICE when K0 proved the batch symmetric (gate_word/gate_mask: pass K0's status word
 * and DGCNN_GRAPH_GENERIC; a skipped bitmap_t stays all-zero).  Both share bmoff.
 * batch / batch32 (optional, at most one: the reference's int64 [N] vector or its int32 form)
 * saves a search per node row.
 *   bitmap  uint32[dgcnn_graph_bitmap_words(N, B, max_nodes)]; graph g owns
 *           np_g * ceil(np_g/32) words at bmoff[g], np_g = n_g rounded up to 16
 *   bmoff   int32[B+1];  gflags int32[B]: bit 0 duplicate edges, bit 1 no bitmap (too large)
 * max_nodes: largest graph in the batch (host knowledge); capped at 1024.
 * ------------------------------------------------------------------------ */
int64_t dgcnn_graph_bitmap_words(int64_t num_nodes, int64_t num_graphs, int64_t max_nodes);
/* Optional second copy in FRAGMENT-MAJOR order for the tensor-core kernels (fragmap may be
 * NULL): graph g with T = np_g/16 row tiles and G = ceil(T/4) column groups owns T*G*32
 * words at fgoff[g]; word (mt, grp, lane) packs the lane's mma.m16n8k16 A-fragment bits of
 * the four 16x16 blocks kt = 4 grp + q (layout: dgcnn_b200/csrc/graph_bitmap.cu).
 *   fragmap uint32[dgcnn_graph_fragmap_words(N, B, max_nodes)];  fgoff int32[B+1]
 * gdesc (optional, needs fragmap; 16-byte aligned int32[4*B]): work descriptors
 *   {graph, first node, nodes, fgoff[graph]} in the order of gorder (K0: descending size;
 *   NULL = natural order) -- what the tensor-core KS kernel reads to deal graphs to SMs. */
int64_t dgcnn_graph_fragmap_words(int64_t num_nodes, int64_t num_graphs, int64_t max_nodes);
int dgcnn_build_bitmaps(const int32_t* rowptr, const int32_t* col,
                        const int32_t* rowptr_t, const int32_t* col_t,
                        const int32_t* gptr, const int64_t* batch, const int32_t* batch32,
                        int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                        uint32_t* bitmap, uint32_t* bitmap_t, int64_t bitmap_words,
                        int32_t* bmoff, int32_t* gflags, int32_t* gflags_t,
                        uint32_t* fragmap, int64_t fragmap_words, int32_t* fgoff,
                        const int32_t* gorder, int32_t* gdesc,
                        const int32_t* gate_word, int32_t gate_mask, void* stream);

/* gptr alone (SortAggregation called without a graph: model.py:35's `batch`). */
int dgcnn_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                    int32_t* gptr, int32_t* status, void* stream);

/* ------------------------------------------------------------------------
 * K1  fused graph convolution, forward.  One launch replaces
 *   model.py:30-33  torch.tanh(self.convN(x, edge_index))
 * i.e. PyG GCNConv.forward (lin -> propagate: gather, scale, scatter-add ->
 * + bias) followed by tanh, and -- through (y, ldy) -- model.py:34's cat.
 *
 *   y[i, :] = act( r_i * sum_{j in in(i) U {i}} c_j * x[j, :] @ W^T + b )
 *   SYM: r = c = dis          RW: r = dis^2, c = 1
 *
 * weight [cout, cin] row-major (PyG's lin.weight), bias [cout] or NULL.
 * Limits: 1 <= cin, cout <= 128.
 * ------------------------------------------------------------------------ */
int dgcnn_graph_conv_fwd(const float* x, int64_t ldx, int32_t cin,
                         const int32_t* rowptr, const int32_t* col, const float* dis,
                         const float* weight, const float* bias,
                         float* y, int64_t ldy, int32_t cout,
                         int64_t num_nodes, int32_t norm, int32_t act, void* stream);

/* ------------------------------------------------------------------------
 * K3  graph convolution, backward (autograd of model.py:30-33, train.py:40).
 *   dpre = dy * (1 - y^2)   (act = TANH; y is the forward OUTPUT)
 *   db   = sum_i dpre[i]           dh = A_hat^T dpre  (walks rowptr_t/col_t)
 *   dw   = dh^T x                  dx = dh W
 * dx may be NULL (first layer).  accumulate_dx != 0 adds into dx instead of
 * overwriting it (lets a layer add its input gradient on top of the pooled
 * gradient already sitting in that slice of the [N,97] gradient buffer).
 * dw [cout,cin] and db [cout] (db may be NULL) are overwritten.
 * ------------------------------------------------------------------------ */
/* K1 with the batch's graph offsets (gptr [B+1], optionally gorder = graph ids by descending
 * size, max_nodes = largest graph or 0 when unknown): the same result as dgcnn_graph_conv_fwd.
 * For 32-channel, 16-byte aligned input rows every graph of up to 1728 nodes is aggregated by
 * ONE CTA from rows staged in shared memory (one coalesced read of the graph's rows instead of
 * 128 B per edge from L2); larger graphs are spread over the device by the row-parallel kernel. */
int dgcnn_graph_conv_fwd_graphs(const float* x, int64_t ldx, int32_t cin, const int32_t* rowptr,
                                const int32_t* col, const float* dis, const int32_t* gptr,
                                const int32_t* gorder, int64_t num_graphs, int64_t max_nodes,
                                const float* weight, const float* bias, float* y, int64_t ldy,
                                int32_t cout, int64_t num_nodes, int32_t norm, int32_t act,
                                void* stream);
/* lin of GCNConv.forward on its own: h[n][32] = x W^T for inputs wider than 32 channels
 * (project first, then aggregate 32-wide rows: dgcnn_graph_conv_fwd with weight == NULL, which
 * is accepted for cin == cout == 32 and 16-byte aligned rows). */
int dgcnn_project_rows(const float* x, int64_t ldx, int32_t cin, const float* weight, float* h,
                       int64_t num_nodes, void* stream);
size_t dgcnn_graph_conv_bwd_workspace_bytes(int64_t num_nodes, int32_t cin, int32_t cout);
int dgcnn_graph_conv_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy,
                         const float* x, int64_t ldx, int32_t cin,
                         const int32_t* rowptr_t, const int32_t* col_t, const float* dis,
                         const float* weight,
                         float* dx, int64_t lddx, int32_t accumulate_dx,
                         float* dw, float* db, int32_t cout,
                         int64_t num_nodes, int32_t norm, int32_t act,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * K2  SortPooling, forward.  Replaces model.py:35 self.sort_pool(x, batch)
 * (PyG SortAggregation.forward: to_dense_batch, sort on the last channel,
 * gather, truncate / pad to k rows, fill -> 0) without the dense [B,Nmax,D]
 * detour.
 *
 * Per graph g: order its rows by x[:, d-1] DESCENDING, ties by ascending node
 * index (torch.sort(stable=True)), -0.0 == +0.0, NaN first; copy the first
 * min(n_g, k) rows to out[g, r, :], zero the rest;
 * perm[g, r] = global node index of that row, or -1 for a padded row.
 * out [B, k*d] fp32, perm [B, k] int32.
 * max_nodes_hint: the largest graph in the batch if the caller knows it
 * (sizes the shared-memory sort buffer), <= 0 if unknown.
 * ------------------------------------------------------------------------ */
size_t dgcnn_sort_pool_workspace_bytes(int64_t num_nodes, int64_t num_graphs);
int dgcnn_sort_pool_fwd(const float* x, int64_t ldx, int32_t d,
                        const int32_t* gptr, int64_t num_nodes, int64_t num_graphs, int32_t k,
                        int64_t max_nodes_hint, float* out, int32_t* perm,
                        void* workspace, size_t workspace_bytes, void* stream);

/* K4  SortPooling, backward: dx[perm[g,r], :] = dout[g, r, :], every other row
 * of dx[0:N, 0:d] is zeroed (truncated nodes get no gradient from the pool). */
int dgcnn_sort_pool_bwd(const float* dout, const int32_t* perm, int64_t num_graphs,
                        int32_t k, int32_t d, float* dx, int64_t lddx, int64_t num_nodes,
                        void* stream);

/* ------------------------------------------------------------------------
 * KS  fused forward of the whole hot path, model.py:28-35, in ONE launch for the
 * model's fixed widths (F -> 32 -> 32 -> 32 -> 1, model.py:13-16):
 *   xcat [N,97] = cat(tanh(conv1..4))  and  (pooled [B,k*97], perm [B,k]) =
 *   SortAggregation(k)(xcat) -- same contracts as K1 and K2 above.
 * One CTA per graph; adjacency, features and sort keys stay in shared memory
 * (see dgcnn_b200/csrc/graph_stack_mma.cu and graph_stack.cu for the two variants).  Needs the size of the largest graph
 * (`max_nodes`, known on the host from the batch's ptr): graphs must fit the
 * shared-memory budget, which dgcnn_stack_fwd_supported() reports (1/0; host arithmetic
 * plus, once per process, an occupancy query: a graph that does not fit ONE CTA is split over
 * the two CTAs of a cluster when the device can co-schedule all pairs -- see
 * dgcnn_stack_fwd_configure -- which lifts the limit from 560 to 608 nodes at F <= 8).
 * When it returns 0 use K1 x 4 + K2.
 * w1 [32,F], w2/w3 [32,32], w4 [1,32] row-major; biases may be NULL.
 * ------------------------------------------------------------------------ */
int dgcnn_stack_fwd_supported(int32_t num_features, int64_t max_nodes);
/* Debug hook (not thread-safe, off by default): when set to a device buffer of
 * int64[num_graphs][16], the tensor-core KS kernel records clock64() at the end of each of
 * its phases per graph (slots 0-8) and (smid << 32 | team threads | n << 12) in slot 15.
 * Pass NULL to switch it off.  Used by scripts/trace_stack_fwd.py. */
void dgcnn_stack_fwd_set_trace(int64_t* device_buffer);
/* Tuning / test hook of the tensor-core KS kernel.  It is launched as thread-block clusters of
 * two CTAs and splits the largest graphs of a batch over a CTA pair (each CTA owns half of the
 * 16-row tiles; planes, keys and ranks are exchanged through distributed shared memory).
 * pairs: 1 clusters (default when the device can co-schedule all pairs), 0 plain launch,
 * -1 probe again (also re-reads DGCNN_KS_PAIRS / DGCNN_KS_SPLIT_PCT); split_pct > 0: split a
 * graph whose cost exceeds this percentage of one SM's fair share of the batch (default 80).
 * Graphs that do not fit one CTA's shared memory are split regardless of the threshold (the
 * *_supported functions answer for the current setting).  The setting also applies to the
 * backward kernel (dgcnn_stack_bwd / dgcnn_stack_bwd_conv5: each CTA of a pair takes half of the
 * row tiles, gradient rows cross through L2).  Forward results are bit-identical either way, the
 * backward's parameter gradients equal up to fp32 summation order (tests/test_gpu_headline.py). */
void dgcnn_stack_fwd_configure(int32_t pairs, int32_t split_pct);
size_t dgcnn_stack_fwd_workspace_bytes(void);
int dgcnn_stack_fwd(const float* x, int64_t ldx, int32_t num_features,
                    const int32_t* rowptr, const int32_t* col, const float* dis,
                    const int32_t* gptr, const int32_t* gorder,
                    const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                    const uint32_t* fragmap, const int32_t* fgoff, const int32_t* gdesc,
                    int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                    const float* w1, const float* b1, const float* w2, const float* b2,
                    const float* w3, const float* b3, const float* w4, const float* b4,
                    float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                    int32_t norm, int32_t variant, int32_t* status,
                    void* workspace, size_t workspace_bytes, void* stream);

/* KS + the head of the dense tail (model.py:36-38 view -> conv5 -> ReLU ->
 * MaxPool1d(2,2)) in the same launch.  conv5 = Conv1d(1,16,97,97) has kernel == stride == 97: it
 * is a per-row 97 -> 16 linear map, so it commutes with SortPooling's row gather.  The kernel
 * accumulates z = W5 x_cat[node] + b5 per NODE in the layer epilogues (tensor cores, hi/lo
 * split) and, once the order is known, emits for the k winners
 *     h1[g][c][j]  = max(relu(z[row 2j][c]), relu(z[row 2j+1][c]))     float [B,16,k/2]
 *     arg[g][c][j] = 0 / 1 the winning row, 2 when the maximum is not positive   uint8 [B,16,k/2]
 * (padding rows have z = b5) -- exactly what dgcnn_tail_fwd's first kernel computes from
 * `pooled`.  `pooled` may be NULL: the [B, k*97] SortPooling output is then never written
 * (-26 MB per COLLAB batch) and dgcnn_tail_fwd is called with pooled == NULL.
 * Tensor-core variant only; dgcnn_stack_fwd_conv5_supported: the graphs need 64 bytes more
 * shared memory per node than for dgcnn_stack_fwd (480 nodes in one CTA, 512 as a CTA pair). */
int dgcnn_stack_fwd_conv5_supported(int32_t num_features, int64_t max_nodes);
int dgcnn_stack_fwd_conv5(const float* x, int64_t ldx, int32_t num_features,
                          const int32_t* rowptr, const int32_t* col, const float* dis,
                          const int32_t* gptr, const int32_t* gorder,
                          const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                          const uint32_t* fragmap, const int32_t* fgoff, const int32_t* gdesc,
                          int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                          const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* w3, const float* b3, const float* w4, const float* b4,
                          const float* w5, const float* b5,
                          float* xcat, int64_t ldc, float* pooled, int32_t* perm, int32_t k,
                          float* h1, uint8_t* arg, int32_t norm, int32_t* status,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * KSB  fused backward of the hot path (autograd of model.py:28-35, train.py:40),
 * companion of dgcnn_stack_fwd: from the gradient of `pooled` to the gradients of
 * the eight GraphConv parameters, one thread team per graph, two launches (per-graph
 * partial gradient vectors + a reduction in graph order: deterministic, no float
 * atomics).  `variant` as for dgcnn_stack_fwd.
 *   grads: flat [dgcnn_stack_num_params(F)] in PyG parameter order
 *     conv1.lin.weight [32,F] | conv1.bias [32] | conv2.lin.weight [32,32] | conv2.bias
 *     | conv3.lin.weight | conv3.bias | conv4.lin.weight [1,32] | conv4.bias [1]
 * Walks the CSR by SOURCE (rowptr_t/col_t).  The gradient w.r.t. x is not produced
 * (use K3/K4 when it is needed).
 * ------------------------------------------------------------------------ */
int dgcnn_stack_bwd_supported(int32_t num_features, int64_t max_nodes);  /* 0 no, 1 MMA, 2 FMA only */
/* Debug hook like dgcnn_stack_fwd_set_trace, for the tensor-core KSB kernel: slots 0-12 =
 * clock64() at start / after phase 0 / after layer 4 / then (dpre, dh+dx, dW) of layers 3, 2, 1
 * / end; slot 15 = (smid << 32 | team threads | n << 12). */
void dgcnn_stack_bwd_set_trace(int64_t* device_buffer);
int64_t dgcnn_stack_num_params(int32_t num_features);
size_t dgcnn_stack_bwd_workspace_bytes(int32_t num_features, int64_t num_graphs,
                                       int64_t num_nodes);
int dgcnn_stack_bwd(const float* dpooled, const int32_t* perm, int32_t k,
                    const float* xcat, int64_t ldc, const float* x, int64_t ldx,
                    int32_t num_features, const int32_t* rowptr_t, const int32_t* col_t,
                    const float* dis, const int32_t* gptr, const int32_t* gorder,
                    const int32_t* gdesc, const uint32_t* fragmap,
                    const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                    const uint32_t* bitmap_t, const int32_t* bmoff_t, const int32_t* gflags_t,
                    int64_t num_nodes, int64_t num_graphs,
                    int64_t max_nodes, const float* w2, const float* w3, const float* w4,
                    int32_t norm, int32_t variant, float* grads, int32_t* status,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * KT  dense tail, model.py:36-43: view -> conv5 Conv1d(1,16,97,97) +
 * ReLU -> MaxPool1d(2,2) -> conv6 Conv1d(16,32,5,1) + ReLU -> flatten -> fc1 Linear(.,128) +
 * ReLU -> Dropout(0.5) -> fc2 Linear(128,C) -> log_softmax.  Parameter layouts are torch's
 * (conv5.weight [16,1,97], conv6.weight [32,16,5], Linear.weight [out,in]).
 * Saved for backward (caller-owned): h1 [B,16,k/2], arg uint8 [B,16,k/2], h2 [B,32*(k/2-4)],
 * h3 [B,128], keep uint8 [B,128].  Dropout decisions come from a counter hash of
 * (seed, *rng_offset, element); the forward increments *rng_offset (device int64) so that a
 * CUDA-graph replay draws a fresh mask.  training == 0: no dropout, rng_offset unused.
 * Limits: num_classes <= 32, k <= ~700 (shared-memory tiles).
 * ------------------------------------------------------------------------ */
size_t dgcnn_tail_workspace_bytes(int64_t num_graphs, int32_t k, int32_t num_classes);
/* pooled == NULL: h1 and arg are INPUTS (written by dgcnn_stack_fwd_conv5); conv5 is skipped. */
int dgcnn_tail_fwd(const float* pooled, int64_t num_graphs, int32_t k,
                   const float* w5, const float* b5, const float* w6, const float* b6,
                   const float* wf1, const float* bf1, const float* wf2, const float* bf2,
                   int32_t num_classes, int32_t training, uint64_t seed, int64_t* rng_offset,
                   float* h1, uint8_t* arg, float* h2, float* h3, uint8_t* keep, float* logp,
                   void* workspace, size_t workspace_bytes, void* stream);
/* autograd of the above (train.py:40): dlogp [B,C] -> dpooled [B,k*97] and the eight
 * parameter gradients (overwritten).  Deterministic: every reduction runs in a fixed order. */
int dgcnn_tail_bwd(const float* dlogp, const float* pooled, int64_t num_graphs, int32_t k,
                   const float* w5, const float* w6, const float* wf1, const float* wf2,
                   int32_t num_classes,
                   const float* h1, const uint8_t* arg, const float* h2, const float* h3,
                   const uint8_t* keep, const float* logp,
                   float* dpooled, float* dw5, float* db5, float* dw6, float* db6,
                   float* dwf1, float* dbf1, float* dwf2, float* dbf2, int32_t overlap,
                   void* workspace, size_t workspace_bytes, void* stream);
/* KSB fed with d(h1) instead of d(pooled): the backward of conv5 + ReLU +
 * MaxPool1d(2,2) (model.py:37-38) runs inside the fused graph backward.  Per graph:
 *   dz[node][c] = dh1[c][r/2] if pooled row r = rank(node) won its pair (arg), else 0
 *   d x_cat[node] = dz[node] W5     (per 32-column slice, on the tensor cores, where the layer's
 *                                    gradient tile is assembled: no [B, k*97] dpooled in HBM)
 *   dW5 = dz^T x_cat, db5 = column sums of dz over ALL k rows (padding rows included)
 * `grads` receives dgcnn_stack_conv5_num_params(F) floats: the eight GraphConv gradients in
 * dgcnn_stack_bwd's order followed by dW5 [16,97] and db5 [16] -- the next two tensors of the
 * model's flat parameter order, so the buffer is one contiguous slice of it.  Bit-reproducible
 * (static plan, ordered reductions).  Workspace: dgcnn_stack_bwd_workspace_bytes. */
int dgcnn_stack_bwd_conv5_supported(int32_t num_features, int64_t max_nodes);
int64_t dgcnn_stack_conv5_num_params(int32_t num_features);
int dgcnn_stack_bwd_conv5(const float* dh1, const uint8_t* arg, const int32_t* perm, int32_t k,
                          const float* xcat, int64_t ldc, const float* x, int64_t ldx,
                          int32_t num_features, const int32_t* rowptr_t, const int32_t* col_t,
                          const float* dis, const int32_t* gptr, const int32_t* gorder,
                          const int32_t* gdesc, const uint32_t* fragmap,
                          const uint32_t* bitmap, const int32_t* bmoff, const int32_t* gflags,
                          const uint32_t* bitmap_t, const int32_t* bmoff_t, const int32_t* gflags_t,
                          int64_t num_nodes, int64_t num_graphs, int64_t max_nodes,
                          const float* w2, const float* w3, const float* w4, const float* w5,
                          int32_t norm, float* grads, int32_t* status,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same backward stopped at d(h1) [B,16,k/2]: conv5's own backward
 * (d x_cat, dw5, db5) runs inside dgcnn_stack_bwd_conv5.  Six parameter gradients. */
int dgcnn_tail_bwd_h1(const float* dlogp, int64_t num_graphs, int32_t k, const float* w6,
                      const float* wf1, const float* wf2, int32_t num_classes,
                      const float* h1, const float* h2, const float* h3, const uint8_t* keep,
                      const float* logp, float* dh1, float* dw6, float* db6, float* dwf1,
                      float* dbf1, float* dwf2, float* dbf2, int32_t overlap,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The training step's variant of the tail: dgcnn_tail_fwd + dgcnn_nll_sum + the first kernel of
 * the backward in ONE call (model.py:36-43 + train.py:39).  fc1's epilogue, fc2, log_softmax, the
 * NLL of the labels `y`, d(sum of NLL)/d(logits) and fc2's row backward are row-local and run as
 * one kernel (one warp per graph) instead of four launches.  Outputs as dgcnn_tail_fwd; dlogit,
 * dz3 and the per-graph loss / hit scalars are left in `workspace_bwd`, which must then be handed
 * to dgcnn_tail_bwd_after_loss: that call skips fc2's row backward, sums the scalars in a fixed
 * order into stats = [sum of NLL, #correct] and advances *rng_offset (pass NULL when not training).
 * dh1 != NULL stops at d(h1) (pooled, arg, w5, dpooled, dw5, db5 may be NULL). */
int dgcnn_tail_fwd_loss(const float* pooled, int64_t num_graphs, int32_t k,
                        const float* w5, const float* b5, const float* w6, const float* b6,
                        const float* wf1, const float* bf1, const float* wf2, const float* bf2,
                        int32_t num_classes, const int64_t* y, int32_t training, uint64_t seed,
                        int64_t* rng_offset, float* h1, uint8_t* arg, float* h2, float* h3,
                        uint8_t* keep, float* logp, void* workspace_fwd, size_t workspace_fwd_bytes,
                        void* workspace_bwd, size_t workspace_bwd_bytes, void* stream);
int dgcnn_tail_bwd_after_loss(const float* pooled, int64_t num_graphs, int32_t k, const float* w5,
                              const float* w6, const float* wf1, const float* wf2,
                              int32_t num_classes, const float* h1, const uint8_t* arg,
                              const float* h2, const float* h3, const uint8_t* keep,
                              const float* logp, float* dpooled, float* dh1, float* dw5, float* db5,
                              float* dw6, float* db6, float* dwf1, float* dbf1, float* dwf2,
                              float* dbf2, float* stats, int64_t* rng_offset, int32_t overlap,
                              void* workspace, size_t workspace_bytes, void* stream);
/* `overlap`: 0 = everything on `stream`.  1 = the parameter-gradient chain (dW/db of fc2, fc1,
 * conv6, conv5) runs on a side stream owned by the library, concurrently with the
 * input-gradient chain that produces dpooled, and is joined into `stream` before the call
 * returns.  2 = as 1 but NOT joined: the caller goes on (e.g. with dgcnn_stack_bwd) and must call
 * dgcnn_tail_bwd_join(stream) before it reads the eight parameter gradients or releases any
 * buffer passed to dgcnn_tail_bwd.  Fork and join are event record/wait pairs: capturable in a
 * CUDA graph.  One side stream per device: with overlap != 0 the call is not re-entrant. */
int dgcnn_tail_bwd_join(void* stream);

/* Adam (train.py:41,99: torch.optim.Adam defaults) on flat fp32 buffers;
 * `step` is a device int64 counter, incremented by the call; gradients are multiplied by
 * grad_scale first (1/B_global after a summed all-reduce). */
int dgcnn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    int64_t n, int64_t* step, float lr, float beta1, float beta2, float eps,
                    float grad_scale, void* stream);

/* NLL on log-probabilities (train.py:39,44-45): stats[0] = -sum_b logp[b, y_b], stats[1] =
 * number of correct argmax predictions; dlogp (optional) [B,C] = d(stats[0] * grad_scale). */
int dgcnn_nll_sum(const float* logp, const int64_t* y, int64_t num_graphs, int32_t num_classes,
                  float grad_scale, float* stats, float* dlogp, void* stream);

/* ------------------------------------------------------------------------
 * X1 + N3  multi-GPU step tail: ONE kernel that all-reduces the flat gradient buffer over
 * NVLink peer memory (one-shot, push: every rank stores its sums into its slot of every rank's
 * exchange buffer -- mapped through CUDA IPC --, signals an arrival counter per rank, waits on
 * its own counter and adds its W local slots in rank order: bit-identical on every rank) and
 * applies the flat Adam update of dgcnn_adam_step.
 * Replaces, for ranks of one node, dist.all_reduce + optimizer.step (train.py:40-42 in the
 * data-parallel harness; the reference itself is single-device).
 *   grads      [n_total]: in = local sums, out = global sums; the first n_params entries
 *              drive Adam (x grad_scale), the rest (loss sum, #correct) just ride along
 *   exchange   HOST array of `world` DEVICE pointers, exchange[r] = rank r's buffer
 *              (dgcnn_exchange_create at [rank], dgcnn_exchange_open of the peers' handles
 *              elsewhere)
 *   step, epoch device int64 counters, bumped by the call; epoch must be equal on all ranks
 *   status     optional; DGCNN_COMM_TIMEOUT if a peer never arrived (no hang)
 * Every rank must make the same sequence of calls.  Capturable in a CUDA graph.
 * ------------------------------------------------------------------------ */
size_t dgcnn_allreduce_adam_exchange_bytes(int64_t n_total, int32_t world);
/* Set-up of the exchange buffers (these four allocate / map; call them once, outside the step):
 * create = cudaMalloc + zero + cudaIpcGetMemHandle on the current device (handle64: 64 bytes to
 * send to the peers); open = cudaIpcOpenMemHandle with the CALLER's device current, which maps
 * the exporter's buffer for kernels of the caller's device (peer access enabled lazily). */
int dgcnn_exchange_create(int64_t n_total, int32_t world, void** local_ptr, unsigned char* handle64);
int dgcnn_exchange_open(const unsigned char* handle64, void** peer_ptr);
int dgcnn_exchange_close(void* peer_ptr);
int dgcnn_exchange_destroy(void* local_ptr);
/* Debug hook: device buffer int64[1024][4]; the exchange kernel of step e records
 * %globaltimer (ns) at [e % 1024]: [0] entered, [1] sums pushed and arrival signalled, [2] all
 * ranks arrived ([2] - [1] = time spent waiting for the slowest rank), [3] sum + Adam done.
 * NULL switches it off.  Used by bench.py --trace-exchange (profiles/r02_exchange_trace_*.json). */
void dgcnn_allreduce_set_trace(int64_t* device_buffer);
int dgcnn_allreduce_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                         int64_t n_params, int64_t n_total, int64_t* step, int64_t* epoch,
                         float lr, float beta1, float beta2, float eps, float grad_scale,
                         void* const* exchange, int32_t world, int32_t rank, int32_t* status,
                         void* stream);

/* ------------------------------------------------------------------------
 * One training step of train.py:35-45 (model(data) -> NLL -> backward -> Adam) as ONE host
 * call: K0 -> K0b -> KS -> KT forward -> NLL -> KT backward (parameter gradients on the side
 * stream) -> KSB -> [dgcnn_allreduce_adam | dgcnn_adam_step], every intermediate buffer carved
 * from `workspace`.  Same kernels, same order and same results as the one-by-one sequence.
 *   edge_index/batch  int64 (index_is_i32 = 0, the reference's format) or int32
 *   params, grads, exp_avg, exp_avg_sq   flat fp32, dgcnn_train_step_num_params() entries in
 *       the order conv1.lin.weight [32,F], conv1.bias, conv2.*, conv3.*, conv4.lin.weight
 *       [1,32], conv4.bias, conv5.weight [16,1,97], conv5.bias, conv6.weight [32,16,5],
 *       conv6.bias, classifier_1.weight [128, 32(k/2-4)], .bias, classifier_2.weight [C,128],
 *       .bias; grads has two more entries: [sum of NLL, #correct] of this call's batch (summed
 *       over ranks when exchange != NULL)
 *   global_batch      the gradient is divided by it (graphs of all ranks)
 *   exchange/world/rank/epoch/comm_status   as for dgcnn_allreduce_adam; exchange NULL or
 *       world <= 1: plain Adam
 *   graph_status      device int32[2]: [0] = DGCNN_GRAPH_* flags of THIS step (reset at its start),
 *                     [1] = OR of the error flags (everything but DGCNN_GRAPH_GENERIC) of all
 *                     earlier steps since the caller last zeroed it: check [0] | [1] once per epoch
 * Needs graphs that fit the fused kernels (dgcnn_stack_fwd_supported / _bwd_supported), else
 * DGCNN_ERR_UNSUPPORTED.  No allocation, no synchronisation: capturable in a CUDA graph.
 * ------------------------------------------------------------------------ */
/* SURVEY.md 8f N2 switch of dgcnn_train_step / dgcnn_train_step_resident: 1 (default, also
 * DGCNN_FUSE_CONV5 unset) runs conv5 + ReLU + max-pool and their backward inside KS / KSB whenever
 * the batch fits (dgcnn_stack_fwd_conv5_supported / dgcnn_stack_bwd_conv5_supported); 0 keeps the
 * unfused sequence; -1 re-reads the environment. */
void dgcnn_train_step_configure(int32_t fuse_conv5);
/* Lazy adjacency maps of dgcnn_train_step (host-fed batches): 1 (default, also DGCNN_LAZY_MAPS unset)
 * lets the fused forward kernel expand every graph's CSR rows into its fragment map in shared memory
 * and export it for the backward kernel, so that K0b shrinks to the offsets / descriptors; 0 runs the
 * full K0b (dgcnn_build_bitmaps) first; -1 re-reads the environment.  Results are bit-identical. */
void dgcnn_train_step_configure_maps(int32_t lazy);
size_t dgcnn_train_step_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t num_graphs,
                                        int32_t num_features, int32_t k, int32_t num_classes,
                                        int64_t max_nodes);
int64_t dgcnn_train_step_num_params(int32_t num_features, int32_t k, int32_t num_classes);
int dgcnn_train_step(const float* x, int64_t ldx, const void* edge_index, int32_t index_is_i32,
                     const void* batch, const int64_t* y, int64_t num_nodes, int64_t num_edges,
                     int64_t num_graphs, int32_t num_features, int32_t k, int32_t num_classes,
                     int64_t max_nodes, int32_t norm, float* params, float* grads,
                     float* exp_avg, float* exp_avg_sq, int64_t* step, float lr, float beta1,
                     float beta2, float eps, int64_t global_batch, int32_t training,
                     uint64_t seed, int64_t* rng_offset, void* const* exchange, int32_t world,
                     int32_t rank, int64_t* epoch, int32_t* comm_status, int32_t* graph_status,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * N1  device-resident data set + collate on the GPU.  Replaces, per step,
 *   train.py:108-109  DataLoader -> PyG Batch.from_data_list (concatenate, shift edge ids by the
 *                     node offset, emit `batch`)        train.py:36  sample.to(device)
 *   model.py:28 + gcn_norm prologue (K0) for that batch
 * The data set is ONE batch of all its graphs that went through dgcnn_build_graph (and,
 * optionally, dgcnn_build_bitmaps with max_nodes = min(largest graph, 1024)) once -- so loops
 * are dropped, rows sorted, the symmetry verdict known -- and stays in HBM:
 *   gptr [G+1], rowptr [Nd+1], col [Ed], dis [Nd]   K0's outputs over the whole data set
 *   rowptr_t, col_t   K0's CSR by source; may be NULL when `symmetric` (K0 did not raise
 *                     DGCNN_GRAPH_GENERIC: both CSRs are equal)
 *   x [Nd, F] (row stride ldx), y [G] int64        features and labels (either may be NULL
 *                     when the matching output is not requested)
 *   bitmap, bmoff [G+1], gflags [G], fragmap, fgoff [G+1]   K0b's outputs over the whole data
 *                     set (NULL: the batch's maps cannot be gathered, run dgcnn_build_bitmaps on
 *                     the gathered CSR instead); bitmap_t, gflags_t [G] for a data set that is
 *                     not symmetric.  K0b's per-graph blocks are relative to the graph's first
 *                     node, so they can be copied as they are.
 * Edges must not cross graphs (true of any PyG data set; dgcnn_dataset_prepare checks it once).
 * The struct itself lives in HOST memory (its members are device pointers) and is only read
 * during the call.
 * ------------------------------------------------------------------------ */
typedef struct dgcnn_dataset {
    int64_t num_graphs, num_nodes, num_edges;   /* G, Nd, Ed = rowptr[Nd] */
    int32_t num_features;                       /* F */
    int32_t symmetric;                          /* 1: rowptr_t/col_t/bitmap_t/gflags_t unused */
    const float* x;
    int64_t ldx;
    const int64_t* y;
    const int32_t* gptr;
    const int32_t* rowptr;
    const int32_t* col;
    const int32_t* rowptr_t;
    const int32_t* col_t;
    const float* dis;
    const uint32_t* bitmap;
    const uint32_t* bitmap_t;
    const int32_t* bmoff;
    const int32_t* gflags;
    const int32_t* gflags_t;
    const uint32_t* fragmap;
    const int32_t* fgoff;
    const int32_t* gext;                        /* [G][4], 16-byte aligned: dgcnn_dataset_prepare */
} dgcnn_dataset;

/* Set-up, once per data set (one launch): validates everything dgcnn_collate then trusts --
 * gptr / rowptr monotone and closed, both CSRs agreeing on every graph's edge span, no edge
 * leaving its graph (status |= DGCNN_GRAPH_BAD_BATCH / DGCNN_GRAPH_BAD_EDGE otherwise) -- and
 * writes gext [G][4] = {first node, nodes, first edge, edges} per graph, which the caller then
 * stores in dataset->gext (the field is not read by this call). */
int dgcnn_dataset_prepare(const dgcnn_dataset* dataset, int32_t* gext, int32_t* status, void* stream);

/* Where a gathered batch goes (HOST struct of device pointers, caller-owned buffers): exactly the
 * outputs of dgcnn_build_graph(_i32) and dgcnn_build_bitmaps on the host-collated batch of the
 * same graphs, plus the collated features / graph ids / labels.
 *   x [N, F] (row stride ldx), batch32 [N], y [B]                 optional
 *   rowptr [N+1], col [E], dis [N], gptr [B+1]                    required;  gorder [B] optional
 *   rowptr_t / col_t   both NULL, or aliases of rowptr / col (fine for a symmetric data set:
 *                      one copy is written), or separate buffers
 *   bitmap .. gdesc    K0b's outputs, sized as for dgcnn_build_bitmaps with max_nodes =
 *                      min(largest graph of the batch, 1024); bitmap NULL skips all of them;
 *                      bitmap_t / gflags_t optional (only written -- bitmap_t -- for a data set
 *                      that is not symmetric); gdesc (16-byte aligned, needs gorder) optional.
 *                      Words past bmoff[B] / fgoff[B] are not touched. */
typedef struct dgcnn_batch_graph {
    float* x;
    int64_t ldx;
    int32_t* batch32;
    int64_t* y;
    int32_t* rowptr;
    int32_t* col;
    int32_t* rowptr_t;
    int32_t* col_t;
    float* dis;
    int32_t* gptr;
    int32_t* gorder;
    uint32_t* bitmap;
    uint32_t* bitmap_t;
    int32_t* bmoff;
    int32_t* gflags;
    int32_t* gflags_t;
    uint32_t* fragmap;
    int32_t* fgoff;
    int32_t* gdesc;
} dgcnn_batch_graph;

/* A batch = ids[0..B) (device int32, graph ids of the data set in batch order, repeats
 * allowed).  `status` is OR-ed with DGCNN_GRAPH_GENERIC when the data set is not symmetric
 * (what K0 would report), DGCNN_GRAPH_BAD_BATCH when an id is outside [0, G) or num_nodes /
 * num_edges (host arithmetic on the per-graph sizes, which size the outputs) disagree with
 * the ids -- then nothing is written.  The data set itself is trusted: check the status of
 * dgcnn_dataset_prepare once.  One launch for B <= 1024 (two beyond); B, N, E < 2^31. */
size_t dgcnn_collate_workspace_bytes(int64_t num_graphs);
int dgcnn_collate(const dgcnn_dataset* dataset, const int32_t* ids, int64_t num_graphs,
                  int64_t num_nodes, int64_t num_edges, const dgcnn_batch_graph* out,
                  int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* dgcnn_train_step with the batch taken from a resident data set: dgcnn_collate replaces the
 * host-to-device copy, K0 and (when the data set carries K0b's maps) K0b; everything after it
 * is the same call sequence, and the result is bit-identical to dgcnn_train_step on the
 * host-collated batch of the same graphs.
 * num_nodes, num_edges, max_nodes: sums / maximum of the per-graph sizes over ids (host
 * arithmetic; the caller keeps the sizes).  Other arguments as for dgcnn_train_step. */
size_t dgcnn_train_step_resident_workspace_bytes(int64_t num_nodes, int64_t num_edges,
                                                 int64_t num_graphs, int32_t num_features, int32_t k,
                                                 int32_t num_classes, int64_t max_nodes);
int dgcnn_train_step_resident(const dgcnn_dataset* dataset, const int32_t* ids, int64_t num_nodes,
                              int64_t num_edges, int64_t num_graphs, int32_t k, int32_t num_classes,
                              int64_t max_nodes, int32_t norm, float* params, float* grads,
                              float* exp_avg, float* exp_avg_sq, int64_t* step, float lr, float beta1,
                              float beta2, float eps, int64_t global_batch, int32_t training,
                              uint64_t seed, int64_t* rng_offset, void* const* exchange, int32_t world,
                              int32_t rank, int64_t* epoch, int32_t* comm_status, int32_t* graph_status,
                              void* workspace, size_t workspace_bytes, void* stream);
/* The same step replayed as a CUDA graph: each call captures the launch sequence (host work only),
 * updates the executable graph kept from the previous call in place (cudaGraphExecUpdate) and launches
 * it -- the product loop's way of running the step without eager-launch gaps (train.py:35-45 once per
 * batch; driver.train_epoch).  Same arguments, same results bit for bit; `stream` must be a real
 * (non-legacy) stream that is not being captured, else the call degrades to dgcnn_train_step_resident.
 * The first call per device runs eagerly. */
int dgcnn_train_step_resident_graphed(const dgcnn_dataset* dataset, const int32_t* ids, int64_t num_nodes,
                              int64_t num_edges, int64_t num_graphs, int32_t k, int32_t num_classes,
                              int64_t max_nodes, int32_t norm, float* params, float* grads,
                              float* exp_avg, float* exp_avg_sq, int64_t* step, float lr, float beta1,
                              float beta2, float eps, int64_t global_batch, int32_t training,
                              uint64_t seed, int64_t* rng_offset, void* const* exchange, int32_t world,
                              int32_t rank, int64_t* epoch, int32_t* comm_status, int32_t* graph_status,
                              void* workspace, size_t workspace_bytes, void* stream);
/* debug: {updated in place, newly instantiated, run eagerly, capture failed} calls of the above */
void dgcnn_train_step_graph_counts(int64_t* out4);

#ifdef __cplusplus
}
#endif
#endif /* DGCNN_B200_H_ */
