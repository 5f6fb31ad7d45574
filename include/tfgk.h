/*
 * tfgk.h - C ABI of the B200 (sm_100a) message-passing kernel backend for tf_geometric's hot path.
 *
 * The reference (CrawlScript/tf_geometric @4539f11) has NO native FFI: its "operator API" for this path is a set
 * of Python functions built on stock TensorFlow ops and the tf_sparse package.  Each entry point below therefore
 * cites the reference Python interface (file:line, relative to /root/reference/tf_geometric) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a tf_geometric maintainer would add at each of those sites.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator in this repo); the library never
 *     allocates or frees device memory: scratch space is queried (`*_workspace_bytes`) and passed in;
 *   - `stream` is a cudaStream_t cast to void* (NULL = default stream); all launches are asynchronous, except
 *     tfgk_csr_build which synchronises `stream` once to report out-of-range node ids;
 *   - outputs are fully overwritten; no global mutable state (calls on distinct streams are thread-safe);
 *   - return value 0 = TFGK_OK, otherwise an error code with a thread-local message in tfgk_last_error();
 *   - float data is IEEE fp32, node ids int32, edge offsets (rowptr) int64; one call handles < 2^31 edges;
 *   - "row" = aggregation target (destination), "col" = neighbour (source): nn/kernel/map_reduce.py:60-70.
 */
#ifndef TFGK_H_
#define TFGK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFGK_ABI_VERSION 5

enum tfgk_status {
    TFGK_OK = 0,
    TFGK_ERR_INVALID_ARGUMENT = 1,
    TFGK_ERR_CUDA = 2,
    TFGK_ERR_WORKSPACE = 3,
    TFGK_ERR_UNSUPPORTED = 4,
    TFGK_ERR_INDEX_OUT_OF_RANGE = 5
};

enum tfgk_reduce { TFGK_REDUCE_SUM = 0, TFGK_REDUCE_MEAN = 1, TFGK_REDUCE_MAX = 2 };
enum tfgk_act { TFGK_ACT_NONE = 0, TFGK_ACT_RELU = 1 };
enum tfgk_deg_power { TFGK_POW_INV_SQRT = 0, TFGK_POW_INV = 1 };
enum tfgk_heads_mode { TFGK_HEADS_SPLIT = 0, TFGK_HEADS_BROADCAST = 1, TFGK_HEADS_REDUCE = 2 };
enum tfgk_edge_flag { TFGK_FLAG_ALL = 0, TFGK_FLAG_UPPER = 1, TFGK_FLAG_MAPPED = 2 };
enum tfgk_bernoulli { TFGK_BERNOULLI_NONE = 0, TFGK_BERNOULLI_DROPOUT = 1, TFGK_BERNOULLI_KEEP = 2 };
enum tfgk_sample_padding { TFGK_SAMPLE_NO_PADDING = 0, TFGK_SAMPLE_PADDING = 1, TFGK_SAMPLE_HEAD = 2 };

int tfgk_version(void);
const char *tfgk_last_error(void);
/* SM count and compute capability of the current device (used by the host side to refuse non-sm_100 parts). */
int tfgk_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- integer edge preprocessing (bit-exact) ------------------------------------------------------------------ */

/* utils/graph_utils.py:350-366 add_self_loop_edge: out[2,E+N] = concat(edge_index[2,E], [[0..N-1],[0..N-1]]). */
int tfgk_self_loops_i32(const int32_t *edge_index, int64_t E, int32_t N, int32_t *out, void *stream);
/* same function, weight half: out[E+N] = concat(w (ones if NULL), fill). */
int tfgk_self_loop_weights_f32(const float *w, int64_t E, int32_t N, float fill, float *out, void *stream);
/* nn/kernel/segment.py:36-40 segment_count: int32 histogram of ids (negative ids dropped like
 * tf.math.unsorted_segment_sum; ids >= N are an error reported by tfgk_csr_build, ignored here). */
int tfgk_segment_count_i32(const int32_t *ids, int64_t E, int32_t N, int32_t *out, void *stream);

/* Destination-sorted CSR of a COO edge list: a STABLE sort by row, so that a left-to-right walk of a row
 * visits its edges in input order (= tf.math.unsorted_segment_sum's CPU summation order).
 *   rowptr[N+1] int64, col_sorted[E] = col[perm], perm[E] (position in the input list).
 * Replaces the implicit scatter of tf.math.unsorted_segment_* (nn/kernel/map_reduce.py:16,28,41) and of
 * tf_sparse.SparseMatrix.matmul (nn/conv/gcn.py:280, gat.py:89, appnp.py:86). */
int tfgk_csr_workspace_bytes(int64_t E, int32_t N, size_t *out_bytes);
int tfgk_csr_build(const int32_t *row, const int32_t *col, int64_t E, int32_t N_rows, int32_t N_cols,
                   int64_t *rowptr, int32_t *col_sorted, int32_t *perm,
                   void *workspace, size_t workspace_bytes, void *stream);

/* utils/graph_utils.py:67-125 merge_duplicated_edge, index half: duplicates of (row, col) collapse onto their FIRST occurrence
 * (tf.unique order on the hash num_nodes*row + col, num_nodes = N).  unique_index is [2, E] row-major with the first
 * *n_unique_host columns valid; unique_of_edge[e] = column of edge e's representative (the segment id the edge properties
 * are merged with).  Bit-exact; synchronises `stream`. */
int tfgk_edge_unique_workspace_bytes(int64_t E, int32_t N, size_t *out_bytes);
int tfgk_edge_unique(const int32_t *row, const int32_t *col, int64_t E, int32_t N, int32_t *unique_index,
                     int32_t *unique_of_edge, int32_t *n_unique_host, void *workspace, size_t workspace_bytes, void *stream);

/* utils/graph_utils.py:181-190 convert_edge_to_directed, index half: out[2, out_ld] = upper edges followed by the mirrored
 * non-self-loop upper edges (in order); lower_src[j] = column of the upper edge mirrored into column U + j (to copy its
 * properties).  upper_index is [2, ld] row-major with U valid columns.  Bit-exact; synchronises `stream`. */
int tfgk_directed_workspace_bytes(int64_t U, size_t *out_bytes);
int tfgk_directed_edges(const int32_t *upper_index, int64_t U, int64_t ld, int32_t *out, int64_t out_ld,
                        int32_t *lower_src, int32_t *n_lower_host, void *workspace, size_t workspace_bytes, void *stream);

/* Work plan of a CSR for the streaming kernels (K1, K3): destination rows are grouped into tasks of at most
 * `rows_per_task` consecutive rows (one warp each), and every row with more than `hub_threshold` edges is cut into slices
 * of `chunk` edges that are reduced by separate warps into `scratch` and merged by a fix-up kernel in slice order
 * (deterministic; a hub row is therefore summed slice by slice instead of strictly left to right).  Without hub rows the
 * kernels need no plan (pass NULL).  The reference has no counterpart: tf.math.unsorted_segment_sum is one sequential loop. */
typedef struct tfgk_plan {
    int32_t n_tasks, n_hubs, n_slots, chunk;
    const int32_t *task_row;    /* [n_tasks] first destination row                                   */
    const int32_t *task_nrows;  /* [n_tasks] number of rows (1 for a hub slice)                      */
    const int64_t *task_e0;     /* [n_tasks] first CSR edge                                          */
    const int64_t *task_e1;     /* [n_tasks] one past the last CSR edge                              */
    const int32_t *task_slot;   /* [n_tasks] -1: whole rows; >= 0: hub slice -> partial in scratch   */
    const int32_t *hub_row;     /* [n_hubs]                                                          */
    const int32_t *hub_slot0;   /* [n_hubs] first scratch slot of the row                            */
    const int32_t *hub_nslots;  /* [n_hubs]                                                          */
    float *scratch;             /* K1: n_slots * D floats;  K3: n_slots * (H*dv + 64) floats         */
    size_t scratch_bytes;
} tfgk_plan;

/* Upper bounds for the plan arrays of a CSR with E edges and N rows. */
int tfgk_plan_capacity(int64_t E, int32_t N, int32_t hub_threshold, int32_t chunk, int32_t rows_per_task,
                       int64_t *max_tasks, int64_t *max_hubs);
int tfgk_plan_workspace_bytes(int32_t N, size_t *out_bytes);
/* Fills the arrays; counts_host[3] = {n_tasks, n_hubs, n_slots} (synchronises `stream` once). */
int tfgk_plan_build(const int64_t *rowptr, int32_t N, int32_t hub_threshold, int32_t chunk, int32_t rows_per_task,
                    int32_t *task_row, int32_t *task_nrows, int64_t *task_e0, int64_t *task_e1, int32_t *task_slot,
                    int32_t *hub_row, int32_t *hub_slot0, int32_t *hub_nslots, int64_t cap_tasks, int64_t cap_hubs,
                    int32_t *counts_host, void *workspace, size_t workspace_bytes, void *stream);

/* dst[i*width + j] = src[perm[i]*width + j]   (COO order -> CSR order) and the inverse scatter. */
int tfgk_permute_f32(const float *src, const int32_t *perm, int64_t E, int32_t width, float *dst, void *stream);
int tfgk_unpermute_f32(const float *src, const int32_t *perm, int64_t E, int32_t width, float *dst, void *stream);

/* ---- GCN normalisation (nn/conv/gcn.py:32-130, utils/graph_utils.py:914-943) -------------------------------- */

/* SparseMatrix.segment_sum(axis=-1) on CSR-ordered values: out[r] = sum of w[rowptr[r]..rowptr[r+1]) in order. */
int tfgk_csr_rowsum_f32(const int64_t *rowptr, const float *w_csr, int32_t N, float *out, void *stream);
/* tf.pow(deg, -0.5 | -1) followed by _remove_inf_and_nan (gcn.py:23-29,81-82,104-105,114-115). */
int tfgk_deg_inv_f32(const float *deg, int32_t N, int power, float *out, void *stream);
/* (diags(dl) @ A) @ diags(dr) on the value array, COO order: out[e] = (dl[row[e]] * w[e]) * dr[col[e]];
 * dl or dr may be NULL (gcn.py:94,109,119). */
int tfgk_scale_edges_f32(const int32_t *row, const int32_t *col, const float *w, int64_t E,
                         const float *dl, const float *dr, float *out, void *stream);

/* ---- K1: gather - edge-apply - segment-reduce ---------------------------------------------------------------- */

/* out[r,:] = epilogue( REDUCE_{e in row r} ( w[e] * h[col[e],:] ) )            r in [0, n_dst)
 *   reduce = SUM  : tf.math.unsorted_segment_sum  (map_reduce.py:15-16), tf_sparse matmul (gcn.py:280)
 *            MEAN : tf.math.unsorted_segment_mean (map_reduce.py:27-28; graph_sage.py:41), empty row -> 0
 *            MAX  : tf.math.unsorted_segment_max  (map_reduce.py:38-42), empty row -> -FLT_MAX
 *   w NULL = identity_mapper (map_reduce.py:7-8), else gcn_mapper (gcn.py:221-222); CSR order.
 *   epilogue: v = agg*alpha + addend*beta (addend NULL -> v = agg; APPNP appnp.py:86-87, sum_updater
 *   map_reduce.py:19-20), then + bias[D] (NULL ok), then activation  (gcn.py:284-288).
 * Deterministic (no atomics); per-row accumulation is sequential in CSR order, except for the hub rows of `plan`
 * (NULL = none), which are accumulated slice by slice. */
int tfgk_spmm_f32(const int64_t *rowptr, const int32_t *col, const float *w,
                  const float *h, int64_t ldh, int32_t n_dst, int32_t D, int reduce,
                  float alpha, const float *addend, int64_t ld_addend, float beta,
                  const float *bias, int act,
                  float *out, int64_t ldo, const tfgk_plan *plan, void *stream);

/* ---- K3: edge softmax and fused GAT ------------------------------------------------------------------------- */

/* nn/kernel/segment.py:26-33 segment_softmax over CSR segments, H interleaved score columns:
 * score/out are [E, H] row-major in CSR order; per (segment, h): exp(s - max) / (sum + 1e-8). */
int tfgk_segment_softmax_f32(const int64_t *rowptr, const float *score, int32_t n_seg, int32_t H,
                             float *out, void *stream);

/* nn/conv/gat.py:73-114 fused: per destination r and head h
 *     s_e = <Q[r,h,:], K[col_e,h,:]> / scale  (scale = sqrt(dqk), gat.py:78-79) ;  a_e = softmax_e(s_e)  (gat.py:83-84,
 *     segment.py:26-33) ;  out[r,h,:] = sum_e a_e V[col_e,h,:]  (gat.py:87-89)
 * Q,K: [N, H*dqk]; V: [N, H*dv]. split_value_heads=1: out[N, H*dv] heads concatenated (gat.py:112);
 * 0: out[N, dv] = mean over heads (gat.py:114).  Then + bias, activation (gat.py:116-120).
 * att: [E, H] CSR order, REQUIRED scratch (raw scores, then exp(s - max)); with write_att != 0 it holds the
 * attention coefficients a_e on return.  The CSR must already contain the self loops gat.py:43 appends. */
int tfgk_gat_fused_f32(const int64_t *rowptr, const int32_t *col,
                       const float *Q, int64_t ldq, const float *K, int64_t ldk, const float *V, int64_t ldv,
                       int32_t N, int32_t H, int32_t dqk, int32_t dv, float scale, int split_value_heads,
                       const float *bias, int act, float *att, int write_att, float *out, int64_t ldo,
                       const tfgk_plan *plan, void *stream);

/* ---- K4: dense projections (gcn.py:272, gat.py:52,61,70, graph_sage.py:43-44, appnp.py:69) ------------------- */

/* C[M,N] = act( op(A) @ op(B) + bias[N] + beta * C ),  op = transpose when the flag is set (backward passes).
 * fp32 in/out. workspace is needed only for split-K (tall-skinny reductions); NULL/0 disables split-K. */
int tfgk_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K, size_t *out_bytes);
int tfgk_gemm_f32(const float *A, int64_t lda, int transA, const float *B, int64_t ldb, int transB,
                  const float *bias, int act, float beta, int32_t M, int32_t N, int32_t K,
                  float *C, int64_t ldc, void *workspace, size_t workspace_bytes, void *stream);

/* Several projections of the same input in ONE launch (round 2): for every column block b < n_blocks
 *   C_b[M, ncols_b] = act_b( A[M, K] @ B_b[K, ncols_b] + bias_b ),   ncols_b <= 128, n_blocks <= 4,
 * e.g. the three projections of gat.py:52,61,70 (Q | K | V) or gcn.py:272 next to them; A is read from HBM once.
 * Same 3xTF32 tcgen05 arithmetic as tfgk_gemm_f32's tensor-core path (bit-identical results).
 * A may live in n_parts <= 8 row blocks of part_rows rows each (a multiple of 128 when n_parts > 1), part i at
 * A_parts[i]: with the other ranks' buffers mapped through tfgk_peer_open this is the fused all-gather -> GEMM of the
 * partitioned path (rows are pulled over NVLink tile by tile while earlier tiles are multiplied); the walk starts at
 * part first_part so that concurrent ranks read from different peers.  max_ctas > 0 bounds the grid (to share the GPU
 * with a kernel on another stream).  Returns TFGK_ERR_UNSUPPORTED when the shape does not qualify (K > 512, unaligned A,
 * W too large for shared memory): the caller then uses tfgk_gemm_f32 per block. */
typedef struct tfgk_proj_block {
    const float *B; int64_t ldb;       /* [K, ncols] row-major weights ([ncols, K] row-major when transB != 0) */
    int32_t ncols;
    int32_t transB;                    /* 0: C = A @ B;  1: C = A @ B^T  (the dX = dY W^T products of the backward pass) */
    const float *bias;                 /* [ncols] or NULL */
    int act;                           /* tfgk_act */
    float *C; int64_t ldc;             /* [M, ncols] output (may be a column slice of a wider buffer) */
} tfgk_proj_block;
int tfgk_gemm_proj_f32(const float *const *A_parts, int32_t n_parts, int64_t part_rows, int64_t lda,
                       int32_t M, int32_t K, const tfgk_proj_block *blocks, int32_t n_blocks,
                       int32_t first_part, int32_t max_ctas, void *stream);

/* ---- K5: peer memory for the partitioned path ---------------------------------------------------
 * The ONE exception to "the library never allocates": buffers that other ranks on the same node read over NVLink
 * must come from cudaMalloc so that they can be exported with CUDA IPC (PyTorch's caching allocator sub-allocates).
 *   tfgk_peer_alloc / _free    device buffer on the current device (zero-initialised)
 *   tfgk_peer_export           64-byte IPC handle of a buffer from tfgk_peer_alloc (sent to the other ranks by the host)
 *   tfgk_peer_open / _close    map another process's buffer into this process (peer access enabled lazily)
 *   tfgk_peer_barrier          device-side barrier over NVLink on `stream`: thread j stores `value` into slot `rank` of
 *                              rank j's flag array (flags[j], uint32[world], from tfgk_peer_open) after a system-scope
 *                              fence, then waits until slot j of the local array reaches `value` (values grow by one per
 *                              barrier).  Traps instead of hanging if a peer does not arrive within timeout_ms.
 * The reference has no counterpart (its multi-GPU demos replicate the graph, demo/demo_distributed_gcn.py:37-57). */
#define TFGK_PEER_HANDLE_BYTES 64
int tfgk_peer_alloc(size_t bytes, void **ptr);
int tfgk_peer_free(void *ptr);
int tfgk_peer_export(void *ptr, void *handle_out);
int tfgk_peer_open(const void *handle, void **ptr);
int tfgk_peer_close(void *ptr);
/* Copies `bytes` (a multiple of 16, both pointers 16-byte aligned) from a peer-mapped buffer into local memory.
 * max_ctas > 0: copy kernel with that many CTAs (wide contiguous loads; ~7.5 GB/s per CTA, 663 GB/s from 148 CTAs on);
 * max_ctas = 0: the kernel on 148 CTAs;  max_ctas < 0: the copy engine (cudaMemcpyAsync on `stream`, 741 GB/s, no SMs). */
int tfgk_peer_pull(const void *src, void *dst, int64_t bytes, int32_t max_ctas, void *stream);
int tfgk_peer_barrier(uint32_t *const *flags, int32_t rank, int32_t world, uint32_t value, int32_t timeout_ms, void *stream);

/* out[c] = sum_r x[r, c]: the bias gradients db = 1^T dY of the backward passes (TensorFlow autodiff's BiasAddGrad under
 * gcn.py:283-284, gat.py:116-117, graph_sage.py:51-52).  Deterministic: per-block partial sums added in block order. */
int tfgk_colsum_workspace_bytes(int64_t n_rows, int32_t D, size_t *out_bytes);
int tfgk_colsum_f32(const float *x, int64_t ldx, int64_t n_rows, int32_t D, float *out, void *workspace,
                    size_t workspace_bytes, void *stream);

/* tf.nn.l2_normalize(x, axis=-1) (graph_sage.py:57-58): out = x * rsqrt(max(sum(x^2), 1e-12)). */
int tfgk_l2_normalize_f32(const float *x, int64_t ldx, int32_t N, int32_t D, float *out, int64_t ldo, void *stream);

/* ---- training-mode extras -------------------------------------------------------------------
 * Randomness is counter-based (Philox4x32-10): draw i of (seed, rng_stream) is a pure function of its arguments,
 * u_i = (philox(counter = (i >> 2, rng_stream, 0), key = seed)[i & 3] >> 8) * 2^-24.  Parity with TensorFlow's own
 * generator is statistical only; parity with oracle/tfg_oracle.py (same generator) is bit-exact. */

/* tf.nn.dropout (gcn.py:262 via tf_sparse dropout on the adjacency values, gat.py:85 on the attention coefficients):
 * out[i] = u_i >= rate ? x[i] * (1 / (1 - rate)) : 0.   x NULL = ones (a scaled keep mask). */
int tfgk_dropout_f32(const float *x, int64_t n, float rate, uint64_t seed, uint32_t rng_stream, float *out, void *stream);

/* Aggregation with one weight per (edge, head): the value half of gat.py:87-114 when the coefficients are already
 * known (attention dropout), and the three scatter-shaped gradients of the fused GAT kernel when run on the
 * transposed CSR.  weight(e, h) = w[pos*H + h] * dropout(pos*H + h), pos = emap ? emap[e] : e.
 *   TFGK_HEADS_SPLIT     out[r, h*dh+u] = alpha * sum_e weight(e,h) * src[col_e, h*dh+u]
 *   TFGK_HEADS_BROADCAST out[r, h*dh+u] = alpha * sum_e weight(e,h) * src[col_e, u]            (src has dh columns)
 *   TFGK_HEADS_REDUCE    out[r, u]      = alpha * sum_h sum_e weight(e,h) * src[col_e, h*dh+u]  (out has dh columns)
 * then + bias, activation.  Sequential in CSR order per row (deterministic). */
int tfgk_spmm_heads_f32(const int64_t *rowptr, const int32_t *col, const int32_t *emap, const float *w,
                        const float *src, int64_t lds, int32_t n_dst, int32_t H, int32_t dh, int mode,
                        float drop_rate, uint64_t seed, uint32_t rng_stream, float alpha,
                        const float *bias, int act, float *out, int64_t ldo, void *stream);

/* Gradient of gat.py:83-114 w.r.t. the scaled scores, given G = dL/d(aggregated rows before bias/activation):
 *   da_e,h = <G[r,h,:], V[col_e,h,:]> * dropout(e*H+h)   (split_value_heads=0: <G[r,:], V[col_e,h,:]> / H)
 *   ds_e,h = a_e,h * (da_e,h - sum_f a_f,h da_f,h)
 * att: [E, H] softmax coefficients BEFORE dropout, CSR order (tfgk_gat_fused_f32 with write_att); ds: [E, H] out.
 * The 1e-8 in segment.py:31 makes d a/d max non-zero by ~1e-8 relative in TF autodiff; that term is dropped. */
int tfgk_gat_softmax_bwd_f32(const int64_t *rowptr, const int32_t *col, const float *att,
                             const float *G, int64_t ldg, const float *V, int64_t ldv,
                             int32_t n_dst, int32_t H, int32_t dv, int split_value_heads,
                             float drop_rate, uint64_t seed, uint32_t rng_stream, float *ds, void *stream);

/* Training without the [E, H] coefficient table (round 2).  The forward pass is the same streaming kernel as
 * tfgk_gat_fused_f32 (heads concatenated, dqk == dv, H * dqk <= 128) and additionally keeps stats[N, 2H]: per (row, head)
 * the softmax maximum and the denominator (+1e-8).  The backward pass recomputes every coefficient from Q, K and stats:
 *   _prepare: GS[N, H*dqk + 32] = [ G * act'(Y) | max (8) | denominator (8) | delta = <G, Y - bias> per head (8) | pad (8) ]
 *   _dst    : dQ[r] = (1/scale) sum_{e in row r} ds_e K[col_e]                     forward CSR
 *   _src    : dK[c] = (1/scale) sum ds_e Q[row_e],  dV[c] = sum a_e G[row_e]       transposed CSR (rows = sources)
 * with a_e = exp(<Q_r, K_c>/scale - max_r) / den_r and ds_e = a_e (<G_r, V_c> - delta_r) per head.  Replaces TensorFlow
 * autodiff over gat.py:73-114.  TFGK_ERR_UNSUPPORTED for H > 8, non power-of-two head sizes or unaligned operands. */
int tfgk_gat_fused_stats_f32(const int64_t *rowptr, const int32_t *col,
                             const float *Q, int64_t ldq, const float *K, int64_t ldk, const float *V, int64_t ldv,
                             int32_t N, int32_t H, int32_t dqk, int32_t dv, float scale,
                             const float *bias, int act, float *out, int64_t ldo, float *stats,
                             const tfgk_plan *plan, void *stream);
int tfgk_gat_bwd_prepare_f32(const float *G, int64_t ldg, const float *Y, int64_t ldy, const float *bias, int act,
                             const float *stats, int32_t N, int32_t H, int32_t dqk, float *GS, int64_t ldgs, void *stream);
int tfgk_gat_bwd_dst_f32(const int64_t *rowptr, const int32_t *col, const float *Q, int64_t ldq,
                         const float *K, int64_t ldk, const float *V, int64_t ldv, const float *GS, int64_t ldgs,
                         int32_t N, int32_t H, int32_t dqk, float scale, float *dQ, int64_t lddq, void *stream);
int tfgk_gat_bwd_src_f32(const int64_t *rowptr_t, const int32_t *col_t, const float *Q, int64_t ldq,
                         const float *K, int64_t ldk, const float *V, int64_t ldv, const float *GS, int64_t ldgs,
                         int32_t N, int32_t H, int32_t dqk, float scale, float *dK, int64_t lddk, float *dV, int64_t lddv,
                         void *stream);

/* ---- device-side edge sampling -------------------------------------------------------------- */

/* flag[e] = structural(e) && bernoulli(e):
 *   structural: ALL | UPPER row<col (drop_edge.py:35 force_undirected) | MAPPED row_map[row]>=0 && col_map[col]>=0
 *               (graph_utils.py:826-832 virtual node filter)
 *   bernoulli : NONE | DROPOUT u_e >= prob (drop_edge.py:36,41: tf.nn.dropout(ones, rate) > 0) | KEEP u_e <= prob
 *               (graph_utils.py:808-809,841-842 UniformNeighborSampler) */
int tfgk_edge_flags_i32(const int32_t *row, const int32_t *col, int64_t E, int mode,
                        const int32_t *row_map, const int32_t *col_map,
                        int bernoulli, float prob, uint64_t seed, uint32_t rng_stream, int32_t *flag, void *stream);

/* tf.boolean_mask(tf.range(n), flag): ascending positions of the non-zero flags; *n_out_host = how many.
 * Synchronises the stream. */
int tfgk_select_workspace_bytes(int64_t n, size_t *out_bytes);
int tfgk_select_flagged_i32(const int32_t *flag, int64_t n, int32_t *out_index, int64_t *n_out_host,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Building blocks of nn/pool/topk_pool.py:31-58 (tf.argsort by source, per-source descending tf.argsort of the scores):
 * keys[i] = order-preserving 32-bit pattern of score[i] (complemented for descending; -0.0 ties with +0.0), and a
 * stable LSD radix argsort of 32-bit patterns read as unsigned numbers (only the low key_bits are examined):
 * perm_out[j] = position of the j-th smallest key, equal keys in input order. */
int tfgk_sort_keys_f32(const float *score, int64_t n, int descending, int32_t *keys, void *stream);
int tfgk_argsort_workspace_bytes(int64_t n, size_t *out_bytes);
int tfgk_stable_argsort_u32(const int32_t *keys, int64_t n, int key_bits, int32_t *perm_out,
                            void *workspace, size_t workspace_bytes, void *stream);

/* RandomNeighborSampler.sample (graph_utils.py:669-776) on a CSR whose rows are the source nodes in ascending order
 * with neighbours in edge order (= the reference's neighbor_dict).  k < 0 and ratio < 0: every neighbour; ratio < 0:
 * k per row (all of them in order when k >= degree and !padding; k draws WITH replacement when padding and
 * k >= degree); otherwise ceil(degree * ratio) without replacement.  Rows without neighbours emit nothing.
 * _count writes the [n_rows + 1] offsets of the sampled edges and their total (synchronises); _fill writes, per sampled
 * edge, its row and the CSR position it was taken from (gather col / weights with tfgk_permute_f32).  Without
 * replacement = reservoir sampling with draws (seed, rng_stream, row << 32 | i).
 * padding = TFGK_SAMPLE_HEAD is the deterministic rule of nn/pool/topk_pool.py:59-82: the FIRST min(k, degree) (or
 * ceil(float(degree) * float(ratio)), float32 like the reference) entries of every row, in order. */
int tfgk_neighbor_sample_workspace_bytes(int32_t n_rows, size_t *out_bytes);
int tfgk_neighbor_sample_count(const int64_t *rowptr, int32_t n_rows, int32_t k, double ratio, int padding,
                               int64_t *out_rowptr, int64_t *total_host, void *workspace, size_t workspace_bytes,
                               void *stream);
int tfgk_neighbor_sample_fill(const int64_t *rowptr, int32_t n_rows, int32_t k, double ratio, int padding,
                              uint64_t seed, uint32_t rng_stream, const int64_t *out_rowptr,
                              int32_t *out_row, int32_t *out_pos, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TFGK_H_ */
