// Device-side edge sampling: structural / Bernoulli edge flags, stable stream compaction and a
// CSR fan-out neighbour sampler.  They replace nn/sampling/drop_edge.py:6-52 (tf.nn.dropout + boolean_mask),
// utils/graph_utils.py:775-846 (UniformNeighborSampler) and the per-node Python loop of RandomNeighborSampler.sample
// (utils/graph_utils.py:669-776).  Integer work, HBM-bound; randomness is counter-based (rng.cuh) so every draw is a
// pure function of (seed, element) and the CPU restatement used by the tests reproduces the output bit for bit.
#include "common.cuh"
#include "scan.cuh"
#include "rng.cuh"

namespace tfgk {
namespace {

__global__ void edge_flags_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ col, int64_t E, int mode,
                                  const int32_t *__restrict__ row_map, const int32_t *__restrict__ col_map,
                                  int bernoulli, float prob, uint64_t seed, uint32_t stream, int32_t *__restrict__ flag) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        bool keep = true;
        if (mode == TFGK_FLAG_UPPER) keep = row[e] < col[e];
        else if (mode == TFGK_FLAG_MAPPED) keep = row_map[row[e]] >= 0 && col_map[col[e]] >= 0;
        if (keep && bernoulli != TFGK_BERNOULLI_NONE) {
            const float u = random_uniform(seed, stream, (uint64_t)e);
            keep = bernoulli == TFGK_BERNOULLI_DROPOUT ? (u >= prob) : (u <= prob);
        }
        flag[e] = keep ? 1 : 0;
    }
}

__global__ void select_emit_kernel(const int32_t *__restrict__ flag, const int32_t *__restrict__ off, int64_t n,
                                   int32_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flag[i]) out[off[i]] = (int32_t)i;
}

enum { kSampleAll = 0, kSampleReplace = 1, kSampleReservoir = 2 };

// how many neighbours row r contributes and by which rule (graph_utils.py:741-756)
__device__ __forceinline__ int sample_rule(int deg, int k, double ratio, int padding, int &num) {
    if (deg == 0) { num = 0; return kSampleAll; }
    if (padding == TFGK_SAMPLE_HEAD) {              // topk_pool.py:59-67: the first node_k entries of the row, in order
        num = ratio < 0.0 ? (k < deg ? k : deg) : (int)ceilf((float)deg * (float)ratio);
        if (num > deg) num = deg;
        if (num < 0) num = 0;
        return kSampleAll;
    }
    if ((k < 0 && ratio < 0.0) || (ratio < 0.0 && !padding && k >= deg)) { num = deg; return kSampleAll; }
    if (ratio < 0.0) { num = k; return (padding && k >= deg) ? kSampleReplace : kSampleReservoir; }
    num = (int)ceil((double)deg * ratio);
    return kSampleReservoir;
}

__global__ void sample_count_kernel(const int64_t *__restrict__ rowptr, int32_t N, int k, double ratio, int padding,
                                    int32_t *__restrict__ cnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int num;
    sample_rule((int)(rowptr[r + 1] - rowptr[r]), k, ratio, padding, num);
    cnt[r] = num;
}

// one thread per row; rows are independent and write disjoint output ranges.  Without replacement: reservoir sampling
// (algorithm R) held directly in the row's output slots.
__global__ void sample_fill_kernel(const int64_t *__restrict__ rowptr, int32_t N, int k, double ratio, int padding,
                                   uint64_t seed, uint32_t stream, const int64_t *__restrict__ out_rowptr,
                                   int32_t *__restrict__ out_row, int32_t *__restrict__ out_pos) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int64_t start = rowptr[r];
    const int deg = (int)(rowptr[r + 1] - start);
    int num;
    const int rule = sample_rule(deg, k, ratio, padding, num);
    const int64_t o = out_rowptr[r];
    const uint64_t base = (uint64_t)r << 32;
    for (int i = 0; i < num; ++i) out_row[o + i] = (int32_t)r;
    if (rule == kSampleReplace) {
        for (int i = 0; i < num; ++i)
            out_pos[o + i] = (int32_t)(start + random_below(seed, stream, base + (uint64_t)i, (uint32_t)deg));
        return;
    }
    for (int i = 0; i < num; ++i) out_pos[o + i] = (int32_t)(start + i);
    if (rule == kSampleReservoir) {
        for (int i = num; i < deg; ++i) {
            const uint32_t j = random_below(seed, stream, base + (uint64_t)i, (uint32_t)(i + 1));
            if (j < (uint32_t)num) out_pos[o + j] = (int32_t)(start + i);
        }
    }
}

}  // namespace
}  // namespace tfgk

using namespace tfgk;

extern "C" {

int tfgk_edge_flags_i32(const int32_t *row, const int32_t *col, int64_t E, int mode,
                        const int32_t *row_map, const int32_t *col_map,
                        int bernoulli, float prob, uint64_t seed, uint32_t rng_stream, int32_t *flag, void *stream) {
    TFGK_CHECK_ARG(E >= 0, "edge_flags: negative E");
    TFGK_CHECK_ARG(mode == TFGK_FLAG_ALL || mode == TFGK_FLAG_UPPER || mode == TFGK_FLAG_MAPPED, "edge_flags: unknown mode %d", mode);
    TFGK_CHECK_ARG(bernoulli == TFGK_BERNOULLI_NONE || bernoulli == TFGK_BERNOULLI_DROPOUT || bernoulli == TFGK_BERNOULLI_KEEP,
                   "edge_flags: unknown bernoulli rule %d", bernoulli);
    if (E == 0) return TFGK_OK;
    TFGK_CHECK_ARG(flag != nullptr, "edge_flags: null output");
    TFGK_CHECK_ARG(mode == TFGK_FLAG_ALL || (row && col), "edge_flags: null edge list");
    TFGK_CHECK_ARG(mode != TFGK_FLAG_MAPPED || (row_map && col_map), "edge_flags: null node map");
    edge_flags_kernel<<<grid_for(E), 256, 0, as_stream(stream)>>>(row, col, E, mode, row_map, col_map, bernoulli, prob, seed,
                                                                  rng_stream, flag);
    TFGK_LAUNCH_CHECK();
    return TFGK_OK;
}

int tfgk_select_workspace_bytes(int64_t n, size_t *out_bytes) {
    TFGK_CHECK_ARG(out_bytes != nullptr && n >= 0 && n < (1ll << 31) - 1, "select_workspace_bytes: bad argument");
    *out_bytes = align_up((size_t)(n + 1) * 4) + scan_scratch_bytes(n + 1) + 256;
    return TFGK_OK;
}

int tfgk_select_flagged_i32(const int32_t *flag, int64_t n, int32_t *out_index, int64_t *n_out_host,
                            void *workspace, size_t workspace_bytes, void *stream) {
    TFGK_CHECK_ARG(n >= 0 && n < (1ll << 31) - 1, "select_flagged: bad size");
    TFGK_CHECK_ARG(n_out_host != nullptr, "select_flagged: null count");
    *n_out_host = 0;
    if (n == 0) return TFGK_OK;
    TFGK_CHECK_ARG(flag && out_index, "select_flagged: null pointer");
    size_t need = 0;
    tfgk_select_workspace_bytes(n, &need);
    if (workspace == nullptr || workspace_bytes < need)
        return set_error(TFGK_ERR_WORKSPACE, "select_flagged: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    cudaStream_t st = as_stream(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *off = reinterpret_cast<int32_t *>(ws);
    int32_t *sums = reinterpret_cast<int32_t *>(ws + align_up((size_t)(n + 1) * 4));
    const int rc = exclusive_scan<int32_t, int32_t>(flag, n, n + 1, off, sums, st);
    if (rc != TFGK_OK) return rc;
    int32_t total = 0;
    TFGK_CUDA(cudaMemcpyAsync(&total, off + n, 4, cudaMemcpyDeviceToHost, st));
    select_emit_kernel<<<grid_for(n), 256, 0, st>>>(flag, off, n, out_index);
    TFGK_LAUNCH_CHECK();
    TFGK_CUDA(cudaStreamSynchronize(st));
    *n_out_host = total;
    return TFGK_OK;
}

int tfgk_neighbor_sample_workspace_bytes(int32_t n_rows, size_t *out_bytes) {
    TFGK_CHECK_ARG(out_bytes != nullptr && n_rows >= 0, "neighbor_sample_workspace_bytes: bad argument");
    *out_bytes = align_up(((size_t)n_rows + 1) * 4) + scan_scratch_bytes((int64_t)n_rows + 1) + 256;
    return TFGK_OK;
}

static int check_sample_args(const char *fn, int32_t n_rows, int32_t k, double ratio) {
    TFGK_CHECK_ARG(n_rows >= 0, "%s: negative row count", fn);
    TFGK_CHECK_ARG(!(k >= 0 && ratio >= 0.0), "%s: k and ratio cannot be provided simultaneously", fn);
    TFGK_CHECK_ARG(ratio <= 1.0, "%s: ratio %g > 1 cannot be sampled without replacement", fn, ratio);
    return TFGK_OK;
}

static int check_sample_mode(const char *fn, int32_t k, double ratio, int padding) {
    TFGK_CHECK_ARG(padding == 0 || padding == 1 || padding == TFGK_SAMPLE_HEAD, "%s: unknown padding mode %d", fn, padding);
    TFGK_CHECK_ARG(padding != TFGK_SAMPLE_HEAD || k >= 0 || ratio >= 0.0, "%s: the head rule needs k or ratio", fn);
    return TFGK_OK;
}

int tfgk_neighbor_sample_count(const int64_t *rowptr, int32_t n_rows, int32_t k, double ratio, int padding,
                               int64_t *out_rowptr, int64_t *total_host, void *workspace, size_t workspace_bytes,
                               void *stream) {
    int rc = check_sample_args("neighbor_sample_count", n_rows, k, ratio);
    if (rc != TFGK_OK) return rc;
    if ((rc = check_sample_mode("neighbor_sample_count", k, ratio, padding)) != TFGK_OK) return rc;
    TFGK_CHECK_ARG(total_host != nullptr && out_rowptr != nullptr, "neighbor_sample_count: null pointer");
    *total_host = 0;
    cudaStream_t st = as_stream(stream);
    if (n_rows == 0) {
        TFGK_CUDA(cudaMemsetAsync(out_rowptr, 0, 8, st));
        return TFGK_OK;
    }
    TFGK_CHECK_ARG(rowptr != nullptr, "neighbor_sample_count: null rowptr");
    size_t need = 0;
    tfgk_neighbor_sample_workspace_bytes(n_rows, &need);
    if (workspace == nullptr || workspace_bytes < need)
        return set_error(TFGK_ERR_WORKSPACE, "neighbor_sample_count: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    char *ws = static_cast<char *>(workspace);
    int32_t *cnt = reinterpret_cast<int32_t *>(ws);
    int64_t *sums = reinterpret_cast<int64_t *>(ws + align_up(((size_t)n_rows + 1) * 4));
    sample_count_kernel<<<(unsigned)ceil_div64(n_rows, 256), 256, 0, st>>>(rowptr, n_rows, k, ratio, padding, cnt);
    TFGK_LAUNCH_CHECK();
    rc = exclusive_scan<int32_t, int64_t>(cnt, n_rows, (int64_t)n_rows + 1, out_rowptr, sums, st);
    if (rc != TFGK_OK) return rc;
    TFGK_CUDA(cudaMemcpyAsync(total_host, out_rowptr + n_rows, 8, cudaMemcpyDeviceToHost, st));
    TFGK_CUDA(cudaStreamSynchronize(st));
    TFGK_CHECK_ARG(*total_host < (1ll << 31) - 1, "neighbor_sample_count: %lld sampled edges exceed int32 positions", (long long)*total_host);
    return TFGK_OK;
}

int tfgk_neighbor_sample_fill(const int64_t *rowptr, int32_t n_rows, int32_t k, double ratio, int padding,
                              uint64_t seed, uint32_t rng_stream, const int64_t *out_rowptr,
                              int32_t *out_row, int32_t *out_pos, void *stream) {
    int rc = check_sample_args("neighbor_sample_fill", n_rows, k, ratio);
    if (rc != TFGK_OK) return rc;
    if ((rc = check_sample_mode("neighbor_sample_fill", k, ratio, padding)) != TFGK_OK) return rc;
    if (n_rows == 0) return TFGK_OK;
    TFGK_CHECK_ARG(rowptr && out_rowptr, "neighbor_sample_fill: null pointer");
    sample_fill_kernel<<<(unsigned)ceil_div64(n_rows, 128), 128, 0, as_stream(stream)>>>(rowptr, n_rows, k, ratio, padding, seed,
                                                                                       rng_stream, out_rowptr, out_row, out_pos);
    TFGK_LAUNCH_CHECK();
    return TFGK_OK;
}

}  // extern "C"
