// K1: gather - edge-apply - segment-reduce over a destination-sorted CSR  (tfgk_spmm_f32).
//
//   out[r,:] = epilogue( REDUCE_{e in [rowptr[r], rowptr[r+1])} w[e] * h[col[e], :] )
//
// HBM-bound (0.25-0.5 flop/byte): no tensor cores.  Mapping: a group of G lanes owns one destination row and
// each lane owns NC vectors of VEC consecutive feature columns, so one gathered source row is one coalesced
// G*VEC*4-byte request (512 B for D=128).  Edge ids/weights are read once, coalesced, G at a time and broadcast
// with shuffles; U gathered rows are kept in flight per group before any of them is consumed.  Accumulation is
// fp32, strictly in CSR (= input) order with separate multiply and add roundings, which makes SUM/MEAN
// bit-identical to tf.math.unsorted_segment_sum's sequential CPU loop (no atomics anywhere => deterministic).
//
// Algorithmic bytes per launch: E*(4*D + 4 [+4 weighted]) + N*(4*D + 8).
#include "common.cuh"

namespace tfgk {

template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float *p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        p[0] = v[0];
    }
}

struct SpmmParams {
    const int64_t *rowptr;
    const int32_t *col;
    const float *w;
    const float *h;
    int64_t ldh;
    int32_t n_dst;
    int32_t D;
    int reduce;
    float alpha;
    const float *addend;
    int64_t ld_addend;
    float beta;
    const float *bias;
    int act;
    float *out;
    int64_t ldo;
};

constexpr int kSpmmThreads = 256;

// G: lanes per row (power of two), NC: vectors per lane, IS_MAX: max-reduce instead of sum/mean, U: rows in flight
template <int VEC, int G, int NC, bool IS_MAX, int U>
__global__ void __launch_bounds__(kSpmmThreads) spmm_kernel(const SpmmParams p) {
    constexpr int RPW = 32 / G;   // rows per warp
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int grp = lane / G;
    const int64_t warp_global = (int64_t)blockIdx.x * (kSpmmThreads / 32) + (threadIdx.x >> 5);
    const int64_t r = warp_global * RPW + grp;
    const bool row_ok = r < p.n_dst;

    int64_t start = 0;
    int deg = 0;
    if (row_ok) {
        start = p.rowptr[r];
        deg = (int)(p.rowptr[r + 1] - start);
    }
    int deg_max = deg;
    if constexpr (RPW > 1) {
#pragma unroll
        for (int off = 16; off >= G; off >>= 1) deg_max = max(deg_max, __shfl_xor_sync(0xffffffffu, deg_max, off));
    }

    int coff[NC];
    bool cok[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        coff[k] = (gl + k * G) * VEC;
        cok[k] = coff[k] < p.D;
    }

    float acc[NC][VEC];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[k][v] = IS_MAX ? -FLT_MAX : 0.0f;

    const bool weighted = p.w != nullptr;
    const float *__restrict__ h = p.h;

    for (int t = 0; t < deg_max; t += G) {
        // coalesced read of up to G (col, w) pairs of this row
        const int e = t + gl;
        int my_c = 0;
        float my_w = 1.0f;
        if (e < deg) {
            my_c = ld_stream_i32(p.col + start + e);
            if (weighted) my_w = ld_stream_f32(p.w + start + e);
        }
        const int nb = min(G, deg - t);           // edges of this row in the batch (may be <= 0)
        const int nb_max = min(G, deg_max - t);   // warp-uniform trip count
        for (int j = 0; j < nb_max; j += U) {
            float v[U][NC][VEC];
            float ww[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = j + u;
                const int c = __shfl_sync(0xffffffffu, my_c, src, G);
                ww[u] = __shfl_sync(0xffffffffu, my_w, src, G);
                const bool ok = src < nb;
                const float *rowp = h + (int64_t)c * p.ldh;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    if (ok && cok[k]) load_vec<VEC>(rowp + coff[k], v[u][k]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (j + u < nb) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        if (cok[k]) {
#pragma unroll
                            for (int x = 0; x < VEC; ++x) {
                                const float m = __fmul_rn(v[u][k][x], ww[u]);   // gcn_mapper rounding
                                acc[k][x] = IS_MAX ? fmaxf(acc[k][x], m) : __fadd_rn(acc[k][x], m);
                            }
                        }
                    }
                }
            }
        }
    }

    if (!row_ok) return;
    const bool is_mean = p.reduce == TFGK_REDUCE_MEAN;
    const float cnt = (float)max(deg, 1);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        if (!cok[k]) continue;
        float o[VEC];
        float ad[VEC];
        float bs[VEC];
        if (p.addend) load_vec<VEC>(p.addend + r * p.ld_addend + coff[k], ad);
        if (p.bias) load_vec<VEC>(p.bias + coff[k], bs);
#pragma unroll
        for (int x = 0; x < VEC; ++x) {
            float a = acc[k][x];
            if (is_mean) a = __fdiv_rn(a, cnt);
            if (p.addend) a = __fadd_rn(__fmul_rn(a, p.alpha), __fmul_rn(ad[x], p.beta));
            else if (p.alpha != 1.0f) a = __fmul_rn(a, p.alpha);
            if (p.bias) a = __fadd_rn(a, bs[x]);
            o[x] = apply_act(a, p.act);
        }
        store_vec<VEC>(p.out + r * p.ldo + coff[k], o);
    }
}

template <int VEC, int G, int NC, int U>
static int launch_spmm(const SpmmParams &p, cudaStream_t st) {
    constexpr int rows_per_block = (kSpmmThreads / 32) * (32 / G);
    const int64_t blocks = ceil_div64(p.n_dst, rows_per_block);
    if (p.reduce == TFGK_REDUCE_MAX)
        spmm_kernel<VEC, G, NC, true, U><<<(unsigned)blocks, kSpmmThreads, 0, st>>>(p);
    else
        spmm_kernel<VEC, G, NC, false, U><<<(unsigned)blocks, kSpmmThreads, 0, st>>>(p);
    TFGK_LAUNCH_CHECK();
    return TFGK_OK;
}

template <int VEC>
static int dispatch_spmm(const SpmmParams &p, int lanes, cudaStream_t st) {
    // lanes = number of VEC-wide vectors in a row (<= 128)
    if (lanes <= 1) return launch_spmm<VEC, 1, 1, 8>(p, st);
    if (lanes <= 2) return launch_spmm<VEC, 2, 1, 8>(p, st);
    if (lanes <= 4) return launch_spmm<VEC, 4, 1, 8>(p, st);
    if (lanes <= 8) return launch_spmm<VEC, 8, 1, 8>(p, st);
    if (lanes <= 16) return launch_spmm<VEC, 16, 1, 8>(p, st);
    if (lanes <= 32) return launch_spmm<VEC, 32, 1, 8>(p, st);
    if (lanes <= 64) return launch_spmm<VEC, 32, 2, 4>(p, st);
    if (lanes <= 96) return launch_spmm<VEC, 32, 3, 2>(p, st);
    return launch_spmm<VEC, 32, 4, 2>(p, st);
}

}  // namespace tfgk

using namespace tfgk;

extern "C" int tfgk_spmm_f32(const int64_t *rowptr, const int32_t *col, const float *w,
                             const float *h, int64_t ldh, int32_t n_dst, int32_t D, int reduce,
                             float alpha, const float *addend, int64_t ld_addend, float beta,
                             const float *bias, int act,
                             float *out, int64_t ldo, void *stream) {
    TFGK_CHECK_ARG(n_dst >= 0 && D >= 0, "spmm: negative size (n_dst=%d, D=%d)", n_dst, D);
    TFGK_CHECK_ARG(reduce >= TFGK_REDUCE_SUM && reduce <= TFGK_REDUCE_MAX, "spmm: unknown reduce %d", reduce);
    TFGK_CHECK_ARG(act == TFGK_ACT_NONE || act == TFGK_ACT_RELU, "spmm: unknown activation %d", act);
    if (n_dst == 0 || D == 0) return TFGK_OK;
    TFGK_CHECK_ARG(rowptr && out && h, "spmm: null pointer");
    TFGK_CHECK_ARG(ldh >= D && ldo >= D && (!addend || ld_addend >= D), "spmm: leading dimension < D");

    const bool vec4 = (D % 4 == 0) && (ldh % 4 == 0) && (ldo % 4 == 0) && aligned16(h) && aligned16(out) &&
                      (!addend || ((ld_addend % 4 == 0) && aligned16(addend))) && (!bias || aligned16(bias));
    const int vec = vec4 ? 4 : 1;
    const int cols_per_launch = 128 * vec;   // 32 lanes x NC<=4 vectors
    for (int c0 = 0; c0 < D; c0 += cols_per_launch) {
        SpmmParams p;
        p.rowptr = rowptr; p.col = col; p.w = w;
        p.h = h + c0; p.ldh = ldh; p.n_dst = n_dst;
        p.D = (D - c0 < cols_per_launch) ? D - c0 : cols_per_launch;
        p.reduce = reduce; p.alpha = alpha;
        p.addend = addend ? addend + c0 : nullptr; p.ld_addend = ld_addend; p.beta = beta;
        p.bias = bias ? bias + c0 : nullptr; p.act = act;
        p.out = out + c0; p.ldo = ldo;
        const int lanes = (p.D + vec - 1) / vec;
        const int rc = vec4 ? dispatch_spmm<4>(p, lanes, as_stream(stream)) : dispatch_spmm<1>(p, lanes, as_stream(stream));
        if (rc != TFGK_OK) return rc;
    }
    return TFGK_OK;
}
