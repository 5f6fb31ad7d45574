// K4 (tensor-core path): C[M,N] = act(A[M,K] @ B[K,N] + bias) for the dense projections x@W of the hot path
// (nn/conv/gcn.py:272, gat.py:52,61,70, graph_sage.py:43-44, appnp.py:69) on the 5th-generation tensor cores.
//
//   * tcgen05.mma.cta_group::1.kind::tf32, UMMA 128 x UN x 8, fp32 accumulators in TMEM (2 x UN columns, double buffered)
//   * fp32-accurate via the 3xTF32 split: a = hi + lo with hi = rna_tf32(a), lo = a - hi (exact);
//     D += Ahi*Bhi + Ahi*Blo + Alo*Bhi  (the dropped lo*lo term is ~2^-22 relative)  => ~1e-6 relative error, well
//     inside the 1e-4 parity gate that single-pass TF32 (~1e-3) would fail (SURVEY.md section 7 "hard parts")
//   * operands are staged by all threads: coalesced 16 B global loads -> hi/lo split in registers -> st.shared in the
//     canonical K-major, no-swizzle UMMA layout (8-row x 16 B core matrices; LBO = 128 B along K, SBO = 1024 B along M/N)
//     W is transposed on the fly (it is [K,N] row-major, the MMA wants K-major) - it is tiny and L2 resident
//   * one elected thread issues the MMAs; tcgen05.commit -> mbarrier releases the smem stage / publishes the accumulator
//   * persistent CTAs: the epilogue of tile i (tcgen05.ld -> +bias -> act -> global) overlaps the MMAs of tile i+1
// The kernel is memory bound (4(MK+MN) bytes for ~6MNK tf32 flops at K~100): the tensor pipe is lightly loaded by design.
#include "common.cuh"
#include <stdlib.h>

namespace tfgk {
namespace tc {

constexpr int BM = 128;            // UMMA_M
constexpr int BK = 32;             // K elements per smem stage (8 core matrices of 16 B along K)
constexpr int UMMA_K = 8;          // tf32
constexpr int kStages = 2;
constexpr int kThreads = 256;
constexpr int kMaxUN = 256;
constexpr int kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 28;

struct Params {
    const float *A; int64_t lda;
    const float *B; int64_t ldb;
    const float *bias; int act;
    int M, N, K;
    float *C; int64_t ldc;
    int un;            // UMMA_N: N rounded up to a multiple of 16
    int tiles_m;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spin > kSpinLimit) __trap();      // never hang the GPU: a lost arrive becomes a launch failure
    }
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_NONE, LBO = 128 B, SBO = 1024 B, descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address,        bits [0,14)
    d |= (uint64_t)(128u >> 4) << 16;                  // leading byte offset,  bits [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;                 // stride byte offset,   bits [32,46)
    d |= (uint64_t)1 << 46;                            // version = 1 (sm_100), bits [46,48)
    return d;                                          // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// 32-bit instruction descriptor: D = F32, A = B = TF32, both K-major, N = un, M = 128
__device__ __forceinline__ uint32_t make_idesc(int un) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format  = F32
    d |= 2u << 7;                       // a_format  = TF32
    d |= 2u << 10;                      // b_format  = TF32
    d |= (uint32_t)(un >> 3) << 17;     // n_dim
    d |= (uint32_t)(BM >> 4) << 24;     // m_dim
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    hi = __uint_as_float(h);
    lo = a - hi;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// smem carve-up (dynamic): per stage  Ahi | Alo (16 KB each) | Bhi | Blo (un*128 B each); then barriers
struct SmemLayout {
    uint32_t a_bytes, b_bytes, stage_bytes, total;
    __host__ __device__ explicit SmemLayout(int un) {
        a_bytes = BM * BK * 4;
        b_bytes = (uint32_t)un * BK * 4;
        stage_bytes = 2 * a_bytes + 2 * b_bytes;
        total = kStages * stage_bytes + 64;
    }
};

template <bool VEC_A>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L(p.un);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * L.stage_bytes);   // [0..kStages): stage free, then 2: acc full
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kStages + 2);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int i = 0; i < kStages + 2; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc(p.un);
    const int nkb = (p.K + BK - 1) / BK;
    const int b_chunks = p.un * (BK / 4);          // 16-byte chunks of one B operand stage

    uint32_t g = 0;                                // global k-block counter (stage ring position)
    int it = 0;                                    // tiles processed by this CTA
    int prev_tile = -1;

    auto epilogue = [&](int tile, int buf, uint32_t use) {
        mbar_wait(&bars[kStages + buf], use & 1u);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const int64_t row = (int64_t)tile * BM + q * 32 + lane;
        const int col_begin = half * (p.un / 2), col_end = (half + 1) * (p.un / 2);
        for (int c0 = col_begin; c0 < col_end; c0 += 8) {
            uint32_t r[8];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kMaxUN + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < p.M) {
                float *dst = p.C + row * p.ldc + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c0 + j < p.N) {
                        float v = __uint_as_float(r[j]);
                        if (p.bias) v += __ldg(p.bias + c0 + j);
                        dst[j] = apply_act(v, p.act);
                    }
                }
            }
        }
        tc_fence_before();
    };

    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t m0 = (int64_t)tile * BM;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
            const uint32_t stage = g % kStages;
            const uint32_t use = g / kStages;
            if (use > 0) mbar_wait(&bars[stage], (use - 1) & 1u);      // MMAs that read this stage have retired
            uint8_t *sbase = smem + stage * L.stage_bytes;
            const uint32_t a_hi = smem_u32(sbase), a_lo = a_hi + L.a_bytes;
            const uint32_t b_hi = a_lo + L.a_bytes, b_lo = b_hi + L.b_bytes;
            const int k0 = kb * BK;

            // ---- A: 128 rows x 32 floats = 1024 chunks of 16 B; chunk c -> smem offset 16*c (rg*1024 + kc*128 + r8*16)
            float4 av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = t + kThreads * i;
                const int r8 = c & 7, kc = (c >> 3) & 7, rg = c >> 6;
                const int64_t row = m0 + rg * 8 + r8;
                const int k = k0 + kc * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < p.M && k < p.K) {
                    const float *src = p.A + row * p.lda + k;
                    if (VEC_A && k + 3 < p.K) {
                        v = __ldg(reinterpret_cast<const float4 *>(src));
                    } else {
                        v.x = __ldg(src);
                        if (k + 1 < p.K) v.y = __ldg(src + 1);
                        if (k + 2 < p.K) v.z = __ldg(src + 2);
                        if (k + 3 < p.K) v.w = __ldg(src + 3);
                    }
                }
                av[i] = v;
            }
            // ---- B: un rows (n) x 32 floats (k), transposed on the fly from W[k][n]
            for (int cb = t; cb < b_chunks; cb += kThreads) {
                const int n8 = cb & 7, kc = (cb >> 3) & 7, ng = cb >> 6;
                const int n = ng * 8 + n8;
                const int k = k0 + kc * 4;
                float w[4] = {0.f, 0.f, 0.f, 0.f};
                if (n < p.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k + i < p.K) w[i] = __ldg(p.B + (int64_t)(k + i) * p.ldb + n);
                }
                float h[4], l[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(w[i], h[i], l[i]);
                st_shared_v4(b_hi + cb * 16, h[0], h[1], h[2], h[3]);
                st_shared_v4(b_lo + cb * 16, l[0], l[1], l[2], l[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = t + kThreads * i;
                float h[4], l[4];
                split_tf32(av[i].x, h[0], l[0]);
                split_tf32(av[i].y, h[1], l[1]);
                split_tf32(av[i].z, h[2], l[2]);
                split_tf32(av[i].w, h[3], l[3]);
                st_shared_v4(a_hi + c * 16, h[0], h[1], h[2], h[3]);
                st_shared_v4(a_lo + c * 16, l[0], l[1], l[2], l[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            tc_fence_before();
            __syncthreads();
            if (t == 0) {
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kMaxUN);
#pragma unroll
                for (int j = 0; j < BK / UMMA_K; ++j) {
                    const uint32_t off = (uint32_t)j * 2u * 128u;          // two 16-byte core matrices along K per MMA
                    const uint64_t dah = make_desc(a_hi + off), dal = make_desc(a_lo + off);
                    const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
                    umma_tf32(d_tmem, dal, dbh, idesc, (kb | j) != 0);     // small terms first, the hi*hi term last
                    umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                    umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                }
                umma_commit(&bars[stage]);                                 // frees the smem stage when these MMAs retire
                if (kb == nkb - 1) umma_commit(&bars[kStages + buf]);      // accumulator of this tile complete
            }
        }
        if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, (uint32_t)((it - 1) >> 1));
        prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1, (uint32_t)((it - 1) >> 1));

    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace tc
}  // namespace tfgk

using namespace tfgk;

extern "C" int tfgk_gemm_tc_f32(const float *A, int64_t lda, const float *B, int64_t ldb, const float *bias, int act,
                                int32_t M, int32_t N, int32_t K, float *C, int64_t ldc, void *stream) {
    // Opt-in (TFGK_GEMM_TC=1) until the TMA-fed version lands: this register-staged variant is correct but latency
    // bound (profiles/r1_kernel_variants.json).  K <= 512: tensor-core accumulation truncates (round-toward-zero), so
    // the error grows ~K^1.5; at K <= 512 it stays below 5e-6 relative, beyond that the exact-fp32 SIMT path is used.
    const char *env = getenv("TFGK_GEMM_TC");
    const bool enabled = env != nullptr && env[0] == '1';
    if (!enabled || N > tc::kMaxUN || M < 1 || K < 1 || K > 512 || (int64_t)M * K < (1 << 14)) return TFGK_ERR_UNSUPPORTED;
    tc::Params p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.bias = bias; p.act = act;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
    p.un = ((N + 15) / 16) * 16;
    p.tiles_m = (int)ceil_div64(M, tc::BM);
    const tc::SmemLayout L(p.un);
    int dev = 0, sms = 0;
    TFGK_CUDA(cudaGetDevice(&dev));
    TFGK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = p.tiles_m < sms ? p.tiles_m : sms;       // persistent: one CTA per SM
    const bool vec_a = (lda % 4 == 0) && aligned16(A);
    if (vec_a) {
        TFGK_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        tc::gemm_tf32x3_kernel<true><<<grid, tc::kThreads, L.total, as_stream(stream)>>>(p);
    } else {
        TFGK_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        tc::gemm_tf32x3_kernel<false><<<grid, tc::kThreads, L.total, as_stream(stream)>>>(p);
    }
    TFGK_LAUNCH_CHECK();
    return TFGK_OK;
}
