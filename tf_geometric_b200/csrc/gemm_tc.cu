// K4 (tensor-core path) placeholder: filled in by the tcgen05 3xTF32 kernel.
#include "common.cuh"
extern "C" int tfgk_gemm_tc_f32(const float *A, int64_t lda, const float *B, int64_t ldb, const float *bias, int act,
                                int32_t M, int32_t N, int32_t K, float *C, int64_t ldc, void *stream) {
    (void)A; (void)lda; (void)B; (void)ldb; (void)bias; (void)act; (void)M; (void)N; (void)K; (void)C; (void)ldc; (void)stream;
    return TFGK_ERR_UNSUPPORTED;
}
