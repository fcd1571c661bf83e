// gemm_tcgen05.cu -- blocked contraction C (+)= A @ B^T on the 5th-gen tensor cores.
// (placeholder entry point: filled in by the tcgen05/TMA kernel; fails loudly until then.)
#include "../../include/b200da.h"
#include <cuda_runtime.h>

extern "C" int b2_set_error_(int code, const char* msg);

extern "C" int b2_gemm_tn(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb,
                          float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                          int accumulate, void* stream) {
    (void)dtype; (void)A; (void)lda; (void)B; (void)ldb; (void)C; (void)ldc; (void)M; (void)N; (void)K;
    (void)accumulate; (void)stream;
    return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn: tcgen05 kernel not built yet");
}
