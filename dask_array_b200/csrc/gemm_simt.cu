// gemm_simt.cu -- exact-arithmetic block GEMM for the dtypes the tensor cores do not serve.
//
// The reference's blocked matmul / tensordot / einsum (linalg/_tensordot.py:20-42,194-249,
// _einsum.py:20-34) call np.matmul / np.tensordot / np.einsum per block for ANY NumPy number type:
// float64 (the dtype of the reference's own test matrix, tests/test_routines.py:321-399) and the
// integer types included.  tcgen05 has no fp64 or 32/64-bit integer MMA, so these run on the CUDA
// cores: C (+)= A @ B^T with A (M,K) and B (N,K) row-major ("TN", the layout of b2_gemm_tn), every
// product and sum in the element type T (fp64: FMA, the same operation np.matmul's BLAS uses;
// integers: wrap-around multiply-add, as NumPy).  64 x 64 output tile per CTA, 16-deep k slices staged
// through shared memory, 4 x 4 accumulators per thread.  Roofline: fp64 FMA pipe (~37 TFLOP/s on
// B200), not HBM -- this kernel is the coverage path, the tensor-core kernel is the fast one.
#include "../../include/b200da.h"

#include <cuda_runtime.h>
#include <cstdint>

extern "C" int b2_set_error_(int code, const char* msg);
extern "C" void b2_count_launch_(void);

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) b2_gemm_tn_simt_kernel(const T* __restrict__ A, long long lda, const T* __restrict__ B,
                                                              long long ldb, T* __restrict__ C, long long ldc, int M, int N,
                                                              int K, int accumulate) {
    __shared__ T sa[TK][TM + 1];
    __shared__ T sb[TK][TN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 x 16 threads, 4 x 4 outputs each
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int k0 = 0; k0 < K; k0 += TK) {
        // 64 rows x 16 k per operand = 1024 elements, 4 per thread; consecutive threads walk k (contiguous)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = threadIdx.x + t * 256;
            const int r = e >> 4, k = e & 15;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + k;
            sa[k][r] = (gm < M && gk < K) ? A[(long long)gm * lda + gk] : T(0);
            sb[k][r] = (gn < N && gk < K) ? B[(long long)gn * ldb + gk] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx + 16 * j;
            if (gn >= N) continue;
            T* c = C + (long long)gm * ldc + gn;
            *c = accumulate ? (T)(*c + acc[i][j]) : acc[i][j];
        }
    }
}

template <typename T>
int launch(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
           int accumulate, void* stream) {
    dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM));
    b2_gemm_tn_simt_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)A, lda, (const T*)B, ldb, (T*)C, ldc, (int)M,
                                                                     (int)N, (int)K, accumulate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(e));
    b2_count_launch_();
    return B2_OK;
}

}  // namespace

extern "C" int b2_gemm_tn_simt(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                               int64_t M, int64_t N, int64_t K, int accumulate, void* stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return b2_set_error_(B2_ERR_INVALID, "b2_gemm_tn_simt: bad argument");
    switch (dtype) {
        case B2_F64: return launch<double>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        case B2_F32: return launch<float>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        case B2_I32: return launch<int>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        case B2_U32: return launch<unsigned>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        case B2_I64: return launch<long long>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        case B2_U64: return launch<unsigned long long>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, stream);
        default: return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_simt: dtype must be f64, f32, i32, u32, i64 or u64");
    }
}
