// gemm_tcgen05.cu -- blocked contraction on the 5th-generation tensor cores (sm_100a).
//
//     C[M,N] (fp32)  (+)=  sum_p  A_p[M,K] . B_p[N,K]^T          (all operands K-major, bf16)
//
// Replaces, for one output block (i, j), the reference's loop
//     reduce(np.add, [np.matmul(a[i,k], b[k,j])[..., None, :] for k in ...])
// (`_matmul` linalg/_tensordot.py:194-213 + `_sum_wo_cat` :216-249): the pair list p runs over
// the contracted block index k, so the (M,1,N) partials the reference materialises (32 GiB at
// BASELINE config 5) never exist -- the k-accumulation stays in TMEM.  The same pair list
// carries the bf16 x 3 splitting used for fp32 operands (6 products per k block).
//
// Kernel: one CTA per 128 x 256 output tile, warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (128B-swizzled 128x64 / 256x64 bf16 boxes) -> 4-stage smem ring
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16);
//                              accumulator = 128 lanes x 256 fp32 columns of TMEM; tcgen05.commit frees stages
//                              (a 128x128 tile needs 32 KiB of operands per 256 MMA cycles = the whole 128 B/clk
//                              shared-memory bandwidth -- ncu: tensor pipe 66 % -- 128x256 needs 94 B/clk)
//   warps 2..5  epilogue       tcgen05.ld 32x32b -> registers -> (+= C) -> global
// SASS evidence: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld).
#include "../../include/b200da.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

extern "C" int b2_set_error_(int code, const char* msg);
extern "C" void b2_count_launch_(void);

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int MAX_PAIRS = 12;             // 24 tensor maps = 3 KiB of kernel parameters
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct GemmMaps {
    CUtensorMap a[MAX_PAIRS];
    CUtensorMap b[MAX_PAIRS];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "B2_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra B2_DONE_%=;\n\t"
        "bra B2_WAIT_%=;\n\t"
        "B2_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, 128-byte swizzle (what the TMA box writes): rows are 128 B apart inside an
// 8-row swizzle atom, atoms 1024 B apart (stride byte offset); descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
__device__ __forceinline__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one output block of a batched launch (device copy of b2_gemm_problem + launch bookkeeping)
struct GemmProblem {
    float* C;
    long long ldc;
    int M, N, K, npairs;
    int accumulate, map_begin;      // maps[2 * (map_begin + p)] = A_p, +1 = B_p
    long long tile_begin;           // exclusive prefix sum of tiles
    int tiles_n, _pad;
};

__device__ __forceinline__ void gemm_tile(const CUtensorMap* __restrict__ amaps, const CUtensorMap* __restrict__ bmaps,
                                          int map_stride, int npairs, float* __restrict__ C, long long ldc,
                                          int M, int N, int K, int accumulate, int m0, int n0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B: 1024 B aligned
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_iters = (K + BK - 1) / BK;
    const int total_iters = npairs * k_iters;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {      // one warp allocates the accumulator: 256 TMEM columns (128 lanes x 256 fp32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(amaps);
            tma_prefetch_desc(bmaps);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                const int p = it / k_iters, kk = it % k_iters;
                if (kk == 0 && p + 1 < npairs) {      // descriptors of the NEXT pair: fetched a whole K loop ahead
                    tma_prefetch_desc(amaps + (size_t)(p + 1) * map_stride);
                    tma_prefetch_desc(bmaps + (size_t)(p + 1) * map_stride);
                }
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                uint8_t* sa = smem + s * STAGE_BYTES;
                tma_load_2d(sa, amaps + (size_t)p * map_stride, kk * BK, m0, &full_bar[s]);
                tma_load_2d(sa + A_BYTES, bmaps + (size_t)p * map_stride, kk * BK, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
                    umma_bf16(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // frees the stage when these MMAs retire
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // epilogue warps 2..5: a warp may only touch TMEM lanes [32*(warp%4), +32)
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        float* crow = C + (long long)row * ldc + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            if (row < M) {
                if (n0 + c + 32 <= N && (ldc % 4 == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        float4* dst = reinterpret_cast<float4*>(crow + c + j);
                        if (accumulate) { float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *dst = v;
                    }
                } else {
                    for (int j = 0; j < 32; ++j) {
                        if (n0 + c + j < N) {
                            float v = __uint_as_float(r[j]);
                            if (accumulate) v += crow[c + j];
                            crow[c + j] = v;
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN));
    }
}

// single problem, tensor maps passed as kernel parameters
__global__ void __launch_bounds__(NUM_THREADS, 1)
b2_gemm_tn_kernel(const __grid_constant__ GemmMaps maps, int npairs, float* __restrict__ C, long long ldc,
                  int M, int N, int K, int accumulate) {
    gemm_tile(maps.a, maps.b, 1, npairs, C, ldc, M, N, K, accumulate, blockIdx.y * BM, blockIdx.x * BN);
}

// all output blocks of a blocked matmul in ONE launch: problems + tensor maps live in global memory
__global__ void __launch_bounds__(NUM_THREADS, 1)
b2_gemm_tn_batched_kernel(const GemmProblem* __restrict__ probs, int nprobs, const CUtensorMap* __restrict__ maps) {
    __shared__ GemmProblem pr;
    if (threadIdx.x == 0) {
        const long long tile = blockIdx.x;
        int lo = 0, hi = nprobs - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (probs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
        }
        pr = probs[lo];
    }
    __syncthreads();
    const int t = (int)((long long)blockIdx.x - pr.tile_begin);
    const int mt = t / pr.tiles_n, nt = t % pr.tiles_n;
    const CUtensorMap* base = maps + 2 * (size_t)pr.map_begin;
    gemm_tile(base, base + 1, 2, pr.npairs, pr.C, pr.ldc, pr.M, pr.N, pr.K, pr.accumulate, mt * BM, nt * BN);
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Two CTAs of one cluster (the two SMs of a TPC) compute ONE 256 x 256 output tile: CTA r owns rows
// [m0 + 128 r, +128) of A and of the accumulator (its own TMEM, 128 lanes x 256 columns) and loads HALF of the
// B tile -- rows [n0 + 128 r, +128) -- into its shared memory; the MMA (M = 256, issued by the leader CTA
// only) reads both halves.  Per CTA and k step that is 16 KiB of A + 16 KiB of B for the tensor work that
// costs the single-CTA kernel 16 + 32 KiB: a third less L2 -> SM traffic and shared-memory fill, which is
// what the power-capped tensor pipe is waiting for, and room for 6 pipeline stages instead of 4.
constexpr int P_BM = 128, P_BN = 256, P_BNH = 128, P_STAGES = 6;
constexpr int P_A_BYTES = P_BM * BK * 2, P_B_BYTES = P_BNH * BK * 2, P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;     // shared::cluster address of the SAME offset in CTA 0 of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void gemm_tile_pair(const CUtensorMap* __restrict__ amaps, const CUtensorMap* __restrict__ bmaps,
                                               int map_stride, int npairs, float* __restrict__ C, long long ldc,
                                               int M, int N, int K, int accumulate, int m0, int n0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + P_STAGES * P_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + P_STAGES;
    uint64_t* tmem_full_bar = empty_bar + P_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int k_iters = (K + BK - 1) / BK;
    const int total_iters = npairs * k_iters;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {      // the same warp of BOTH CTAs allocates: 256 TMEM columns in each SM
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(P_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(amaps);
            tma_prefetch_desc(bmaps);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % P_STAGES;
                const uint32_t ph = (uint32_t)(it / P_STAGES) & 1u;
                const int p = it / k_iters, kk = it % k_iters;
                if (kk == 0 && p + 1 < npairs) {
                    tma_prefetch_desc(amaps + (size_t)(p + 1) * map_stride);
                    tma_prefetch_desc(bmaps + (size_t)(p + 1) * map_stride);
                }
                mbar_wait(&empty_bar[s], ph ^ 1u);                               // own CTA's copy: freed by the multicast commit
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * P_STAGE_BYTES);  // the leader counts both CTAs' bytes
                uint8_t* sa = smem + s * P_STAGE_BYTES;
                tma_load_2d_pair(sa, amaps + (size_t)p * map_stride, kk * BK, m0 + (int)rank * P_BM, &full_bar[s]);
                tma_load_2d_pair(sa + P_A_BYTES, bmaps + (size_t)p * map_stride, kk * BK, n0 + (int)rank * P_BNH, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * P_BM, P_BN);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % P_STAGES;
                const uint32_t ph = (uint32_t)(it / P_STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * P_STAGE_BYTES);
                const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + P_A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_bf16_pair(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
                umma_commit_pair(&empty_bar[s]);        // frees the stage in BOTH CTAs when these MMAs retire
            }
            umma_commit_pair(tmem_full_bar);            // accumulators of both CTAs complete
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = m0 + (int)rank * P_BM + q * 32 + lane;
        float* crow = C + (long long)row * ldc + n0;
#pragma unroll 1
        for (int c = 0; c < P_BN; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            if (row < M) {
                if (n0 + c + 32 <= N && (ldc % 4 == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        float4* dst = reinterpret_cast<float4*>(crow + c + j);
                        if (accumulate) { float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *dst = v;
                    }
                } else {
                    for (int j = 0; j < 32; ++j) {
                        if (n0 + c + j < N) {
                            float v = __uint_as_float(r[j]);
                            if (accumulate) v += crow[c + j];
                            crow[c + j] = v;
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // neither CTA may free TMEM / exit while the peer still uses it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(P_BN));
    }
}

// every output block of a blocked matmul, one CTA pair per 256 x 256 tile.  Tiles of a block are walked in
// bands of 8 tile rows, column by column inside a band, so that the ~74 pairs resident at a time cover a
// near-square 8 x 9 patch of the output (operand rows fetched per wave: 8 x 256 + 9 x 256).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
b2_gemm_tn_pair_kernel(const GemmProblem* __restrict__ probs, int nprobs, const CUtensorMap* __restrict__ maps) {
    __shared__ GemmProblem pr;
    const long long tile = blockIdx.x >> 1;
    if (threadIdx.x == 0) {
        int lo = 0, hi = nprobs - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (probs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
        }
        pr = probs[lo];
    }
    __syncthreads();
    const int t = (int)(tile - pr.tile_begin);
    const int tiles_m = (pr.M + 2 * P_BM - 1) / (2 * P_BM);
    constexpr int BAND = 8;
    const int per_band = BAND * pr.tiles_n;
    const int band = t / per_band, in_band = t - band * per_band;
    const int rows_here = (tiles_m - band * BAND < BAND) ? (tiles_m - band * BAND) : BAND;
    const int mt = band * BAND + in_band % rows_here, nt = in_band / rows_here;
    const CUtensorMap* base = maps + 2 * (size_t)pr.map_begin;
    gemm_tile_pair(base, base + 1, 2, pr.npairs, pr.C, pr.ldc, pr.M, pr.N, pr.K, pr.accumulate, mt * 2 * P_BM, nt * P_BN);
}

// fp32 -> three bf16 planes with hi + mid + lo == x to ~2^-24 (x - hi and the next residual are exact)
__global__ void __launch_bounds__(256) b2_split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo,
                                                        long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(m);
        hi[i] = h; mid[i] = m; lo[i] = __float2bfloat16_rn(r2);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (h) fn = (EncodeTiledFn)dlsym(h, "cuTensorMapEncodeTiled");
    });
    return fn;
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return b2_set_error_(B2_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no NVIDIA driver): no CPU fallback");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r,
                 (long long)rows, (long long)K, (long long)ld);
        return b2_set_error_(B2_ERR_CUDA, msg);
    }
    return B2_OK;
}

}  // namespace

extern "C" int b2_gemm_tn_pairs(int dtype, const void* const* A, const void* const* B, int npairs,
                                int64_t lda, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                int accumulate, void* stream) {
    if (dtype != B2_BF16) return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_pairs: operands must be bf16 planes (split fp32 with b2_split3_bf16)");
    if (!A || !B || !C || npairs <= 0 || M <= 0 || N <= 0 || K <= 0) return b2_set_error_(B2_ERR_INVALID, "b2_gemm_tn_pairs: bad argument");
    if (lda % 8 || ldb % 8) return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_pairs: lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(b2_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(attr_err));
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    for (int p0 = 0; p0 < npairs; p0 += MAX_PAIRS) {
        const int np = (npairs - p0 < MAX_PAIRS) ? npairs - p0 : MAX_PAIRS;
        GemmMaps maps;
        memset(&maps, 0, sizeof maps);
        for (int p = 0; p < np; ++p) {
            if (((uintptr_t)A[p0 + p] | (uintptr_t)B[p0 + p]) % 16) return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_pairs: operand not 16-byte aligned");
            int rc = make_map(&maps.a[p], A[p0 + p], M, K, lda, BM);
            if (rc) return rc;
            rc = make_map(&maps.b[p], B[p0 + p], N, K, ldb, BN);
            if (rc) return rc;
        }
        b2_gemm_tn_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
            maps, np, C, (long long)ldc, (int)M, (int)N, (int)K, (accumulate || p0 > 0) ? 1 : 0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(e));
        b2_count_launch_();
    }
    return B2_OK;
}

extern "C" int b2_gemm_tn_batched(int dtype, const b2_gemm_problem* problems, int nproblems,
                                  void* workspace, size_t workspace_bytes, size_t* needed, void* stream) {
    if (dtype != B2_BF16) return b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_batched: operands must be bf16 planes");
    if (!problems || nproblems <= 0) return b2_set_error_(B2_ERR_INVALID, "b2_gemm_tn_batched: bad argument");
    size_t npairs_total = 0;
    for (int i = 0; i < nproblems; ++i) npairs_total += (size_t)problems[i].npairs;
    const size_t maps_bytes = npairs_total * 2 * sizeof(CUtensorMap);
    const size_t probs_off = (maps_bytes + 255) / 256 * 256;
    const size_t need = probs_off + (size_t)nproblems * sizeof(GemmProblem);
    if (needed) *needed = need;
    if (!workspace) return B2_OK;                       // size query
    if (workspace_bytes < need || ((uintptr_t)workspace % 64)) return b2_set_error_(B2_ERR_WORKSPACE, "b2_gemm_tn_batched: workspace too small or not 64-byte aligned");
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(b2_gemm_tn_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(b2_gemm_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(attr_err));
    // CTA-pair kernel (256 x 256 tiles) unless B2_GEMM_2CTA=0 asks for the single-CTA one (A/B comparison)
    static const bool pair = [] { const char* e = getenv("B2_GEMM_2CTA"); return !(e && e[0] == '0'); }();
    const int box_b = pair ? P_BNH : BN;
    const int tile_m = pair ? 2 * P_BM : BM;
    void* host = nullptr;
    if (posix_memalign(&host, 64, need)) return b2_set_error_(B2_ERR_INVALID, "out of host memory");
    memset(host, 0, need);
    CUtensorMap* hm = (CUtensorMap*)host;
    GemmProblem* hp = (GemmProblem*)((char*)host + probs_off);
    long long tiles = 0;
    int map_begin = 0, rc = B2_OK;
    for (int i = 0; i < nproblems && rc == B2_OK; ++i) {
        const b2_gemm_problem& q = problems[i];
        if (!q.A || !q.B || !q.C || q.npairs <= 0 || q.M <= 0 || q.N <= 0 || q.K <= 0) { rc = b2_set_error_(B2_ERR_INVALID, "b2_gemm_tn_batched: bad problem"); break; }
        if (q.lda % 8 || q.ldb % 8) { rc = b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_batched: lda/ldb must be multiples of 8 elements"); break; }
        for (int p = 0; p < q.npairs && rc == B2_OK; ++p) {
            if (((uintptr_t)q.A[p] | (uintptr_t)q.B[p]) % 16) { rc = b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_batched: operand not 16-byte aligned"); break; }
            const int64_t kp = q.Kpair ? q.Kpair[p] : q.K;
            const int64_t la = q.Kpair ? kp : q.lda, lb = q.Kpair ? kp : q.ldb;
            if (kp <= 0 || kp > q.K || la % 8 || lb % 8) { rc = b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_batched: per-pair K must be in (0, K] and a multiple of 8"); break; }
            rc = make_map(&hm[2 * (map_begin + p)], q.A[p], q.M, kp, la, BM);
            if (rc == B2_OK) rc = make_map(&hm[2 * (map_begin + p) + 1], q.B[p], q.N, kp, lb, box_b);
        }
        GemmProblem& g = hp[i];
        g.C = q.C; g.ldc = q.ldc; g.M = (int)q.M; g.N = (int)q.N; g.K = (int)q.K; g.npairs = q.npairs;
        g.accumulate = q.accumulate; g.map_begin = map_begin; g.tile_begin = tiles;
        g.tiles_n = (int)((q.N + BN - 1) / BN);
        tiles += (long long)((q.M + tile_m - 1) / tile_m) * g.tiles_n;
        map_begin += q.npairs;
    }
    if (rc == B2_OK && tiles > 0x3fffffffLL) rc = b2_set_error_(B2_ERR_UNSUPPORTED, "b2_gemm_tn_batched: too many tiles");
    if (rc == B2_OK) {
        // pageable source: the copy is staged before the call returns, so `host` can be freed
        cudaError_t e = cudaMemcpyAsync(workspace, host, need, cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) rc = b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(e));
    }
    free(host);
    if (rc != B2_OK) return rc;
    if (pair)
        b2_gemm_tn_pair_kernel<<<(unsigned)(2 * tiles), NUM_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(
            (const GemmProblem*)((char*)workspace + probs_off), nproblems, (const CUtensorMap*)workspace);
    else
        b2_gemm_tn_batched_kernel<<<(unsigned)tiles, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
            (const GemmProblem*)((char*)workspace + probs_off), nproblems, (const CUtensorMap*)workspace);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(e));
    b2_count_launch_();
    return B2_OK;
}

extern "C" int b2_gemm_tn(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb,
                          float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                          int accumulate, void* stream) {
    const void* a[1] = {A};
    const void* b[1] = {B};
    return b2_gemm_tn_pairs(dtype, a, b, 1, lda, ldb, C, ldc, M, N, K, accumulate, stream);
}

extern "C" int b2_split3_bf16(const float* src, void* hi, void* mid, void* lo, int64_t n, void* stream) {
    if (!src || !hi || !mid || !lo || n < 0) return b2_set_error_(B2_ERR_INVALID, "b2_split3_bf16: bad argument");
    if (n == 0) return B2_OK;
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    b2_split3_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)mid,
                                                                  (__nv_bfloat16*)lo, (long long)n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b2_set_error_(B2_ERR_CUDA, cudaGetErrorString(e));
    b2_count_launch_();
    return B2_OK;
}
