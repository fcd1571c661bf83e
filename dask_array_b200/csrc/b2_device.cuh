// b2_device.cuh -- device-side kernel templates of libb200da (sm_100a).
//
// This header is compiled two ways:
//   * by nvcc, ahead of time, into libb200da.so (b2_abi.cu);
//   * by NVRTC at run time, as the body of every fused kernel: the host-side generator
//     (dask_array_b200/_codegen.py) emits a `Chain` functor for one FusedBlockwise
//     expression plus a handful of #defines and then includes this file.
// It therefore uses no host or libstdc++ headers.
//
// What it replaces in the reference (paths under /root/reference/dask_array/):
//   * the per-block NumPy calls run by FusedBlockwise._task (_blockwise.py:1697-1728):
//     one temporary per operator becomes registers of one kernel;
//   * the reduction chunk step fused behind the chain (reductions/_reduction.py:154-226,
//     chunk kernels reductions/_common.py:92-105,270-281,368-404,704-732).
//
// Canonical block view: (B, R, C), C innermost.  A CTA owns one tile =
// (B2_RPT rows) x (B2_TX*B2_V columns) of one block [modes EW, R, RC] or
// (B2_RPT rows) x (all columns) [mode C].  Thread (tx, ty) walks rows ty, ty+TY, ...
// of the tile with B2_V consecutive columns, B2_U row-loads in flight.
#pragma once

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

#define B2_MAX_IN 6
#define B2_MAX_ND 4

// layout-identical to `b2_block` / `b2_scalars` in include/b200da.h (checked by
// static_asserts in b2_abi.cu)
struct B2Block {
    const void* in[B2_MAX_IN];
    i64 in_sb[B2_MAX_IN];
    i64 in_sr[B2_MAX_IN];
    i64 in_sc[B2_MAX_IN];
    void* out0;
    void* out1;
    i64 B, R, C;
    i64 tile_begin;
    i64 tiles_r, tiles_c;
    void* work;
    u32* counter;
    i64 arg_offset;
    int arg_ndim;
    int mirror;              // b2_run_ewt_sym: table index of the transposed partner block
    i64 arg_shape[B2_MAX_ND];
    i64 arg_start[B2_MAX_ND];
    i64 arg_total[B2_MAX_ND];
};

struct B2Scalars {
    double f[8];
    i64 i[8];
};

enum { B2M_EW = 0, B2M_R = 1, B2M_C = 2, B2M_RC = 3, B2M_SR = 4, B2M_SC = 5 };
enum { B2R_NONE = 0, B2R_SUM = 1, B2R_MIN = 2, B2R_MAX = 3, B2R_ARGMIN = 4, B2R_ARGMAX = 5,
       B2R_MOMENT = 6, B2R_PROD = 7, B2R_ANY = 8, B2R_ALL = 9, B2R_NANMIN = 10, B2R_NANMAX = 11 };

// ------------------------------------------------------------------ loads / stores
// 128-bit (or narrower) read-only streaming loads: every input byte is touched once,
// so keep it out of L1.
__device__ __forceinline__ void b2_ldg16(const void* p, u32 (&w)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
}
__device__ __forceinline__ void b2_ldg8(const void* p, u32 (&w)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];"
                 : "=r"(w[0]), "=r"(w[1]) : "l"(p));
}
__device__ __forceinline__ u32 b2_ldg4(const void* p) {
    u32 w;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(w) : "l"(p));
    return w;
}

// scalar read-only load of any element type (bool / 8-bit types go through unsigned char)
template <typename T> __device__ __forceinline__ T b2_ld(const T* p) {
    if constexpr (sizeof(T) == 1) {
        union { unsigned char c; T t; } u;
        u.c = __ldg(reinterpret_cast<const unsigned char*>(p));
        return u.t;
    } else {
        return __ldg(p);
    }
}

// V consecutive elements of T starting at p (p is V*sizeof(T)-aligned by host contract)
template <typename T, int V>
__device__ __forceinline__ void b2_load_vec(const T* p, T (&dst)[V]) {
    constexpr int BYTES = V * (int)sizeof(T);
    if constexpr (BYTES == 16) {
        union { u32 w[4]; T t[V]; } u;
        b2_ldg16(p, u.w);
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = u.t[v];
    } else if constexpr (BYTES == 8) {
        union { u32 w[2]; T t[V]; } u;
        b2_ldg8(p, u.w);
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = u.t[v];
    } else if constexpr (BYTES == 4) {
        union { u32 w; T t[V]; } u;
        u.w = b2_ldg4(p);
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = u.t[v];
    } else if constexpr (BYTES == 32) {
        union { u32 w[8]; T t[V]; } u;
        u32 a[4], b[4];
        b2_ldg16(p, a);
        b2_ldg16((const char*)p + 16, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { u.w[i] = a[i]; u.w[4 + i] = b[i]; }
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = u.t[v];
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = b2_ld(p + v);
    }
}
// one element broadcast to the V lanes (column stride 0)
template <typename T, int V>
__device__ __forceinline__ void b2_load_bcast(const T* p, T (&dst)[V]) {
    T x = b2_ld(p);
#pragma unroll
    for (int v = 0; v < V; ++v) dst[v] = x;
}
// arbitrary column stride (transposed views)
template <typename T, int V>
__device__ __forceinline__ void b2_load_strided(const T* p, i64 sc, T (&dst)[V]) {
#pragma unroll
    for (int v = 0; v < V; ++v) dst[v] = b2_ld(p + v * sc);
}

template <typename T, int V>
__device__ __forceinline__ void b2_store_vec(T* p, const T (&src)[V]) {
    constexpr int BYTES = V * (int)sizeof(T);
    if constexpr (BYTES == 16) {
        union { uint4 q; T t[V]; } u;
#pragma unroll
        for (int v = 0; v < V; ++v) u.t[v] = src[v];
        *reinterpret_cast<uint4*>(p) = u.q;
    } else if constexpr (BYTES == 8) {
        union { uint2 q; T t[V]; } u;
#pragma unroll
        for (int v = 0; v < V; ++v) u.t[v] = src[v];
        *reinterpret_cast<uint2*>(p) = u.q;
    } else if constexpr (BYTES == 4) {
        union { u32 q; T t[V]; } u;
#pragma unroll
        for (int v = 0; v < V; ++v) u.t[v] = src[v];
        *reinterpret_cast<u32*>(p) = u.q;
    } else if constexpr (BYTES == 32) {
        union { uint4 q[2]; T t[V]; } u;
#pragma unroll
        for (int v = 0; v < V; ++v) u.t[v] = src[v];
        reinterpret_cast<uint4*>(p)[0] = u.q[0];
        reinterpret_cast<uint4*>(p)[1] = u.q[1];
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) p[v] = src[v];
    }
}

// ------------------------------------------------------------------ NumPy scalar semantics
template <typename T> struct b2_is_float { static constexpr bool value = false; };
template <> struct b2_is_float<float> { static constexpr bool value = true; };
template <> struct b2_is_float<double> { static constexpr bool value = true; };

template <typename T> __device__ __forceinline__ bool b2_isnan(T v) {
    if constexpr (b2_is_float<T>::value) return v != v; else return false;
}

// Value conversion as the reference's NumPy performs it on x86-64: float -> int64 of a NaN or
// out-of-range value yields INT64_MIN (cvttsd2si "integer indefinite"), where CUDA's cvt saturates.
template <typename TO, typename FROM> __device__ __forceinline__ TO b2_cast(FROM v) {
    if constexpr (b2_is_float<FROM>::value && sizeof(TO) == 8 && !b2_is_float<TO>::value && (TO)(-1) < (TO)0) {
        return (v >= (FROM)-9223372036854775808.0 && v < (FROM)9223372036854775808.0) ? (TO)v : (TO)(-9223372036854775807LL - 1);
    } else {
        return (TO)v;
    }
}

// Python/NumPy floor division and modulo (sign follows the divisor; x // 0 == 0 for ints)
template <typename T> __device__ __forceinline__ T b2_floordiv_int(T a, T b) {
    if (b == 0) return 0;
    T q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}
template <typename T> __device__ __forceinline__ T b2_mod_int(T a, T b) {
    if (b == 0) return 0;
    T r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) r += b;
    return r;
}
template <typename T> __device__ __forceinline__ T b2_floordiv_uint(T a, T b) { return b == 0 ? 0 : a / b; }
template <typename T> __device__ __forceinline__ T b2_mod_uint(T a, T b) { return b == 0 ? 0 : a % b; }
__device__ __forceinline__ double b2_mod_f(double a, double b) {
    double r = fmod(a, b);
    if (b == 0.0) return r;               // nan
    if (r != 0.0) { if ((b < 0.0) != (r < 0.0)) r += b; }
    else r = copysign(0.0, b);
    return r;
}
__device__ __forceinline__ float b2_mod_f(float a, float b) {
    float r = fmodf(a, b);
    if (b == 0.0f) return r;
    if (r != 0.0f) { if ((b < 0.0f) != (r < 0.0f)) r += b; }
    else r = copysignf(0.0f, b);
    return r;
}
__device__ __forceinline__ double b2_floordiv_f(double a, double b) {
    if (b == 0.0) return a / b;
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div != 0.0) { double fl = floor(div); if (div - fl > 0.5) fl += 1.0; return fl; }
    return copysign(0.0, a / b);
}
__device__ __forceinline__ float b2_floordiv_f(float a, float b) {
    if (b == 0.0f) return a / b;
    float mod = fmodf(a, b);
    float div = (a - mod) / b;
    if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div -= 1.0f;
    if (div != 0.0f) { float fl = floorf(div); if (div - fl > 0.5f) fl += 1.0f; return fl; }
    return copysignf(0.0f, a / b);
}
// integer power by squaring, wrap-around like NumPy
template <typename T> __device__ __forceinline__ T b2_ipow(T base, i64 e) {
    T r = 1;
    while (e > 0) { if (e & 1) r *= base; base *= base; e >>= 1; }
    return r;
}
// np.maximum / np.minimum propagate NaN
template <typename T> __device__ __forceinline__ T b2_np_max(T a, T b) {
    if constexpr (b2_is_float<T>::value) return (a != a) ? a : ((b != b) ? b : (a >= b ? a : b));
    else return a >= b ? a : b;
}
template <typename T> __device__ __forceinline__ T b2_np_min(T a, T b) {
    if constexpr (b2_is_float<T>::value) return (a != a) ? a : ((b != b) ? b : (a <= b ? a : b));
    else return a <= b ? a : b;
}
template <typename T> __device__ __forceinline__ T b2_sign(T a) {
    if constexpr (b2_is_float<T>::value) return (a != a) ? a : (T)((a > 0) - (a < 0));
    else return (T)((a > 0) - (a < 0));
}

// ------------------------------------------------------------------ fp32 sin / cos
// The chain `sin(x)*2 + x**2` costs ~32 issue slots per element with libdevice's sinf
// (F2I range reduction, a branch and an inlined Payne-Hanek path PER ELEMENT), which caps a
// streaming kernel at ~50 % of HBM.  These versions keep libdevice's algorithm class --
// 3-term Cody-Waite reduction + minimax polynomial, no MUFU approximation -- in 13 slots:
//   * quotient by magic-number rounding (no F2I/I2F), reduction by pi so that ONE odd
//     polynomial on [-pi/2, pi/2] serves every quadrant and the sign is an XOR;
//   * arguments beyond the Cody-Waite range (|x| > 105615, inf) are detected once per
//     vector by the caller (max of |x|) and that vector is redone with libdevice's sinf/cosf.
// Max error 1.64 ulp vs the correctly rounded result on the fast path (tests/test_gpu_kernels.py
// checks it exhaustively against fp64); NumPy's own float32 sin is specified to < 1.5 ulp.
#define B2_SINCOS_BIG 105615.0f
__device__ __forceinline__ float b2_sin_poly(float r) {          // |r| <= pi/2
    const float s = r * r;
    float p = 2.605750751172309e-06f;
    p = fmaf(p, s, -0.00019809573132079095f);
    p = fmaf(p, s, 0.00833306647837162f);
    p = fmaf(p, s, -0.16666659712791443f);
    return fmaf(r * s, p, r);
}
__device__ __forceinline__ float b2_sinf_fast(float x, float& big) {
    big = fmaxf(big, fabsf(x));
    const float jm = fmaf(x, 0.318309886183790672f, 12582912.0f);   // 1.5 * 2^23: low bits = round(x/pi)
    const float j = jm - 12582912.0f;
    float r = fmaf(j, -3.1415925025939941406f, x);
    r = fmaf(j, -1.5099578831723192707e-07f, r);
    r = fmaf(j, -1.0780605906948476785e-14f, r);
    return __int_as_float(__float_as_int(b2_sin_poly(r)) ^ (__float_as_int(jm) << 31));
}
__device__ __forceinline__ float b2_cosf_fast(float x, float& big) {
    big = fmaxf(big, fabsf(x));
    // cos(x) = sin(x + pi/2):  j = round(x/pi + 1/2),  r = x - (j - 1/2) pi,  cos(x) = (-1)^j sin(r)
    const float jm = fmaf(x, 0.318309886183790672f, 0.5f) + 12582912.0f;
    const float j = (jm - 12582912.0f) - 0.5f;
    float r = fmaf(j, -3.1415925025939941406f, x);
    r = fmaf(j, -1.5099578831723192707e-07f, r);
    r = fmaf(j, -1.0780605906948476785e-14f, r);
    return __int_as_float(__float_as_int(b2_sin_poly(r)) ^ (__float_as_int(jm) << 31));
}
// ---- packed fp32 x 2 (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot for two lanes.  The chain
// kernels are issue-bound (ncu: issue-active ~80 %), so halving the slots of the polynomial and
// of the +,-,* operators moves them back under the HBM roofline.  Values are (lo, hi) pairs in one
// 64-bit register; consecutive elements of a 128-bit load are already such pairs.
typedef unsigned long long b2f2;
__device__ __forceinline__ b2f2 b2_pk(float a, float b) { b2f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void b2_upk(b2f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ b2f2 b2_bc(float c) { return b2_pk(c, c); }
__device__ __forceinline__ b2f2 b2_fma2(b2f2 a, b2f2 b, b2f2 c) { b2f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ b2f2 b2_add2(b2f2 a, b2f2 b) { b2f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ b2f2 b2_sub2(b2f2 a, b2f2 b) { b2f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ b2f2 b2_mul2(b2f2 a, b2f2 b) { b2f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ b2f2 b2_neg2(b2f2 a) { return a ^ 0x8000000080000000ULL; }
__device__ __forceinline__ b2f2 b2_abs2(b2f2 a) { return a & 0x7fffffff7fffffffULL; }
__device__ __forceinline__ float b2_maxabs2(b2f2 v) { float a, b; b2_upk(v, a, b); return fmaxf(fabsf(a), fabsf(b)); }
__device__ __forceinline__ b2f2 b2_sin_poly2(b2f2 r) {
    const b2f2 s = b2_mul2(r, r);
    b2f2 p = b2_fma2(b2_bc(2.605750751172309e-06f), s, b2_bc(-0.00019809573132079095f));
    p = b2_fma2(p, s, b2_bc(0.00833306647837162f));
    p = b2_fma2(p, s, b2_bc(-0.16666659712791443f));
    return b2_fma2(b2_mul2(r, s), p, r);
}
__device__ __forceinline__ b2f2 b2_sinf_fast2(b2f2 x, float& big) {
    big = fmaxf(big, b2_maxabs2(x));
    const b2f2 jm = b2_fma2(x, b2_bc(0.318309886183790672f), b2_bc(12582912.0f));
    const b2f2 j = b2_add2(jm, b2_bc(-12582912.0f));
    b2f2 r = b2_fma2(j, b2_bc(-3.1415925025939941406f), x);
    r = b2_fma2(j, b2_bc(-1.5099578831723192707e-07f), r);
    r = b2_fma2(j, b2_bc(-1.0780605906948476785e-14f), r);
    return b2_sin_poly2(r) ^ ((jm << 31) & 0x8000000080000000ULL);       // (-1)^j per lane
}
__device__ __forceinline__ b2f2 b2_cosf_fast2(b2f2 x, float& big) {
    big = fmaxf(big, b2_maxabs2(x));
    const b2f2 jm = b2_add2(b2_fma2(x, b2_bc(0.318309886183790672f), b2_bc(0.5f)), b2_bc(12582912.0f));
    const b2f2 j = b2_add2(b2_add2(jm, b2_bc(-12582912.0f)), b2_bc(-0.5f));
    b2f2 r = b2_fma2(j, b2_bc(-3.1415925025939941406f), x);
    r = b2_fma2(j, b2_bc(-1.5099578831723192707e-07f), r);
    r = b2_fma2(j, b2_bc(-1.0780605906948476785e-14f), r);
    return b2_sin_poly2(r) ^ ((jm << 31) & 0x8000000080000000ULL);
}
__device__ __noinline__ float b2_sinf_slow(float x) { return sinf(x); }   // libdevice, full range
__device__ __noinline__ float b2_cosf_slow(float x) { return cosf(x); }

// ------------------------------------------------------------------ accumulators
// An accumulator `A` offers: init(), add(value, index), merge(other) [other comes LATER
// in index order unless the op is commutative], and Packed <-> A for the two-stage
// partials.  Merges are deterministic: fixed shuffle pattern, fixed tile order.

template <typename T> __device__ __forceinline__ T b2_shfl_down(T v, int off, int width) {
    return __shfl_down_sync(0xffffffffu, v, off, width);
}
__device__ __forceinline__ unsigned char b2_shfl_down(unsigned char v, int off, int width) {
    return (unsigned char)__shfl_down_sync(0xffffffffu, (int)v, off, width);
}
__device__ __forceinline__ signed char b2_shfl_down(signed char v, int off, int width) {
    return (signed char)__shfl_down_sync(0xffffffffu, (int)v, off, width);
}
__device__ __forceinline__ short b2_shfl_down(short v, int off, int width) {
    return (short)__shfl_down_sync(0xffffffffu, (int)v, off, width);
}
__device__ __forceinline__ unsigned short b2_shfl_down(unsigned short v, int off, int width) {
    return (unsigned short)__shfl_down_sync(0xffffffffu, (int)v, off, width);
}
__device__ __forceinline__ bool b2_shfl_down(bool v, int off, int width) {
    return (bool)__shfl_down_sync(0xffffffffu, (int)v, off, width);
}

template <typename T, typename ACC>
struct B2AccSum {   // np.sum(x, dtype=ACC)  (_chunk.py:172; ints widen to 64 bit)
    typedef ACC Packed;
    ACC s;
    __device__ __forceinline__ void init() { s = (ACC)0; }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T v, int, int) { s += b2_cast<ACC>(v); }
    __device__ __forceinline__ void merge(const B2AccSum& o) { s += o.s; }
    __device__ __forceinline__ void shfl(int off, int width) { B2AccSum o; o.s = b2_shfl_down(s, off, width); merge(o); }
    __device__ __forceinline__ void lane_merge(const B2AccSum& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { return s; }
    __device__ __forceinline__ void unpack(const Packed& p) { s = p; }
};
template <typename T, typename ACC>
struct B2AccProd {
    typedef ACC Packed;
    ACC s;
    __device__ __forceinline__ void init() { s = (ACC)1; }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T v, int, int) { s *= b2_cast<ACC>(v); }
    __device__ __forceinline__ void merge(const B2AccProd& o) { s *= o.s; }
    __device__ __forceinline__ void shfl(int off, int width) { B2AccProd o; o.s = b2_shfl_down(s, off, width); merge(o); }
    __device__ __forceinline__ void lane_merge(const B2AccProd& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { return s; }
    __device__ __forceinline__ void unpack(const Packed& p) { s = p; }
};
template <typename T, bool ALL>
struct B2AccAnyAll {   // np.any / np.all -> bool
    typedef unsigned char Packed;
    unsigned char s;
    __device__ __forceinline__ void init() { s = ALL ? 1 : 0; }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T v, int, int) { bool t = (v != (T)0); s = ALL ? (s & (unsigned char)t) : (s | (unsigned char)t); }
    __device__ __forceinline__ void merge(const B2AccAnyAll& o) { s = ALL ? (s & o.s) : (s | o.s); }
    __device__ __forceinline__ void shfl(int off, int width) { B2AccAnyAll o; o.s = b2_shfl_down(s, off, width); merge(o); }
    __device__ __forceinline__ void lane_merge(const B2AccAnyAll& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { return s; }
    __device__ __forceinline__ void unpack(const Packed& p) { s = p; }
};
// np.nanmin / np.nanmax (_chunk.py:189-190): NaNs are skipped, an all-NaN slice gives NaN
template <typename T> __device__ __forceinline__ T b2_nan_max(T a, T b) {
    if constexpr (b2_is_float<T>::value) return (a != a) ? b : ((b != b) ? a : (a >= b ? a : b));
    else return a >= b ? a : b;
}
template <typename T> __device__ __forceinline__ T b2_nan_min(T a, T b) {
    if constexpr (b2_is_float<T>::value) return (a != a) ? b : ((b != b) ? a : (a <= b ? a : b));
    else return a <= b ? a : b;
}
template <typename T, bool ISMAX>
struct B2AccNanMinMax {
    struct Packed { T m; int has; };
    T m; bool has;
    __device__ __forceinline__ void init() { has = false; m = (T)0; }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T v, int, int) {
        if (!has) { m = v; has = true; return; }
        m = ISMAX ? b2_nan_max(m, v) : b2_nan_min(m, v);
    }
    __device__ __forceinline__ void merge(const B2AccNanMinMax& o) {
        if (!o.has) return;
        if (!has) { m = o.m; has = true; return; }
        m = ISMAX ? b2_nan_max(m, o.m) : b2_nan_min(m, o.m);
    }
    __device__ __forceinline__ void shfl(int off, int width) {
        B2AccNanMinMax o; o.m = b2_shfl_down(m, off, width); o.has = b2_shfl_down(has, off, width); merge(o);
    }
    __device__ __forceinline__ void lane_merge(const B2AccNanMinMax& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { Packed p; p.m = m; p.has = has ? 1 : 0; return p; }
    __device__ __forceinline__ void unpack(const Packed& p) { m = p.m; has = (p.has != 0); }
};
template <typename T, bool ISMAX>
struct B2AccMinMax {   // np.min / np.max: NaN propagates (chunk_min/chunk_max _common.py:92-105)
    struct Packed { T m; int has; };
    T m; bool has;
    __device__ __forceinline__ void init() { has = false; m = (T)0; }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T v, int, int) {
        if (!has) { m = v; has = true; return; }
        m = ISMAX ? b2_np_max(m, v) : b2_np_min(m, v);
    }
    __device__ __forceinline__ void merge(const B2AccMinMax& o) {
        if (!o.has) return;
        if (!has) { m = o.m; has = true; return; }
        m = ISMAX ? b2_np_max(m, o.m) : b2_np_min(m, o.m);
    }
    __device__ __forceinline__ void shfl(int off, int width) {
        B2AccMinMax o; o.m = b2_shfl_down(m, off, width); o.has = b2_shfl_down(has, off, width); merge(o);
    }
    __device__ __forceinline__ void lane_merge(const B2AccMinMax& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { Packed p; p.m = m; p.has = has ? 1 : 0; return p; }
    __device__ __forceinline__ void unpack(const Packed& p) { m = p.m; has = (p.has != 0); }
};
// identities for min / max style folds
template <typename T> struct b2_limits;
template <> struct b2_limits<float> { __device__ static float lowest() { return __int_as_float(0xff800000); } __device__ static float highest() { return __int_as_float(0x7f800000); } };
template <> struct b2_limits<double> { __device__ static double lowest() { return __longlong_as_double(0xfff0000000000000LL); } __device__ static double highest() { return __longlong_as_double(0x7ff0000000000000LL); } };
template <> struct b2_limits<bool> { __device__ static bool lowest() { return false; } __device__ static bool highest() { return true; } };
template <> struct b2_limits<signed char> { __device__ static signed char lowest() { return -128; } __device__ static signed char highest() { return 127; } };
template <> struct b2_limits<unsigned char> { __device__ static unsigned char lowest() { return 0; } __device__ static unsigned char highest() { return 255; } };
template <> struct b2_limits<short> { __device__ static short lowest() { return -32768; } __device__ static short highest() { return 32767; } };
template <> struct b2_limits<unsigned short> { __device__ static unsigned short lowest() { return 0; } __device__ static unsigned short highest() { return 65535; } };
template <> struct b2_limits<int> { __device__ static int lowest() { return -2147483647 - 1; } __device__ static int highest() { return 2147483647; } };
template <> struct b2_limits<unsigned int> { __device__ static unsigned int lowest() { return 0u; } __device__ static unsigned int highest() { return 4294967295u; } };
template <> struct b2_limits<long long> { __device__ static long long lowest() { return -9223372036854775807LL - 1; } __device__ static long long highest() { return 9223372036854775807LL; } };
template <> struct b2_limits<unsigned long long> { __device__ static unsigned long long lowest() { return 0ULL; } __device__ static unsigned long long highest() { return 18446744073709551615ULL; } };

// index-carrying argmin/argmax with np.argmax semantics: first occurrence wins ties,
// the first NaN wins outright (arg_chunk _common.py:704-732; keepdims_wrapper _chunk.py:137).
// Thread-local phase: a thread visits its elements in increasing index order, so "first
// occurrence" is simply "replace only when strictly better" -- branch-free selects on a 32-bit
// step counter; the 64-bit index is rebuilt once in finish_arg().  Merged phase: (v, i) with
// lexicographic (value, index) merges.
template <typename T, bool ISMAX>
struct B2AccArg {
    struct Packed { T v; i64 i; };
    T v; i64 i;                 // merged phase; i < 0: empty
    int bk, bl; bool bnan;      // thread-local phase: step / lane of the best element, best-is-NaN flag
    __device__ __forceinline__ void init() {
        v = ISMAX ? b2_limits<T>::lowest() : b2_limits<T>::highest();
        i = -1; bk = -1; bl = 0; bnan = false;
    }
    __device__ __forceinline__ void prime(T) {}
    __device__ __forceinline__ void add(T x, int k, int lane) {
        bool better = ISMAX ? (x > v) : (x < v);
        bool xnan = false;
        if constexpr (b2_is_float<T>::value) { xnan = (x != x); better = better || xnan; }
        const bool take = better && !bnan;
        v = take ? x : v;
        bk = take ? k : bk;
        bl = take ? lane : bl;
        if constexpr (b2_is_float<T>::value) bnan = bnan || xnan;
    }
    // any: the thread saw at least one element; element (k, lane) has index idx0 + k * istep + lane
    __device__ __forceinline__ void finish_arg(bool any, i64 idx0, i64 istep) {
        // bk < 0 with elements seen: every element equalled the identity -> the first one wins
        i = any ? (bk < 0 ? idx0 : idx0 + (i64)bk * istep + bl) : -1;
    }
    // `a` strictly better than `b` (both valid)?
    __device__ __forceinline__ static bool better(T av, i64 ai, T bv, i64 bi) {
        bool an = b2_isnan(av), bn = b2_isnan(bv);
        if (an || bn) { if (an && bn) return ai < bi; return an; }
        if (av == bv) return ai < bi;
        return ISMAX ? (av > bv) : (av < bv);
    }
    __device__ __forceinline__ void merge(const B2AccArg& o) {
        if (o.i < 0) return;
        if (i < 0 || better(o.v, o.i, v, i)) { v = o.v; i = o.i; }
    }
    __device__ __forceinline__ void shfl(int off, int width) {
        B2AccArg o; o.v = b2_shfl_down(v, off, width); o.i = b2_shfl_down(i, off, width); merge(o);
    }
    __device__ __forceinline__ void lane_merge(const B2AccArg& o) { merge(o); }
    __device__ __forceinline__ void lane_shfl(int off, int width) { shfl(off, width); }
    __device__ __forceinline__ void lane_finish() {}
    __device__ __forceinline__ Packed pack() const { Packed p; p.v = v; p.i = i; return p; }
    __device__ __forceinline__ void unpack(const Packed& p) { v = p.v; i = p.i; bk = -1; bl = 0; bnan = false; }
};
// Single-pass second moment.  Per thread: sums of (x-K) and (x-K)^2 around a pivot K
// taken from the data (kills the catastrophic cancellation of the naive sum of squares);
// across threads / tiles / blocks: Chan's pairwise merge of (n, mean, M2) in fp64 -- the
// same algebra as moment_combine (_common.py:415-453), which the reference applies
// between blocks after a two-pass moment_chunk (:368-404).
template <typename T, typename W>   // W: working type (float for fp32 data, double otherwise)
struct B2AccMoment {
    struct Packed { double n, mean, m2; };
    W K, s1, s2; int cnt;        // thread-local phase
    double n, mean, m2;          // merged phase
    __device__ __forceinline__ void init() { K = (W)0; s1 = (W)0; s2 = (W)0; cnt = 0; n = 0.0; mean = 0.0; m2 = 0.0; }
    __device__ __forceinline__ void prime(T v) { K = (W)v; }
    __device__ __forceinline__ void add(T v, int, int) { W d = (W)v - K; s1 += d; s2 = fma(d, d, s2); }
    // fold the thread-local sums (over `count` elements, known from the loop bounds) into (n, mean, M2)
    __device__ __forceinline__ void finish_local(i64 count) {
        cnt = (int)count;
        if (cnt > 0) {
            double dn = (double)cnt, a = (double)s1, b = (double)s2;
            Packed p; p.n = dn; p.mean = (double)K + a / dn; p.m2 = b - a * a / dn;
            if (p.m2 < 0.0) p.m2 = 0.0;
            chan(p.n, p.mean, p.m2);
            s1 = (W)0; s2 = (W)0; cnt = 0;
        }
    }
    __device__ __forceinline__ void chan(double on, double omean, double om2) {
        if (on == 0.0) return;
        if (n == 0.0) { n = on; mean = omean; m2 = om2; return; }
        double tot = n + on, delta = omean - mean;
        mean = mean + delta * (on / tot);
        m2 = m2 + om2 + delta * delta * (n * on / tot);
        n = tot;
    }
    __device__ __forceinline__ void merge(const B2AccMoment& o) { chan(o.n, o.mean, o.m2); }
    // Lanes that share ONE pivot K (a whole tile in mode RC, a whole row in mode C) reduce the raw
    // shifted sums -- plain adds, no divisions -- and convert once: (n, mean, m2) temporarily hold
    // (count, sum(x-K), sum((x-K)^2)).
    __device__ __forceinline__ void to_raw(i64 count) { n = (double)count; mean = (double)s1; m2 = (double)s2; }
    __device__ __forceinline__ void lane_merge(const B2AccMoment& o) { n += o.n; mean += o.mean; m2 += o.m2; }
    __device__ __forceinline__ void lane_shfl(int off, int width) {
        n += b2_shfl_down(n, off, width); mean += b2_shfl_down(mean, off, width); m2 += b2_shfl_down(m2, off, width);
    }
    __device__ __forceinline__ void lane_finish() {
        if (n > 0.0) {
            const double a = mean, bsum = m2;
            mean = (double)K + a / n;
            m2 = bsum - a * a / n;
            if (m2 < 0.0) m2 = 0.0;
        } else { mean = 0.0; m2 = 0.0; }
    }
    __device__ __forceinline__ void shfl(int off, int width) {
        double on = b2_shfl_down(n, off, width), om = b2_shfl_down(mean, off, width), o2 = b2_shfl_down(m2, off, width);
        chan(on, om, o2);
    }
    __device__ __forceinline__ Packed pack() const { Packed p; p.n = n; p.mean = mean; p.m2 = m2; return p; }
    __device__ __forceinline__ void unpack(const Packed& p) { init(); n = p.n; mean = p.mean; m2 = p.m2; }
};

template <int REDOP, typename T, typename ACC> struct B2AccSel;
template <typename T, typename ACC> struct B2AccSel<B2R_SUM, T, ACC> { typedef B2AccSum<T, ACC> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_PROD, T, ACC> { typedef B2AccProd<T, ACC> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_MIN, T, ACC> { typedef B2AccMinMax<T, false> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_MAX, T, ACC> { typedef B2AccMinMax<T, true> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_NANMIN, T, ACC> { typedef B2AccNanMinMax<T, false> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_NANMAX, T, ACC> { typedef B2AccNanMinMax<T, true> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_ARGMIN, T, ACC> { typedef B2AccArg<T, false> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_ARGMAX, T, ACC> { typedef B2AccArg<T, true> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_MOMENT, T, ACC> { typedef B2AccMoment<T, ACC> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_ANY, T, ACC> { typedef B2AccAnyAll<T, false> type; };
template <typename T, typename ACC> struct B2AccSel<B2R_ALL, T, ACC> { typedef B2AccAnyAll<T, true> type; };

// final store of one reduced element
template <int REDOP, typename T, typename ACC, typename A>
__device__ __forceinline__ void b2_store_result(const B2Block& blk, i64 o, A& a, i64 idx_fix) {
    if constexpr (REDOP == B2R_SUM || REDOP == B2R_PROD) {
        ((ACC*)blk.out0)[o] = a.s;
    } else if constexpr (REDOP == B2R_MIN || REDOP == B2R_MAX || REDOP == B2R_NANMIN || REDOP == B2R_NANMAX) {
        ((T*)blk.out0)[o] = a.m;
    } else if constexpr (REDOP == B2R_ARGMIN || REDOP == B2R_ARGMAX) {
        ((T*)blk.out0)[o] = a.v;
        ((i64*)blk.out1)[o] = a.i + idx_fix;
    } else if constexpr (REDOP == B2R_MOMENT) {
        double* q = (double*)blk.out0 + 3 * o;
        q[0] = a.n; q[1] = a.mean; q[2] = a.m2;
    } else {
        ((unsigned char*)blk.out0)[o] = a.s;
    }
}

// local flat index of a ravelled arg reduction -> index into the WHOLE array
// (arg_chunk ravel branch, _common.py:709-713: unravel_index + offset + ravel_multi_index)
__device__ __forceinline__ i64 b2_ravel_fix(const B2Block& blk, i64 local) {
    i64 coords[B2_MAX_ND];
    i64 rem = local;
    for (int d = blk.arg_ndim - 1; d >= 0; --d) { coords[d] = rem % blk.arg_shape[d]; rem /= blk.arg_shape[d]; }
    i64 g = 0;
    for (int d = 0; d < blk.arg_ndim; ++d) g = g * blk.arg_total[d] + (coords[d] + blk.arg_start[d]);
    return g;
}

// ------------------------------------------------------------------ tile lookup
__device__ __forceinline__ int b2_find_block(const B2Block* __restrict__ blocks, int nblocks, i64 tile) {
    if (nblocks == 1) return 0;
    {   // uniform launches (all blocks tiled alike -- the common case): one division finds the block,
        // two independent loads confirm it, instead of log2(nblocks) dependent L2 round trips
        const i64 per = blocks[1].tile_begin;
        i64 g = tile / per;
        if (g > nblocks - 1) g = nblocks - 1;
        const i64 lo_t = blocks[g].tile_begin;
        const i64 hi_t = (g + 1 < nblocks) ? blocks[g + 1].tile_begin : tile + 1;
        if (lo_t <= tile && tile < hi_t) return (int)g;
    }
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (blocks[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------ the fused kernel body
// partials written by other CTAs: read through L2 (never a stale L1 line)
template <typename P> __device__ __forceinline__ P b2_load_cg(const P* p) {
    union { P v; u32 w[(sizeof(P) + 3) / 4]; unsigned char c[sizeof(P)]; } u;
    if constexpr (sizeof(P) % 4 == 0) {
#pragma unroll
        for (int i = 0; i < (int)(sizeof(P) / 4); ++i) u.w[i] = __ldcg(reinterpret_cast<const u32*>(p) + i);
    } else {
#pragma unroll
        for (int i = 0; i < (int)sizeof(P); ++i) u.c[i] = __ldcg(reinterpret_cast<const unsigned char*>(p) + i);
    }
    return u.v;
}

// Streaming loop shared by every mode: walk `n` V-wide vectors reachable through `P`
// (one step = one vector), evaluate the chain and hand each result to `consume(state, k, o)`.
//  * Rolling prefetch: a thread keeps U vector loads in flight at all times -- the registers
//    of vector u are refilled with the NEXT batch's vector u right after the chain consumed
//    them -- so HBM latency is covered by this thread's own compute, not only by other warps.
//  * Chains with fp32 sin/cos run the fast versions and track max|arg|.  If an argument of
//    the batch left the Cody-Waite range (rare) the state is rolled back to the checkpoint
//    taken at the start of the batch and the batch is redone from memory with libdevice:
//    one predictable branch per U*V elements instead of one per element.
template <typename Chain, int V, int U, typename S, typename F>
__device__ __forceinline__ void b2_stream(typename Chain::Ptrs& P, const int n, const B2Scalars& sc, S& st, F consume) {
    typedef typename Chain::out_t T;
    int it = 0;
    if (n >= U) {
        typename Chain::Regs g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) Chain::load(P, u, g[u]);
        Chain::advance(P, U);
        for (; it + U <= n; it += U) {
            const bool more = it + 2 * U <= n;
            S saved;
            if constexpr (Chain::HAS_SLOW) saved = st;
            float big = 0.0f;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                T o[V];
                Chain::compute(g[u], sc, o, big);
                if (more) Chain::load(P, u, g[u]);
                consume(st, it + u, o);
            }
            if (more) Chain::advance(P, U);
            if constexpr (Chain::HAS_SLOW) {
                if (big > B2_SINCOS_BIG) {
                    st = saved;
                    const int back = more ? 2 * U : U;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        typename Chain::Regs gs; T o[V];
                        Chain::load(P, u - back, gs);
                        Chain::compute_slow(gs, sc, o);
                        consume(st, it + u, o);
                    }
                }
            }
        }
    }
    for (; it < n; ++it) {       // tail (< U vectors): exact chain, nothing speculative
        typename Chain::Regs gs; T o[V];
        Chain::load(P, 0, gs); Chain::advance(P, 1);
        Chain::compute_slow(gs, sc, o);
        consume(st, it, o);
    }
}

struct B2NoState {};
template <typename A, int V> struct B2AccState { A acc[V]; };
// packed thread-local state of the fp32 moment accumulators: lanes (v, v+1) share one register pair
template <int V> struct B2MomState2 { b2f2 K[V / 2], s1[V / 2], s2[V / 2]; };

// Chain supplies:
//   out_t; Regs; Ptrs (one typed pointer + element step per input); HAS_SLOW;
//   setup_rows(blk, b, r, c, rstep, Ptrs&)  -- point at (b, r, c), one step = rstep rows down;
//   setup_cols(blk, b, r, c, cstep, Ptrs&)  -- point at (b, r, c), one step = cstep columns right;
//   load(Ptrs, k, Regs&)   -- the V-wide load k steps ahead;   advance(Ptrs&, n);
//   compute(Regs, scalars, out_t(&)[V], float& big) [fast], compute_slow(Regs, scalars, out_t(&)[V]).
template <typename Chain, int MODE, int REDOP, int V, int TX, int TY, int RPT, int U, typename ACC>
__device__ __forceinline__ void b2_run(const B2Block* __restrict__ blocks, int nblocks, const B2Scalars& sc) {
    typedef typename Chain::out_t T;
    constexpr int NT = TX * TY;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;

    __shared__ B2Block sblk;
    {
        const i64 tile = blockIdx.x;
        const int bi = b2_find_block(blocks, nblocks, tile);
        const int nw = (int)(sizeof(B2Block) / 4);
        const u32* src = reinterpret_cast<const u32*>(blocks + bi);
        u32* dst = reinterpret_cast<u32*>(&sblk);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        __syncthreads();
    }
    const B2Block& blk = sblk;
    i64 t = (i64)blockIdx.x - blk.tile_begin;
    const i64 tc = t % blk.tiles_c; t /= blk.tiles_c;
    const i64 tr = t % blk.tiles_r; t /= blk.tiles_r;
    const i64 b = t;
    const i64 R = blk.R, C = blk.C;
    const i64 r0 = tr * RPT;
    const i64 rend = (r0 + RPT < R) ? (r0 + RPT) : R;

    if constexpr (MODE == B2M_EW) {
        const i64 c = (tc * TX + tx) * V;
        const i64 first = r0 + ty;
        if (c < C && first < rend) {
            const int nrows = (int)((rend - first + TY - 1) / TY);
            T* const outp = (T*)blk.out0 + (b * R + first) * C + c;
            const i64 ostep = (i64)TY * C;
            typename Chain::Ptrs P;
            Chain::setup_rows(blk, b, first, c, TY, P);
            B2NoState none;
            b2_stream<Chain, V, U>(P, nrows, sc, none,
                [&](B2NoState&, int k, const T (&o)[V]) { b2_store_vec<T, V>(outp + k * ostep, o); });
        }
        return;
    } else {
        typedef typename B2AccSel<REDOP, T, ACC>::type A;
        typedef typename A::Packed P_t;
        __shared__ bool last;
        constexpr bool WANT_IDX = (REDOP == B2R_ARGMIN || REDOP == B2R_ARGMAX);

        if constexpr (MODE == B2M_R || MODE == B2M_RC) {
            // ---- accumulate this thread's rows of the tile
            const i64 c = (tc * TX + tx) * V;
            const i64 first = r0 + ty;
            B2AccState<A, V> st;
            A (&acc)[V] = st.acc;
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v].init();
            const int nrows = (c < C && first < rend) ? (int)((rend - first + TY - 1) / TY) : 0;
            if (nrows > 0) {
                typename Chain::Ptrs P;
                Chain::setup_rows(blk, b, first, c, TY, P);
                if constexpr (REDOP == B2R_MOMENT) {
                    // pivot: mode R -- the first element of each column lane; mode RC -- ONE pivot
                    // for the whole tile (its first element), so the CTA reduces plain sums
                    typename Chain::Regs g0; T o0[V];
                    if constexpr (MODE == B2M_R) {
                        Chain::load(P, 0, g0);
                    } else {
                        typename Chain::Ptrs P0;
                        Chain::setup_rows(blk, b, r0, tc * TX * V, TY, P0);
                        Chain::load(P0, 0, g0);
                    }
                    Chain::compute_slow(g0, sc, o0);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v].prime(MODE == B2M_R ? o0[v] : o0[0]);
                }
                if constexpr (REDOP == B2R_MOMENT && sizeof(T) == 4 && sizeof(ACC) == 4 && b2_is_float<T>::value && V % 2 == 0) {
                    // fp32 moments: d = v - K, s1 += d, s2 += d*d as FADD2 / FADD2 / FFMA2 on lane pairs
                    B2MomState2<V> ms;
#pragma unroll
                    for (int h = 0; h < V / 2; ++h) {
                        ms.K[h] = b2_pk(acc[2 * h].K, acc[2 * h + 1].K);
                        ms.s1[h] = 0ULL; ms.s2[h] = 0ULL;
                    }
                    b2_stream<Chain, V, U>(P, nrows, sc, ms,
                        [&](B2MomState2<V>& s_, int, const T (&o)[V]) {
#pragma unroll
                            for (int h = 0; h < V / 2; ++h) {
                                const b2f2 d = b2_sub2(b2_pk(o[2 * h], o[2 * h + 1]), s_.K[h]);
                                s_.s1[h] = b2_add2(s_.s1[h], d);
                                s_.s2[h] = b2_fma2(d, d, s_.s2[h]);
                            }
                        });
#pragma unroll
                    for (int h = 0; h < V / 2; ++h) {
                        b2_upk(ms.s1[h], acc[2 * h].s1, acc[2 * h + 1].s1);
                        b2_upk(ms.s2[h], acc[2 * h].s2, acc[2 * h + 1].s2);
                    }
                } else {
                    b2_stream<Chain, V, U>(P, nrows, sc, st,
                        [&](B2AccState<A, V>& s_, int k, const T (&o)[V]) {
#pragma unroll
                            for (int v = 0; v < V; ++v) s_.acc[v].add(o[v], k, 0);
                        });
                }
            }
            if constexpr (WANT_IDX) {
                // element k of accumulator v: row first + k*TY (mode R) / flat (first + k*TY)*C + c + v (mode RC)
#pragma unroll
                for (int v = 0; v < V; ++v)
                    acc[v].finish_arg(nrows > 0, (MODE == B2M_R) ? first : (first * C + c + v),
                                      (MODE == B2M_R) ? (i64)TY : (i64)TY * C);
            }
            if constexpr (REDOP == B2R_MOMENT) {
#pragma unroll
                for (int v = 0; v < V; ++v) { if constexpr (MODE == B2M_R) acc[v].finish_local(nrows); else acc[v].to_raw(nrows); }
            }

            if constexpr (MODE == B2M_R) {
                // ---- fold the TY row-lanes of each column (fixed order), then tiles_r partials
                if constexpr (TY > 1) {
                    __shared__ P_t sm[NT * V];
#pragma unroll
                    for (int v = 0; v < V; ++v) sm[(ty * TX + tx) * V + v] = acc[v].pack();
                    __syncthreads();
                    if (ty == 0) {
                        const int ny = (rend - r0 < TY) ? (int)(rend - r0) : TY;   // lanes that saw a row
                        for (int y = 1; y < ny; ++y) {
#pragma unroll
                            for (int v = 0; v < V; ++v) { A o; o.unpack(sm[(y * TX + tx) * V + v]); acc[v].merge(o); }
                        }
                    }
                }
                const i64 tiles_r = blk.tiles_r;
                const i64 Cpad = blk.tiles_c * TX * V;
                if (tiles_r == 1) {
                    if (ty == 0 && c < C) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            i64 fx = blk.arg_offset;
                            if constexpr (WANT_IDX) { if (blk.arg_ndim > 0) fx = b2_ravel_fix(blk, acc[v].i) - acc[v].i; }
                            b2_store_result<REDOP, T, ACC>(blk, b * C + c + v, acc[v], fx);
                        }
                    }
                    return;
                }
                P_t* work = (P_t*)blk.work;
                if (ty == 0 && c < C) {
#pragma unroll
                    for (int v = 0; v < V; ++v) work[(b * tiles_r + tr) * Cpad + c + v] = acc[v].pack();
                    __threadfence();
                }
                __syncthreads();
                if (tid == 0) {
                    u32* ctr = blk.counter + (b * blk.tiles_c + tc);
                    u32 prev = atomicAdd(ctr, 1u);
                    last = (prev == (u32)(tiles_r - 1));
                    if (last) *ctr = 0u;      // self-reset: the workspace is reusable
                }
                __syncthreads();
                if (!last) return;
                __threadfence();
                if (ty == 0 && c < C) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        // partials are fetched eight at a time (independent L2 loads), merged in tile order:
                        // the fold is ONE thread deep per column, so its duration is the load chain
                        // (52 row tiles x ~0.5 us when fetched one by one)
                        A tot; tot.unpack(b2_load_cg(&work[(b * tiles_r) * Cpad + c + v]));
                        for (i64 k0 = 1; k0 < tiles_r; k0 += 8) {
                            P_t raw[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (k0 + u < tiles_r) raw[u] = b2_load_cg(&work[(b * tiles_r + k0 + u) * Cpad + c + v]);
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (k0 + u < tiles_r) { A part; part.unpack(raw[u]); tot.merge(part); }
                        }
                        b2_store_result<REDOP, T, ACC>(blk, b * C + c + v, tot, blk.arg_offset);
                    }
                }
                return;
            } else {
                // ---- MODE RC: one value per tile, then all tiles of (block, b) in order
                A a = acc[0];
#pragma unroll
                for (int v = 1; v < V; ++v) a.lane_merge(acc[v]);
                // warp reduce (lower lane = earlier), then across warps in order
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) a.lane_shfl(off, 32);
                constexpr int NW = (NT + 31) / 32;
                __shared__ A smw[NW];
                const int lane = tid & 31, wid = tid >> 5;
                if constexpr (NW > 1) {
                    if (lane == 0) smw[wid] = a;
                    __syncthreads();
                    if (tid == 0) { for (int w = 1; w < NW; ++w) a.lane_merge(smw[w]); }
                }
                if (tid == 0) a.lane_finish();
                const i64 ntile = blk.tiles_r * blk.tiles_c;
                const i64 tslot = tr * blk.tiles_c + tc;
                i64 fix = 0;
                if (ntile == 1) {
                    if (tid == 0) {
                        if constexpr (WANT_IDX) fix = (blk.arg_ndim > 0) ? (b2_ravel_fix(blk, a.i) - a.i) : blk.arg_offset;
                        b2_store_result<REDOP, T, ACC>(blk, b, a, fix);
                    }
                    return;
                }
                P_t* work = (P_t*)blk.work;
                if (tid == 0) {
                    work[b * ntile + tslot] = a.pack();
                    __threadfence();
                    u32* ctr = blk.counter + b;
                    u32 prev = atomicAdd(ctr, 1u);
                    last = (prev == (u32)(ntile - 1));
                    if (last) *ctr = 0u;
                }
                __syncthreads();
                if (!last) return;
                __threadfence();
                // last CTA: fold the tile partials.  Lanes take contiguous runs of tiles so
                // the fold order stays "earlier tiles first".
                A tot; tot.init();
                {
                    const i64 per = (ntile + NT - 1) / NT;
                    const i64 k0 = (i64)tid * per;
                    const i64 k1 = (k0 + per < ntile) ? (k0 + per) : ntile;
                    for (i64 kb = k0; kb < k1; kb += 4) {      // four independent loads, merged in order
                        P_t raw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) if (kb + u < k1) raw[u] = b2_load_cg(&work[b * ntile + kb + u]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) if (kb + u < k1) { A part; part.unpack(raw[u]); tot.merge(part); }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) tot.shfl(off, 32);
                if constexpr (NW > 1) {
                    __syncthreads();
                    if (lane == 0) smw[wid] = tot;
                    __syncthreads();
                    if (tid == 0) { for (int w = 1; w < NW; ++w) tot.merge(smw[w]); }
                }
                if (tid == 0) {
                    if constexpr (WANT_IDX) fix = (blk.arg_ndim > 0) ? (b2_ravel_fix(blk, tot.i) - tot.i) : blk.arg_offset;
                    b2_store_result<REDOP, T, ACC>(blk, b, tot, fix);
                }
                return;
            }
        } else {
            // ---- MODE C: every row of the tile is reduced over ALL columns by its TX lanes
            static_assert(MODE == B2M_C, "mode");
            constexpr int WPR = (TX + 31) / 32;     // warps per row
            constexpr int SW = TX < 32 ? TX : 32;   // shuffle width
            __shared__ A smc[TY * WPR];
            const i64 c0 = (i64)tx * V;
            const int ncol = (c0 < C) ? (int)((C - c0 + (i64)TX * V - 1) / ((i64)TX * V)) : 0;   // steps of this lane
            for (i64 rr = r0; rr < rend; rr += TY) {       // uniform trip count (barriers inside)
                const i64 r = rr + ty;
                const bool rok = (r < rend) && ncol > 0;
                B2AccState<A, 1> st;
                A& acc = st.acc[0];
                acc.init();
                if (rok) {
                    typename Chain::Ptrs P;
                    Chain::setup_cols(blk, b, r, c0, (i64)TX * V, P);
                    if constexpr (REDOP == B2R_MOMENT) {
                        typename Chain::Regs g0; T o0[V];
                        typename Chain::Ptrs P0;                       // ONE pivot per row: its first element
                        Chain::setup_cols(blk, b, r, 0, (i64)TX * V, P0);
                        Chain::load(P0, 0, g0); Chain::compute_slow(g0, sc, o0); acc.prime(o0[0]);
                    }
                    b2_stream<Chain, V, U>(P, ncol, sc, st,
                        [&](B2AccState<A, 1>& s_, int k, const T (&o)[V]) {
#pragma unroll
                            for (int v = 0; v < V; ++v) s_.acc[0].add(o[v], k, v);
                        });
                    if constexpr (WANT_IDX) acc.finish_arg(true, c0, (i64)TX * V);
                    if constexpr (REDOP == B2R_MOMENT) acc.to_raw((i64)ncol * V);
                }
                if constexpr (REDOP == B2R_MOMENT) {
                    // lanes without work must still agree on the row's pivot: take lane 0's
                    acc.K = __shfl_sync(0xffffffffu, acc.K, (tid & 31) & ~(SW - 1), 32);
                }
#pragma unroll
                for (int off = SW / 2; off > 0; off >>= 1) acc.lane_shfl(off, SW);
                if constexpr (WPR > 1) {
                    const int lane = tid & 31, w = tx >> 5;
                    __syncthreads();
                    if (lane == 0) smc[ty * WPR + w] = acc;
                    __syncthreads();
                    if (tx == 0) { for (int k = 1; k < WPR; ++k) acc.lane_merge(smc[ty * WPR + k]); }
                }
                if (tx == 0) acc.lane_finish();
                if (tx == 0 && r < rend) {
                    i64 fx = blk.arg_offset;      // a ravelled arg reduction over a block whose other dims are 1
                    if constexpr (WANT_IDX) { if (blk.arg_ndim > 0) fx = b2_ravel_fix(blk, acc.i) - acc.i; }
                    b2_store_result<REDOP, T, ACC>(blk, b * R + r, acc, fx);
                }
            }
            return;
        }
    }
}

// ------------------------------------------------------------------ element-wise with transposed operands
// `x.T + x` style chains (manipulation/_transpose.py:14-75 feeding an Elemwise): an operand whose
// contiguous dimension is the OUTPUT's row dimension would be read with a stride by b2_run.
// Here a CTA owns a 64 x 64 output tile; every transposed operand is first staged through a
// padded shared-memory tile -- read along ITS contiguous dimension (coalesced 256 B runs), written
// transposed with a 65-element pitch (bank-conflict free) -- and the chain then reads it like any
// other operand.  Chain supplies, besides the b2_run interface:
//   TBYTES (shared bytes of all staged tiles), stage(blk, b, r0, c0, smem, tid, nthreads),
//   load_t(blk, Ptrs, smem, lr, lc, Regs&)  -- normal operands through Ptrs, staged ones from smem.
#define B2_TT 64           // tile edge
#define B2_TP 65           // padded pitch (elements)
template <typename Chain, int V, int TX, int TY>
__device__ __forceinline__ void b2_run_ewt(const B2Block* __restrict__ blocks, int nblocks, const B2Scalars& sc) {
    typedef typename Chain::out_t T;
    constexpr int NT = TX * TY;
    static_assert(TX * V == B2_TT, "tile width");
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    __shared__ B2Block sblk;
    __shared__ __align__(16) unsigned char tiles[Chain::TBYTES];
    {
        const i64 tile = blockIdx.x;
        const int bi = b2_find_block(blocks, nblocks, tile);
        const int nw = (int)(sizeof(B2Block) / 4);
        const u32* src = reinterpret_cast<const u32*>(blocks + bi);
        u32* dst = reinterpret_cast<u32*>(&sblk);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        __syncthreads();
    }
    const B2Block& blk = sblk;
    i64 t = (i64)blockIdx.x - blk.tile_begin;
    const i64 tc = t % blk.tiles_c; t /= blk.tiles_c;
    const i64 tr = t % blk.tiles_r; t /= blk.tiles_r;
    const i64 b = t;
    const i64 R = blk.R, C = blk.C;
    const i64 r0 = tr * B2_TT, c0 = tc * B2_TT;
    const int lc = tx * V;
    const i64 c = c0 + lc;
    constexpr int ITER = B2_TT / TY;
    typename Chain::Ptrs P;
    typename Chain::Regs g[ITER];
    Chain::setup_rows(blk, b, r0, (c < C) ? c : 0, 1, P);
    // 1. staged operands: all loads of the tile in flight, then the transposed smem writes
    Chain::template stage<NT>(blk, b, r0, c0, tiles, tid);
    // 2. the other operands' loads are issued before the barrier so they overlap the staging
    if (c < C) {
#pragma unroll
        for (int m = 0; m < ITER; ++m) { const int lr = ty + m * TY; if (r0 + lr < R) Chain::load_n(P, lr, g[m]); }
    }
    __syncthreads();
    if (c >= C) return;
    T* outp = (T*)blk.out0 + (b * R + r0) * C + c;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
        const int lr = ty + m * TY;
        if (r0 + lr < R) {
            T o[V];
            Chain::load_s(tiles, lr, lc, g[m]);
            Chain::compute_slow(g[m], sc, o);
            b2_store_vec<T, V>(outp + (i64)lr * C, o);
        }
    }
}

// ------------------------------------------------------------------ mirror-pair element-wise kernel
// `x.T + x` (and every chain f(x, x.T) over ONE array, tests/test_collection.py `a + a.T`):
// output block (i, j) reads x[i, j] and x[j, i].T, output block (j, i) reads x[j, i] and x[i, j].T --
// the same two input blocks.  b2_run_ewt moves 3 N bytes for that (every element is read twice);
// here a CTA owns a PAIR of tiles, A = tile (tr, tc) of block d and B = tile (tc, tr) of its mirror
// block d', loads each once with 128-bit loads, parks both in swizzled shared memory and produces
//   out [d ] tile (tr, tc) = f(A, B^T)     and     out[d'] tile (tc, tr) = f(B, A^T),
// so the launch moves the algorithmic minimum of 2 N bytes.  The host lists the primary
// block of every pair first (only those are tiled) and the partners behind them; `mirror` is the
// table index of the partner (== own index for diagonal blocks, whose tiles below the diagonal exit).
//
// Shared layout of a TT x TT tile (rows of 256 B = 16 chunks of 16 B): element (r, c) lives at
//   r * TT + ((c / V) ^ ((r / V) & 7)) * V + c % V
// -- row-wise 16 B stores are conflict free, the transposed scalar reads are 2-way conflicted (no
// 16 B-preserving swizzle does better for 16 chunk columns over 32 banks); shared bandwidth is not
// the limit, issue slots are: 1 LDG.128 + 1 STS.128 + 4 LDS + chain + 1 STG.128 per 4 elements.
// Chain supplies: sym_t, sym_put(tile, off, Regs) [the normal operand's V lanes -> 16 B store],
// sym_get(tile, base, Regs) [V lanes of the transposed operand, TT apart].
template <typename Chain, int V, int TT>
__device__ __forceinline__ void b2_run_ewt_sym(const B2Block* __restrict__ blocks, int nblocks, const B2Scalars& sc) {
    typedef typename Chain::out_t T;
    typedef typename Chain::sym_t E;
    constexpr int NT = 256, TX = TT / V, TY = NT / TX, ITER = TT / TY;
    static_assert(TX == 16, "a tile row is 16 chunks of 16 bytes");
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    __shared__ B2Block sblk[2];
    __shared__ __align__(16) E tiles[2][TT * TT];
    {
        const int bi = b2_find_block(blocks, nblocks, (i64)blockIdx.x);
        const int mi = blocks[bi].mirror;
        constexpr int nw = (int)(sizeof(B2Block) / 4);
        const u32* s0 = reinterpret_cast<const u32*>(blocks + bi);
        const u32* s1 = reinterpret_cast<const u32*>(blocks + mi);
        u32* dst = reinterpret_cast<u32*>(&sblk[0]);
        for (int i = tid; i < 2 * nw; i += NT) dst[i] = (i < nw) ? s0[i] : s1[i - nw];
        __syncthreads();
    }
    const B2Block& blk = sblk[0];
    const B2Block& blk2 = sblk[1];
    const int t = (int)((i64)blockIdx.x - blk.tile_begin);
    const int tc = t % (int)blk.tiles_c, tr = t / (int)blk.tiles_c;
    const bool diag = (blk.out0 == blk2.out0);
    if (diag && tr > tc) return;
    const bool self = diag && tr == tc;
    const int R = (int)blk.R, C = (int)blk.C;             // partner block: C rows x R columns
    const int r0 = tr * TT, c0 = tc * TT;
    const int lc = tx * V;
    const bool colA = (c0 + lc < C), colB = (r0 + lc < R);
    typename Chain::Ptrs PA, PB;
    typename Chain::Regs gA[ITER], gB[ITER];
    Chain::setup_rows(blk, 0, r0, colA ? c0 + lc : 0, 1, PA);
    Chain::setup_rows(blk2, 0, c0, colB ? r0 + lc : 0, 1, PB);
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
        const int lr = ty + m * TY;
        if (colA && r0 + lr < R) Chain::load_n(PA, lr, gA[m]);
    }
    if (!self) {
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
            const int lr = ty + m * TY;
            if (colB && c0 + lr < C) Chain::load_n(PB, lr, gB[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
        const int lr = ty + m * TY;
        const int off = lr * TT + ((tx ^ ((lr / V) & 7)) * V);
        if (colA && r0 + lr < R) Chain::sym_put(tiles[0], off, gA[m]);
        if (!self && colB && c0 + lr < C) Chain::sym_put(tiles[1], off, gB[m]);
    }
    __syncthreads();
    const E* const tB = self ? tiles[0] : tiles[1];
    if (colA) {
        T* outp = (T*)blk.out0 + (i64)r0 * C + c0 + lc;
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
            const int lr = ty + m * TY;
            if (r0 + lr < R) {
                T o[V];
                Chain::sym_get(tB, lc * TT + (((lr / V) ^ (tx & 7)) * V) + lr % V, gA[m]);
                Chain::compute_slow(gA[m], sc, o);
                b2_store_vec<T, V>(outp + (i64)lr * C, o);
            }
        }
    }
    if (!self && colB) {
        T* outp = (T*)blk2.out0 + (i64)c0 * R + r0 + lc;
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
            const int lr = ty + m * TY;
            if (c0 + lr < C) {
                T o[V];
                Chain::sym_get(tiles[0], lc * TT + (((lr / V) ^ (tx & 7)) * V) + lr % V, gB[m]);
                Chain::compute_slow(gB[m], sc, o);
                b2_store_vec<T, V>(outp + (i64)lr * R, o);
            }
        }
    }
}

// ------------------------------------------------------------------ cumulative scans
// cumsum / cumprod (reductions/_cumulative.py:100-265, CumReduction): the reference scans every block
// with np.cumsum, walks the blocks along the axis carrying the running total (`extra`) and adds it
// to each block in a third pass -- 4 N bytes plus temporaries.  Here the per-segment totals come
// from the ordinary reduction kernels (1 read), their scan is a tiny launch of THIS kernel, and one
// pass then writes  out = carry (+|*) local_scan(x)  -- 3 N bytes in total.
//   mode SR: scan along rows; a thread owns V columns and walks all R rows (running total in
//            registers, U loads in flight);
//   mode SC: scan along the contiguous dim; one warp per row walks it in steps of 32*V elements:
//            lane-local scan, warp shuffle scan of the lane totals, running total in a register.
template <int REDOP, typename ACC> struct B2ScanOp {
    __device__ __forceinline__ static ACC ident() { return REDOP == B2R_PROD ? (ACC)1 : (ACC)0; }
    __device__ __forceinline__ static ACC op(ACC a, ACC b) { return REDOP == B2R_PROD ? (ACC)(a * b) : (ACC)(a + b); }
};
template <typename ACC, int V> struct B2ScanState { ACC run[V]; };

template <typename Chain, int MODE, int REDOP, int V, int TX, int TY, int RPT, int U, typename ACC>
__device__ __forceinline__ void b2_run_scan(const B2Block* __restrict__ blocks, int nblocks, const B2Scalars& sc) {
    typedef typename Chain::out_t T;
    typedef B2ScanOp<REDOP, ACC> OP;
    static_assert(!Chain::HAS_SLOW, "scan kernels take exact chains only");
    static_assert(sizeof(ACC) == 4 || sizeof(ACC) == 8, "scan accumulators are 32 or 64 bit");
    constexpr int NT = TX * TY;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    __shared__ B2Block sblk;
    {
        const int bi = b2_find_block(blocks, nblocks, (i64)blockIdx.x);
        const int nw = (int)(sizeof(B2Block) / 4);
        const u32* src = reinterpret_cast<const u32*>(blocks + bi);
        u32* dst = reinterpret_cast<u32*>(&sblk);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        __syncthreads();
    }
    const B2Block& blk = sblk;
    i64 t = (i64)blockIdx.x - blk.tile_begin;
    const i64 tc = t % blk.tiles_c; t /= blk.tiles_c;
    const i64 tr = t % blk.tiles_r; t /= blk.tiles_r;
    const i64 b = t;
    const i64 R = blk.R, C = blk.C;
    const ACC* const carry = (const ACC*)blk.out1;

    // Chained launches (single pass, 2 N bytes): the blocks along the scanned axis form a chain -- `mirror`
    // holds the table index of the NEXT block of the chain (0 = last; heads come first in the table, so 0 is
    // never a successor) -- and the thread / warp that owns a column strip / row keeps walking from block to
    // block with its running total in registers: no per-block totals pass, no carry table.
    if constexpr (MODE == B2M_SR) {
        static_assert(TY == 1, "row scans: one thread per column strip");
        const i64 c = (tc * TX + tx) * V;
        if (c >= C) return;
        ACC cv[V];
#pragma unroll
        for (int v = 0; v < V; ++v) cv[v] = carry ? carry[b * C + c + v] : OP::ident();
        B2ScanState<ACC, V> st;
#pragma unroll
        for (int v = 0; v < V; ++v) st.run[v] = OP::ident();
        const bool has = carry != nullptr;
        bool started = false;
        const B2Block* cur = &blk;
        for (;;) {
            const i64 Rc = cur->R;
            typename Chain::Ptrs P;
            Chain::setup_rows(*cur, b, 0, c, 1, P);
            ACC* const outp = (ACC*)cur->out0 + (b * Rc) * C + c;
            const bool fresh = !started;
            b2_stream<Chain, V, U>(P, (int)Rc, sc, st,
                [&](B2ScanState<ACC, V>& s, int k, const T (&o)[V]) {
                    ACC w[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        s.run[v] = (fresh && k == 0) ? b2_cast<ACC>(o[v]) : OP::op(s.run[v], b2_cast<ACC>(o[v]));
                        w[v] = has ? OP::op(cv[v], s.run[v]) : s.run[v];
                    }
                    b2_store_vec<ACC, V>(outp + (i64)k * C, w);
                });
            if (Rc > 0) started = true;
            const int nxt = cur->mirror;
            if (nxt == 0) break;
            cur = blocks + nxt;
        }
    } else {
        static_assert(MODE == B2M_SC && TX == 32, "column scans: one warp per row");
        const int lane = tx;
        const i64 r0 = tr * RPT;
        const i64 rend = (r0 + RPT < R) ? (r0 + RPT) : R;
        const i64 step = (i64)TX * V;
        for (i64 r = r0 + ty; r < rend; r += TY) {         // a warp owns its row: uniform control flow
            const ACC cv = carry ? carry[b * R + r] : OP::ident();
            const bool has = carry != nullptr;
            B2ScanState<ACC, 1> st;
            st.run[0] = OP::ident();
            bool first = true;
            const B2Block* cur = &blk;
            for (;;) {
                const i64 Cc = cur->C;
                const int nfull = (int)(Cc / step);                 // steps every lane takes
                const i64 ctail = (i64)nfull * step + (i64)lane * V; // this lane's columns in the ragged last step
                ACC* const outp = (ACC*)cur->out0 + (b * R + r) * Cc + (i64)lane * V;
                auto stepfn = [&](B2ScanState<ACC, 1>& s, i64 k, const T (&o)[V], bool valid) {
                    ACC w[V];
                    w[0] = valid ? b2_cast<ACC>(o[0]) : OP::ident();
#pragma unroll
                    for (int v = 1; v < V; ++v) w[v] = valid ? OP::op(w[v - 1], b2_cast<ACC>(o[v])) : OP::ident();
                    ACC inc = w[V - 1];
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const ACC up = __shfl_up_sync(0xffffffffu, inc, off);
                        if (lane >= off) inc = OP::op(up, inc);
                    }
                    ACC excl = __shfl_up_sync(0xffffffffu, inc, 1);
                    const ACC total = __shfl_sync(0xffffffffu, inc, 31);
                    // prefix of everything before this lane's chunk: previous steps, then the lanes to the left
                    const bool have_base = !first || lane > 0;
                    ACC base = first ? excl : (lane > 0 ? OP::op(s.run[0], excl) : s.run[0]);
                    if (valid) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            ACC x = have_base ? OP::op(base, w[v]) : w[v];
                            w[v] = has ? OP::op(cv, x) : x;
                        }
                        b2_store_vec<ACC, V>(outp + k * step, w);
                    }
                    s.run[0] = first ? total : OP::op(s.run[0], total);
                    first = false;
                };
                if (nfull > 0) {
                    typename Chain::Ptrs P;
                    Chain::setup_cols(*cur, b, r, (i64)lane * V, step, P);
                    b2_stream<Chain, V, U>(P, nfull, sc, st,
                        [&](B2ScanState<ACC, 1>& s, int k, const T (&o)[V]) { stepfn(s, (i64)k, o, true); });
                }
                if ((i64)nfull * step < Cc) {
                    const bool valid = ctail < Cc;
                    T o[V];
                    if (valid) {
                        typename Chain::Ptrs P;
                        typename Chain::Regs g;
                        Chain::setup_cols(*cur, b, r, ctail, step, P);
                        Chain::load(P, 0, g);
                        Chain::compute_slow(g, sc, o);
                    }
                    stepfn(st, (i64)nfull, o, valid);
                }
                const int nxt = cur->mirror;
                if (nxt == 0) break;
                cur = blocks + nxt;
            }
        }
    }
}
