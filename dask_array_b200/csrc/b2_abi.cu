// b2_abi.cu -- the C ABI of libb200da.so (declared in include/b200da.h).
//
// Host side of the B200 backend: error handling, NVRTC JIT of generated fused kernels,
// module loading through the CUDA driver API (resolved with dlopen so the library also
// loads -- for symbol / JIT-compile checks -- on a machine without a GPU driver), launch
// planning, and the small ahead-of-time kernels (tree-level combine, tiled gather, fill).
//
// Reference interfaces replaced (paths under /root/reference/dask_array/): see the
// per-function comments in include/b200da.h.
#include "../../include/b200da.h"
#include "b2_device.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

// the device header text is embedded at build time (see Makefile: b2_device_text.inc)
static const char kDeviceHeader[] = {
#include "b2_device_text.inc"
    , 0};

// ---- layout checks: B2Block (device) == b2_block (ABI)
static_assert(sizeof(B2Block) == sizeof(b2_block), "B2Block/b2_block size mismatch");
static_assert(offsetof(B2Block, out0) == offsetof(b2_block, out0), "out0");
static_assert(offsetof(B2Block, B) == offsetof(b2_block, B), "B");
static_assert(offsetof(B2Block, tile_begin) == offsetof(b2_block, tile_begin), "tile_begin");
static_assert(offsetof(B2Block, work) == offsetof(b2_block, work), "work");
static_assert(offsetof(B2Block, counter) == offsetof(b2_block, counter), "counter");
static_assert(offsetof(B2Block, arg_offset) == offsetof(b2_block, arg_offset), "arg_offset");
static_assert(offsetof(B2Block, arg_shape) == offsetof(b2_block, arg_shape), "arg_shape");
static_assert(offsetof(B2Block, arg_total) == offsetof(b2_block, arg_total), "arg_total");
static_assert(sizeof(B2Scalars) == sizeof(b2_scalars), "scalars");
static_assert(sizeof(B2Block) % 4 == 0, "descriptor copied as u32 words");

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail(B2_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
    } while (0)

extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }
extern "C" const char* b2_last_error(void) { return g_err.c_str(); }
extern "C" int64_t b2_launch_count(void) { return g_launches.load(); }
extern "C" void b2_free(void* p) { free(p); }
extern "C" const char* b2_device_header(void) { return kDeviceHeader; }

extern "C" int b2_device_sm_count(int* out) {
    if (!out) return fail(B2_ERR_INVALID, "out is NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return B2_OK;
}

// ------------------------------------------------------------------ driver API (dlopen)
struct DriverApi {
    void* handle = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*MemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;
    bool ok = false;
    std::string why;
};
static DriverApi& driver() {
    static DriverApi d;
    static std::once_flag once;
    std::call_once(once, [] {
        d.handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!d.handle) { d.why = "libcuda.so.1 not found (no NVIDIA driver): there is no CPU fallback"; return; }
#define LOAD(field, sym)                                                    \
        *(void**)(&d.field) = dlsym(d.handle, sym);                         \
        if (!d.field) { d.why = std::string("missing driver symbol ") + sym; return; }
        LOAD(ModuleLoadData, "cuModuleLoadData");
        LOAD(ModuleUnload, "cuModuleUnload");
        LOAD(ModuleGetFunction, "cuModuleGetFunction");
        LOAD(LaunchKernel, "cuLaunchKernel");
        LOAD(GetErrorString, "cuGetErrorString");
        LOAD(FuncGetAttribute, "cuFuncGetAttribute");
        LOAD(MemGetAddressRange, "cuMemGetAddressRange_v2");
#undef LOAD
        d.ok = true;
    });
    return d;
}
static const char* cu_err(CUresult r) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return s ? s : "unknown CUDA driver error";
}

// ------------------------------------------------------------------ JIT
extern "C" int b2_jit_compile(const char* source, const char* name, void** cubin, size_t* cubin_size) {
    if (!source || !cubin || !cubin_size) return fail(B2_ERR_INVALID, "NULL argument");
    *cubin = nullptr;
    *cubin_size = 0;
    nvrtcProgram prog;
    const char* hdr_src[] = {kDeviceHeader};
    const char* hdr_names[] = {"b2_device.cuh"};
    nvrtcResult r = nvrtcCreateProgram(&prog, source, name ? name : "b2_fused.cu", 1, hdr_src, hdr_names);
    if (r != NVRTC_SUCCESS) return fail(B2_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    // --fmad=false: a*b+c stays two roundings, as in NumPy's separate ufunc loops, so float
    // chains are bit-identical to the reference; the kernels call fma() where they want one.
    // Element-wise kernels keep two roundings for a*b+c (as NumPy's separate ufunc loops) so that
    // float chains are bit-identical to the reference; a kernel whose chain is only observable
    // through a reduction (rtol contract) opts in with a first line "// b2-options: fmad".
    const bool fmad = strncmp(source, "// b2-options: fmad", 19) == 0;
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--restrict",
                          "-default-device", fmad ? "--fmad=true" : "--fmad=false"};
    r = nvrtcCompileProgram(prog, (int)(sizeof opts / sizeof *opts), opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        if (log.size() > 3500) log.resize(3500);
        return fail(B2_ERR_NVRTC, "nvrtcCompileProgram: %s\n%s", nvrtcGetErrorString(r), log.c_str());
    }
    size_t n = 0;
    r = nvrtcGetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(B2_ERR_NVRTC, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    void* buf = malloc(n);
    if (!buf) { nvrtcDestroyProgram(&prog); return fail(B2_ERR_INVALID, "out of host memory"); }
    r = nvrtcGetCUBIN(prog, (char*)buf);
    nvrtcDestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) { free(buf); return fail(B2_ERR_NVRTC, "nvrtcGetCUBIN: %s", nvrtcGetErrorString(r)); }
    *cubin = buf;
    *cubin_size = n;
    return B2_OK;
}

struct b2_kernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    b2_geom g{};
};

extern "C" int b2_kernel_load(const void* cubin, size_t cubin_size, const char* entry,
                              const b2_geom* geom, b2_kernel** out) {
    if (!cubin || !cubin_size || !entry || !geom || !out) return fail(B2_ERR_INVALID, "NULL argument");
    if (geom->vec < 1 || geom->tx < 1 || geom->ty < 1 || geom->tx * geom->ty > 1024 || geom->rpt < 1)
        return fail(B2_ERR_INVALID, "bad geometry vec=%d tx=%d ty=%d rpt=%d", geom->vec, geom->tx, geom->ty, geom->rpt);
    DriverApi& d = driver();
    if (!d.ok) return fail(B2_ERR_CUDA, "%s", d.why.c_str());
    CUDA_TRY(cudaFree(0));   // make sure the primary context of the caller's device is current
    b2_kernel* k = new b2_kernel();
    CUresult r = d.ModuleLoadData(&k->mod, cubin);
    if (r != CUDA_SUCCESS) { delete k; return fail(B2_ERR_CUDA, "cuModuleLoadData: %s", cu_err(r)); }
    r = d.ModuleGetFunction(&k->fn, k->mod, entry);
    if (r != CUDA_SUCCESS) { d.ModuleUnload(k->mod); delete k; return fail(B2_ERR_CUDA, "cuModuleGetFunction(%s): %s", entry, cu_err(r)); }
    k->g = *geom;
    *out = k;
    return B2_OK;
}

extern "C" int b2_kernel_free(b2_kernel* k) {
    if (!k) return B2_OK;
    if (k->mod && driver().ok) driver().ModuleUnload(k->mod);
    delete k;
    return B2_OK;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

extern "C" int b2_fused_plan(const b2_kernel* k, b2_block* blocks, int nblocks,
                             void* workspace, size_t workspace_bytes, size_t* needed,
                             int64_t* total_tiles) {
    if (!k || !blocks || nblocks <= 0 || !total_tiles) return fail(B2_ERR_INVALID, "bad argument");
    const b2_geom& g = k->g;
    const int64_t tile_cols = (int64_t)g.tx * g.vec;
    int64_t tiles = 0;
    size_t ncounters = 0, work_bytes = 0;
    std::vector<size_t> ctr_off(nblocks, 0), work_off(nblocks, 0);
    for (int i = 0; i < nblocks; ++i) {
        b2_block& b = blocks[i];
        if (b.B <= 0 || b.R <= 0 || b.C <= 0)
            return fail(B2_ERR_INVALID, "block %d has an empty extent (%lld,%lld,%lld): skip it on the host",
                        i, (long long)b.B, (long long)b.R, (long long)b.C);
        b.tiles_r = cdiv(b.R, g.rpt);
        b.tiles_c = (g.mode == B2_MODE_C || g.mode == B2_MODE_SC) ? 1 : cdiv(b.C, tile_cols);
        b.tile_begin = tiles;
        tiles += b.B * b.tiles_r * b.tiles_c;
        ctr_off[i] = ncounters;
        work_off[i] = work_bytes;
        if (g.mode == B2_MODE_R && b.tiles_r > 1) {
            ncounters += (size_t)(b.B * b.tiles_c);
            work_bytes += align_up((size_t)(b.B * b.tiles_r * b.tiles_c * tile_cols) * g.packed_bytes, 256);
        } else if (g.mode == B2_MODE_RC && b.tiles_r * b.tiles_c > 1) {
            ncounters += (size_t)b.B;
            work_bytes += align_up((size_t)(b.B * b.tiles_r * b.tiles_c) * g.packed_bytes, 256);
        }
    }
    if (tiles > 0x7fffffffLL) return fail(B2_ERR_UNSUPPORTED, "launch needs %lld tiles (> 2^31-1)", (long long)tiles);
    const size_t ctr_bytes = align_up(ncounters * sizeof(unsigned), 256);
    const size_t need = ctr_bytes + work_bytes;
    if (needed) *needed = need;
    *total_tiles = tiles;
    if (need > 0 && (!workspace || workspace_bytes < need)) {
        if (!workspace) return B2_OK;   // size query
        return fail(B2_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    }
    for (int i = 0; i < nblocks; ++i) {
        blocks[i].counter = need ? (unsigned*)workspace + ctr_off[i] : nullptr;
        blocks[i].work = need ? (char*)workspace + ctr_bytes + work_off[i] : nullptr;
    }
    return B2_OK;
}

extern "C" int b2_fused_launch(const b2_kernel* k, const b2_block* d_blocks, int nblocks,
                               int64_t total_tiles, const b2_scalars* scalars, void* stream) {
    if (!k || !d_blocks || nblocks <= 0 || total_tiles <= 0) return fail(B2_ERR_INVALID, "bad argument");
    DriverApi& d = driver();
    if (!d.ok) return fail(B2_ERR_CUDA, "%s", d.why.c_str());
    b2_scalars zero;
    memset(&zero, 0, sizeof zero);
    const b2_scalars* sc = scalars ? scalars : &zero;
    void* params[] = {(void*)&d_blocks, (void*)&nblocks, (void*)sc};
    CUresult r = d.LaunchKernel(k->fn, (unsigned)total_tiles, 1, 1, (unsigned)(k->g.tx * k->g.ty), 1, 1,
                                0, (CUstream)stream, params, nullptr);
    if (r != CUDA_SUCCESS) return fail(B2_ERR_CUDA, "cuLaunchKernel: %s", cu_err(r));
    g_launches.fetch_add(1);
    return B2_OK;
}

// ------------------------------------------------------------------ tree-level combine (AOT)
template <typename T, typename OUT>
__device__ __forceinline__ void b2_post_store(void* out0, i64 e, T total, int post, double count) {
    if (post == B2_POST_MEAN) {      // divide(total, n, dtype): in the output type (integer dtype= truncates like NumPy's cast)
        if constexpr (b2_is_float<OUT>::value) ((OUT*)out0)[e] = (OUT)total / (OUT)count;
        else ((OUT*)out0)[e] = (OUT)((double)total / count);
    } else {
        ((T*)out0)[e] = total;
    }
}

template <typename T, int OP, typename OUT>
__device__ __forceinline__ void b2_combine_one(const void* const* __restrict__ parts, const void* const* __restrict__ parts1,
                                               int fanin, i64 e, void* out0, void* out1, int post, double count, double ddof) {
    // The partials are fetched EIGHT AT A TIME before they are folded (pointer and value loads of a batch are
    // independent, the fold keeps the given order): a level with fan-in 16 costs two memory round trips instead
    // of sixteen -- these launches are a few threads wide, so their duration IS that dependent-load chain
    // (ncu, C2 std() tree: 11.4 us -> see profiles/).
    if constexpr (OP == B2R_SUM || OP == B2R_PROD) {
        T acc = ((const T*)parts[0])[e];
        for (int g0 = 1; g0 < fanin; g0 += 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (g0 + u < fanin) v[u] = ((const T*)parts[g0 + u])[e];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (g0 + u < fanin) acc = (OP == B2R_SUM) ? (T)(acc + v[u]) : (T)(acc * v[u]);
        }
        b2_post_store<T, OUT>(out0, e, acc, post, count);
    } else if constexpr (OP == B2R_MIN || OP == B2R_MAX) {
        T acc = ((const T*)parts[0])[e];
        for (int g0 = 1; g0 < fanin; g0 += 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (g0 + u < fanin) v[u] = ((const T*)parts[g0 + u])[e];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (g0 + u < fanin) acc = (OP == B2R_MAX) ? b2_np_max(acc, v[u]) : b2_np_min(acc, v[u]);
        }
        ((T*)out0)[e] = acc;
    } else if constexpr (OP == B2R_NANMIN || OP == B2R_NANMAX) {
        T acc = ((const T*)parts[0])[e];
        for (int g = 1; g < fanin; ++g) { T v = ((const T*)parts[g])[e]; acc = (OP == B2R_NANMAX) ? b2_nan_max(acc, v) : b2_nan_min(acc, v); }
        ((T*)out0)[e] = acc;
    } else if constexpr (OP == B2R_ANY || OP == B2R_ALL) {
        unsigned char acc = ((const unsigned char*)parts[0])[e];
        for (int g = 1; g < fanin; ++g) { unsigned char v = ((const unsigned char*)parts[g])[e]; acc = (OP == B2R_ALL) ? (acc & v) : (acc | v); }
        ((unsigned char*)out0)[e] = acc;
    } else if constexpr (OP == B2R_ARGMIN || OP == B2R_ARGMAX) {
        // _arg_combine (_common.py:675-701) re-applies np.argmax/np.argmin to the concatenated
        // per-block `vals`: the EARLIEST partial in nesting order wins a tie (and the first NaN
        // wins outright) -- for axis=None that is block order, not global index order.
        T bv = ((const T*)parts[0])[e];
        i64 bi = ((const i64*)parts1[0])[e];
        for (int g = 1; g < fanin; ++g) {
            T v = ((const T*)parts[g])[e];
            bool take = !b2_isnan(bv) && (b2_isnan(v) || ((OP == B2R_ARGMAX) ? (v > bv) : (v < bv)));
            if (take) { bv = v; bi = ((const i64*)parts1[g])[e]; }
        }
        ((T*)out0)[e] = bv;
        ((i64*)out1)[e] = bi;
    } else {   // MOMENT: packed (n, mean, M2) fp64 triples
        B2AccMoment<double, double> acc; acc.init();
        for (int g0 = 0; g0 < fanin; g0 += 8) {
            double q[8][3];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (g0 + u < fanin) {
                    const double* p = (const double*)parts[g0 + u] + 3 * e;
                    q[u][0] = p[0]; q[u][1] = p[1]; q[u][2] = p[2];
                }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (g0 + u < fanin) acc.chan(q[u][0], q[u][1], q[u][2]);
        }
        if (post == B2_POST_NONE) {
            double* q = (double*)out0 + 3 * e;
            q[0] = acc.n; q[1] = acc.mean; q[2] = acc.m2;
        } else {
            double den = acc.n - ddof;
            double var = (den < 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : acc.m2 / den;
            if (post == B2_POST_STD) var = sqrt(var);
            ((OUT*)out0)[e] = (OUT)var;
        }
    }
}

template <typename T, int OP, typename OUT>
__global__ void __launch_bounds__(256)
b2_combine_kernel(const void* const* __restrict__ parts, const void* const* __restrict__ parts1, int fanin,
                  i64 nelem, void* out0, void* out1, int post, double count, double ddof) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nelem) return;
    b2_combine_one<T, OP, OUT>(parts, parts1, fanin, e, out0, out1, post, count, ddof);
}

template <typename T, int OP, typename OUT>
__global__ void __launch_bounds__(256)
b2_combine_groups_kernel(const b2_group* __restrict__ groups, int ngroups, i64 total) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int lo = 0, hi = ngroups - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (groups[mid].elem_begin <= e) lo = mid; else hi = mid - 1;
    }
    const b2_group g = groups[lo];
    b2_combine_one<T, OP, OUT>(g.parts, g.parts1, g.fanin, e - g.elem_begin, g.out0, g.out1, g.post, g.count, g.ddof);
}

template <typename T, typename OUT>
static int launch_groups_t(int redop, const b2_group* g, int ng, int64_t total, cudaStream_t st) {
    const unsigned grid = (unsigned)cdiv(total, 256);
#define B2_CASE(OPC)                                                                     \
    case OPC:                                                                            \
        b2_combine_groups_kernel<T, OPC, OUT><<<grid, 256, 0, st>>>(g, ng, total);       \
        break;
    switch (redop) {
        B2_CASE(B2R_SUM) B2_CASE(B2R_PROD) B2_CASE(B2R_MIN) B2_CASE(B2R_MAX)
        B2_CASE(B2R_ARGMIN) B2_CASE(B2R_ARGMAX) B2_CASE(B2R_ANY) B2_CASE(B2R_ALL)
        B2_CASE(B2R_NANMIN) B2_CASE(B2R_NANMAX)
        default: return fail(B2_ERR_UNSUPPORTED, "combine: redop %d", redop);
    }
#undef B2_CASE
    return B2_OK;
}

template <typename T, typename OUT>
static int launch_combine_t(int redop, const void* const* p0, const void* const* p1, int fanin, int64_t n,
                            void* o0, void* o1, int post, double count, double ddof, cudaStream_t st) {
    const unsigned grid = (unsigned)cdiv(n, 256);
#define B2_CASE(OPC)                                                                                  \
    case OPC:                                                                                         \
        b2_combine_kernel<T, OPC, OUT><<<grid, 256, 0, st>>>(p0, p1, fanin, n, o0, o1, post, count, ddof); \
        break;
    switch (redop) {
        B2_CASE(B2R_SUM) B2_CASE(B2R_PROD) B2_CASE(B2R_MIN) B2_CASE(B2R_MAX)
        B2_CASE(B2R_ARGMIN) B2_CASE(B2R_ARGMAX) B2_CASE(B2R_ANY) B2_CASE(B2R_ALL)
        B2_CASE(B2R_NANMIN) B2_CASE(B2R_NANMAX)
        default: return fail(B2_ERR_UNSUPPORTED, "combine: redop %d", redop);
    }
#undef B2_CASE
    return B2_OK;
}

extern "C" int b2_combine(int redop, int dtype, const void* const* d_parts, const void* const* d_parts1,
                          int fanin, int64_t nelem, void* out0, void* out1,
                          int post, int out_dtype, double count, double ddof, void* stream) {
    if (!d_parts || fanin <= 0 || nelem < 0 || !out0) return fail(B2_ERR_INVALID, "bad argument");
    if (nelem == 0) return B2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = B2_OK;
    if (redop == B2R_MOMENT) {
        const unsigned grid = (unsigned)cdiv(nelem, 256);
        if (post != B2_POST_NONE && out_dtype == B2_F32)
            b2_combine_kernel<double, B2R_MOMENT, float><<<grid, 256, 0, st>>>(d_parts, d_parts1, fanin, nelem, out0, out1, post, count, ddof);
        else
            b2_combine_kernel<double, B2R_MOMENT, double><<<grid, 256, 0, st>>>(d_parts, d_parts1, fanin, nelem, out0, out1, post, count, ddof);
    } else {
        const bool mean = (post == B2_POST_MEAN);
        if (mean && redop != B2R_SUM) return fail(B2_ERR_INVALID, "POST_MEAN needs SUM");
        if (mean && out_dtype != B2_F32 && out_dtype != B2_F64) return fail(B2_ERR_UNSUPPORTED, "b2_combine: mean output must be f32/f64 (use b2_combine_groups)");
#define B2_T(DT, T)                                                                                                   \
    case DT:                                                                                                          \
        rc = (mean && out_dtype == B2_F32)                                                                            \
                 ? launch_combine_t<T, float>(redop, d_parts, d_parts1, fanin, nelem, out0, out1, post, count, ddof, st)   \
                 : launch_combine_t<T, double>(redop, d_parts, d_parts1, fanin, nelem, out0, out1, post, count, ddof, st); \
        break;
        switch (dtype) {
            B2_T(B2_BOOL, unsigned char) B2_T(B2_I8, signed char) B2_T(B2_U8, unsigned char)
            B2_T(B2_I16, short) B2_T(B2_U16, unsigned short) B2_T(B2_I32, int) B2_T(B2_U32, unsigned)
            B2_T(B2_I64, long long) B2_T(B2_U64, unsigned long long) B2_T(B2_F32, float) B2_T(B2_F64, double)
            default: return fail(B2_ERR_UNSUPPORTED, "combine: dtype %d", dtype);
        }
#undef B2_T
    }
    if (rc != B2_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

extern "C" int b2_combine_groups(int redop, int dtype, int out_dtype, const b2_group* d_groups, int ngroups,
                                 int64_t total_elems, void* stream) {
    if (!d_groups || ngroups <= 0 || total_elems < 0) return fail(B2_ERR_INVALID, "bad argument");
    if (total_elems == 0) return B2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = B2_OK;
    const unsigned grid = (unsigned)cdiv(total_elems, 256);
    if (redop == B2R_MOMENT) {
        if (out_dtype == B2_F32)
            b2_combine_groups_kernel<double, B2R_MOMENT, float><<<grid, 256, 0, st>>>(d_groups, ngroups, total_elems);
        else if (out_dtype == B2_I64)
            b2_combine_groups_kernel<double, B2R_MOMENT, long long><<<grid, 256, 0, st>>>(d_groups, ngroups, total_elems);
        else
            b2_combine_groups_kernel<double, B2R_MOMENT, double><<<grid, 256, 0, st>>>(d_groups, ngroups, total_elems);
    } else {
        const bool f32out = (out_dtype == B2_F32), i64out = (out_dtype == B2_I64);
#define B2_T(DT, T)                                                                                     \
    case DT:                                                                                            \
        rc = f32out ? launch_groups_t<T, float>(redop, d_groups, ngroups, total_elems, st)              \
           : i64out ? launch_groups_t<T, long long>(redop, d_groups, ngroups, total_elems, st)          \
                    : launch_groups_t<T, double>(redop, d_groups, ngroups, total_elems, st);            \
        break;
        switch (dtype) {
            B2_T(B2_BOOL, unsigned char) B2_T(B2_I8, signed char) B2_T(B2_U8, unsigned char)
            B2_T(B2_I16, short) B2_T(B2_U16, unsigned short) B2_T(B2_I32, int) B2_T(B2_U32, unsigned)
            B2_T(B2_I64, long long) B2_T(B2_U64, unsigned long long) B2_T(B2_F32, float) B2_T(B2_F64, double)
            default: return fail(B2_ERR_UNSUPPORTED, "combine: dtype %d", dtype);
        }
#undef B2_T
    }
    if (rc != B2_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// ------------------------------------------------------------------ tiled gather (AOT)
// One CTA copies one tile = tile_rows x (<= B2_GATHER_COL_BYTES) of one rectangle; a warp
// takes a row, lanes move the widest aligned word.
template <typename W>
__device__ __forceinline__ void b2_copy_row(const char* s, char* d, i64 nbytes, int lane) {
    const i64 n = nbytes / (i64)sizeof(W);
    const W* sw = reinterpret_cast<const W*>(s);
    W* dw = reinterpret_cast<W*>(d);
    for (i64 i = lane; i < n; i += 32) dw[i] = sw[i];
}

__global__ void __launch_bounds__(256) b2_gather_kernel(const b2_copy* __restrict__ copies, int n) {
    __shared__ b2_copy cp;
    {
        const i64 tile = blockIdx.x;
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (copies[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
        }
        const unsigned* src = reinterpret_cast<const unsigned*>(copies + lo);
        unsigned* dst = reinterpret_cast<unsigned*>(&cp);
        for (int i = threadIdx.x; i < (int)(sizeof(b2_copy) / 4); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const i64 t = (i64)blockIdx.x - cp.tile_begin;
    const i64 tc = t % cp.tiles_c, tr = t / cp.tiles_c;
    const i64 r0 = tr * cp.tile_rows;
    const i64 r1 = (r0 + cp.tile_rows < cp.rows) ? r0 + cp.tile_rows : cp.rows;
    const i64 c0 = tc * B2_GATHER_COL_BYTES;
    const i64 cb = (c0 + B2_GATHER_COL_BYTES < cp.row_bytes) ? B2_GATHER_COL_BYTES : cp.row_bytes - c0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (i64 r = r0 + warp; r < r1; r += nwarp) {
        const char* s = (const char*)cp.src + r * cp.src_pitch + c0;
        char* d = (char*)cp.dst + r * cp.dst_pitch + c0;
        switch (cp.vec_bytes) {
            case 16: b2_copy_row<uint4>(s, d, cb, lane); break;
            case 8: b2_copy_row<uint2>(s, d, cb, lane); break;
            case 4: b2_copy_row<unsigned>(s, d, cb, lane); break;
            case 2: b2_copy_row<unsigned short>(s, d, cb, lane); break;
            default: b2_copy_row<unsigned char>(s, d, cb, lane); break;
        }
    }
}

// Bulk-async variant (TMA, cp.async.bulk): one elected thread issues a global->shared bulk copy per
// row segment of the tile (all tracked by one mbarrier), then a shared->global bulk store per row.
// No register staging, ~64 KiB in flight per CTA.  Used when every address / pitch / length is a
// multiple of 16 bytes (vec_bytes == 16 for all rectangles).
__device__ __forceinline__ unsigned b2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) b2_gather_bulk_kernel(const b2_copy* __restrict__ copies, int n) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ b2_copy cp;
    __shared__ __align__(8) unsigned long long mbar;
    {
        const i64 tile = blockIdx.x;
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (copies[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
        }
        const unsigned* src = reinterpret_cast<const unsigned*>(copies + lo);
        unsigned* dst = reinterpret_cast<unsigned*>(&cp);
        for (int i = threadIdx.x; i < (int)(sizeof(b2_copy) / 4); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b2_smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const i64 t = (i64)blockIdx.x - cp.tile_begin;
    const i64 tc = t % cp.tiles_c, tr = t / cp.tiles_c;
    const i64 r0 = tr * cp.tile_rows;
    const int nr = (int)((r0 + cp.tile_rows < cp.rows) ? cp.tile_rows : cp.rows - r0);
    const i64 c0 = tc * B2_GATHER_COL_BYTES;
    const unsigned cb = (unsigned)((c0 + B2_GATHER_COL_BYTES < cp.row_bytes) ? B2_GATHER_COL_BYTES : cp.row_bytes - c0);
    const char* s = (const char*)cp.src + r0 * cp.src_pitch + c0;
    char* d = (char*)cp.dst + r0 * cp.dst_pitch + c0;
    const unsigned bar = b2_smem_u32(&mbar);
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb * (unsigned)nr) : "memory");
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(b2_smem_u32(bulk_smem + (size_t)r * cb)), "l"(s + (i64)r * cp.src_pitch), "r"(cb), "r"(bar) : "memory");
    }
    // everyone waits for the tile to land (phase 0)
    asm volatile(
        "{\n\t.reg .pred p;\n\tB2G_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra B2G_DONE_%=;\n\tbra B2G_WAIT_%=;\n\tB2G_DONE_%=:\n\t}" ::"r"(bar) : "memory");
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(d + (i64)r * cp.dst_pitch), "r"(b2_smem_u32(bulk_smem + (size_t)r * cb)), "r"(cb) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // smem may be released once read
}

extern "C" int b2_gather_plan(b2_copy* copies, int n, int64_t* total_tiles) {
    if (!copies || n <= 0 || !total_tiles) return fail(B2_ERR_INVALID, "bad argument");
    int64_t tiles = 0;
    for (int i = 0; i < n; ++i) {
        b2_copy& c = copies[i];
        if (c.rows <= 0 || c.row_bytes <= 0) return fail(B2_ERR_INVALID, "copy %d is empty: skip it on the host", i);
        uintptr_t bits = (uintptr_t)c.src | (uintptr_t)c.dst | (uintptr_t)c.row_bytes;
        if (c.rows > 1) bits |= (uintptr_t)c.src_pitch | (uintptr_t)c.dst_pitch;
        c.vec_bytes = (bits % 16 == 0) ? 16 : (bits % 8 == 0) ? 8 : (bits % 4 == 0) ? 4 : (bits % 2 == 0) ? 2 : 1;
        const int64_t colb = c.row_bytes < B2_GATHER_COL_BYTES ? c.row_bytes : B2_GATHER_COL_BYTES;
        int64_t tr = 65536 / colb;            // ~64 KiB per CTA
        if (tr < 8) tr = 8;
        if (tr > 256) tr = 256;
        if (tr > c.rows) tr = c.rows;
        c.tile_rows = (int32_t)tr;
        c.tiles_c = (int32_t)cdiv(c.row_bytes, B2_GATHER_COL_BYTES);
        c.tile_begin = tiles;
        tiles += cdiv(c.rows, tr) * c.tiles_c;
    }
    if (tiles > 0x7fffffffLL) return fail(B2_ERR_UNSUPPORTED, "gather needs %lld tiles", (long long)tiles);
    *total_tiles = tiles;
    return B2_OK;
}

extern "C" int b2_gather_launch_bulk(const b2_copy* d_copies, int n, int64_t total_tiles, void* stream) {
    if (!d_copies || n <= 0 || total_tiles <= 0) return fail(B2_ERR_INVALID, "bad argument");
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(b2_gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    });
    if (attr_err != cudaSuccess) return fail(B2_ERR_CUDA, "%s", cudaGetErrorString(attr_err));
    b2_gather_bulk_kernel<<<(unsigned)total_tiles, 128, 65536, (cudaStream_t)stream>>>(d_copies, n);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

extern "C" int b2_gather_launch(const b2_copy* d_copies, int n, int64_t total_tiles, void* stream) {
    if (!d_copies || n <= 0 || total_tiles <= 0) return fail(B2_ERR_INVALID, "bad argument");
    b2_gather_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>(d_copies, n);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

extern "C" int b2_memcpy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                           int64_t row_bytes, int64_t rows, int kind, void* stream) {
    if (!dst || !src || row_bytes < 0 || rows < 0) return fail(B2_ERR_INVALID, "bad argument");
    if (row_bytes == 0 || rows == 0) return B2_OK;
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (rows == 1 || (dst_pitch == row_bytes && src_pitch == row_bytes))
        CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)(row_bytes * rows), k, (cudaStream_t)stream));
    else
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes,
                                   (size_t)rows, k, (cudaStream_t)stream));
    return B2_OK;
}

// ------------------------------------------------------------------ top-k (AOT)
// strict weak order "a is better than b": NaN counts as the largest value (np.sort / np.partition put it
// last), invalid (padding, idx < 0) entries are the worst, ties go to the smaller index.
template <typename T, bool LARGEST>
__device__ __forceinline__ bool b2_topk_better(T av, int ai, T bv, int bi) {
    if (ai < 0 || bi < 0) return bi < 0 && ai >= 0;
    const bool an = b2_isnan(av), bn = b2_isnan(bv);
    if (an || bn) {
        if (an && bn) return ai < bi;
        return LARGEST ? an : bn;
    }
    if (av == bv) return ai < bi;
    return LARGEST ? (av > bv) : (av < bv);
}

template <typename T, bool LARGEST>
__global__ void __launch_bounds__(256) b2_topk_kernel(const T* __restrict__ src, i64 n, i64 src_pitch, int seg, int nseg,
                                                      int kk, T* __restrict__ out_vals, i64* __restrict__ out_idx,
                                                      i64 out_pitch, const i64* __restrict__ in_idx, i64 idx_offset) {
    extern __shared__ __align__(16) unsigned char b2_topk_smem[];
    T* sv = reinterpret_cast<T*>(b2_topk_smem);
    int* si = reinterpret_cast<int*>(b2_topk_smem + (((size_t)seg * sizeof(T) + 15) & ~(size_t)15));
    const i64 row = blockIdx.x / nseg;
    const int s = (int)(blockIdx.x % nseg);
    const i64 e0 = (i64)s * seg;
    const int len = (int)((n - e0 < seg) ? (n - e0) : seg);
    const T* rp = src + row * src_pitch + e0;
    for (int i = threadIdx.x; i < seg; i += blockDim.x) {
        // candidates of an earlier level carry idx = -1 where their segment ran out of elements
        const bool ok = i < len && !(in_idx && in_idx[row * n + e0 + i] < 0);
        if (ok) { sv[i] = rp[i]; si[i] = i; } else { sv[i] = T(); si[i] = -1; }
    }
    __syncthreads();
    for (int size = 2; size <= seg; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < seg / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);                 // this sub-sequence sorts best-first
                const T a = sv[lo], b = sv[hi];
                const int ia = si[lo], ib = si[hi];
                const bool swap = up ? b2_topk_better<T, LARGEST>(b, ib, a, ia) : b2_topk_better<T, LARGEST>(a, ia, b, ib);
                if (swap) { sv[lo] = b; sv[hi] = a; si[lo] = ib; si[hi] = ia; }
            }
            __syncthreads();
        }
    }
    T* ov = out_vals + row * out_pitch + (i64)s * kk;
    i64* oi = out_idx + row * out_pitch + (i64)s * kk;
    for (int j = threadIdx.x; j < kk; j += blockDim.x) {
        const int li = si[j];
        ov[j] = sv[j];
        oi[j] = (li < 0) ? (i64)-1 : (in_idx ? in_idx[row * n + e0 + li] : idx_offset + e0 + li);
    }
}

template <typename T>
static int launch_topk(const void* src, int64_t rows, int64_t n, int64_t src_pitch, int seg, int k, int largest,
                       void* out_vals, int64_t* out_idx, int64_t out_pitch, const int64_t* in_idx, int64_t idx_offset,
                       cudaStream_t st) {
    const int nseg = (int)cdiv(n, (int64_t)seg);
    const int kk = k < seg ? k : seg;
    const size_t smem = (((size_t)seg * sizeof(T) + 15) & ~(size_t)15) + (size_t)seg * sizeof(int);
    const unsigned grid = (unsigned)(rows * nseg);
    if (largest)
        b2_topk_kernel<T, true><<<grid, 256, smem, st>>>((const T*)src, (i64)n, (i64)src_pitch, seg, nseg, kk, (T*)out_vals,
                                                        (i64*)out_idx, (i64)out_pitch, (const i64*)in_idx, (i64)idx_offset);
    else
        b2_topk_kernel<T, false><<<grid, 256, smem, st>>>((const T*)src, (i64)n, (i64)src_pitch, seg, nseg, kk, (T*)out_vals,
                                                         (i64*)out_idx, (i64)out_pitch, (const i64*)in_idx, (i64)idx_offset);
    return B2_OK;
}

extern "C" int b2_topk_rows(int dtype, const void* src, int64_t rows, int64_t n, int64_t src_pitch, int seg, int k,
                            int largest, void* out_vals, int64_t* out_idx, int64_t out_pitch, const int64_t* in_idx,
                            int64_t idx_offset, void* stream) {
    if (!src || !out_vals || !out_idx || rows < 0 || n < 0 || k <= 0) return fail(B2_ERR_INVALID, "bad argument");
    if (rows == 0 || n == 0) return B2_OK;
    if (seg < 2 || (seg & (seg - 1)) != 0) return fail(B2_ERR_INVALID, "topk: segment length %d is not a power of two", seg);
    if (rows * cdiv(n, (int64_t)seg) > 0x7fffffffLL) return fail(B2_ERR_UNSUPPORTED, "topk: too many segments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = B2_OK;
#define B2_T(DT, T)                                                                                                     \
    case DT:                                                                                                            \
        if ((size_t)seg * (sizeof(T) + 4) > B2_TOPK_SEG_BYTES + 16384)                                                  \
            return fail(B2_ERR_INVALID, "topk: segment of %d elements does not fit shared memory", seg);                \
        rc = launch_topk<T>(src, rows, n, src_pitch, seg, k, largest, out_vals, out_idx, out_pitch, in_idx, idx_offset, st); \
        break;
    switch (dtype) {
        B2_T(B2_BOOL, unsigned char) B2_T(B2_I8, signed char) B2_T(B2_U8, unsigned char)
        B2_T(B2_I16, short) B2_T(B2_U16, unsigned short) B2_T(B2_I32, int) B2_T(B2_U32, unsigned)
        B2_T(B2_I64, long long) B2_T(B2_U64, unsigned long long) B2_T(B2_F32, float) B2_T(B2_F64, double)
        default: return fail(B2_ERR_UNSUPPORTED, "topk: dtype %d", dtype);
    }
#undef B2_T
    if (rc != B2_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// ------------------------------------------------------------------ peer memory (CUDA IPC)
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
extern "C" int b2_ipc_export(const void* ptr, b2_ipc_handle* out) {
    if (!ptr || !out) return fail(B2_ERR_INVALID, "NULL argument");
    DriverApi& d = driver();
    if (!d.ok) return fail(B2_ERR_CUDA, "%s", d.why.c_str());
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult r = d.MemGetAddressRange(&base, &size, (CUdeviceptr)ptr);
    if (r != CUDA_SUCCESS) return fail(B2_ERR_CUDA, "cuMemGetAddressRange: %s", cu_err(r));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, (void*)base));
    memcpy(out->reserved, &h, 64);
    out->offset = (int64_t)((CUdeviceptr)ptr - base);
    out->size = (int64_t)size;
    return B2_OK;
}

extern "C" int b2_ipc_open(const b2_ipc_handle* h, void** ptr) {
    if (!h || !ptr) return fail(B2_ERR_INVALID, "NULL argument");
    // an allocation can be opened once per process: cache the mapping per (device, handle)
    static std::mutex mu;
    static std::map<std::string, void*> mapped;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::string key((const char*)h->reserved, 64);
    key.push_back((char)dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = mapped.find(key);
    void* base = nullptr;
    if (it != mapped.end()) {
        base = it->second;
    } else {
        cudaIpcMemHandle_t ch;
        memcpy(&ch, h->reserved, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&base, ch, cudaIpcMemLazyEnablePeerAccess));
        mapped[key] = base;
    }
    *ptr = (char*)base + h->offset;
    return B2_OK;
}

__global__ void __launch_bounds__(32) b2_peer_barrier_kernel(unsigned long long* const* __restrict__ sig, int me, int world,
                                                             unsigned long long epoch, unsigned long long timeout_ns) {
    for (int p = threadIdx.x; p < world; p += blockDim.x) {
        __threadfence_system();
        unsigned long long* remote = sig[p] + me;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    }
    for (int p = threadIdx.x; p < world; p += blockDim.x) {
        const unsigned long long* mine = sig[me] + p;
        unsigned long long v, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) __trap();
            __nanosleep(100);
        }
    }
}

extern "C" int b2_peer_barrier(void* const* d_sig_table, int me, int world, uint64_t epoch, void* stream) {
    if (!d_sig_table || world <= 0 || me < 0 || me >= world) return fail(B2_ERR_INVALID, "bad argument");
    b2_peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)d_sig_table, me, world,
                                                               (unsigned long long)epoch, 120ull * 1000000000ull);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// Device-epoch variants: the epoch lives in a device counter that the kernel itself advances, so the
// launch arguments never change and the launch can sit inside a CUDA graph that is replayed.
__device__ __forceinline__ void b2_signal_all(unsigned long long* const* sig, int me, int world, unsigned long long epoch) {
    for (int p = threadIdx.x; p < world; p += blockDim.x) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig[p] + me), "l"(epoch) : "memory");
    }
}
__device__ __forceinline__ void b2_wait_all(unsigned long long* const* sig, int me, int world, unsigned long long epoch,
                                            unsigned long long timeout_ns) {
    for (int p = threadIdx.x; p < world; p += blockDim.x) {
        const unsigned long long* mine = sig[me] + p;
        unsigned long long v, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) __trap();
            __nanosleep(40);
        }
    }
}

__global__ void __launch_bounds__(32) b2_peer_barrier_dev_kernel(unsigned long long* const* __restrict__ sig,
                                                                 unsigned long long* epoch_ctr, int me, int world,
                                                                 unsigned long long timeout_ns) {
    const unsigned long long epoch = *epoch_ctr + 1;     // stream order serialises the launches that touch it
    b2_signal_all(sig, me, world, epoch);
    b2_wait_all(sig, me, world, epoch, timeout_ns);
    __syncwarp();
    if (threadIdx.x == 0) *epoch_ctr = epoch;
}

extern "C" int b2_peer_barrier_dev(void* const* d_sig_table, uint64_t* d_epoch, int me, int world, void* stream) {
    if (!d_sig_table || !d_epoch || world <= 0 || me < 0 || me >= world) return fail(B2_ERR_INVALID, "bad argument");
    b2_peer_barrier_dev_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)d_sig_table,
                                                                   (unsigned long long*)d_epoch, me, world,
                                                                   120ull * 1000000000ull);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// All-gather of small per-rank records through peer memory in ONE launch:
//   barrier (every rank has consumed the previous contents of the receive windows)
//   -> store my `nbytes` at slot `me` of every rank's window (16-byte vectors, NVLink stores)
//   -> barrier (every record has landed and is visible).
// One CTA: the payloads are the per-block partials of a tree reduction (bytes to a few hundred KiB).
__global__ void __launch_bounds__(1024) b2_peer_allgather_kernel(unsigned long long* const* __restrict__ sig,
                                                                 unsigned long long* epoch_ctr,
                                                                 unsigned char* const* __restrict__ windows,
                                                                 const unsigned char* __restrict__ send, long long nbytes,
                                                                 long long slot_bytes, int me, int world,
                                                                 unsigned long long timeout_ns) {
    const unsigned long long e0 = *epoch_ctr + 1;
    if (threadIdx.x < 32) {
        b2_signal_all(sig, me, world, e0);
        b2_wait_all(sig, me, world, e0, timeout_ns);
    }
    __syncthreads();
    const long long nvec = nbytes >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(send);
    for (int k = 0; k < world; ++k) {
        const int p = (me + 1 + k) % world;          // start at the right-hand neighbour: a permutation at any moment
        uint4* d4 = reinterpret_cast<uint4*>(windows[p] + (long long)me * slot_bytes);
        for (long long i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = s4[i];
        unsigned char* d1 = windows[p] + (long long)me * slot_bytes;
        for (long long i = (nvec << 4) + threadIdx.x; i < nbytes; i += blockDim.x) d1[i] = send[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < 32) {
        b2_signal_all(sig, me, world, e0 + 1);
        b2_wait_all(sig, me, world, e0 + 1, timeout_ns);
        __syncwarp();
        if (threadIdx.x == 0) *epoch_ctr = e0 + 1;
    }
}

extern "C" int b2_peer_allgather(void* const* d_sig_table, uint64_t* d_epoch, void* const* d_windows, const void* send,
                                 int64_t nbytes, int64_t slot_bytes, int me, int world, void* stream) {
    if (!d_sig_table || !d_epoch || !d_windows || world <= 0 || me < 0 || me >= world || nbytes < 0 || nbytes > slot_bytes
        || (nbytes && !send))
        return fail(B2_ERR_INVALID, "bad argument");
    if (((uintptr_t)send & 15) || (slot_bytes & 15)) return fail(B2_ERR_INVALID, "peer_allgather: 16-byte alignment");
    b2_peer_allgather_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        (unsigned long long* const*)d_sig_table, (unsigned long long*)d_epoch, (unsigned char* const*)d_windows,
        (const unsigned char*)send, (long long)nbytes, (long long)slot_bytes, me, world, 120ull * 1000000000ull);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// ------------------------------------------------------------------ fill (AOT)
template <typename W>
__global__ void __launch_bounds__(256) b2_fill_kernel(W* dst, i64 n, W value) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = value;
}

extern "C" int b2_fill(void* dst, int64_t nelem, int itemsize, const void* value, void* stream) {
    if (!dst || !value || nelem < 0) return fail(B2_ERR_INVALID, "bad argument");
    if (nelem == 0) return B2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = cdiv(nelem, 256);
    if (g > 148 * 16) g = 148 * 16;
    switch (itemsize) {
        case 1: { unsigned char v; memcpy(&v, value, 1); b2_fill_kernel<<<(unsigned)g, 256, 0, st>>>((unsigned char*)dst, nelem, v); break; }
        case 2: { unsigned short v; memcpy(&v, value, 2); b2_fill_kernel<<<(unsigned)g, 256, 0, st>>>((unsigned short*)dst, nelem, v); break; }
        case 4: { unsigned v; memcpy(&v, value, 4); b2_fill_kernel<<<(unsigned)g, 256, 0, st>>>((unsigned*)dst, nelem, v); break; }
        case 8: { unsigned long long v; memcpy(&v, value, 8); b2_fill_kernel<<<(unsigned)g, 256, 0, st>>>((unsigned long long*)dst, nelem, v); break; }
        default: return fail(B2_ERR_UNSUPPORTED, "fill: itemsize %d", itemsize);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// ------------------------------------------------------------------ sliding-window reductions (AOT)
// reduction(sliding_window_view(x, w), axis=window_axis) -- SlidingWindowReduction
// (reductions/_sliding_window.py:405-560): the reference combines, per block, a SUFFIX scan of the block, the
// totals of the blocks a window covers whole and a PREFIX scan of the band its right edge sweeps
// (_sliding_window_banded_reduce :96-160).  The same decomposition with segments of exactly `w` elements
// (van Herk / Gil-Werman) needs no totals:  out[t] = suffix_seg(t)[t]  (+)  prefix_seg(t)+1[t + w - 1],
// any associative (+), O(1) operations per element whatever the window, every input element read twice
// (once per scan direction; the second read is an L1 / L2 hit) and every output written once.
template <typename T, int OP> __device__ __forceinline__ T b2w_op(T a, T b) {
    if constexpr (OP == B2R_SUM) return (T)(a + b);
    else if constexpr (OP == B2R_PROD) return (T)(a * b);
    else if constexpr (OP == B2R_MIN) return b2_np_min(a, b);
    else return b2_np_max(a, b);
}

// window along ROWS of a (B, R, C) block: a thread owns V adjacent columns of one segment of w output rows.
// U row loads are issued before they are consumed (a thread keeps U x 16 B in flight: the scans are serial in
// the rows, not in the loads).
struct B2WindowJob {          // == b2_window_job (ABI)
    const void* src;
    void* dst;
    i64 B, R, C;
    i64 src_pitch;
    i64 tile_begin, col_tiles, row_tiles;
};
__device__ __forceinline__ int b2w_find_job(const B2WindowJob* __restrict__ jobs, int njobs, i64 tile) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename T, int OP, int V>
__global__ void __launch_bounds__(256, 3) b2_window_rows_kernel(const B2WindowJob* __restrict__ jobs, int njobs, i64 w, int mean) {
    constexpr int U = 4;      // (U = 8 needs 172 registers: one CTA per SM, latency-bound at 2 TB/s on B200;
                              //  U = 4 at two CTAs per SM: 2.7 TB/s; three CTAs per SM keep 48 KiB per SM in flight)
    const B2WindowJob job = jobs[b2w_find_job(jobs, njobs, (i64)blockIdx.x)];
    const T* __restrict__ src = (const T*)job.src;
    T* __restrict__ dst = (T*)job.dst;
    const i64 B = job.B, R = job.R, C = job.C, col_tiles = job.col_tiles, seg_tiles = job.row_tiles;
    const i64 Rout = R - w + 1, nseg = (Rout + w - 1) / w;
    i64 t = (i64)blockIdx.x - job.tile_begin;
    const i64 ct = t % col_tiles; t /= col_tiles;
    const i64 stile = t % seg_tiles; const i64 b = t / seg_tiles;
    const i64 c = (ct * 32 + threadIdx.x) * V, s = stile * 8 + threadIdx.y;
    if (c >= C || s >= nseg || b >= B) return;
    const T* x = src + (b * R) * C + c;
    T* o = dst + (b * Rout) * C + c;
    const i64 base = s * w;
    const i64 hi = (base + w < R) ? base + w : R;
    const T wdiv = (T)w;
    T h[V];
    bool first = true;
    for (i64 i0 = hi - 1; i0 >= base; i0 -= U) {             // suffix scan of this segment, last row first
        T v[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 - u >= base) b2_load_vec<T, V>(x + (i0 - u) * C, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 i = i0 - u;
            if (i < base) break;
            T out[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                h[k] = first ? v[u][k] : b2w_op<T, OP>(h[k], v[u][k]);
                out[k] = (mean && i == base) ? (T)(h[k] / wdiv) : h[k];
            }
            first = false;
            if (i < Rout) b2_store_vec<T, V>(o + i * C, out);
        }
    }
    i64 jmax = w - 1;                                        // rows of the next segment that complete a window
    if (R - (base + w) < jmax) jmax = R - (base + w);
    if (Rout - (base + 1) < jmax) jmax = Rout - (base + 1);
    T g[V];
    first = true;
    for (i64 j0 = 0; j0 < jmax; j0 += U) {
        T v[U][V], prev[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j0 + u < jmax) {
                b2_load_vec<T, V>(x + (base + w + j0 + u) * C, v[u]);
                const T* op_ = o + (base + j0 + u + 1) * C;   // written by THIS thread in the first phase
#pragma unroll
                for (int k = 0; k < V; ++k) prev[u][k] = op_[k];
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j0 + u >= jmax) break;
            T out[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                g[k] = first ? v[u][k] : b2w_op<T, OP>(g[k], v[u][k]);
                const T res = b2w_op<T, OP>(prev[u][k], g[k]);
                out[k] = mean ? (T)(res / wdiv) : res;
            }
            first = false;
            b2_store_vec<T, V>(o + (base + j0 + u + 1) * C, out);
        }
    }
}

// window along the CONTIGUOUS axis of (rows, C): a tile of RT rows x L columns is staged in shared memory
// (coalesced loads, a warp per row), scanned per (row, segment) in both directions, combined and stored
// coalesced.  All index arithmetic inside the tile is 32-bit.
template <typename T, int OP, int RT>
__global__ void __launch_bounds__(256) b2_window_cols_kernel(const B2WindowJob* __restrict__ jobs, int njobs, i64 w64,
                                                             i64 tile_out, int mean) {
    extern __shared__ __align__(16) unsigned char b2w_smem[];
    const B2WindowJob job = jobs[b2w_find_job(jobs, njobs, (i64)blockIdx.x)];
    const T* __restrict__ src = (const T*)job.src;
    T* __restrict__ dst = (T*)job.dst;
    const i64 rows = job.B * job.R, C = job.C, col_tiles = job.col_tiles;
    const i64 pitch = job.src_pitch ? job.src_pitch : C;       // elements between consecutive source rows
    const i64 Cout = C - w64 + 1;
    const i64 tl = (i64)blockIdx.x - job.tile_begin;
    const i64 ct = tl % col_tiles, rt = tl / col_tiles;
    const i64 t0 = ct * tile_out, row0 = rt * RT;
    const int w = (int)w64;
    const int nout = (int)((t0 + tile_out <= Cout) ? tile_out : Cout - t0);   // outputs of this tile
    const int L = nout + w - 1;                                               // input columns it needs
    const int Lp = L | 1;                                                     // odd pitch: conflict-free column walks
    T* H = reinterpret_cast<T*>(b2w_smem);
    T* G = H + (size_t)RT * Lp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nrows = (rows - row0 < RT) ? (int)(rows - row0) : RT;
    // stage the tile: every element is one asynchronous global -> shared copy (4- and 8-byte types), so a
    // thread has its whole share of the tile in flight at once instead of a few register loads
    for (int r = warp; r < RT; r += 8) {
        const T* p = src + (row0 + r) * pitch + t0;
        T* hrow = H + r * Lp;
        if (r < nrows) {
            if constexpr (sizeof(T) == 4 || sizeof(T) == 8) {
                for (int k = lane; k < L; k += 32) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(hrow + k);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(p + k), "n"((int)sizeof(T)) : "memory");
                }
            } else {
                for (int k = lane; k < L; k += 32) hrow[k] = b2_ld(p + k);
            }
        } else {
            for (int k = lane; k < L; k += 32) hrow[k] = T(0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int r = warp; r < RT; r += 8)
        for (int k = lane; k < L; k += 32) G[r * Lp + k] = H[r * Lp + k];
    __syncthreads();
    const int nsg = (L + w - 1) / w;
    const int ntask = 2 * RT * nsg;
    for (int task = tid; task < ntask; task += 256) {
        const int r = task % RT, q = task / RT;
        const int sg = q % nsg, dir = q / nsg;
        const int lo = sg * w, hi = (lo + w < L) ? lo + w : L;
        if (dir == 0) {                                     // suffix scan within the segment
            T* p = H + r * Lp;
            T acc = p[hi - 1];
            for (int i = hi - 2; i >= lo; --i) { acc = b2w_op<T, OP>(acc, p[i]); p[i] = acc; }
        } else {                                            // prefix scan within the segment
            T* p = G + r * Lp;
            T acc = p[lo];
            for (int i = lo + 1; i < hi; ++i) { acc = b2w_op<T, OP>(acc, p[i]); p[i] = acc; }
        }
    }
    __syncthreads();
    const T wdiv = (T)w;
    for (int r = warp; r < nrows; r += 8) {
        T* q = dst + (row0 + r) * Cout + t0;
        int kmod = lane % w;
        for (int k = lane; k < nout; k += 32) {
            T res = H[r * Lp + k];
            if (kmod != 0) res = b2w_op<T, OP>(res, G[r * Lp + k + w - 1]);   // (k % w == 0: the window IS a segment)
            q[k] = mean ? (T)(res / wdiv) : res;
            kmod = (kmod + 32) % w;
        }
    }
}

static_assert(sizeof(B2WindowJob) == sizeof(b2_window_job), "B2WindowJob / b2_window_job");

template <typename T, int OP>
static int b2_window_launch(b2_window_job* jobs, int n, void* d_jobs, i64 w, int along_cols, int mean, cudaStream_t st) {
    i64 tiles = 0;
    if (!along_cols) {
        constexpr int VMAX = 16 / (int)sizeof(T);
        bool vec = VMAX > 1;
        for (int i = 0; i < n; ++i)
            vec = vec && jobs[i].C % VMAX == 0 && ((uintptr_t)jobs[i].src % 16 == 0) && ((uintptr_t)jobs[i].dst % 16 == 0);
        for (int i = 0; i < n; ++i) {
            const i64 Rout = jobs[i].R - w + 1, nseg = cdiv(Rout, w);
            jobs[i].col_tiles = cdiv(jobs[i].C, 32 * (vec ? VMAX : 1));
            jobs[i].row_tiles = cdiv(nseg, 8);
            jobs[i].tile_begin = tiles;
            tiles += jobs[i].col_tiles * jobs[i].row_tiles * jobs[i].B;
        }
        if (tiles > 0x7fffffffLL) return fail(B2_ERR_UNSUPPORTED, "window_reduce: too many tiles");
        CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs, (size_t)n * sizeof(b2_window_job), cudaMemcpyHostToDevice, st));
        if (vec)
            b2_window_rows_kernel<T, OP, VMAX><<<(unsigned)tiles, dim3(32, 8), 0, st>>>((const B2WindowJob*)d_jobs, n, w, mean);
        else
            b2_window_rows_kernel<T, OP, 1><<<(unsigned)tiles, dim3(32, 8), 0, st>>>((const B2WindowJob*)d_jobs, n, w, mean);
    } else {
        // outputs per tile: SEVEN whole segments (the tile then holds 8 segments: 2 scan directions x 16 rows x 8
        // segments = 256 scan tasks, one per thread; halo overhead 1/7) unless the window is short, then about
        // 256 columns; rows per tile: 16, fewer when shared memory (<= 96 KiB: two CTAs per SM) demands it
        i64 tile_out = 7 * w;
        if (tile_out < 256) tile_out = w * cdiv(256, w);
        const i64 L = tile_out + w - 1, Lp = L | 1;
        const size_t per_row = 2 * (size_t)Lp * sizeof(T);
        int rt = 16;
        while (rt > 1 && per_row * rt > 96 * 1024) rt /= 2;
        if (per_row * rt > 200 * 1024)
            return fail(B2_ERR_UNSUPPORTED, "window_reduce: a window of %lld elements along the contiguous axis does not fit shared memory", (long long)w);
        const size_t smem = per_row * rt;
        for (int i = 0; i < n; ++i) {
            jobs[i].col_tiles = cdiv(jobs[i].C - w + 1, tile_out);
            jobs[i].row_tiles = cdiv(jobs[i].B * jobs[i].R, rt);
            jobs[i].tile_begin = tiles;
            tiles += jobs[i].col_tiles * jobs[i].row_tiles;
        }
        if (tiles > 0x7fffffffLL) return fail(B2_ERR_UNSUPPORTED, "window_reduce: too many tiles");
        CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs, (size_t)n * sizeof(b2_window_job), cudaMemcpyHostToDevice, st));
#define B2W_COLS(RT_)                                                                                                         \
        {                                                                                                                      \
            CUDA_TRY(cudaFuncSetAttribute(b2_window_cols_kernel<T, OP, RT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
            b2_window_cols_kernel<T, OP, RT_><<<(unsigned)tiles, 256, smem, st>>>((const B2WindowJob*)d_jobs, n, w, tile_out, mean); \
        }
        switch (rt) {
            case 16: B2W_COLS(16) break;
            case 8: B2W_COLS(8) break;
            case 4: B2W_COLS(4) break;
            case 2: B2W_COLS(2) break;
            default: B2W_COLS(1) break;
        }
#undef B2W_COLS
    }
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

template <typename T>
static int b2_window_dispatch_op(int redop, b2_window_job* jobs, int n, void* d_jobs, i64 w, int along_cols, int mean, cudaStream_t st) {
    switch (redop) {
        case B2_RED_SUM: return b2_window_launch<T, B2R_SUM>(jobs, n, d_jobs, w, along_cols, mean, st);
        case B2_RED_PROD: return b2_window_launch<T, B2R_PROD>(jobs, n, d_jobs, w, along_cols, mean, st);
        case B2_RED_MIN: return b2_window_launch<T, B2R_MIN>(jobs, n, d_jobs, w, along_cols, mean, st);
        case B2_RED_MAX: return b2_window_launch<T, B2R_MAX>(jobs, n, d_jobs, w, along_cols, mean, st);
        default: return fail(B2_ERR_UNSUPPORTED, "window_reduce: redop %d (sum, prod, min, max)", redop);
    }
}

extern "C" int b2_window_reduce_batched(int redop, int dtype, b2_window_job* jobs, int njobs, void* d_jobs,
                                        int64_t window, int along_cols, int mean, void* stream) {
    if (!jobs || !d_jobs || njobs <= 0 || window <= 0) return fail(B2_ERR_INVALID, "bad argument");
    for (int i = 0; i < njobs; ++i) {
        const b2_window_job& j = jobs[i];
        if (!j.src || !j.dst || j.B <= 0 || j.R <= 0 || j.C <= 0) return fail(B2_ERR_INVALID, "window_reduce: bad job %d", i);
        if ((along_cols ? j.C : j.R) < window) return fail(B2_ERR_INVALID, "window_reduce: window longer than the axis (job %d)", i);
        if (j.src_pitch && (!along_cols || j.src_pitch < j.C)) return fail(B2_ERR_INVALID, "window_reduce: src_pitch is for along_cols jobs and >= C (job %d)", i);
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case B2_F32: return b2_window_dispatch_op<float>(redop, jobs, njobs, d_jobs, window, along_cols, mean, st);
        case B2_F64: return b2_window_dispatch_op<double>(redop, jobs, njobs, d_jobs, window, along_cols, mean, st);
        case B2_I32: return b2_window_dispatch_op<int>(redop, jobs, njobs, d_jobs, window, along_cols, mean, st);
        case B2_I64: return b2_window_dispatch_op<long long>(redop, jobs, njobs, d_jobs, window, along_cols, mean, st);
        case B2_U8: case B2_BOOL: return b2_window_dispatch_op<unsigned char>(redop, jobs, njobs, d_jobs, window, along_cols, mean, st);
        default: return fail(B2_ERR_UNSUPPORTED, "window_reduce: dtype %d (f32, f64, i32, i64, u8/bool)", dtype);
    }
}

// ------------------------------------------------------------------ take (AOT)
// Integer-array gather of the arg-reduction combine step (_arg_combine, reductions/_common.py:687-697):
//   inner == 0 : out[j] = src[idx[j]]                              (vals.ravel()[local_args])
//   inner  > 0 : out[o, i] = src[o, idx[o, i], i], src (outer, n, inner) contiguous   (vals[ogrid.., local_args, ..])
// Negative indices wrap like NumPy; out-of-range ones are clamped (no fault on garbage input).
template <typename W>
__global__ void __launch_bounds__(256) b2_take_kernel(const W* __restrict__ src, const long long* __restrict__ idx,
                                                      W* __restrict__ out, i64 count, i64 n, i64 inner) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (i64)gridDim.x * blockDim.x) {
        long long k = idx[j];
        if (k < 0) k += n;
        k = k < 0 ? 0 : (k >= n ? n - 1 : k);
        if (inner == 0) out[j] = src[k];
        else { const i64 o = j / inner, i = j - o * inner; out[j] = src[(o * n + k) * inner + i]; }
    }
}

extern "C" int b2_take(int itemsize, const void* src, const int64_t* idx, void* out, int64_t count, int64_t n,
                       int64_t inner, void* stream) {
    if (count < 0 || n <= 0 || inner < 0 || (count && (!src || !idx || !out))) return fail(B2_ERR_INVALID, "bad argument");
    if (count == 0) return B2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = cdiv(count, 256);
    if (g > 148 * 16) g = 148 * 16;
    const long long* ix = (const long long*)idx;
    switch (itemsize) {
        case 1: b2_take_kernel<<<(unsigned)g, 256, 0, st>>>((const unsigned char*)src, ix, (unsigned char*)out, count, n, inner); break;
        case 2: b2_take_kernel<<<(unsigned)g, 256, 0, st>>>((const unsigned short*)src, ix, (unsigned short*)out, count, n, inner); break;
        case 4: b2_take_kernel<<<(unsigned)g, 256, 0, st>>>((const unsigned*)src, ix, (unsigned*)out, count, n, inner); break;
        case 8: b2_take_kernel<<<(unsigned)g, 256, 0, st>>>((const unsigned long long*)src, ix, (unsigned long long*)out, count, n, inner); break;
        default: return fail(B2_ERR_UNSUPPORTED, "take: itemsize %d", itemsize);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return B2_OK;
}

// shared with the other translation units of the library
extern "C" int b2_set_error_(int code, const char* msg) { return fail(code, "%s", msg); }
extern "C" void b2_count_launch_(void) { g_launches.fetch_add(1); }
