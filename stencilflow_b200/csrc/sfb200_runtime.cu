// libsfb200.so -- program-independent B200 runtime behind StencilFlow's `cuda` mode.
// C ABI declared in include/sfb200.h (the drop-in boundary; see the citations there).
//
// Linking: the CUDA runtime is linked statically and the driver API / NVRTC are resolved at run time
// (cudaGetDriverEntryPoint, dlopen), so the library loads on a machine without a GPU or driver --
// the CPU test-suite checks the exported symbols that way -- while every device entry point fails
// loudly with SFB_ERR_NO_DEVICE / SFB_ERR_CUDA there.  There is no CPU fallback of any kind.

#include "../../include/sfb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;
int g_device = -1;

int fail(int code, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define SFB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? SFB_ERR_OOM                        \
                        : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)       \
                              ? SFB_ERR_NO_DEVICE                                               \
                              : SFB_ERR_CUDA;                                                   \
            (void)cudaGetLastError(); /* this call reports its own failure: leave nothing behind */ \
            return fail(code_, "%s failed: %s (%s)", #call, cudaGetErrorString(e_),             \
                        cudaGetErrorName(e_));                                                  \
        }                                                                                       \
    } while (0)

// ---- driver API through the runtime's entry-point lookup (no link-time libcuda) -------------
struct DriverApi {
    bool loaded = false;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    CUresult (*StreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;
    CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
} g_drv;

template <typename F>
int load_entry(const char* name, F* out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
        return fail(SFB_ERR_NOT_FOUND, "driver entry point %s unavailable (%s)", name,
                    cudaGetErrorString(e));
    *out = reinterpret_cast<F>(p);
    return SFB_OK;
}

int load_driver() {
    if (g_drv.loaded) return SFB_OK;
    int rc;
#define SFB_ENTRY(field, sym) \
    if ((rc = load_entry(sym, &g_drv.field)) != SFB_OK) return rc;
    SFB_ENTRY(ModuleLoadData, "cuModuleLoadData")
    SFB_ENTRY(ModuleUnload, "cuModuleUnload")
    SFB_ENTRY(ModuleGetFunction, "cuModuleGetFunction")
    SFB_ENTRY(LaunchKernel, "cuLaunchKernel")
    SFB_ENTRY(FuncSetAttribute, "cuFuncSetAttribute")
    SFB_ENTRY(FuncGetAttribute, "cuFuncGetAttribute")
    SFB_ENTRY(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
    SFB_ENTRY(TensorMapEncodeTiled, "cuTensorMapEncodeTiled")
    SFB_ENTRY(StreamWriteValue32, "cuStreamWriteValue32")
    SFB_ENTRY(StreamWaitValue32, "cuStreamWaitValue32")
    SFB_ENTRY(GetErrorString, "cuGetErrorString")
#undef SFB_ENTRY
    g_drv.loaded = true;
    return SFB_OK;
}

int drv_fail(const char* what, CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    return fail(SFB_ERR_CUDA, "%s failed: %s (CUresult %d)", what, s ? s : "?", (int)r);
}

#define SFB_DRV(call)                                   \
    do {                                                \
        int rc_ = load_driver();                        \
        if (rc_ != SFB_OK) return rc_;                  \
        CUresult r_ = g_drv.call;                       \
        if (r_ != CUDA_SUCCESS) return drv_fail(#call, r_); \
    } while (0)

int require_init() {
    if (g_device < 0) return fail(SFB_ERR_INVALID, "sfb_init has not been called");
    return SFB_OK;
}

// ---- NVRTC, resolved lazily ------------------------------------------------------------------
struct Nvrtc {
    void* lib = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(void*, int, const char* const*);
    int (*GetCUBINSize)(void*, size_t*);
    int (*GetCUBIN)(void*, char*);
    int (*GetProgramLogSize)(void*, size_t*);
    int (*GetProgramLog)(void*, char*);
    int (*DestroyProgram)(void**);
    const char* (*GetErrorString)(int);
} g_nvrtc;

int load_nvrtc() {
    if (g_nvrtc.lib) return SFB_OK;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) return fail(SFB_ERR_NOT_FOUND, "cannot load libnvrtc: %s", dlerror());
#define SFB_SYM(field, sym)                                                         \
    *(void**)(&g_nvrtc.field) = dlsym(lib, sym);                                    \
    if (!g_nvrtc.field) return fail(SFB_ERR_NOT_FOUND, "libnvrtc lacks %s", sym);
    SFB_SYM(CreateProgram, "nvrtcCreateProgram")
    SFB_SYM(CompileProgram, "nvrtcCompileProgram")
    SFB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SFB_SYM(GetCUBIN, "nvrtcGetCUBIN")
    SFB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SFB_SYM(GetProgramLog, "nvrtcGetProgramLog")
    SFB_SYM(DestroyProgram, "nvrtcDestroyProgram")
    SFB_SYM(GetErrorString, "nvrtcGetErrorString")
#undef SFB_SYM
    g_nvrtc.lib = lib;
    return SFB_OK;
}

// ---- built-in kernels ------------------------------------------------------------------------
__host__ __device__ inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// u in [0,1) with 24 random bits; mirrored in stencilflow_b200/synthetic.py
__host__ __device__ inline double hash_unit(uint64_t idx, uint64_t seed) {
    uint32_t s = mix32((uint32_t)seed + 0x9e3779b9U * (uint32_t)(idx >> 32) +
                       0x85ebca6bU * (uint32_t)(seed >> 32));
    uint32_t h = mix32((uint32_t)idx ^ s);
    return (double)(h >> 8) * (1.0 / 16777216.0);
}

template <typename T>
__global__ void k_fill_constant(T* __restrict__ p, uint64_t n, T v) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

template <typename T>
__global__ void k_fill_hash(T* __restrict__ p, uint64_t n, uint64_t seed, double lo, double hi,
                            uint64_t off) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double span = hi - lo;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u = hash_unit(i + off, seed);
        p[i] = (T)__dadd_rn(lo, __dmul_rn(span, u));   // no FMA contraction: host-reproducible
    }
}

template <typename T, typename W>
__global__ void k_checksum(const T* __restrict__ p, uint64_t n, double* sum, unsigned long long* bits) {
    double s = 0.0;
    unsigned long long b = 0;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T v = p[i];
        s += (double)v;
        W w;
        memcpy(&w, &v, sizeof(W));
        b += (unsigned long long)w;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum, s);
        atomicAdd(bits, b);
    }
}

template <typename T>
__global__ void k_compare(const T* __restrict__ ref, const T* __restrict__ res, uint64_t n, double eps,
                          double tol, unsigned long long* max_bits, unsigned long long* bad) {
    double m = 0.0;
    unsigned long long nb = 0;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = (double)ref[i], b = (double)res[i];
        double r = fabs(a - b) / (fmax(fabs(a), fabs(b)) + eps);
        if (!(r <= tol)) nb++;          // NaN counts as a mismatch
        if (!(r == r)) r = INFINITY;
        m = fmax(m, r);
    }
    for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
        nb += __shfl_down_sync(0xffffffffu, nb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(max_bits, (unsigned long long)__double_as_longlong(m));  // m >= 0: bit order == value order
        if (nb) atomicAdd(bad, nb);
    }
}

unsigned grid_for(uint64_t n, unsigned block) {
    uint64_t g = (n + block - 1) / block;
    uint64_t cap = 148ull * 16ull;   // 148 SMs x 16 resident CTAs of 128..256 threads
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int sfb_abi_version(void) { return SFB_ABI_VERSION; }

const char* sfb_last_error(void) { return g_error.c_str(); }

int sfb_device_count(int* count) {
    if (!count) return fail(SFB_ERR_INVALID, "count is NULL");
    *count = 0;
    SFB_CUDA(cudaGetDeviceCount(count));
    return SFB_OK;
}

int sfb_init(int device) {
    int n = 0;
    SFB_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) return fail(SFB_ERR_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= n) return fail(SFB_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    SFB_CUDA(cudaSetDevice(device));
    SFB_CUDA(cudaFree(0));
    g_device = device;
    return load_driver();
}

int sfb_shutdown(void) {
    if (g_device >= 0) {
        cudaDeviceSynchronize();
        g_device = -1;
    }
    return SFB_OK;
}

int sfb_current_device(int* device) {
    if (!device) return fail(SFB_ERR_INVALID, "device is NULL");
    *device = g_device;
    return g_device >= 0 ? SFB_OK : fail(SFB_ERR_INVALID, "sfb_init has not been called");
}

int sfb_device_properties(int device, sfb_device_props* out) {
    if (!out) return fail(SFB_ERR_INVALID, "out is NULL");
    cudaDeviceProp p;
    SFB_CUDA(cudaGetDeviceProperties(&p, device));
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%.127s", p.name);
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    out->max_smem_per_block_optin = (int)p.sharedMemPerBlockOptin;
    out->l2_bytes = p.l2CacheSize;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device) == cudaSuccess) out->clock_khz = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, device) == cudaSuccess) out->mem_clock_khz = v;
    out->total_mem = p.totalGlobalMem;
    if (device == g_device) {
        size_t f = 0, t = 0;
        if (cudaMemGetInfo(&f, &t) == cudaSuccess) out->free_mem = f;
    }
    return SFB_OK;
}

int sfb_device_synchronize(void) {
    SFB_CUDA(cudaDeviceSynchronize());
    return SFB_OK;
}

// ---- compilation -----------------------------------------------------------------------------
int sfb_compile(const char* source, const char* file_name, int num_options, const char* const* options,
                void** image, size_t* image_size, char** log) {
    if (!source || !image || !image_size) return fail(SFB_ERR_INVALID, "NULL argument");
    *image = nullptr;
    *image_size = 0;
    if (log) *log = nullptr;
    int rc = load_nvrtc();
    if (rc != SFB_OK) return rc;
    void* prog = nullptr;
    int r = g_nvrtc.CreateProgram(&prog, source, file_name ? file_name : "sfb200_program.cu", 0, nullptr, nullptr);
    if (r != 0) return fail(SFB_ERR_COMPILE, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    int cr = g_nvrtc.CompileProgram(prog, num_options, options);
    size_t log_size = 0;
    g_nvrtc.GetProgramLogSize(prog, &log_size);
    std::string text(log_size ? log_size : 1, '\0');
    if (log_size) g_nvrtc.GetProgramLog(prog, &text[0]);
    if (log) {
        *log = (char*)malloc(text.size() + 1);
        if (*log) {
            memcpy(*log, text.data(), text.size());
            (*log)[text.size()] = 0;
        }
    }
    if (cr != 0) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(SFB_ERR_COMPILE, "nvrtcCompileProgram: %s\n%.1500s", g_nvrtc.GetErrorString(cr), text.c_str());
    }
    size_t sz = 0;
    r = g_nvrtc.GetCUBINSize(prog, &sz);
    if (r != 0 || sz == 0) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(SFB_ERR_COMPILE, "nvrtcGetCUBINSize: %s (use a real architecture, e.g. -arch=sm_100a)",
                    g_nvrtc.GetErrorString(r));
    }
    char* buf = (char*)malloc(sz);
    if (!buf) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(SFB_ERR_OOM, "malloc(%zu)", sz);
    }
    r = g_nvrtc.GetCUBIN(prog, buf);
    g_nvrtc.DestroyProgram(&prog);
    if (r != 0) {
        free(buf);
        return fail(SFB_ERR_COMPILE, "nvrtcGetCUBIN: %s", g_nvrtc.GetErrorString(r));
    }
    *image = buf;
    *image_size = sz;
    return SFB_OK;
}

void sfb_free_host(void* p) { free(p); }

// ---- modules ---------------------------------------------------------------------------------
int sfb_module_load(const void* image, size_t image_size, void** module) {
    (void)image_size;
    if (!image || !module) return fail(SFB_ERR_INVALID, "NULL argument");
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    CUmodule m = nullptr;
    SFB_DRV(ModuleLoadData(&m, image));
    *module = m;
    return SFB_OK;
}

int sfb_module_unload(void* module) {
    if (!module) return SFB_OK;
    SFB_DRV(ModuleUnload((CUmodule)module));
    return SFB_OK;
}

int sfb_module_get_function(void* module, const char* name, void** function) {
    if (!module || !name || !function) return fail(SFB_ERR_INVALID, "NULL argument");
    int rc = load_driver();
    if (rc != SFB_OK) return rc;
    CUfunction f = nullptr;
    CUresult r = g_drv.ModuleGetFunction(&f, (CUmodule)module, name);
    if (r == CUDA_ERROR_NOT_FOUND) return fail(SFB_ERR_NOT_FOUND, "kernel %s not found in module", name);
    if (r != CUDA_SUCCESS) return drv_fail("cuModuleGetFunction", r);
    *function = f;
    return SFB_OK;
}

int sfb_function_set_max_dynamic_smem(void* function, int bytes) {
    if (!function) return fail(SFB_ERR_INVALID, "NULL function");
    SFB_DRV(FuncSetAttribute((CUfunction)function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes));
    return SFB_OK;
}

int sfb_function_attributes(void* function, int* num_regs, int* static_smem, int* local_bytes,
                            int* max_threads) {
    if (!function) return fail(SFB_ERR_INVALID, "NULL function");
    int v = 0;
    if (num_regs) { SFB_DRV(FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, (CUfunction)function)); *num_regs = v; }
    if (static_smem) { SFB_DRV(FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, (CUfunction)function)); *static_smem = v; }
    if (local_bytes) { SFB_DRV(FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, (CUfunction)function)); *local_bytes = v; }
    if (max_threads) { SFB_DRV(FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, (CUfunction)function)); *max_threads = v; }
    return SFB_OK;
}

int sfb_occupancy(void* function, int block_threads, size_t dynamic_smem, int* blocks_per_sm) {
    if (!function || !blocks_per_sm) return fail(SFB_ERR_INVALID, "NULL argument");
    SFB_DRV(OccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, (CUfunction)function, block_threads, dynamic_smem));
    return SFB_OK;
}

int sfb_launch(void* function, const unsigned grid[3], const unsigned block[3], unsigned dynamic_smem,
               void* stream, void** kernel_params) {
    if (!function || !grid || !block) return fail(SFB_ERR_INVALID, "NULL argument");
    if (!grid[0] || !grid[1] || !grid[2] || !block[0] || !block[1] || !block[2])
        return fail(SFB_ERR_INVALID, "empty launch grid=(%u,%u,%u) block=(%u,%u,%u)", grid[0], grid[1], grid[2],
                    block[0], block[1], block[2]);
    SFB_DRV(LaunchKernel((CUfunction)function, grid[0], grid[1], grid[2], block[0], block[1], block[2],
                         dynamic_smem, (CUstream)stream, kernel_params, nullptr));
    return SFB_OK;
}

// ---- TMA descriptors -------------------------------------------------------------------------
int sfb_tensor_map_tiled(void* out_map, int dtype, int rank, void* global_address, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int l2_promotion_bytes) {
    if (!out_map || !global_address || !dims || !box) return fail(SFB_ERR_INVALID, "NULL argument");
    if (rank < 1 || rank > 5) return fail(SFB_ERR_INVALID, "rank %d not in 1..5", rank);
    if (((uintptr_t)out_map & 63) != 0) return fail(SFB_ERR_INVALID, "out_map must be 64-byte aligned");
    CUtensorMapDataType dt;
    switch (dtype) {
        case SFB_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
        case SFB_F64: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; break;
        case SFB_I32: dt = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
        case SFB_I64: dt = CU_TENSOR_MAP_DATA_TYPE_INT64; break;
        default: return fail(SFB_ERR_INVALID, "unsupported dtype %d", dtype);
    }
    cuuint64_t gdims[5], gstrides[4];
    cuuint32_t gbox[5], estr[5];
    for (int d = 0; d < rank; ++d) {
        gdims[d] = dims[d];
        gbox[d] = box[d];
        estr[d] = 1;
        if (d > 0) gstrides[d - 1] = strides_bytes[d - 1];
    }
    CUtensorMapL2promotion promo = l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                   : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                   : l2_promotion_bytes >= 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                               : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    SFB_DRV(TensorMapEncodeTiled((CUtensorMap*)out_map, dt, (cuuint32_t)rank, global_address, gdims, gstrides,
                                 gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    return SFB_OK;
}

// ---- memory ----------------------------------------------------------------------------------
int sfb_malloc(void** dptr, size_t bytes) {
    if (!dptr) return fail(SFB_ERR_INVALID, "dptr is NULL");
    *dptr = nullptr;
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    SFB_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return SFB_OK;
}

int sfb_free(void* dptr) {
    if (!dptr) return SFB_OK;
    SFB_CUDA(cudaFree(dptr));
    return SFB_OK;
}

int sfb_memset(void* dptr, int byte_value, size_t bytes, void* stream) {
    SFB_CUDA(cudaMemsetAsync(dptr, byte_value, bytes, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_host_alloc(void** hptr, size_t bytes) {
    if (!hptr) return fail(SFB_ERR_INVALID, "hptr is NULL");
    *hptr = nullptr;
    SFB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SFB_OK;
}

int sfb_host_free(void* hptr) {
    if (!hptr) return SFB_OK;
    SFB_CUDA(cudaFreeHost(hptr));
    return SFB_OK;
}

int sfb_host_register(void* hptr, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    {   // huge pages: fewer, larger DMA descriptors -- registered 4 KiB pages were measured 12 % (copies in both
        // directions at once) to 25 % (the overlapped host call) slower than cudaHostAlloc memory; advice only
        uintptr_t lo = ((uintptr_t)hptr + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)hptr + bytes) & ~(uintptr_t)4095;
        if (hi > lo) {
            (void)madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
#ifndef MADV_COLLAPSE
#define MADV_COLLAPSE 25                /* Linux >= 6.1: collapse what is already resident, synchronously */
#endif
            if (!getenv("SFB200_NO_COLLAPSE")) (void)madvise((void*)lo, hi - lo, MADV_COLLAPSE);
        }
    }
#endif
    cudaError_t e = cudaHostRegister(hptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();       // page-locked already (sfb_host_alloc memory, an earlier registration)
        return 1;
    }
    SFB_CUDA(e);
    return SFB_OK;
}

int sfb_host_unregister(void* hptr) {
    SFB_CUDA(cudaHostUnregister(hptr));
    return SFB_OK;
}

int sfb_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream) {
    SFB_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream) {
    SFB_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    SFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_mem_info(uint64_t* free_bytes, uint64_t* total_bytes) {
    size_t f = 0, t = 0;
    SFB_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SFB_OK;
}

// ---- streams, events, graphs -----------------------------------------------------------------
int sfb_stream_create(void** stream) {
    if (!stream) return fail(SFB_ERR_INVALID, "stream is NULL");
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    cudaStream_t s;
    SFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return SFB_OK;
}

int sfb_stream_destroy(void* stream) {
    if (!stream) return SFB_OK;
    SFB_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return SFB_OK;
}

int sfb_stream_synchronize(void* stream) {
    SFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return SFB_OK;
}

int sfb_event_create(void** event, int timing) {
    if (!event) return fail(SFB_ERR_INVALID, "event is NULL");
    cudaEvent_t e;
    SFB_CUDA(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
    *event = e;
    return SFB_OK;
}

int sfb_event_destroy(void* event) {
    if (!event) return SFB_OK;
    SFB_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return SFB_OK;
}

int sfb_event_record(void* event, void* stream) {
    SFB_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_event_synchronize(void* event) {
    SFB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return SFB_OK;
}

int sfb_stream_wait_event(void* stream, void* event) {
    SFB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return SFB_OK;
}

int sfb_event_elapsed_ms(void* start, void* stop, float* ms) {
    if (!ms) return fail(SFB_ERR_INVALID, "ms is NULL");
    SFB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return SFB_OK;
}

int sfb_graph_begin_capture(void* stream) {
    SFB_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return SFB_OK;
}

int sfb_graph_end_capture(void* stream, void** graph_exec) {
    if (!graph_exec) return fail(SFB_ERR_INVALID, "graph_exec is NULL");
    cudaGraph_t g = nullptr;
    SFB_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t err = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
    *graph_exec = e;
    return SFB_OK;
}

int sfb_graph_launch(void* graph_exec, void* stream) {
    SFB_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return SFB_OK;
}

int sfb_graph_destroy(void* graph_exec) {
    if (!graph_exec) return SFB_OK;
    SFB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return SFB_OK;
}

// ---- built-in utilities ----------------------------------------------------------------------
int sfb_fill_constant(void* dptr, uint64_t n, int dtype, double value, void* stream) {
    if (n == 0) return SFB_OK;
    if (!dptr) return fail(SFB_ERR_INVALID, "dptr is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = grid_for(n, 256);
    (void)cudaGetLastError();           // the check below is about this launch only
    switch (dtype) {
        case SFB_F32: k_fill_constant<float><<<g, 256, 0, s>>>((float*)dptr, n, (float)value); break;
        case SFB_F64: k_fill_constant<double><<<g, 256, 0, s>>>((double*)dptr, n, value); break;
        case SFB_I32: k_fill_constant<int><<<g, 256, 0, s>>>((int*)dptr, n, (int)value); break;
        case SFB_I64: k_fill_constant<long long><<<g, 256, 0, s>>>((long long*)dptr, n, (long long)value); break;
        default: return fail(SFB_ERR_INVALID, "unsupported dtype %d", dtype);
    }
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
}

int sfb_fill_hash(void* dptr, uint64_t n, int dtype, uint64_t seed, double lo, double hi,
                  uint64_t index_offset, void* stream) {
    if (n == 0) return SFB_OK;
    if (!dptr) return fail(SFB_ERR_INVALID, "dptr is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = grid_for(n, 256);
    (void)cudaGetLastError();           // the check below is about this launch only
    switch (dtype) {
        case SFB_F32: k_fill_hash<float><<<g, 256, 0, s>>>((float*)dptr, n, seed, lo, hi, index_offset); break;
        case SFB_F64: k_fill_hash<double><<<g, 256, 0, s>>>((double*)dptr, n, seed, lo, hi, index_offset); break;
        default: return fail(SFB_ERR_INVALID, "sfb_fill_hash supports float32/float64 only");
    }
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
}

int sfb_checksum(const void* dptr, uint64_t n, int dtype, double* sum, uint64_t* bits) {
    if (!sum || !bits) return fail(SFB_ERR_INVALID, "NULL argument");
    *sum = 0.0;
    *bits = 0;
    if (n == 0) return SFB_OK;
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    void* scratch = nullptr;
    SFB_CUDA(cudaMalloc(&scratch, 16));
    cudaMemset(scratch, 0, 16);
    double* dsum = (double*)scratch;
    unsigned long long* dbits = (unsigned long long*)((char*)scratch + 8);
    unsigned g = grid_for(n, 256);
    switch (dtype) {
        case SFB_F32: k_checksum<float, uint32_t><<<g, 256>>>((const float*)dptr, n, dsum, dbits); break;
        case SFB_F64: k_checksum<double, uint64_t><<<g, 256>>>((const double*)dptr, n, dsum, dbits); break;
        case SFB_I32: k_checksum<int, uint32_t><<<g, 256>>>((const int*)dptr, n, dsum, dbits); break;
        case SFB_I64: k_checksum<long long, uint64_t><<<g, 256>>>((const long long*)dptr, n, dsum, dbits); break;
        default: cudaFree(scratch); return fail(SFB_ERR_INVALID, "unsupported dtype %d", dtype);
    }
    char host[16];
    cudaError_t e = cudaMemcpy(host, scratch, 16, cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_checksum: %s", cudaGetErrorString(e));
    memcpy(sum, host, 8);
    memcpy(bits, host + 8, 8);
    return SFB_OK;
}

int sfb_compare(const void* ref, const void* res, uint64_t n, int dtype, double tolerance,
                double* max_rel_err, uint64_t* num_bad) {
    if (!max_rel_err || !num_bad) return fail(SFB_ERR_INVALID, "NULL argument");
    *max_rel_err = 0.0;
    *num_bad = 0;
    if (n == 0) return SFB_OK;
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    void* scratch = nullptr;
    SFB_CUDA(cudaMalloc(&scratch, 16));
    cudaMemset(scratch, 0, 16);
    unsigned long long* dmax = (unsigned long long*)scratch;
    unsigned long long* dbad = dmax + 1;
    unsigned g = grid_for(n, 256);
    switch (dtype) {
        case SFB_F32: k_compare<float><<<g, 256>>>((const float*)ref, (const float*)res, n, (double)FLT_EPSILON, tolerance, dmax, dbad); break;
        case SFB_F64: k_compare<double><<<g, 256>>>((const double*)ref, (const double*)res, n, DBL_EPSILON, tolerance, dmax, dbad); break;
        default: cudaFree(scratch); return fail(SFB_ERR_INVALID, "sfb_compare supports float32/float64 only");
    }
    unsigned long long host[2];
    cudaError_t e = cudaMemcpy(host, scratch, 16, cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_compare: %s", cudaGetErrorString(e));
    memcpy(max_rel_err, &host[0], 8);
    *num_bad = host[1];
    return SFB_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------
int sfb_ipc_get_handle(void* dptr, void* handle_out) {
    if (!dptr || !handle_out) return fail(SFB_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SFB_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    SFB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle_out, &h, sizeof(h));
    return SFB_OK;
}

int sfb_ipc_open_handle(const void* handle, void** peer_dptr) {
    if (!handle || !peer_dptr) return fail(SFB_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SFB_CUDA(cudaIpcOpenMemHandle(peer_dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SFB_OK;
}

int sfb_ipc_close_handle(void* peer_dptr) {
    if (!peer_dptr) return SFB_OK;
    SFB_CUDA(cudaIpcCloseMemHandle(peer_dptr));
    return SFB_OK;
}

int sfb_enable_peer_access(int peer_device) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return SFB_OK;
    }
    SFB_CUDA(e);
    return SFB_OK;
}

int sfb_stream_write_flag(void* stream, void* flag_dptr, uint32_t value) {
    if (!flag_dptr) return fail(SFB_ERR_INVALID, "flag is NULL");
    SFB_DRV(StreamWriteValue32((CUstream)stream, (CUdeviceptr)flag_dptr, value, CU_STREAM_WRITE_VALUE_DEFAULT));
    return SFB_OK;
}

int sfb_stream_wait_flag(void* stream, void* flag_dptr, uint32_t value) {
    if (!flag_dptr) return fail(SFB_ERR_INVALID, "flag is NULL");
    SFB_DRV(StreamWaitValue32((CUstream)stream, (CUdeviceptr)flag_dptr, value, CU_STREAM_WAIT_VALUE_GEQ));
    return SFB_OK;
}

// ---- per-program handle ----------------------------------------------------------------------
// What DaCe generates per program -- __dace_init_<name> -> handle, __program_<name>(handle, args...),
// __dace_exit_<name>(handle) (dace/dace/codegen/compiled_sdfg.py:182-185,256-294) -- as one object of this
// program-independent library: the loaded module, the device fields, every launch of the plan with its
// parameter block (device pointers, TMA descriptors, work tables patched in here), the caller's host arrays.
struct sfb_program {
    struct Buffer {
        std::string name;
        void* dptr = nullptr;
        size_t bytes = 0;
        int storage = -1;          // index of the buffer that owns the memory (itself if it does)
        void* host = nullptr;      // bound host array (caller-owned)
        size_t host_bytes = 0;
        int is_output = 0;
    };
    struct alignas(64) TensorMap { unsigned char bytes[128]; };
    struct Launch {
        CUfunction fn = nullptr;
        unsigned grid[3] = {1, 1, 1}, block[3] = {1, 1, 1}, smem = 0;
        std::vector<std::vector<unsigned char>> values;    // one entry per parameter (by value)
        std::vector<TensorMap*> maps;                       // 64-byte aligned descriptors live here
        std::vector<void*> tables;                          // device memory of work tables
        std::vector<void*> ptrs;                            // cuLaunchKernel's kernelParams
    };
    CUmodule module = nullptr;
    std::vector<Buffer> buffers;
    std::vector<Launch*> launches;
};

static int program_find(sfb_program* p, const char* field) {
    for (size_t i = 0; i < p->buffers.size(); ++i)
        if (p->buffers[i].name == field) return (int)i;
    return -1;
}

static void program_free_launch(sfb_program::Launch* l) {
    for (auto* m : l->maps) delete m;
    for (void* t : l->tables) cudaFree(t);
    delete l;
}

int sfb_program_create(const void* image, size_t image_size, sfb_program** out) {
    (void)image_size;
    if (!image || !out) return fail(SFB_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = require_init();
    if (rc != SFB_OK) return rc;
    CUmodule m = nullptr;
    SFB_DRV(ModuleLoadData(&m, image));
    sfb_program* p = new sfb_program();
    p->module = m;
    *out = p;
    return SFB_OK;
}

int sfb_program_add_buffer(sfb_program* p, const char* field, size_t bytes, int share_with, int* index) {
    if (!p || !field) return fail(SFB_ERR_INVALID, "NULL argument");
    if (program_find(p, field) >= 0) return fail(SFB_ERR_INVALID, "field %s was added already", field);
    sfb_program::Buffer b;
    b.name = field;
    b.bytes = bytes;
    if (share_with >= 0) {
        if (share_with >= (int)p->buffers.size()) return fail(SFB_ERR_INVALID, "share_with %d out of range", share_with);
        int owner = p->buffers[share_with].storage;
        if (p->buffers[owner].bytes < bytes)
            return fail(SFB_ERR_INVALID, "field %s (%zu bytes) does not fit the storage of %s (%zu bytes)", field, bytes,
                        p->buffers[owner].name.c_str(), p->buffers[owner].bytes);
        b.dptr = p->buffers[owner].dptr;
        b.storage = owner;
    } else {
        SFB_CUDA(cudaMalloc(&b.dptr, bytes ? bytes : 1));
        b.storage = (int)p->buffers.size();
    }
    p->buffers.push_back(b);
    if (index) *index = (int)p->buffers.size() - 1;
    return SFB_OK;
}

int sfb_program_buffer(sfb_program* p, const char* field, void** dptr, size_t* bytes) {
    if (!p || !field) return fail(SFB_ERR_INVALID, "NULL argument");
    int i = program_find(p, field);
    if (i < 0) return fail(SFB_ERR_NOT_FOUND, "program has no field %s", field);
    if (dptr) *dptr = p->buffers[i].dptr;
    if (bytes) *bytes = p->buffers[i].bytes;
    return SFB_OK;
}

int sfb_program_add_launch(sfb_program* p, const char* kernel, const unsigned grid[3], const unsigned block[3],
                           unsigned dynamic_smem, int num_params, const sfb_launch_param* params) {
    if (!p || !kernel || !grid || !block || (num_params > 0 && !params)) return fail(SFB_ERR_INVALID, "NULL argument");
    if (!grid[0] || !grid[1] || !grid[2] || !block[0] || !block[1] || !block[2])
        return fail(SFB_ERR_INVALID, "empty launch of %s", kernel);
    int rc = load_driver();
    if (rc != SFB_OK) return rc;
    CUfunction f = nullptr;
    CUresult r = g_drv.ModuleGetFunction(&f, p->module, kernel);
    if (r == CUDA_ERROR_NOT_FOUND) return fail(SFB_ERR_NOT_FOUND, "kernel %s not found in module", kernel);
    if (r != CUDA_SUCCESS) return drv_fail("cuModuleGetFunction", r);
    if (dynamic_smem > 48 * 1024)
        SFB_DRV(FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_smem));
    sfb_program::Launch* l = new sfb_program::Launch();
    l->fn = f;
    for (int d = 0; d < 3; ++d) { l->grid[d] = grid[d]; l->block[d] = block[d]; }
    l->smem = dynamic_smem;
    l->values.resize(num_params);
    for (int k = 0; k < num_params; ++k) {
        const sfb_launch_param& q = params[k];
        int status = SFB_OK;
        switch (q.kind) {
            case SFB_PARAM_BYTES:
                if (!q.data || !q.size) status = fail(SFB_ERR_INVALID, "parameter %d of %s has no value", k, kernel);
                else l->values[k].assign((const unsigned char*)q.data, (const unsigned char*)q.data + q.size);
                break;
            case SFB_PARAM_BUFFER: {
                if (q.buffer < 0 || q.buffer >= (int)p->buffers.size()) { status = fail(SFB_ERR_INVALID, "parameter %d of %s: no buffer %d", k, kernel, q.buffer); break; }
                unsigned char* addr = (unsigned char*)p->buffers[q.buffer].dptr + q.offset;
                l->values[k].assign((unsigned char*)&addr, (unsigned char*)&addr + sizeof(addr));
                break;
            }
            case SFB_PARAM_TMAP: {
                if (q.buffer < 0 || q.buffer >= (int)p->buffers.size()) { status = fail(SFB_ERR_INVALID, "parameter %d of %s: no buffer %d", k, kernel, q.buffer); break; }
                sfb_program::TensorMap* m = new sfb_program::TensorMap();
                l->maps.push_back(m);
                status = sfb_tensor_map_tiled(m->bytes, q.dtype, q.rank, (unsigned char*)p->buffers[q.buffer].dptr + q.offset,
                                              q.dims, q.strides_bytes, q.box, 128);
                break;
            }
            case SFB_PARAM_TABLE: {
                if (!q.data || !q.size) { status = fail(SFB_ERR_INVALID, "parameter %d of %s has no table", k, kernel); break; }
                void* t = nullptr;
                cudaError_t e = cudaMalloc(&t, q.size);
                if (e == cudaSuccess) e = cudaMemcpy(t, q.data, q.size, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) {
                    (void)cudaGetLastError();
                    if (t) cudaFree(t);
                    status = fail(SFB_ERR_CUDA, "work table of %s: %s", kernel, cudaGetErrorString(e));
                    break;
                }
                l->tables.push_back(t);
                l->values[k].assign((unsigned char*)&t, (unsigned char*)&t + sizeof(t));
                break;
            }
            default:
                status = fail(SFB_ERR_INVALID, "parameter %d of %s: unknown kind %d", k, kernel, q.kind);
        }
        if (status != SFB_OK) {
            program_free_launch(l);
            return status;
        }
    }
    // kernelParams: pointers to the values; tensor maps are passed by value from their aligned storage
    size_t next_map = 0;
    for (int k = 0; k < num_params; ++k) {
        if (params[k].kind == SFB_PARAM_TMAP) l->ptrs.push_back(l->maps[next_map++]->bytes);
        else l->ptrs.push_back(l->values[k].data());
    }
    p->launches.push_back(l);
    return SFB_OK;
}

int sfb_program_clear_launches(sfb_program* p) {
    if (!p) return fail(SFB_ERR_INVALID, "NULL program");
    for (auto* l : p->launches) program_free_launch(l);
    p->launches.clear();
    return SFB_OK;
}

int sfb_program_num_launches(sfb_program* p, int* count) {
    if (!p || !count) return fail(SFB_ERR_INVALID, "NULL argument");
    *count = (int)p->launches.size();
    return SFB_OK;
}

int sfb_program_bind(sfb_program* p, const char* field, void* host_ptr, size_t bytes, int is_output) {
    if (!p || !field) return fail(SFB_ERR_INVALID, "NULL argument");
    int i = program_find(p, field);
    if (i < 0) return fail(SFB_ERR_NOT_FOUND, "program has no field %s", field);
    if (host_ptr && bytes > p->buffers[i].bytes)
        return fail(SFB_ERR_INVALID, "host array of %s has %zu bytes, the field %zu", field, bytes, p->buffers[i].bytes);
    p->buffers[i].host = host_ptr;
    p->buffers[i].host_bytes = host_ptr ? bytes : 0;
    p->buffers[i].is_output = is_output;
    return SFB_OK;
}

int sfb_program_run(sfb_program* p, int repetitions, void* stream, float* ms_out) {
    if (!p) return fail(SFB_ERR_INVALID, "NULL program");
    if (repetitions < 1) return fail(SFB_ERR_INVALID, "repetitions must be positive");
    if (p->launches.empty()) return fail(SFB_ERR_INVALID, "program has no launches");
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) {
        SFB_CUDA(cudaEventCreate(&e0));
        SFB_CUDA(cudaEventCreate(&e1));
        SFB_CUDA(cudaEventRecord(e0, s));
    }
    for (int rep = 0; rep < repetitions; ++rep)
        for (auto* l : p->launches)
            SFB_DRV(LaunchKernel(l->fn, l->grid[0], l->grid[1], l->grid[2], l->block[0], l->block[1], l->block[2],
                                 l->smem, (CUstream)stream, l->ptrs.data(), nullptr));
    if (ms_out) {
        SFB_CUDA(cudaEventRecord(e1, s));
        SFB_CUDA(cudaEventSynchronize(e1));
        SFB_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return SFB_OK;
}

int sfb_program_call(sfb_program* p, void* stream) {
    if (!p) return fail(SFB_ERR_INVALID, "NULL program");
    cudaStream_t s = (cudaStream_t)stream;
    for (auto& b : p->buffers)
        if (b.host && !b.is_output)
            SFB_CUDA(cudaMemcpyAsync(b.dptr, b.host, b.host_bytes, cudaMemcpyHostToDevice, s));
    int rc = sfb_program_run(p, 1, stream, nullptr);
    if (rc != SFB_OK) return rc;
    for (auto& b : p->buffers)
        if (b.host && b.is_output)
            SFB_CUDA(cudaMemcpyAsync(b.host, b.dptr, b.host_bytes, cudaMemcpyDeviceToHost, s));
    SFB_CUDA(cudaStreamSynchronize(s));
    return SFB_OK;
}

int sfb_program_destroy(sfb_program* p) {
    if (!p) return SFB_OK;
    for (auto* l : p->launches) program_free_launch(l);
    for (size_t i = 0; i < p->buffers.size(); ++i)
        if (p->buffers[i].storage == (int)i && p->buffers[i].dptr) cudaFree(p->buffers[i].dptr);
    if (p->module && load_driver() == SFB_OK) g_drv.ModuleUnload(p->module);
    delete p;
    (void)cudaGetLastError();
    return SFB_OK;
}

}  // extern "C"
