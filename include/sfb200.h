/*
 * sfb200.h -- C ABI of libsfb200.so, the B200 (sm_100a) runtime behind StencilFlow's `cuda` mode.
 *
 * This is the drop-in boundary for the stencil-chain execution path.  In the reference the same
 * boundary is DaCe's ctypes loader for a compiled SDFG:
 *     dace/dace/codegen/compiled_sdfg.py:182-185   __dace_init_<name>  -> handle (NULL = failure)
 *     dace/dace/codegen/compiled_sdfg.py:286-294   __program_<name>(handle, args...)
 *     dace/dace/codegen/compiled_sdfg.py:256-267   __dace_exit_<name>(handle)
 *     dace/dace/codegen/tools/dacestub.cpp:1-86    load_library / get_symbol / unload_library
 * and the program object `sdfg.compile()` returns to stencilflow/run_program.py:123,164-172.
 * Where DaCe generates one shared object per program, this library is program-independent:
 * it compiles the generated CUDA C++ of a program (NVRTC), loads it, owns device memory, builds TMA
 * descriptors, launches the kernels, and moves halos between the GPUs of one box.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative sfb_status;
 *     nothing throws or aborts; sfb_last_error() gives the text for the calling thread.
 *   - host buffers are owned by the caller (numpy), device buffers by whoever called sfb_malloc.
 *   - strings are UTF-8, NUL-terminated, borrowed for the duration of the call.
 *   - one device per process (sfb_init); calls are made from one thread at a time.
 *   - functions taking a stream are asynchronous with respect to the host unless stated otherwise.
 */
#ifndef SFB200_H
#define SFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 1

#if defined(__GNUC__)
#define SFB_API __attribute__((visibility("default")))
#else
#define SFB_API
#endif

typedef enum sfb_status {
    SFB_OK = 0,
    SFB_ERR_INVALID = -1,      /* bad argument */
    SFB_ERR_CUDA = -2,         /* CUDA runtime/driver call failed */
    SFB_ERR_COMPILE = -3,      /* NVRTC rejected the source (see log) */
    SFB_ERR_NOT_FOUND = -4,    /* symbol / kernel / library not found */
    SFB_ERR_NO_DEVICE = -5,    /* no usable GPU or driver */
    SFB_ERR_OOM = -6           /* allocation failed */
} sfb_status;

typedef enum sfb_dtype {
    SFB_F32 = 0,
    SFB_F64 = 1,
    SFB_I32 = 2,
    SFB_I64 = 3
} sfb_dtype;

typedef struct sfb_device_props {
    char name[128];
    int cc_major, cc_minor;
    int sm_count;
    int max_smem_per_block_optin;   /* bytes */
    int l2_bytes;
    int clock_khz, mem_clock_khz;
    uint64_t total_mem, free_mem;   /* bytes */
} sfb_device_props;

/* ---- library / device ------------------------------------------------------------------- */
SFB_API int sfb_abi_version(void);
SFB_API const char* sfb_last_error(void);
SFB_API int sfb_device_count(int* count);
/* Binds the calling process to `device`, creates the context and the default stream. */
SFB_API int sfb_init(int device);
SFB_API int sfb_shutdown(void);
SFB_API int sfb_current_device(int* device);
SFB_API int sfb_device_properties(int device, sfb_device_props* out);
SFB_API int sfb_device_synchronize(void);

/* ---- compilation: generated CUDA C++ -> cubin (replaces DaCe's cmake/g++/aoc step,
 *      dace/dace/codegen/compiler.py:27,103) ------------------------------------------------ */
/* `image`/`log` are malloc'd by the library; release with sfb_free_host.  `log` may be NULL. */
SFB_API int sfb_compile(const char* source, const char* file_name, int num_options, const char* const* options,
                void** image, size_t* image_size, char** log);
SFB_API void sfb_free_host(void* p);

/* ---- modules and kernels ---------------------------------------------------------------- */
SFB_API int sfb_module_load(const void* image, size_t image_size, void** module);
SFB_API int sfb_module_unload(void* module);
SFB_API int sfb_module_get_function(void* module, const char* name, void** function);
SFB_API int sfb_function_set_max_dynamic_smem(void* function, int bytes);
SFB_API int sfb_function_attributes(void* function, int* num_regs, int* static_smem, int* local_bytes,
                            int* max_threads);
SFB_API int sfb_occupancy(void* function, int block_threads, size_t dynamic_smem, int* blocks_per_sm);
/* kernel_params: array of `num_params` pointers to the argument values (cuLaunchKernel style). */
SFB_API int sfb_launch(void* function, const unsigned grid[3], const unsigned block[3], unsigned dynamic_smem,
               void* stream, void** kernel_params);

/* ---- TMA descriptors (cuTensorMapEncodeTiled) ------------------------------------------- */
/* Writes a 128-byte CUtensorMap to `out_map` (64-byte aligned).  dims/box are innermost-first;
 * strides_bytes has rank-1 entries (stride of dim 1.., the innermost is dense).  Out-of-bounds
 * elements of a box are filled with zeros. */
SFB_API int sfb_tensor_map_tiled(void* out_map, int dtype, int rank, void* global_address,
                         const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                         int l2_promotion_bytes);

/* ---- memory ------------------------------------------------------------------------------ */
SFB_API int sfb_malloc(void** dptr, size_t bytes);
SFB_API int sfb_free(void* dptr);
SFB_API int sfb_memset(void* dptr, int byte_value, size_t bytes, void* stream);
SFB_API int sfb_host_alloc(void** hptr, size_t bytes);          /* pinned host memory */
SFB_API int sfb_host_free(void* hptr);
/* Page-locks caller-owned memory (numpy arrays of the reference driver, stencilflow/run_program.py:145-159).
 * Returns 1 (not an error) when the range is page-locked already; only a 0 return has to be undone. */
SFB_API int sfb_host_register(void* hptr, size_t bytes);
SFB_API int sfb_host_unregister(void* hptr);
SFB_API int sfb_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream);
SFB_API int sfb_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream);
SFB_API int sfb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);  /* peer-capable */
SFB_API int sfb_mem_info(uint64_t* free_bytes, uint64_t* total_bytes);

/* ---- streams, events, graphs ------------------------------------------------------------- */
SFB_API int sfb_stream_create(void** stream);
SFB_API int sfb_stream_destroy(void* stream);
SFB_API int sfb_stream_synchronize(void* stream);
SFB_API int sfb_event_create(void** event, int timing);
SFB_API int sfb_event_destroy(void* event);
SFB_API int sfb_event_record(void* event, void* stream);
SFB_API int sfb_event_synchronize(void* event);
SFB_API int sfb_stream_wait_event(void* stream, void* event);
SFB_API int sfb_event_elapsed_ms(void* start, void* stop, float* ms);
SFB_API int sfb_graph_begin_capture(void* stream);
SFB_API int sfb_graph_end_capture(void* stream, void** graph_exec);
SFB_API int sfb_graph_launch(void* graph_exec, void* stream);
SFB_API int sfb_graph_destroy(void* graph_exec);

/* ---- built-in device utilities (kernels compiled into the library) ----------------------- */
/* field[n] = value */
SFB_API int sfb_fill_constant(void* dptr, uint64_t n, int dtype, double value, void* stream);
/* field[idx] = lo + (hi-lo) * u(idx+index_offset, seed), u in [0,1): counter-based hash, reproducible on the
 * host (stencilflow_b200/synthetic.py) -- used for synthetic fields that are generated in HBM. */
SFB_API int sfb_fill_hash(void* dptr, uint64_t n, int dtype, uint64_t seed, double lo, double hi,
                  uint64_t index_offset, void* stream);
/* out[0] = sum (double), out[1] = bitwise checksum (wrapping sum of the raw words as uint64, order independent).
 * Synchronous. */
SFB_API int sfb_checksum(const void* dptr, uint64_t n, int dtype, double* sum, uint64_t* bits);
/* max over elements of |ref-res| / (max(|ref|,|res|) + eps(dtype)) and the count of elements above `tolerance`.
 * This is stencilflow/helper.py:261-276 (arrays_are_equal) evaluated on the device.  Synchronous. */
SFB_API int sfb_compare(const void* ref, const void* res, uint64_t n, int dtype, double tolerance,
                double* max_rel_err, uint64_t* num_bad);

/* ---- multi-GPU: slab halos over NVLink (replaces the SMI channels of the multi-FPGA path,
 *      stencilflow/sdfg_generator.py:846-853, bin/run_distributed_program.py:193-202) ------- */
#define SFB_IPC_HANDLE_BYTES 64
SFB_API int sfb_ipc_get_handle(void* dptr, void* handle_out /* 64 bytes */);
SFB_API int sfb_ipc_open_handle(const void* handle /* 64 bytes */, void** peer_dptr);
SFB_API int sfb_ipc_close_handle(void* peer_dptr);
SFB_API int sfb_enable_peer_access(int peer_device);
/* *flag = value after all prior work of `stream` (release semantics); flag may live on a peer GPU. */
SFB_API int sfb_stream_write_flag(void* stream, void* flag_dptr, uint32_t value);
/* blocks `stream` (not the host) until *flag >= value. */
SFB_API int sfb_stream_wait_flag(void* stream, void* flag_dptr, uint32_t value);

/* ---- per-program handle ------------------------------------------------------------------
 * The counterpart of the trio DaCe generates per program and the reference's driver calls through ctypes:
 *     __dace_init_<name>(...) -> handle     dace/dace/codegen/compiled_sdfg.py:182-185
 *     __program_<name>(handle, args...)     dace/dace/codegen/compiled_sdfg.py:286-294
 *     __dace_exit_<name>(handle)            dace/dace/codegen/compiled_sdfg.py:256-267
 * Here the library is program-independent, so the handle is *built*: the compiled image, the fields, and
 * every launch of the plan with a description of its parameters; the library owns the module, the device
 * memory, the TMA descriptors and the work tables, and a step is one call (sfb_program_run / _call).
 * A host in any language binds these nine functions and needs nothing else of this header. */
typedef struct sfb_program sfb_program;

typedef enum sfb_param_kind {
    SFB_PARAM_BYTES = 0,    /* `size` bytes at `data`, passed by value (scalars, plane ranges) */
    SFB_PARAM_BUFFER = 1,   /* device address of field `buffer` + `offset` bytes */
    SFB_PARAM_TMAP = 2,     /* 128-byte tiled TMA descriptor of field `buffer` (+ `offset`), by value */
    SFB_PARAM_TABLE = 3     /* device address of a copy of the `size` bytes at `data` (work lists) */
} sfb_param_kind;

typedef struct sfb_launch_param {
    int32_t kind;
    int32_t buffer;
    uint64_t offset;
    const void* data;
    uint32_t size;
    int32_t dtype, rank;            /* SFB_PARAM_TMAP: sfb_dtype, 1..5; dims/box innermost first */
    uint64_t dims[5];
    uint64_t strides_bytes[4];
    uint32_t box[5];
} sfb_launch_param;

/* Loads the compiled image (cubin from sfb_compile).  *out is NULL on failure. */
SFB_API int sfb_program_create(const void* image, size_t image_size, sfb_program** out);
/* Adds a field of `bytes` bytes of device memory; share_with >= 0 places it in the storage of that
 * earlier field instead (intermediates that are never live together). */
SFB_API int sfb_program_add_buffer(sfb_program* p, const char* field, size_t bytes, int share_with, int* index);
SFB_API int sfb_program_buffer(sfb_program* p, const char* field, void** dptr, size_t* bytes);
/* Appends a launch; parameters are resolved now and kept by the handle. */
SFB_API int sfb_program_add_launch(sfb_program* p, const char* kernel, const unsigned grid[3], const unsigned block[3],
                           unsigned dynamic_smem, int num_params, const sfb_launch_param* params);
SFB_API int sfb_program_clear_launches(sfb_program* p);
SFB_API int sfb_program_num_launches(sfb_program* p, int* count);
/* Caller-owned host array of an input (is_output = 0) or output field; NULL unbinds. */
SFB_API int sfb_program_bind(sfb_program* p, const char* field, void* host_ptr, size_t bytes, int is_output);
/* All launches, `repetitions` times, on `stream`.  ms_out != NULL: blocks and returns the device time
 * (CUDA events); ms_out == NULL: asynchronous. */
SFB_API int sfb_program_run(sfb_program* p, int repetitions, void* stream, float* ms_out);
/* __program_<name>: bound inputs host->device, all launches, bound outputs device->host; blocking. */
SFB_API int sfb_program_call(sfb_program* p, void* stream);
SFB_API int sfb_program_destroy(sfb_program* p);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
