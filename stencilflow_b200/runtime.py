"""ctypes binding of ``libsfb200.so`` (C ABI: ``include/sfb200.h``).

This is the only place the Python host code touches CUDA.  It mirrors the role of DaCe's
``CompiledSDFG`` loader in the reference (``dace/dace/codegen/compiled_sdfg.py:20-150``): load the
shared object, bind the entry points, translate return codes into exceptions.  There is no fallback:
if the library is missing or no GPU is usable the calls raise.
"""

import ctypes
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libsfb200.so")

SFB_OK = 0
STATUS_NAMES = {
    0: "SFB_OK", -1: "SFB_ERR_INVALID", -2: "SFB_ERR_CUDA", -3: "SFB_ERR_COMPILE",
    -4: "SFB_ERR_NOT_FOUND", -5: "SFB_ERR_NO_DEVICE", -6: "SFB_ERR_OOM",
}
DTYPE_CODES = {"float32": 0, "float64": 1, "int32": 2, "int64": 3}


class SfbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("{}: {}".format(STATUS_NAMES.get(status, status), message))
        self.status = status


class DeviceProps(ctypes.Structure):
    _fields_ = [
        ("name", ctypes.c_char * 128),
        ("cc_major", ctypes.c_int), ("cc_minor", ctypes.c_int),
        ("sm_count", ctypes.c_int),
        ("max_smem_per_block_optin", ctypes.c_int),
        ("l2_bytes", ctypes.c_int),
        ("clock_khz", ctypes.c_int), ("mem_clock_khz", ctypes.c_int),
        ("total_mem", ctypes.c_uint64), ("free_mem", ctypes.c_uint64),
    ]


_vp = ctypes.c_void_p
_vpp = ctypes.POINTER(ctypes.c_void_p)
_u3 = ctypes.c_uint * 3

PARAM_BYTES, PARAM_BUFFER, PARAM_TMAP, PARAM_TABLE = 0, 1, 2, 3


class LaunchParam(ctypes.Structure):
    """``sfb_launch_param`` (include/sfb200.h)."""
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("buffer", ctypes.c_int32),
        ("offset", ctypes.c_uint64),
        ("data", ctypes.c_void_p),
        ("size", ctypes.c_uint32),
        ("dtype", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("dims", ctypes.c_uint64 * 5),
        ("strides_bytes", ctypes.c_uint64 * 4),
        ("box", ctypes.c_uint32 * 5),
    ]

# name -> (restype, argtypes); the table doubles as the list the CPU tests check against sfb200.h
PROTOTYPES = {
    "sfb_abi_version": (ctypes.c_int, []),
    "sfb_last_error": (ctypes.c_char_p, []),
    "sfb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "sfb_init": (ctypes.c_int, [ctypes.c_int]),
    "sfb_shutdown": (ctypes.c_int, []),
    "sfb_current_device": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "sfb_device_properties": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(DeviceProps)]),
    "sfb_device_synchronize": (ctypes.c_int, []),
    "sfb_compile": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_char_p), _vpp,
                                   ctypes.POINTER(ctypes.c_size_t), _vpp]),
    "sfb_free_host": (None, [_vp]),
    "sfb_module_load": (ctypes.c_int, [_vp, ctypes.c_size_t, _vpp]),
    "sfb_module_unload": (ctypes.c_int, [_vp]),
    "sfb_module_get_function": (ctypes.c_int, [_vp, ctypes.c_char_p, _vpp]),
    "sfb_function_set_max_dynamic_smem": (ctypes.c_int, [_vp, ctypes.c_int]),
    "sfb_function_attributes": (ctypes.c_int, [_vp] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "sfb_occupancy": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]),
    "sfb_launch": (ctypes.c_int, [_vp, _u3, _u3, ctypes.c_uint, _vp, _vpp]),
    "sfb_tensor_map_tiled": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp,
                                            ctypes.POINTER(ctypes.c_uint64),
                                            ctypes.POINTER(ctypes.c_uint64),
                                            ctypes.POINTER(ctypes.c_uint32), ctypes.c_int]),
    "sfb_malloc": (ctypes.c_int, [_vpp, ctypes.c_size_t]),
    "sfb_free": (ctypes.c_int, [_vp]),
    "sfb_memset": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_size_t, _vp]),
    "sfb_host_alloc": (ctypes.c_int, [_vpp, ctypes.c_size_t]),
    "sfb_host_free": (ctypes.c_int, [_vp]),
    "sfb_host_register": (ctypes.c_int, [_vp, ctypes.c_size_t]),
    "sfb_host_unregister": (ctypes.c_int, [_vp]),
    "sfb_memcpy_h2d": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "sfb_memcpy_d2h": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "sfb_memcpy_d2d": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "sfb_mem_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64)] * 2),
    "sfb_stream_create": (ctypes.c_int, [_vpp]),
    "sfb_stream_destroy": (ctypes.c_int, [_vp]),
    "sfb_stream_synchronize": (ctypes.c_int, [_vp]),
    "sfb_event_create": (ctypes.c_int, [_vpp, ctypes.c_int]),
    "sfb_event_destroy": (ctypes.c_int, [_vp]),
    "sfb_event_record": (ctypes.c_int, [_vp, _vp]),
    "sfb_event_synchronize": (ctypes.c_int, [_vp]),
    "sfb_stream_wait_event": (ctypes.c_int, [_vp, _vp]),
    "sfb_event_elapsed_ms": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "sfb_graph_begin_capture": (ctypes.c_int, [_vp]),
    "sfb_graph_end_capture": (ctypes.c_int, [_vp, _vpp]),
    "sfb_graph_launch": (ctypes.c_int, [_vp, _vp]),
    "sfb_graph_destroy": (ctypes.c_int, [_vp]),
    "sfb_fill_constant": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_double, _vp]),
    "sfb_fill_hash": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_uint64, _vp]),
    "sfb_checksum": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "sfb_compare": (ctypes.c_int, [_vp, _vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_double,
                                   ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "sfb_ipc_get_handle": (ctypes.c_int, [_vp, _vp]),
    "sfb_ipc_open_handle": (ctypes.c_int, [_vp, _vpp]),
    "sfb_ipc_close_handle": (ctypes.c_int, [_vp]),
    "sfb_enable_peer_access": (ctypes.c_int, [ctypes.c_int]),
    "sfb_stream_write_flag": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32]),
    "sfb_stream_wait_flag": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32]),
    "sfb_program_create": (ctypes.c_int, [_vp, ctypes.c_size_t, _vpp]),
    "sfb_program_add_buffer": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_int)]),
    "sfb_program_buffer": (ctypes.c_int, [_vp, ctypes.c_char_p, _vpp, ctypes.POINTER(ctypes.c_size_t)]),
    "sfb_program_add_launch": (ctypes.c_int, [_vp, ctypes.c_char_p, _u3, _u3, ctypes.c_uint, ctypes.c_int,
                                              ctypes.POINTER(LaunchParam)]),
    "sfb_program_clear_launches": (ctypes.c_int, [_vp]),
    "sfb_program_num_launches": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int)]),
    "sfb_program_bind": (ctypes.c_int, [_vp, ctypes.c_char_p, _vp, ctypes.c_size_t, ctypes.c_int]),
    "sfb_program_run": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_float)]),
    "sfb_program_call": (ctypes.c_int, [_vp, _vp]),
    "sfb_program_destroy": (ctypes.c_int, [_vp]),
}

_lib = None


def load_library(path=None):
    """Load ``libsfb200.so`` and bind every entry point.  Raises if it has not been built
    (``python -c 'import __graft_entry__ as g; g.build()'`` or ``stencilflow_b200.build.build_native()``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise SfbError(-4, "{} not found: build the CUDA runtime first "
                           "(stencilflow_b200.build.build_native()); there is no CPU fallback".format(path))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sfb_abi_version() != 1:
        raise SfbError(-1, "ABI version mismatch")
    _lib = lib
    return lib


def _check(status):
    if status != SFB_OK:
        raise SfbError(status, (_lib.sfb_last_error() or b"").decode("utf-8", "replace"))


def dtype_code(dtype):
    name = dtype.name if hasattr(dtype, "name") and not isinstance(dtype, np.dtype) else np.dtype(dtype).name
    try:
        return DTYPE_CODES[name]
    except KeyError:
        raise SfbError(-1, "dtype {} is not supported on the device".format(name))


class Runtime:
    """One process, one GPU.  Thin object wrapper over the C ABI."""

    _instance = None

    def __init__(self, device=None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("SFB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _check(self.lib.sfb_init(device))
        self.device = device
        self.props = self.device_properties(device)
        self.stream = self.stream_create()

    @classmethod
    def get(cls, device=None):
        if cls._instance is None:
            cls._instance = Runtime(device)
        elif device is not None and device != cls._instance.device:
            raise SfbError(-1, "runtime already bound to device {}".format(cls._instance.device))
        return cls._instance

    # -- device
    def device_count(self):
        n = ctypes.c_int(0)
        _check(self.lib.sfb_device_count(ctypes.byref(n)))
        return n.value

    def device_properties(self, device=None):
        p = DeviceProps()
        _check(self.lib.sfb_device_properties(self.device if device is None else device, ctypes.byref(p)))
        return p

    def synchronize(self):
        _check(self.lib.sfb_device_synchronize())

    def mem_info(self):
        f, t = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(self.lib.sfb_mem_info(ctypes.byref(f), ctypes.byref(t)))
        return f.value, t.value

    # -- compile / modules
    def compile(self, source, file_name, options):
        return compile_source(source, file_name, options)

    def module_load(self, image):
        buf = ctypes.create_string_buffer(image, len(image))
        mod = _vp()
        _check(self.lib.sfb_module_load(buf, len(image), ctypes.byref(mod)))
        return mod

    def module_unload(self, module):
        _check(self.lib.sfb_module_unload(module))

    def get_function(self, module, name):
        fn = _vp()
        _check(self.lib.sfb_module_get_function(module, name.encode(), ctypes.byref(fn)))
        return fn

    def set_max_dynamic_smem(self, function, nbytes):
        _check(self.lib.sfb_function_set_max_dynamic_smem(function, int(nbytes)))

    def function_attributes(self, function):
        vals = [ctypes.c_int(0) for _ in range(4)]
        _check(self.lib.sfb_function_attributes(function, *[ctypes.byref(v) for v in vals]))
        return dict(zip(("num_regs", "static_smem", "local_bytes", "max_threads"), (v.value for v in vals)))

    def occupancy(self, function, block_threads, dynamic_smem):
        n = ctypes.c_int(0)
        _check(self.lib.sfb_occupancy(function, block_threads, dynamic_smem, ctypes.byref(n)))
        return n.value

    def launch(self, function, grid, block, smem, params, stream=None):
        """``params``: ctypes array of void* built by :func:`pack_params`."""
        _check(self.lib.sfb_launch(function, _u3(*grid), _u3(*block), int(smem),
                                   self.stream if stream is None else stream, params))

    def tensor_map(self, dptr, dtype, dims, strides_bytes, box, l2_promotion=128):
        """128-byte TMA descriptor (innermost-first dims/box) as an aligned ctypes buffer."""
        raw = ctypes.create_string_buffer(128 + 64)
        addr = (ctypes.addressof(raw) + 63) & ~63
        rank = len(dims)
        _check(self.lib.sfb_tensor_map_tiled(
            _vp(addr), dtype_code(dtype), rank, _vp(dptr),
            (ctypes.c_uint64 * rank)(*dims),
            (ctypes.c_uint64 * max(1, rank - 1))(*strides_bytes),
            (ctypes.c_uint32 * rank)(*box), l2_promotion))
        return raw, addr

    # -- memory
    def malloc(self, nbytes):
        p = _vp()
        _check(self.lib.sfb_malloc(ctypes.byref(p), int(nbytes)))
        return p.value

    def free(self, dptr):
        if dptr:
            _check(self.lib.sfb_free(_vp(dptr)))

    def memset(self, dptr, value, nbytes, stream=None):
        _check(self.lib.sfb_memset(_vp(dptr), value, int(nbytes), self.stream if stream is None else stream))

    def host_alloc(self, shape, dtype):
        """Pinned host array (numpy view over cudaHostAlloc memory); free with host_free."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if len(shape) else 1
        p = _vp()
        _check(self.lib.sfb_host_alloc(ctypes.byref(p), max(1, n * dtype.itemsize)))
        buf = (ctypes.c_char * (n * dtype.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        return arr, p.value

    def host_free(self, hptr):
        _check(self.lib.sfb_host_free(_vp(hptr)))

    def host_register(self, arr):
        """True if this call page-locked the array (and ``host_unregister`` has to undo it), False if
        it was page-locked already."""
        status = self.lib.sfb_host_register(_vp(arr.ctypes.data), arr.nbytes)
        if status == 1:
            return False
        _check(status)
        return True

    def host_unregister(self, arr):
        _check(self.lib.sfb_host_unregister(_vp(arr.ctypes.data)))

    def h2d(self, dptr, arr, stream=None, nbytes=None):
        assert arr.flags["C_CONTIGUOUS"]
        _check(self.lib.sfb_memcpy_h2d(_vp(dptr), _vp(arr.ctypes.data),
                                       arr.nbytes if nbytes is None else nbytes,
                                       self.stream if stream is None else stream))

    def d2h(self, arr, dptr, stream=None, nbytes=None):
        assert arr.flags["C_CONTIGUOUS"] and arr.flags["WRITEABLE"]
        _check(self.lib.sfb_memcpy_d2h(_vp(arr.ctypes.data), _vp(dptr),
                                       arr.nbytes if nbytes is None else nbytes,
                                       self.stream if stream is None else stream))

    def d2d(self, dst, src, nbytes, stream=None):
        _check(self.lib.sfb_memcpy_d2d(_vp(dst), _vp(src), int(nbytes), self.stream if stream is None else stream))

    # -- streams / events / graphs
    def stream_create(self):
        s = _vp()
        _check(self.lib.sfb_stream_create(ctypes.byref(s)))
        return s

    def stream_synchronize(self, stream=None):
        _check(self.lib.sfb_stream_synchronize(self.stream if stream is None else stream))

    def event_create(self, timing=True):
        e = _vp()
        _check(self.lib.sfb_event_create(ctypes.byref(e), 1 if timing else 0))
        return e

    def event_destroy(self, event):
        _check(self.lib.sfb_event_destroy(event))

    def event_record(self, event, stream=None):
        _check(self.lib.sfb_event_record(event, self.stream if stream is None else stream))

    def event_synchronize(self, event):
        _check(self.lib.sfb_event_synchronize(event))

    def stream_wait_event(self, stream, event):
        """``stream`` None = the runtime's launch stream (not the legacy NULL stream)."""
        _check(self.lib.sfb_stream_wait_event(self.stream if stream is None else stream, event))

    def elapsed_ms(self, start, stop):
        ms = ctypes.c_float(0)
        _check(self.lib.sfb_event_elapsed_ms(start, stop, ctypes.byref(ms)))
        return ms.value

    def graph_begin(self, stream=None):
        _check(self.lib.sfb_graph_begin_capture(self.stream if stream is None else stream))

    def graph_end(self, stream=None):
        g = _vp()
        _check(self.lib.sfb_graph_end_capture(self.stream if stream is None else stream, ctypes.byref(g)))
        return g

    def graph_launch(self, graph, stream=None):
        _check(self.lib.sfb_graph_launch(graph, self.stream if stream is None else stream))

    def graph_destroy(self, graph):
        _check(self.lib.sfb_graph_destroy(graph))

    # -- utilities
    def fill_constant(self, dptr, n, dtype, value, stream=None):
        _check(self.lib.sfb_fill_constant(_vp(dptr), n, dtype_code(dtype), float(value),
                                          self.stream if stream is None else stream))

    def fill_hash(self, dptr, n, dtype, seed, lo=0.0, hi=1.0, index_offset=0, stream=None):
        _check(self.lib.sfb_fill_hash(_vp(dptr), n, dtype_code(dtype), seed, lo, hi, index_offset,
                                      self.stream if stream is None else stream))

    def checksum(self, dptr, n, dtype):
        self.stream_synchronize()
        s, b = ctypes.c_double(0), ctypes.c_uint64(0)
        _check(self.lib.sfb_checksum(_vp(dptr), n, dtype_code(dtype), ctypes.byref(s), ctypes.byref(b)))
        return s.value, b.value

    def compare(self, ref_dptr, res_dptr, n, dtype, tolerance):
        self.stream_synchronize()
        m, bad = ctypes.c_double(0), ctypes.c_uint64(0)
        _check(self.lib.sfb_compare(_vp(ref_dptr), _vp(res_dptr), n, dtype_code(dtype), tolerance,
                                    ctypes.byref(m), ctypes.byref(bad)))
        return m.value, bad.value

    # -- per-program handle (sfb_program_*: module, fields, launches and their parameters live in the library)
    def program_create(self, image):
        buf = ctypes.create_string_buffer(image, len(image))
        h = _vp()
        _check(self.lib.sfb_program_create(buf, len(image), ctypes.byref(h)))
        return h

    def program_add_buffer(self, handle, field, nbytes, share_with=-1):
        idx = ctypes.c_int(-1)
        _check(self.lib.sfb_program_add_buffer(handle, field.encode(), int(nbytes), int(share_with), ctypes.byref(idx)))
        return idx.value

    def program_buffer(self, handle, field):
        p, n = _vp(), ctypes.c_size_t(0)
        _check(self.lib.sfb_program_buffer(handle, field.encode(), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def program_add_launch(self, handle, kernel, grid, block, smem, params, keep):
        """``params``: list of LaunchParam; ``keep``: objects the params point into (alive for the call)."""
        arr = (LaunchParam * max(1, len(params)))(*params)
        _check(self.lib.sfb_program_add_launch(handle, kernel.encode(), _u3(*grid), _u3(*block), int(smem),
                                               len(params), arr))

    def program_clear_launches(self, handle):
        _check(self.lib.sfb_program_clear_launches(handle))

    def program_num_launches(self, handle):
        n = ctypes.c_int(0)
        _check(self.lib.sfb_program_num_launches(handle, ctypes.byref(n)))
        return n.value

    def program_bind(self, handle, field, arr, is_output):
        if arr is None:
            _check(self.lib.sfb_program_bind(handle, field.encode(), None, 0, int(is_output)))
        else:
            assert arr.flags["C_CONTIGUOUS"]
            _check(self.lib.sfb_program_bind(handle, field.encode(), _vp(arr.ctypes.data), arr.nbytes, int(is_output)))

    def program_run(self, handle, repetitions=1, stream=None, timed=False):
        """All launches ``repetitions`` times in one call; ``timed``: blocks and returns the device time (ms)."""
        ms = ctypes.c_float(0)
        _check(self.lib.sfb_program_run(handle, int(repetitions), self.stream if stream is None else stream,
                                        ctypes.byref(ms) if timed else None))
        return ms.value if timed else None

    def program_call(self, handle, stream=None):
        _check(self.lib.sfb_program_call(handle, self.stream if stream is None else stream))

    def program_destroy(self, handle):
        _check(self.lib.sfb_program_destroy(handle))

    # -- multi-GPU
    def ipc_get_handle(self, dptr):
        buf = ctypes.create_string_buffer(64)
        _check(self.lib.sfb_ipc_get_handle(_vp(dptr), buf))
        return buf.raw

    def ipc_open_handle(self, handle):
        p = _vp()
        _check(self.lib.sfb_ipc_open_handle(ctypes.create_string_buffer(handle, 64), ctypes.byref(p)))
        return p.value

    def ipc_close_handle(self, dptr):
        _check(self.lib.sfb_ipc_close_handle(_vp(dptr)))

    def write_flag(self, stream, flag_dptr, value):
        _check(self.lib.sfb_stream_write_flag(stream, _vp(flag_dptr), value))

    def wait_flag(self, stream, flag_dptr, value):
        _check(self.lib.sfb_stream_wait_flag(stream, _vp(flag_dptr), value))


def compile_source(source, file_name, options):
    """NVRTC: CUDA C++ -> cubin bytes.  Works without a GPU (used by the build check)."""
    lib = load_library()
    opts = (ctypes.c_char_p * len(options))(*[o.encode() for o in options])
    image, size, log = _vp(), ctypes.c_size_t(0), _vp()
    status = lib.sfb_compile(source.encode(), file_name.encode(), len(options), opts,
                             ctypes.byref(image), ctypes.byref(size), ctypes.byref(log))
    log_text = ctypes.string_at(log.value).decode("utf-8", "replace") if log.value else ""
    if log.value:
        lib.sfb_free_host(log)
    if status != SFB_OK:
        raise SfbError(status, (lib.sfb_last_error() or b"").decode("utf-8", "replace") + "\n" + log_text)
    data = ctypes.string_at(image.value, size.value)
    lib.sfb_free_host(image)
    return data, log_text


class ParamPack:
    """Keeps kernel arguments alive and exposes the ``void**`` cuLaunchKernel wants."""

    def __init__(self, values):
        self.values = values            # ctypes objects (by value) or (buffer, address) for 128-byte maps
        ptrs = []
        for v in values:
            if isinstance(v, tuple):    # (owner buffer, aligned address)
                ptrs.append(v[1])
            else:
                ptrs.append(ctypes.addressof(v))
        self.array = (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)


def pack_params(values):
    return ParamPack(values)
