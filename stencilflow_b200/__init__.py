"""stencilflow_b200 -- B200-native execution backend for StencilFlow stencil programs.

The package keeps StencilFlow's JSON program format and front-end API
(``KernelChainGraph``, ``run_program``, the ``helper`` functions) and executes the
operator chain as generated sm_100a CUDA kernels behind a ctypes C-ABI
(``libsfb200.so``, see ``include/sfb200.h``).
"""

from .helper import *  # noqa: F401,F403
from .helper import (ITERATORS, aligned, arrays_are_equal, convert_3d_to_1d, dim_to_abs_val,
                     list_add_cwise, list_subtract_cwise, load_array, load_input_arrays,
                     max_dict_entry_key, parse_json, save_array, save_output_arrays,
                     str_to_dtype, unique, OpCounter)
from .log_level import LogLevel
from .bounded_queue import BoundedQueue
from .kernel import Kernel
from .kernel_chain_graph import KernelChainGraph

__version__ = "0.1.0"


def __getattr__(name):
    # heavier pieces (they pull in the native library) are imported on first use
    if name == "run_program":
        from .run_program import run_program
        return run_program
    if name == "CudaProgram":
        from .cuda_program import CudaProgram
        return CudaProgram
    raise AttributeError(name)
