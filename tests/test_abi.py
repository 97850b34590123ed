"""The C-ABI boundary without a GPU: the library builds, loads, exports every symbol the header
declares, compiles generated programs with NVRTC, and refuses loudly to compute without a device."""
import os
import re

import pytest

from conftest import ROOT, all_programs, program_path


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "sfb200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"SFB_API\s+[\w\s\*]+?\b(sfb_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    assert len(syms) >= 50
    for required in ("sfb_init", "sfb_compile", "sfb_launch", "sfb_tensor_map_tiled",
                     "sfb_ipc_open_handle", "sfb_stream_wait_flag", "sfb_last_error"):
        assert required in syms


def test_library_exports_every_declared_symbol(native_lib):
    from stencilflow_b200 import runtime
    for name in _declared_symbols():
        assert hasattr(native_lib, name), name
        assert name in runtime.PROTOTYPES, "ctypes binding lacks " + name
    assert set(runtime.PROTOTYPES) == set(_declared_symbols())
    assert native_lib.sfb_abi_version() == 1


def test_library_does_not_link_torch_or_driver(native_lib):
    import subprocess
    from stencilflow_b200 import runtime
    out = subprocess.run(["ldd", runtime.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcuda.so" not in out


@pytest.mark.parametrize("name", all_programs())
def test_every_program_compiles_to_sm100a(native_lib, name, tmp_path, monkeypatch):
    monkeypatch.setenv("SFB200_CACHE", str(tmp_path))
    from stencilflow_b200.cuda_program import CudaProgram
    prog = CudaProgram(program_path(name), allocate=False)
    assert prog.image[:4] == b"\x7fELF"
    assert os.path.isfile(os.path.join(prog.cache_dir, "kernel.cu"))
    assert os.path.isfile(os.path.join(prog.cache_dir, "plan.json"))
    again = CudaProgram(program_path(name), allocate=False)
    assert again.was_cached


def test_compile_error_is_reported(native_lib):
    from stencilflow_b200 import runtime
    with pytest.raises(runtime.SfbError) as err:
        runtime.compile_source("__global__ void k() { this is not c++; }", "bad.cu", ["-arch=sm_100a"])
    assert err.value.status == -3


def test_no_cpu_fallback_without_device(native_lib):
    """On a machine without a GPU the product path must raise, not compute on the CPU."""
    import ctypes
    from stencilflow_b200 import runtime
    n = ctypes.c_int(-1)
    status = native_lib.sfb_device_count(ctypes.byref(n))
    if status == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(runtime.SfbError):
        runtime.Runtime.get()
    from stencilflow_b200.run_program import run_program
    with pytest.raises(runtime.SfbError):
        run_program(program_path("ref_simulator"), "cuda", compare_to_reference=True, log_level=0)


def test_fpga_modes_are_rejected():
    from stencilflow_b200.run_program import run_program
    with pytest.raises(ValueError):
        run_program(program_path("ref_simulator"), "emulation")
    with pytest.raises(ValueError, match="Unrecognized execution mode"):
        run_program(program_path("ref_simulator"), "bogus")
