"""Command-line surfaces that need no GPU: bin/report.py (the reference's static model, bin/report.py:11-57,
plus the pass plan), bin/synthesize.py, the argument contracts of the drivers."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, program_path


def _run(args, **kw):
    return subprocess.run([sys.executable] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=300, **kw)


def test_report_cli_prints_model_and_plan(native_lib):
    res = _run([os.path.join(ROOT, "bin", "report.py"), program_path("ref_jacobi3d_32x32x32_8itr_8vec"), "300"])
    assert res.returncode == 0, res.stdout
    out = res.stdout
    # the reference's three sections ...
    assert "Compute performance" in out and "Total:" in out and "300" in out
    # ... with the reference's numbers for this program (262144 B minimum volume)
    assert "262144" in out or "0.262144" in out or "256.0" in out
    # ... and the plan that replaces the FPGA buffer placement
    # (a 32^3 grid is launch-bound: eight one-operator launches; the benchmark program gets fused passes)
    assert "pass 7: general" in out and "Off-chip volume of the plan" in out
    res = _run([os.path.join(ROOT, "bin", "report.py"), os.path.join(ROOT, "programs", "jacobi3d_1024_8itr_f32.json"), "300"])
    assert res.returncode == 0, res.stdout
    assert "pass 1: streamed [4 operator(s): b4, b5, b6, b7]" in res.stdout and "tile 72x64" in res.stdout


def test_report_cli_without_plan(native_lib):
    res = _run([os.path.join(ROOT, "bin", "report.py"), program_path("ref_jacobi2d_128x128"), "250", "-no-plan"])
    assert res.returncode == 0, res.stdout
    assert "Compute performance" in res.stdout and "streamed" not in res.stdout.lower()


def test_run_program_cli_rejects_unknown_mode():
    res = _run([os.path.join(ROOT, "bin", "run_program.py"), program_path("ref_simulator"), "verilog"])
    assert res.returncode != 0 and "invalid choice" in res.stdout


def test_run_program_cli_keeps_the_reference_flags():
    res = _run([os.path.join(ROOT, "bin", "run_program.py"), "-h"])
    assert res.returncode == 0
    for flag in ("-run-simulation", "-compare-to-reference", "-input-directory", "-skip-execution", "-plot",
                 "-log-level", "-print-result", "-halo", "-repetitions"):
        assert flag in res.stdout, flag
    assert "cuda" in res.stdout and "emulation" in res.stdout and "hardware" in res.stdout


def test_run_distributed_program_cli_contract():
    res = _run([os.path.join(ROOT, "bin", "run_distributed_program.py"), "-h"])
    assert res.returncode == 0
    for flag in ("-gpus", "-compare-to-reference", "-halo", "-repetitions", "-input-directory"):
        assert flag in res.stdout, flag


def test_cuda_mode_fails_loudly_without_a_device(native_lib, tmp_path):
    """No CPU fallback: on a box without a GPU `run_program ... cuda` must fail, not compute elsewhere."""
    import ctypes
    n = ctypes.c_int(0)
    native_lib.sfb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    res = _run([os.path.join(ROOT, "bin", "run_program.py"), program_path("ref_simulator"), "cuda",
                "-compare-to-reference"], cwd=str(tmp_path))
    assert res.returncode != 0
    assert "NO_DEVICE" in res.stdout or "no usable GPU" in res.stdout or "CUDA" in res.stdout


def test_plain_c_host_example_builds(native_lib, tmp_path):
    """examples/run_sfbplan.c needs nothing but include/sfb200.h and libsfb200.so."""
    from stencilflow_b200 import runtime
    libdir = os.path.dirname(runtime.LIB_PATH)
    exe = str(tmp_path / "run_sfbplan")
    res = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "run_sfbplan.c"), "-o", exe, "-L", libdir, "-lsfb200",
                          "-Wl,-rpath," + libdir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 2 and "usage" in res.stdout
    out = subprocess.run(["ldd", exe], stdout=subprocess.PIPE, text=True).stdout
    assert "libsfb200" in out and "python" not in out.lower() and "torch" not in out.lower()
