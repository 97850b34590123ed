"""Builds the native pieces in-tree: ``libsfb200.so`` (nvcc, sm_100a) next to this file."""

import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(_PKG, "csrc", "sfb200_runtime.cu")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "sfb200.h")
LIBRARY = os.path.join(_PKG, "libsfb200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static"]


def build_native(force=False, verbose=False):
    """Compile the C-ABI runtime if it is missing or older than its sources."""
    if not force and os.path.isfile(LIBRARY):
        newest = max(os.path.getmtime(SOURCE), os.path.getmtime(HEADER))
        if os.path.getmtime(LIBRARY) >= newest:
            return LIBRARY
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = LIBRARY + ".tmp{}".format(os.getpid())
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, SOURCE, "-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    os.replace(tmp, LIBRARY)
    return LIBRARY
