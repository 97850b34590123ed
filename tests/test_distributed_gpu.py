"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): slabs + NVLink halo pushes must
match the oracle and reproduce the single-GPU result bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpu_count(native_lib):
    import ctypes
    n = ctypes.c_int(0)
    native_lib.sfb_device_count(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("name,fuse", [
    ("ref_jacobi3d_32x32x32_8itr_8vec", True),
    ("ref_jacobi3d_32x32x32_8itr_8vec", False),
    ("jacobi2d_96x128_6itr_shrink_f64", True),
    ("hdiff_24x28x16", True),
    ("fork_join_20x16x24", True),
    ("lowdim3d_20x24x48_3st_f32", True),
    ("ref_varying_dimensionality", True),
    ("chain3d:160x64x128", True),
])
def test_two_gpus_match_oracle_and_single_gpu(native_lib, name, fuse):
    if _gpu_count(native_lib) < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), name, "1" if fuse else "0"]
    env = dict(os.environ)
    if name.startswith("chain3d"):
        env["SFB200_PIPELINE_PIECES"] = "4"         # 88 planes per rank: pieces of 22 planes
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-4000:]
    lines = [json.loads(l[len("RESULT "):]) for l in res.stdout.splitlines() if l.startswith("RESULT ")]
    assert len(lines) == 2 and all(l["ok"] for l in lines)
    if name.startswith("chain3d"):
        # wide halo (accumulated reach 8) -> the host-array call ran as the overlapped exchange-free schedule
        assert all(l["halo"] == 8 and l["report"]["call_pipelined"] for l in lines), lines
