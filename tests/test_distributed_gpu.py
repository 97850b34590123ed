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


def _run_group(cmd, env, timeout):
    """Runs ``cmd`` in its own process group and kills the whole group on a timeout, so that a hung
    rank cannot keep a GPU busy behind the test's back."""
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                            start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout)
        return out, proc.returncode
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, _ = proc.communicate()
        return (out or "") + "\n[timeout after {} s: process group killed]".format(timeout), -9


def _results(out):
    """Every ``RESULT {json}`` record in the merged output of the ranks (two ranks finishing together may
    share a line)."""
    dec, found, at = json.JSONDecoder(), [], 0
    while True:
        at = out.find("RESULT {", at)
        if at < 0:
            return found
        obj, end = dec.raw_decode(out, at + len("RESULT "))
        found.append(obj)
        at = end


@pytest.mark.parametrize("name,fuse", [
    ("ref_jacobi3d_32x32x32_8itr_8vec", True),
    ("ref_jacobi3d_32x32x32_8itr_8vec", False),
    ("jacobi2d_96x128_6itr_shrink_f64", True),
    ("hdiff_24x28x16", True),
    ("fork_join_20x16x24", True),
    ("lowdim3d_20x24x48_3st_f32", True),
    ("ref_varying_dimensionality", True),
    ("chain3d:160x64x128", True),
    ("upwind3d_fwd_24x16x32_4st", True),           # one-sided reach: rank 0 only receives, rank 1 only sends
    ("upwind3d_fwd_24x16x32_4st", False),
    ("upwind3d_bwd_20x12x32_5st_f64", True),
    ("ref_jacobi3d_32x32x32_8itr_8vec:copy", True),    # SFB200_PEER_PUSH=0 (the default): copy pushes for streamed passes too
    ("upwind3d_fwd_24x16x32_4st:copy", True),
    ("chain3d:160x64x128:d2", True),                    # 4 passes, ping-pong storage
    ("ref_jacobi3d_32x32x32_8itr_8vec:kernel", True),  # SFB200_PEER_PUSH=1: the streamed kernels push their edge planes themselves
    ("upwind3d_fwd_24x16x32_4st:kernel", True),
    ("chain3d:160x64x128:d2:kernel", True),
    ("jacobi3d_24x20x40_4itr_shrink_f64:kernel", True),
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
    if name.endswith(":copy"):
        name = name[:-len(":copy")]
        env["SFB200_PEER_PUSH"] = "0"
    if name.endswith(":kernel"):
        name = name[:-len(":kernel")]
        env["SFB200_PEER_PUSH"] = "1"
    if name.endswith(":d2"):
        name = name[:-len(":d2")]
        env["SFB200_MAX_DEPTH"] = "2"
    if name.startswith("upwind3d") and fuse:
        env["SFB200_MAX_DEPTH"] = "2"               # two passes: the one-sided halo really is exchanged
    cmd[-2] = name
    if name.startswith("chain3d"):
        env["SFB200_PIPELINE_PIECES"] = "4"         # 88 planes per rank: pieces of 22 planes
    out, code = _run_group(cmd, env, 240)
    assert code == 0, out[-4000:]
    lines = _results(out)
    assert len(lines) == 2 and all(l["ok"] for l in lines)
    if name.startswith("chain3d") and "SFB200_MAX_DEPTH" not in env:
        # wide halo (accumulated reach 8) -> the host-array call ran as the overlapped exchange-free schedule
        assert all(l["halo"] == 8 and l["report"]["call_pipelined"] for l in lines), lines
    if name.startswith("upwind3d"):
        assert sorted(l["sends"] for l in lines) != [0, 0]          # the halo really was exchanged
        assert min(l["sends"] for l in lines) == 0                  # ... by one side only



@pytest.mark.parametrize("name,halo", [("ref_jacobi3d_32x32x32_8itr_8vec", 0), ("hdiff_24x28x16", 2)])
def test_run_distributed_program_cli(native_lib, name, halo, tmp_path):
    """``bin/run_distributed_program.py prog.json cuda -gpus 2 -compare-to-reference``: the command line starts
    its own ranks (TCP rendezvous, no torchrun), the last rank verifies against the CPU program and the
    exit code says so (reference bin/run_distributed_program.py:283-341)."""
    if _gpu_count(native_lib) < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import program_path
    cmd = [sys.executable, os.path.join(ROOT, "bin", "run_distributed_program.py"), program_path(name), "cuda",
           "-gpus", "2", "-compare-to-reference", "-halo", str(halo), "-repetitions", "2",
           "-input-directory", os.path.dirname(program_path(name))]
    env = dict(os.environ, SFB200_MAX_DEPTH="4")
    env.pop("RANK", None)
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                            start_new_session=True, cwd=str(tmp_path))
    try:
        out, _ = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(proc.pid, signal.SIGKILL)
        out, _ = proc.communicate()
        raise AssertionError("timeout\n" + (out or "")[-3000:])
    assert proc.returncode == 0, out[-4000:]
    assert "Results verified." in out and "halo pushes per execution" in out
    assert os.path.isdir(tmp_path / "results" / name)
