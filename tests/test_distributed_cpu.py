"""N > 1 host logic on CPU: two gloo ranks run the slab decomposition + halo schedule of real plans,
with the numpy oracle standing in for the kernels, and must reproduce the single-domain oracle
bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT, program_path
from stencilflow_b200 import distributed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_world(path, fuse, world=2, seed=5):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen(
            [sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), path, "1" if fuse else "0", str(seed)],
            env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    results = []
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out[-3000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
        results.append(json.loads(line[len("RESULT "):]))
    return results


@pytest.mark.parametrize("name,fuse", [
    ("ref_jacobi3d_32x32x32_8itr_8vec", False),
    ("ref_jacobi3d_32x32x32_8itr_8vec", True),
    ("jacobi2d_96x128_6itr_shrink_f64", True),
    ("hdiff_24x28x16", True),
    ("fork_join_20x16x24", False),
    ("box3d_10x12x16", False),
    ("lowdim3d_20x24x48_3st_f32", True),
    ("ref_varying_dimensionality", True),
])
def test_two_ranks_reproduce_single_domain(native_lib, name, fuse):
    results = _run_world(program_path(name), fuse)
    assert all(r["ok"] for r in results)
    assert results[0]["max"] == 1.5            # max-reduction over ranks (used for timings)
    assert results[0]["slab"][1] == results[1]["slab"][0]
    if name.startswith("ref_jacobi3d") and not fuse:
        assert results[0]["halo"] == 1 and results[0]["sends"] == 7    # every pass but the last


def test_slab_partition():
    slabs = [distributed.Slab(r, 4, 1030, 4) for r in range(4)]
    assert slabs[0].begin == 0 and slabs[-1].end == 1030
    assert all(a.end == b.begin for a, b in zip(slabs, slabs[1:]))
    assert slabs[0].alloc_begin == 0 and slabs[0].alloc_end == slabs[0].end + 4
    assert slabs[2].alloc_begin == slabs[2].begin - 4
    assert slabs[3].alloc_end == 1030
    with pytest.raises(ValueError):
        distributed.Slab(0, 8, 16, 4)
