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
    ("upwind3d_fwd_24x16x32_4st", False),          # one-sided reach: rank 0 only receives
    ("upwind3d_bwd_20x12x32_5st_f64", True),
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


@pytest.mark.parametrize("world,pieces", [(1, 16), (2, 8), (4, 4), (8, 4)])
def test_exchange_free_call_schedule(native_lib, world, pieces):
    """The overlapped host-array call on a slab (CudaProgram._pipeline_ranges): every launch produces
    exactly the planes later launches and the owned output need, in order, never ahead of its inputs,
    and only owned output planes travel back -- with no halo exchange, given a halo of the accumulated
    reach (distributed.total_reach)."""
    from stencilflow_b200 import programs
    from stencilflow_b200.cuda_program import CudaProgram
    name, prog, _ = programs.baseline_config(1)
    path = programs.write_program(prog, name)
    probe = CudaProgram(path, allocate=False)
    n = 1024
    wide = distributed.total_reach(probe.lowered)
    assert wide == 8 and distributed.halo_depth(probe.lowered) == 4       # two passes of four operators
    for rank in range(world):
        slab = distributed.Slab(rank, world, n, wide) if world > 1 else None
        p = CudaProgram(path, allocate=False, slab=slab)
        steps = p._pipeline_ranges(pieces)
        assert steps is not None
        own = (slab.begin, slab.end) if slab else (0, n)
        produced = {}                      # launch -> [begin, end) so far
        frontier = {}                      # field -> planes final so far (exclusive)
        lowest = {}
        sent_back = []
        for step in steps:
            for (f, b, e) in step["h2d"]:
                assert b == frontier.get(f, b) and e >= b
                lowest.setdefault(f, b)
                frontier[f] = e
            for (idx, b, e) in step["launch"]:
                l = p.lowered.launches[idx]
                if idx in produced:
                    assert b == produced[idx][1]
                produced[idx] = (produced.get(idx, (b, b))[0], e)
                for f, (back, fwd) in distributed.launch_reach(p.lowered, idx).items():
                    assert min(e - 1 + fwd, n - 1) < frontier[f], (idx, f, e, frontier[f])
                    assert max(b - back, 0) >= lowest[f]
                for f in l.writes:
                    lowest.setdefault(f, b)
                    frontier[f] = e
            sent_back += step["d2h"]
        assert produced[len(p.lowered.launches) - 1] == own
        lo0, hi0 = produced[0]
        assert lo0 == max(0, own[0] - 4) and hi0 == min(n, own[1] + 4)     # widened by the second pass's reach
        assert [b for (_, b, _) in sent_back][0] == own[0] and sent_back[-1][2] == own[1]
        assert all(a[2] == b[1] for a, b in zip(sent_back, sent_back[1:]))
    # a halo of one pass only is not enough: the call falls back to execute() with halo pushes
    if world > 1:
        thin = CudaProgram(path, allocate=False, slab=distributed.Slab(0, world, n, 4))
        assert thin._pipeline_ranges(pieces) is None
