"""Multi-GPU host logic without GPUs: every rank's ``SlabProgram.execute`` is recorded on a fake runtime
(tests/fake_runtime.py) and all ranks' streams are then played under random interleavings.  A launch
must find exactly the version of every field -- own planes and pushed halos -- that the program order
prescribes; anything else is a missing wait, a premature overwrite or a deadlock."""
import pytest

from conftest import program_path
from fake_runtime import build_slab_programs, simulate
from stencilflow_b200 import distributed
from stencilflow_b200.planner import PlanOptions

CASES = [
    # (program, plan options, world)
    ("ref_jacobi3d_32x32x32_8itr_8vec", dict(max_depth=4), 2),
    ("ref_jacobi3d_32x32x32_8itr_8vec", dict(max_depth=2), 4),      # 4 passes share 2 ping-pong buffers
    ("ref_jacobi3d_32x32x32_8itr_8vec", dict(fuse=False), 3),       # one-operator kernels: copy pushes
    ("upwind3d_fwd_24x16x32_4st", dict(max_depth=2), 2),            # one-sided reach: only "down" pushes
    ("upwind3d_fwd_24x16x32_4st", dict(fuse=False), 3),
    ("upwind3d_bwd_20x12x32_5st_f64", dict(max_depth=2), 4),        # only "up" pushes, odd number of exchanges
    ("hdiff_24x28x16", dict(fuse=False), 2),
    ("fork_join_20x16x24", dict(fuse=False), 2),
    ("jacobi2d_96x128_6itr_shrink_f64", dict(max_depth=2), 4),
]


@pytest.mark.parametrize("peer_push", [True, False], ids=["kernel_push", "copy_push"])
@pytest.mark.parametrize("name,opts,world", CASES, ids=lambda v: v if isinstance(v, str) else None)
def test_recorded_exchange_is_ordered_and_deadlock_free(native_lib, name, opts, world, peer_push):
    reps = 3
    progs, fakes, fw = build_slab_programs(program_path(name), world, lambda: PlanOptions(**opts),
                                           peer_push=peer_push, reps=reps)
    kinds = {op[0] for f in fakes for s in f.streams.values() for op in s}
    if peer_push and opts.get("fuse", True):
        assert any(l.info.get("peer_push") for l in progs[0].lowered.launches)
        assert "d2d" not in kinds                  # streamed passes push from inside the kernel
    elif progs[0].sends or progs[-1].sends:
        assert "d2d" in kinds
    for seed in range(12):
        simulate(progs, fakes, fw, reps, seed=seed)


def test_exchange_plan_is_rank_independent_and_handles_one_sided_reach(native_lib):
    from stencilflow_b200.cuda_program import CudaProgram
    p = CudaProgram(program_path("upwind3d_fwd_24x16x32_4st"), allocate=False, plan_options=PlanOptions(max_depth=2))
    x = distributed.ExchangePlan(p.lowered, p.plan.buffer_assignment())
    # b[i+1] taps: the *lower* neighbour needs my bottom planes; nothing ever goes up
    assert x.up_events == [] and x.down_events == [0]
    assert x.down[0] == [("b1", 2)]
    # rank 0 of 2 has nothing to send but still receives from above and waits for it
    s0, s1 = distributed.Slab(0, 2, 24, 2), distributed.Slab(1, 2, 24, 2)
    assert x.for_rank(s0) == [] and len(x.for_rank(s1)) == 1
    assert x.wait_values(0, 1) == (0, 1) and x.wait_values(3, 1) == (0, 4)
    assert x.wait_values(0, 0) == (0, 0)
    # the pass that writes b1 again in the next repetition must wait for the reader of the previous one
    assert x.war[0] == (-1, 1) and x.war_value(0, 0) == 0 and x.war_value(1, 0) == 2
    assert x.progress_points == [1]


def test_simulator_catches_a_missing_wait(native_lib):
    """The checker itself: drop the halo waits from one rank's recording and it must fail."""
    progs, fakes, fw = build_slab_programs(program_path("ref_jacobi3d_32x32x32_8itr_8vec"), 2,
                                           lambda: PlanOptions(max_depth=4), reps=2)
    flags = progs[1].flags
    fakes[1].streams[0] = [op for op in fakes[1].streams[0]
                           if not (op[0] == "wait_flag" and op[1] in (flags + 0, flags + 4))]
    with pytest.raises(AssertionError):
        for seed in range(40):
            simulate(progs, fakes, fw, 2, seed=seed)


@pytest.mark.parametrize("peer_push", [True, False], ids=["kernel_push", "copy_push"])
def test_simulator_catches_a_missing_write_after_read_guard(native_lib, peer_push):
    """Without the progress waits a fast rank overwrites halo planes its neighbour has not read yet
    (the race the round-1 protocol had across repetitions)."""
    opts = (lambda: PlanOptions(max_depth=4)) if peer_push else (lambda: PlanOptions(fuse=False))
    progs, fakes, fw = build_slab_programs(program_path("ref_jacobi3d_32x32x32_8itr_8vec"), 2, opts,
                                           peer_push=peer_push, reps=4)
    for seed in range(10):
        simulate(progs, fakes, fw, 4, seed=seed)          # intact: fine
    removed = 0
    for r, f in enumerate(fakes):
        guard = (progs[r].flags + 8, progs[r].flags + 12)
        for s in f.streams:
            kept = [op for op in f.streams[s] if not (op[0] == "wait_flag" and op[1] in guard)]
            removed += len(f.streams[s]) - len(kept)
            f.streams[s] = kept
    assert removed > 0
    with pytest.raises(AssertionError):
        for seed in range(200):
            simulate(progs, fakes, fw, 4, seed=seed)
