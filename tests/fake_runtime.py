"""A recording stand-in for ``stencilflow_b200.runtime.Runtime`` plus a simulator of what it recorded.

``SlabProgram.execute`` only *enqueues* work: launches, peer copies, counter writes / waits, events.
With this runtime every rank (a thread) records its streams instead of driving a GPU; ``simulate``
then plays all streams of all ranks under a random interleaving that respects stream order, counter
waits and events, and checks on the way that every launch finds exactly the data it expects --
its own fields and the halo planes the neighbours pushed -- which fails on a missing wait (stale
halo), a premature overwrite (write after read) or a deadlock.  No GPU, no CUDA call.
"""
import ctypes
import random
import threading

import numpy as np


class FakeProps:
    sm_count = 148
    name = b"fake B200"


class FakeEvent:
    def __init__(self):
        self.count = 0          # records enqueued
        self.done = 0           # records executed by the simulator


class World:
    """State shared by the fake runtimes of all ranks: one address space, one set of counters."""

    def __init__(self):
        self.lock = threading.Lock()
        self.next_addr = 1 << 20
        self.allocs = {}        # base address -> (rank, nbytes)

    def malloc(self, rank, nbytes):
        with self.lock:
            addr = self.next_addr
            self.next_addr += (int(nbytes) + 4095) // 4096 * 4096 + 4096
            self.allocs[addr] = (rank, int(nbytes))
            return addr

    def owner(self, addr):
        for base, (rank, n) in self.allocs.items():
            if base <= addr < base + n:
                return rank, base
        raise KeyError(hex(addr))


class FakeRuntime:
    def __init__(self, world, rank):
        self.world, self.rank = world, rank
        self.props = FakeProps()
        self.device = rank
        self.streams = {0: []}
        self.stream = 0
        self._next_stream = 1

    # -- what CudaProgram.load / _pack_launch need
    def module_load(self, image):
        return object()

    def module_unload(self, module):
        pass

    def get_function(self, module, name):
        return name

    def set_max_dynamic_smem(self, fn, n):
        pass

    def occupancy(self, fn, threads, smem):
        return 1

    def tensor_map(self, dptr, dtype, dims, strides, box, l2_promotion=128):
        raw = ctypes.create_string_buffer(192)
        return raw, (ctypes.addressof(raw) + 63) & ~63

    def malloc(self, nbytes):
        return self.world.malloc(self.rank, nbytes)

    def free(self, dptr):
        pass

    def memset(self, dptr, value, nbytes, stream=None):
        pass

    def h2d(self, dptr, arr, stream=None, nbytes=None):
        pass

    def d2h(self, arr, dptr, stream=None, nbytes=None):
        pass

    def stream_synchronize(self, stream=None):
        pass

    def ipc_get_handle(self, dptr):
        return int(dptr).to_bytes(8, "little") + bytes(56)

    def ipc_open_handle(self, handle):
        return int.from_bytes(bytes(handle)[:8], "little")

    def ipc_close_handle(self, dptr):
        pass

    # -- what gets recorded
    def _s(self, stream):
        return self.stream if stream is None else stream

    def stream_create(self):
        sid = self._next_stream
        self._next_stream += 1
        self.streams[sid] = []
        return sid

    def event_create(self, timing=True):
        return FakeEvent()

    def event_destroy(self, ev):
        pass

    def event_record(self, ev, stream=None):
        ev.count += 1
        self.streams[self._s(stream)].append(("record", ev, ev.count))

    def stream_wait_event(self, stream, ev):
        self.streams[self._s(stream)].append(("wait_event", ev, ev.count))

    def launch(self, fn, grid, block, smem, params, stream=None):
        self.streams[self._s(stream)].append(("launch", fn))

    def d2d(self, dst, src, nbytes, stream=None):
        self.streams[self._s(stream)].append(("d2d", int(dst), int(src), int(nbytes)))

    def write_flag(self, stream, addr, value):
        self.streams[self._s(stream)].append(("write_flag", int(addr), int(value)))

    def wait_flag(self, stream, addr, value):
        self.streams[self._s(stream)].append(("wait_flag", int(addr), int(value)))


class ThreadComm:
    """Rendezvous between the rank threads of one test process."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.shared = rank, world, shared

    def allgather(self, obj):
        self.shared["slots"][self.rank] = obj
        self.shared["barrier"].wait()
        out = list(self.shared["slots"])
        self.shared["barrier"].wait()
        return out

    def barrier(self):
        self.shared["barrier"].wait()

    def max_float(self, x):
        return max(self.allgather(float(x)))

    def close(self):
        pass


def build_slab_programs(path, world, plan_options_fn=None, peer_push=True, reps=2):
    """One ``SlabProgram`` per rank on fake runtimes, each executed ``reps`` times (recorded only)."""
    import os
    from stencilflow_b200 import distributed, runtime
    shared = {"slots": [None] * world, "barrier": threading.Barrier(world)}
    fake_world = World()
    fakes = [FakeRuntime(fake_world, r) for r in range(world)]
    by_thread = {}
    original = runtime.Runtime.get
    runtime.Runtime.get = classmethod(lambda cls, device=None: by_thread[threading.get_ident()])
    old_env = os.environ.get("SFB200_PEER_PUSH")
    os.environ["SFB200_PEER_PUSH"] = "1" if peer_push else "0"
    progs, errors = [None] * world, []

    def run(rank):
        try:
            by_thread[threading.get_ident()] = fakes[rank]
            comm = ThreadComm(rank, world, shared)
            opts = plan_options_fn() if plan_options_fn else None
            p = distributed.SlabProgram(path, comm, device=rank, plan_options=opts)
            for _ in range(reps):
                p.execute()
            progs[rank] = p
        except Exception as exc:        # noqa: BLE001 - reported by the test
            errors.append((rank, repr(exc)))
            shared["barrier"].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    try:
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        runtime.Runtime.get = original
        if old_env is None:
            os.environ.pop("SFB200_PEER_PUSH", None)
        else:
            os.environ["SFB200_PEER_PUSH"] = old_env
    if errors:
        raise RuntimeError(errors)
    return progs, fakes, fake_world


def simulate(progs, fakes, fake_world, reps, seed=0, max_steps=10 ** 6):
    """Plays the recorded streams of all ranks in a random order that respects their dependencies.
    Returns the number of operations executed; raises AssertionError on a stale or prematurely
    overwritten halo, a wrong counter, or a deadlock."""
    from stencilflow_b200 import distributed
    rng = random.Random(seed)
    world = len(progs)
    lowered = progs[0].lowered
    n = len(lowered.launches)
    fields = progs[0].program.fields
    reach = [distributed.launch_reach(lowered, i) for i in range(n)]
    writer = {}
    for i, l in enumerate(lowered.launches):
        for f in l.writes:
            writer[f] = i
    xplan = progs[0].xplan
    flags = {}                          # address -> value
    ver = {}                            # (rank, buffer base, region) -> (rep, launch) of the content
    heads = {(r, s): 0 for r in range(world) for s in fakes[r].streams}
    for f in fakes:                     # events keep their state between runs of the simulator
        for ops in f.streams.values():
            for op in ops:
                if op[0] in ("record", "wait_event"):
                    op[1].done = 0
    launches_done = [0] * world
    executed = 0

    def storage(rank, field):
        return progs[rank].buffers[field].dptr

    def ready(rank, op):
        if op[0] == "wait_flag":
            return flags.get(op[1], 0) >= op[2]
        if op[0] == "wait_event":
            return op[1].done >= op[2]
        return True

    def run(rank, op):
        if op[0] == "write_flag":
            assert op[2] >= flags.get(op[1], 0), "counter moved backwards"
            flags[op[1]] = op[2]
        elif op[0] == "record":
            op[1].done = op[2]
        elif op[0] == "launch":
            g = launches_done[rank]
            launches_done[rank] += 1
            rep, idx = divmod(g, n)
            l = lowered.launches[idx]
            for f in l.reads:
                if fields[f].is_scalar or f not in writer:
                    continue                     # program inputs are loaded with their halos
                want = (rep, writer[f])
                got = ver.get((rank, storage(rank, f), "own"))
                assert got == want, "rank {} launch {} rep {} reads own {}: {} != {}".format(rank, idx, rep, f, got, want)
                back, fwd = reach[idx].get(f, (0, 0))
                if back and rank > 0:
                    got = ver.get((rank, storage(rank, f), "lo"))
                    assert got == want, "rank {} launch {} rep {}: lower halo of {} is {} not {}".format(
                        rank, idx, rep, f, got, want)
                if fwd and rank + 1 < world:
                    got = ver.get((rank, storage(rank, f), "hi"))
                    assert got == want, "rank {} launch {} rep {}: upper halo of {} is {} not {}".format(
                        rank, idx, rep, f, got, want)
            for f in l.writes:
                ver[(rank, storage(rank, f), "own")] = (rep, idx)
            if l.info.get("peer_push"):
                for (f, _) in xplan.up[idx]:
                    if rank + 1 < world:
                        ver[(rank + 1, storage(rank + 1, f), "lo")] = (rep, idx)
                for (f, _) in xplan.down[idx]:
                    if rank > 0:
                        ver[(rank - 1, storage(rank - 1, f), "hi")] = (rep, idx)
        elif op[0] == "d2d":
            dst_rank, dst_base = fake_world.owner(op[1])
            src_rank, src_base = fake_world.owner(op[2])
            assert src_rank == rank and abs(dst_rank - rank) == 1
            region = "lo" if dst_rank > rank else "hi"
            ver[(dst_rank, dst_base, region)] = ver.get((rank, src_base, "own"))

    while executed < max_steps:
        pending = [(r, s) for (r, s), h in heads.items() if h < len(fakes[r].streams[s])]
        if not pending:
            break
        can = [(r, s) for (r, s) in pending if ready(r, fakes[r].streams[s][heads[(r, s)]])]
        assert can, "deadlock: {}".format([(r, s, fakes[r].streams[s][heads[(r, s)]][:1] + tuple(
            x for x in fakes[r].streams[s][heads[(r, s)]][1:] if isinstance(x, int))) for (r, s) in pending])
        r, s = rng.choice(can)
        run(r, fakes[r].streams[s][heads[(r, s)]])
        heads[(r, s)] += 1
        executed += 1
    assert all(d == reps * n for d in launches_done), launches_done
    return executed
