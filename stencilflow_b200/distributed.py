"""Multi-GPU execution: slab decomposition of the outermost dimension with per-pass halo exchange.

The reference scales across devices by cutting the *operator pipeline* in two and streaming whole
fields through SMI channels (``split_sdfg``, ``stencilflow/sdfg_generator.py:680-1000``;
``bin/run_distributed_program.py``).  On a box of NVLink-connected GPUs the natural decomposition is
spatial instead: rank g owns planes ``[g*N/G, (g+1)*N/G)`` of the outermost dimension (i for 3-D, j for
2-D programs); row-major layout makes every halo one contiguous block.  After a pass has written a
field that a later pass reads with an offset along the slab axis, each rank *pushes* its edge planes
straight into the halo region of its neighbours' buffers (peer ``cudaMemcpyAsync`` over NVLink on
IPC-mapped memory) and raises a flag in the neighbour's memory (``cuStreamWriteValue32``); the
neighbour's stream waits on that flag (``cuStreamWaitValue32``) before launching the consumer.  No host
synchronisation and no collective is involved; results are bit-identical to the single-GPU run
because every cell is computed by the same instruction sequence.

One process per GPU.  Rendezvous (exchange of IPC handles, barriers, max-reduction of timings) goes
through a tiny ``Comm`` interface; ``TorchComm`` implements it with ``torch.distributed`` (gloo), which
is how ``torchrun`` launches ``bench.py``.  Torch never touches device memory or the data path.
"""

import os
import pickle
from typing import Dict, List, Optional

import numpy as np


class Slab:
    """Ownership of the slab axis for one rank: owned planes [begin, end), allocated planes
    [alloc_begin, alloc_end) = owned + halo, clipped to the domain."""

    def __init__(self, rank, world, n, halo):
        if n < world:
            raise ValueError("cannot split {} planes over {} ranks".format(n, world))
        self.rank, self.world, self.n, self.halo = rank, world, n, halo
        self.begin = (n * rank) // world
        self.end = (n * (rank + 1)) // world
        self.alloc_begin = max(0, self.begin - halo)
        self.alloc_end = min(n, self.end + halo)
        if world > 1 and (self.end - self.begin) < halo:
            raise ValueError("slab of {} planes is thinner than the halo {}".format(self.end - self.begin, halo))

    @staticmethod
    def of_rank(rank, world, n, halo):
        return Slab(rank, world, n, halo)

    def __repr__(self):
        return "Slab(rank {}/{}: own [{}, {}), alloc [{}, {}))".format(
            self.rank, self.world, self.begin, self.end, self.alloc_begin, self.alloc_end)


class HaloSend:
    """After launch ``launch``: planes [src_begin, src_end) of ``field`` go to rank ``peer``."""

    def __init__(self, launch, field, peer, src_begin, src_end):
        self.launch, self.field, self.peer = launch, field, peer
        self.src_begin, self.src_end = src_begin, src_end

    def __repr__(self):
        return "HaloSend(after launch {}: {}[{}:{}] -> rank {})".format(
            self.launch, self.field, self.src_begin, self.src_end, self.peer)


def launch_reach(lowered, launch_index):
    """(back, fwd): how many planes below / above its output range a launch reads along the slab
    axis, per field it reads: {field: (back, fwd)}."""
    l = lowered.launches[launch_index]
    program = lowered.program
    axis = lowered.slab_axis
    reach = {}
    if axis is None:
        return reach
    it = "ijk"[axis]
    if l.family == "streamed":
        for f, (b, fw) in l.info["reach"].items():
            reach[f] = (b, fw)
        return reach
    op = next(o for o in program.ops if o.name == l.ops[0])
    pos = "ijk".index(it)
    for f in l.reads:
        if it not in program.fields[f].dims:
            continue
        lo = hi = 0
        for off in op.offsets3(f):
            if off[pos] is not None:
                lo, hi = min(lo, off[pos]), max(hi, off[pos])
        reach[f] = (-lo, hi)
    return reach


def halo_depth(lowered):
    """Planes of halo every slab-decomposed buffer is allocated with."""
    h = 0
    for idx in range(len(lowered.launches)):
        for (b, f) in launch_reach(lowered, idx).values():
            h = max(h, b, f)
    return h


def total_reach(lowered):
    """Planes below / above the owned range that the *whole program* reaches for along the slab
    axis: the halo a rank needs to run all its launches without any exchange, recomputing the
    intermediate halo planes itself (what the overlapped host-array call does)."""
    need_back = need_fwd = 0
    acc = {}                       # field -> (back, fwd) still needed of it beyond the owned range
    for idx in range(len(lowered.launches) - 1, -1, -1):
        l = lowered.launches[idx]
        b0 = max([acc.get(w, (0, 0))[0] for w in l.writes] + [0])
        f0 = max([acc.get(w, (0, 0))[1] for w in l.writes] + [0])
        for f, (b, fw) in launch_reach(lowered, idx).items():
            old = acc.get(f, (0, 0))
            acc[f] = (max(old[0], b0 + b), max(old[1], f0 + fw))
            need_back, need_fwd = max(need_back, acc[f][0]), max(need_fwd, acc[f][1])
    return max(need_back, need_fwd)


def halo_schedule(lowered, slab: Slab) -> List[HaloSend]:
    """Which planes this rank must push to which neighbour after which launch.

    Field F written by launch l and read by a later launch with reach (back, fwd) along the slab
    axis: the upper neighbour needs my top ``back`` planes as its lower halo, the lower neighbour my
    bottom ``fwd`` planes as its upper halo.  Program inputs are loaded with their halos and never
    exchanged."""
    sends = []
    if lowered.slab_axis is None or slab.world == 1:
        return sends
    n_launch = len(lowered.launches)
    for idx, l in enumerate(lowered.launches):
        for field in l.writes:
            back = fwd = 0
            for later in range(idx + 1, n_launch):
                r = launch_reach(lowered, later).get(field)
                if r:
                    back, fwd = max(back, r[0]), max(fwd, r[1])
            if back and slab.rank + 1 < slab.world:
                sends.append(HaloSend(idx, field, slab.rank + 1, slab.end - back, slab.end))
            if fwd and slab.rank > 0:
                sends.append(HaloSend(idx, field, slab.rank - 1, slab.begin, slab.begin + fwd))
    return sends


# ------------------------------------------------------------------------------------ rendezvous


class Comm:
    rank = 0
    world = 1

    def allgather(self, obj):
        return [obj]

    def barrier(self):
        pass

    def max_float(self, x):
        return x

    def close(self):
        pass


class TorchComm(Comm):
    """``torch.distributed`` (gloo) as rendezvous: reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""

    def __init__(self, backend="gloo"):
        import torch.distributed as dist
        self.dist = dist
        self.owns = not dist.is_initialized()
        if self.owns:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend)
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self.dist.barrier()

    def max_float(self, x):
        return max(self.allgather(float(x)))

    def close(self):
        if self.owns and self.dist.is_initialized():
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------ device side


def _cuda_program_base():
    from .cuda_program import CudaProgram
    return CudaProgram


class SlabProgram:
    """A stencil program executed on this rank's slab.  Same execution interface as ``CudaProgram``
    (``execute``, ``launch_count``, ``plan``, ``buffers`` ...) plus slab-aware upload/download."""

    def __init__(self, stencil_file, comm: Comm, device=None, plan_options=None):
        from .cuda_program import CudaProgram
        from .kernel_chain_graph import KernelChainGraph
        from .stencil_op import make_program
        from . import planner
        self.comm = comm
        chain = KernelChainGraph(stencil_file)
        probe = planner.plan_program(make_program(chain), options=plan_options)
        axis = probe.lowered.slab_axis
        if axis is None:
            raise ValueError("1-D programs do not shard")
        n = probe.program.shape3[axis]
        self.halo = halo_depth(probe.lowered)
        # buffers wide enough for the exchange-free host-array call (``__call__``) when that costs
        # little: the accumulated reach of the program, at most a quarter of the thinnest slab
        wide = total_reach(probe.lowered)
        if os.environ.get("SFB200_SLAB_WIDE_HALO", "1") != "0" and wide <= (n // comm.world) // 4:
            self.halo = max(self.halo, wide)
        self.slab = Slab(comm.rank, comm.world, n, self.halo)
        self.inner = CudaProgram(chain=chain, device=device, plan_options=plan_options, slab=self.slab)
        self.rt = self.inner.rt
        self.program = self.inner.program
        self.plan = self.inner.plan
        self.lowered = self.inner.lowered
        self.buffers = self.inner.buffers
        self.sends = halo_schedule(self.lowered, self.slab)
        self.comm_stream = self.rt.stream
        self._setup_peers()
        self.exchange_seq = 0

    # -- forwarding
    def local_shape(self, name):
        return self.inner.local_shape(name)

    @property
    def launch_count(self):
        return self.inner.launch_count

    @property
    def launches_per_execution(self):
        return self.inner.launches_per_execution

    def set_scalars(self, values):
        self.inner.set_scalars(values)

    def input_index_offsets(self):
        """Global flat index of the first allocated element of every slab-decomposed input."""
        offs = {}
        it = "ijk"[self.lowered.slab_axis]
        for name, f in self.program.fields.items():
            if f.kind == "input" and not f.is_scalar and it in f.dims:
                plane = int(np.prod(f.shape[1:])) if f.dims[0] == it else None
                if plane is None:
                    raise ValueError("slab axis must be the outermost dimension of {}".format(name))
                offs[name] = self.slab.alloc_begin * plane
        return offs

    def _plane_elems(self, name):
        f = self.program.fields[name]
        return int(np.prod(f.shape[1:]))

    def _is_sharded(self, name):
        f = self.program.fields[name]
        return (not f.is_scalar) and ("ijk"[self.lowered.slab_axis] in f.dims)

    # -- peers
    def _setup_peers(self):
        rt = self.rt
        self.flags = rt.malloc(256)                      # [0]: halo from lower, [1]: halo from upper,
        rt.memset(self.flags, 0, 256)                    # [2]: lower neighbour done, [3]: upper neighbour done
        rt.stream_synchronize()
        storage = {}
        for name, buf in self.buffers.items():
            storage.setdefault(buf.dptr, rt.ipc_get_handle(buf.dptr))
        mine = {"flags": rt.ipc_get_handle(self.flags),
                "buffers": {name: storage[buf.dptr] for name, buf in self.buffers.items()},
                "alloc_begin": self.slab.alloc_begin}
        everyone = self.comm.allgather(pickle.dumps(mine))
        self.peers = {}
        opened = {}
        for peer in (self.comm.rank - 1, self.comm.rank + 1):
            if 0 <= peer < self.comm.world:
                info = pickle.loads(everyone[peer])
                ptrs = {}
                for name, handle in info["buffers"].items():
                    if handle not in opened:
                        opened[handle] = rt.ipc_open_handle(handle)
                    ptrs[name] = opened[handle]
                self.peers[peer] = {"flags": rt.ipc_open_handle(info["flags"]), "buffers": ptrs,
                                    "alloc_begin": info["alloc_begin"]}
        self._opened = list(opened.values()) + [p["flags"] for p in self.peers.values()]
        self.comm.barrier()

    # -- execution
    def execute(self):
        """All launches of the program on the owned slab, halos pushed to the neighbours after every
        launch whose result a later launch reads across the slab boundary."""
        inner, rt = self.inner, self.rt
        if inner._packs is None:
            inner._build_packs()
        stream = rt.stream
        rank, world = self.comm.rank, self.comm.world
        by_launch: Dict[int, List[HaloSend]] = {}
        for s in self.sends:
            by_launch.setdefault(s.launch, []).append(s)
        lower, upper = self.peers.get(rank - 1), self.peers.get(rank + 1)
        for idx, (l, fn, grid, pack) in enumerate(inner._packs):
            rt.launch(fn, grid, l.block, l.smem, pack.array, stream)
            sends = by_launch.get(idx)
            if sends is None:
                continue
            self.exchange_seq += 1
            seq = self.exchange_seq
            # the neighbours must have finished everything that could still read the halo regions
            # about to be overwritten: they acknowledge the previous exchange before we push
            if seq > 1:
                if lower is not None:
                    rt.wait_flag(stream, self.flags + 8, seq - 1)
                if upper is not None:
                    rt.wait_flag(stream, self.flags + 12, seq - 1)
            for s in sends:
                f = self.program.fields[s.field]
                plane_bytes = self._plane_elems(s.field) * f.data_type.bytes
                peer = self.peers[s.peer]
                src = self.buffers[s.field].dptr + (s.src_begin - self.slab.alloc_begin) * plane_bytes
                dst = peer["buffers"][s.field] + (s.src_begin - peer["alloc_begin"]) * plane_bytes
                rt.d2d(dst, src, (s.src_end - s.src_begin) * plane_bytes, stream)
            # tell the neighbours their halos are in place, then wait for ours
            if upper is not None:
                rt.write_flag(stream, upper["flags"] + 0, seq)      # I am their lower neighbour
            if lower is not None:
                rt.write_flag(stream, lower["flags"] + 4, seq)      # I am their upper neighbour
            if lower is not None:
                rt.wait_flag(stream, self.flags + 0, seq)
            if upper is not None:
                rt.wait_flag(stream, self.flags + 4, seq)
            # acknowledge: everything I launched before this point has consumed its halos once the
            # stream reaches here, so the neighbours may overwrite them at the next exchange
            if lower is not None:
                rt.write_flag(stream, lower["flags"] + 12, seq)     # I am their upper neighbour
            if upper is not None:
                rt.write_flag(stream, upper["flags"] + 8, seq)      # I am their lower neighbour
        inner.launch_count += len(inner._packs)

    # -- the reference-facing call with host arrays
    def __call__(self, **kwargs):
        """Run once with this rank's planes as host arrays (``local_shape``: owned planes + halo):
        inputs are copied in, the program runs, owned output planes are copied back in place.  When
        the buffers carry the accumulated reach of the program the call is cut into pieces whose
        copies and passes overlap and needs no halo exchange at all (``CudaProgram._pipeline_schedule``);
        otherwise: copy in, ``execute()`` with NVLink halo pushes, copy out."""
        inner, rt = self.inner, self.rt
        arrays, scalars = inner._split_call_args(kwargs)
        if scalars:
            inner.set_scalars(scalars)
        pieces = int(os.environ.get("SFB200_PIPELINE_PIECES", "16"))
        fields = self.program.fields
        in_out = [n for n, f in fields.items() if not f.is_scalar and f.kind in ("input", "output")]
        if pieces > 1 and all(n in arrays for n in in_out) and inner._call_pipelined(arrays, pieces):
            self.last_call_bytes = inner.last_call_bytes
            return
        h2d = d2h = 0
        for name, f in fields.items():
            if f.kind == "input" and not f.is_scalar:
                arr = np.ascontiguousarray(np.asarray(arrays[name], dtype=f.data_type.type))
                rt.h2d(self.buffers[name].dptr, arr.reshape(-1))
                h2d += arr.nbytes
        self.execute()
        for name in self.program.outputs:
            if name in arrays:
                out = arrays[name].reshape(-1)
                rt.d2h(out, self.buffers[name].dptr)
                d2h += out.nbytes
        rt.stream_synchronize()
        self.last_call_bytes = (h2d, d2h)

    # -- host data movement (parity runs; the benchmark generates its fields in HBM)
    def upload_global(self, name, array):
        """Copy this rank's allocated planes of a *global* host array to the device."""
        f = self.program.fields[name]
        arr = np.ascontiguousarray(np.asarray(array, dtype=f.data_type.type).reshape(f.shape))
        if self._is_sharded(name):
            arr = np.ascontiguousarray(arr[self.slab.alloc_begin:self.slab.alloc_end])
        self.rt.h2d(self.buffers[name].dptr, arr)
        self.rt.stream_synchronize()

    def download_owned(self, name):
        """Owned planes of a field as a host array."""
        local = self.inner.download(name)
        if not self._is_sharded(name):
            return local
        lo = self.slab.begin - self.slab.alloc_begin
        return np.ascontiguousarray(local[lo:lo + (self.slab.end - self.slab.begin)])

    def gather(self, name):
        """Full field on every rank (parity checks at reduced size only)."""
        parts = self.comm.allgather(self.download_owned(name))
        return np.concatenate(parts, axis=0) if self._is_sharded(name) else parts[0]

    def checksum_owned(self, name):
        """(sum, bit checksum) of the owned planes, computed on the device."""
        f = self.program.fields[name]
        plane = self._plane_elems(name)
        lo = (self.slab.begin - self.slab.alloc_begin) * plane if self._is_sharded(name) else 0
        n = (self.slab.end - self.slab.begin) * plane if self._is_sharded(name) else f.size
        return self.rt.checksum(self.buffers[name].dptr + lo * f.data_type.bytes, n, f.data_type.type)

    def close(self):
        self.rt.stream_synchronize()
        self.comm.barrier()
        for p in self._opened:
            try:
                self.rt.ipc_close_handle(p)
            except Exception:
                pass
        self.rt.free(self.flags)
        self.inner.close()
