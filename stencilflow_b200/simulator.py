"""Cycle-level simulation of the streaming dataflow design of a stencil program (``-run-simulation``).

Counterpart of the reference's ``stencilflow/simulator.py`` + ``Kernel.try_read/try_execute/try_write``
(``stencilflow/kernel.py:634-738``): the program DAG is run as a network of processing elements that
exchange one word per cycle over bounded FIFO channels, which (a) yields every output a second,
fully independent way -- ``run_program`` compares it with the device result like the reference
compares it with the FPGA result (``run_program.py:232-248``) -- and (b) checks the *analysis*: the
delay buffers ``KernelChainGraph.compute_delay_buffer`` sized for every edge and the sliding windows
``Kernel.setup_internal_buffers`` sized for every field must suffice, otherwise the network
dead-locks and the simulation says on which channel.

This is a new model, not the reference's code: same interface (constructor arguments, ``simulate``,
``get_result``, ``step_execution``, ``all_done``) and the same machine, formulated directly as a
Kahn process network:

* every node streams its field in row-major order; a channel (src -> dst) holds the words
  ``[released, arrived)`` of that stream and has room for ``window + delay`` words, ``window`` being
  the span between the furthest-ahead and furthest-behind access of ``dst`` on that field (the sum of
  the internal-buffer chunks + 1) and ``delay`` the depth computed by ``compute_delay_buffer``;
* a processing element computes cell ``p`` in the cycle in which, for every in-bounds access, the
  word ``p + flat(offset)`` has arrived on the field's channel; out-of-bounds accesses read the
  boundary value (``constant``; ``copy`` = the centre access, ``shrink`` = the junk value the
  compiled paths use); the result leaves ``max_latency`` cycles later (``ComputeGraph.calculate_latency``)
  and is written to all successor channels at once -- a full successor stalls the whole element
  (back-pressure), exactly one word per cycle and edge;
* inputs emit one word per cycle; 0-D and lower-dimensional inputs are not streams but small
  memories every element can read at any time (``helper.load_array``'s full-shape arrays, indexed by
  the dimensions the input has);
* arithmetic is done in Python floats and rounded to the operator's data type per result, like
  ``Calculator.eval_expr`` + ``data_type(...)`` in the reference (``kernel.py:716-718``).

For programs inside the envelope of the reference's simulator the results agree with it to rounding
and the cycle counts to a fraction of a percent (tests/test_simulator.py uses the outputs and cycle
counts of the reference's simulator stored in tests/golden/reference_sim.*).
"""
import functools
import math
import operator
import os
from typing import Dict, List

import numpy as np

from . import expr as ex
from . import helper
from .kernel import Kernel
from .log_level import LogLevel

JUNK_VAL = -100000.0          # reference stencilflow/stencil/_common.py:8

_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "sinh": math.sinh, "cosh": math.cosh,
          "tanh": math.tanh, "sqrt": math.sqrt, "exp": math.exp, "log": math.log, "fabs": math.fabs,
          "abs": math.fabs, "floor": math.floor, "ceil": math.ceil, "min": min, "max": max, "pow": math.pow}
_BIN = {"+": operator.add, "-": operator.sub, "*": operator.mul, "/": operator.truediv}
_CMP = {"<": operator.lt, "<=": operator.le, ">": operator.gt, ">=": operator.ge, "==": operator.eq,
        "!=": operator.ne}


class SimulationDeadlock(RuntimeError):
    pass


class _Channel:
    """Words ``[released, arrived)`` of the producer's row-major stream, at most ``capacity`` of them."""

    def __init__(self, name, capacity, total):
        self.name = name
        self.capacity = capacity
        self.data = np.zeros(total, dtype=np.float64)
        self.arrived = 0
        self.released = 0
        self.max_occupancy = 0

    def has_room(self):
        return self.arrived - self.released < self.capacity

    def push(self, value):
        self.data[self.arrived] = value
        self.arrived += 1
        self.max_occupancy = max(self.max_occupancy, self.arrived - self.released)


class _Element:
    """One operator as a pipelined processing element."""

    def __init__(self, kernel: Kernel, dims, memories, scalars):
        self.kernel = kernel
        self.name = kernel.name
        self.dims = dims
        self.total = functools.reduce(operator.mul, dims, 1)
        self.memories = memories          # lower-dimensional inputs: name -> (ndarray, dims present)
        self.scalars = scalars            # 0-D inputs and program constants: name -> value
        self.latency = int(kernel.graph.max_latency)
        self.pipeline: List = [None] * self.latency      # results in flight, head = next to leave
        self.channels: Dict[str, _Channel] = {}          # streamed field -> channel
        self.successors: List[_Channel] = []
        self.pc = 0                                      # cells computed
        self.sent = 0                                    # results that left the pipeline
        self.stall_cycles = 0
        self.first_cycle = None
        self.last_cycle = None
        self.strides = [functools.reduce(operator.mul, dims[d + 1:], 1) for d in range(len(dims))]
        # accesses per streamed field: (offset tuple, flat offset)
        self.taps: Dict[str, List] = {}
        for stmt in kernel.statements:
            for node in ex.walk(stmt.value):
                if isinstance(node, ex.Tap) and node.field not in memories:
                    flat = sum((o or 0) * s for o, s in zip(node.offset, self.strides))
                    entry = (tuple(node.offset), flat)
                    if entry not in self.taps.setdefault(node.field, []):
                        self.taps[node.field].append(entry)
        self.behind = {f: min(0, min(fl for _, fl in t)) for f, t in self.taps.items()}

    def window(self, field):
        flats = [fl for _, fl in self.taps[field]]
        return max(max(flats), 0) - min(min(flats), 0) + 1

    # ---------------------------------------------------------------- cell evaluation
    def _coords(self, p):
        out = []
        for s in self.strides:
            out.append(p // s)
            p -= out[-1] * s
        return out

    def _in_bounds(self, coords, offset):
        for c, o, n in zip(coords, offset, self.dims):
            if o is not None and not 0 <= c + o < n:
                return False
        return True

    def ready(self, p):
        """All words cell ``p`` reads have arrived."""
        coords = self._coords(p)
        for field, taps in self.taps.items():
            ch = self.channels[field]
            for offset, flat in taps:
                if self._in_bounds(coords, offset) and p + flat >= ch.arrived:
                    return False
        return True

    def evaluate(self, p):
        coords = self._coords(p)
        kernel = self.kernel
        env = dict(self.scalars)

        def tap(node: ex.Tap):
            if node.field in self.memories:
                arr, present = self.memories[node.field]
                idx = []
                for d in present:
                    c = coords[d] + (node.offset[d] or 0)
                    if not 0 <= c < self.dims[d]:
                        return self._boundary(node, coords, tap)
                    idx.append(c)
                return float(arr[tuple(idx)])
            if not self._in_bounds(coords, node.offset):
                return self._boundary(node, coords, tap)
            flat = sum((o or 0) * s for o, s in zip(node.offset, self.strides))
            return float(self.channels[node.field].data[p + flat])

        def go(e):
            if isinstance(e, ex.Const):
                return float(e.value)
            if isinstance(e, ex.Tap):
                return tap(e)
            if isinstance(e, ex.Var):
                return env[e.name]
            if isinstance(e, ex.Bin):
                return _BIN[e.op](go(e.a), go(e.b))
            if isinstance(e, ex.Neg):
                return -go(e.a)
            if isinstance(e, ex.Cmp):
                return _CMP[e.op](go(e.a), go(e.b))
            if isinstance(e, ex.Logic):
                if e.op == "not":
                    return not go(e.args[0])
                if e.op == "and":
                    return all(go(a) for a in e.args)
                return any(go(a) for a in e.args)
            if isinstance(e, ex.Select):
                return go(e.a) if go(e.cond) else go(e.b)
            if isinstance(e, ex.Call):
                return _FUNCS[e.fn](*[go(a) for a in e.args])
            raise NotImplementedError("expression node {}".format(type(e).__name__))

        value = 0.0
        for stmt in kernel.statements:
            value = go(stmt.value)
            env[stmt.target] = value
        return kernel.data_type.type(value)

    def _boundary(self, node, coords, tap):
        bc = self.kernel.boundary_conditions.get(node.field)
        if bc is None:
            raise ValueError("operator {} reads {} out of bounds without a boundary condition".format(
                self.name, node.field))
        if bc["type"] == "constant":
            return float(bc["value"])
        if bc["type"] == "copy":
            centre = ex.Tap(node.field, tuple(None if o is None else 0 for o in node.offset))
            return tap(centre)
        if bc["type"] == "shrink":
            return JUNK_VAL
        raise NotImplementedError("boundary condition type {}".format(bc["type"]))

    def release(self):
        """Words behind the furthest-behind access of the next cell are never read again."""
        for field, ch in self.channels.items():
            ch.released = max(ch.released, min(self.pc + self.behind[field], ch.arrived))
            if self.pc >= self.total:
                ch.released = ch.arrived


class Simulator:
    def __init__(self, program_name: str, program_description: Dict, input_nodes: Dict, kernel_nodes: Dict,
                 output_nodes: Dict, dimensions: List[int], write_output: bool, log_level=LogLevel.NO_LOG) -> None:
        self.program_name = program_name
        self.program_description = program_description
        self.dimensions = list(dimensions)
        self.input_nodes = input_nodes
        self.kernel_nodes = kernel_nodes
        self.output_nodes = output_nodes
        self.write_output = write_output
        self.log_level = LogLevel(log_level) if isinstance(log_level, int) else log_level
        self.total = functools.reduce(operator.mul, self.dimensions, 1)
        self.cycles = 0
        self.elements: Dict[str, _Element] = {}
        self.streams: Dict[str, np.ndarray] = {}       # streamed inputs, flat
        self.input_pc: Dict[str, int] = {}
        self.input_channels: Dict[str, List[_Channel]] = {}
        self.results: Dict[str, np.ndarray] = {}
        self.result_count: Dict[str, int] = {}
        self._initialized = False

    # ---------------------------------------------------------------- set-up
    def initialize(self):
        desc = self.program_description
        prefix = desc.get("path") or os.getcwd()
        if os.path.isfile(prefix):
            prefix = os.path.dirname(prefix)
        ndim_json = len(desc["dimensions"])
        iterators = list(helper.ITERATORS[3 - ndim_json:])
        arrays = helper.load_input_arrays(desc["inputs"], prefix=prefix, shape=desc["dimensions"])
        memories, scalars = {}, {}
        for name, cfg in desc["inputs"].items():
            dims = cfg.get("input_dims", cfg.get("dimensions"))
            val = arrays[name]
            if dims is not None and len(dims) == 0:
                scalars[name] = float(val)
            elif dims is not None and list(dims) != iterators:
                # full-shape allocation, indexed by the dimensions the input has (leading elements)
                present = [["i", "j", "k"].index(d) for d in dims]
                arr = np.asarray(val)
                if arr.ndim != len(present):
                    arr = arr[tuple(slice(None) if it in dims else 0 for it in iterators)] \
                        if arr.ndim == len(iterators) else arr.reshape([self.dimensions[p] for p in present])
                memories[name] = (arr, present)
            else:
                self.streams[name] = np.ravel(np.asarray(val)).astype(np.float64)
                if self.streams[name].size != self.total:
                    raise ValueError("input {} has {} elements, the program {}".format(
                        name, self.streams[name].size, self.total))
        for name, cfg in (desc.get("constants") or {}).items():
            scalars[name] = float(cfg["value"])
        for name, kernel in self.kernel_nodes.items():
            self.elements[name] = _Element(kernel, self.dimensions, memories, scalars)
        for name, el in self.elements.items():
            for field in el.taps:
                delay = el.kernel.delay_buffer.get(field)
                depth = delay.maxsize if delay is not None else 1
                ch = _Channel("{}_{}".format(field, name), el.window(field) + max(depth, 1), self.total)
                el.channels[field] = ch
                if field in self.elements:
                    self.elements[field].successors.append(ch)
                elif field in self.streams:
                    self.input_channels.setdefault(field, []).append(ch)
                else:
                    raise ValueError("operator {} reads unknown field {}".format(name, field))
        for name in self.streams:
            self.input_pc[name] = 0
            self.input_channels.setdefault(name, [])
        for name in self.output_nodes:
            dt = self.kernel_nodes[name].data_type.type
            self.results[name] = np.zeros(self.total, dtype=dt)
            self.result_count[name] = 0
        self._initialized = True

    # ---------------------------------------------------------------- one cycle
    def step_execution(self):
        """Decisions are taken on the state at the start of the cycle (registered hand-shakes): first
        every element tries to compute a cell and to retire its oldest result, then inputs emit."""
        progressed = False
        room = {id(ch): ch.has_room() for el in self.elements.values() for ch in el.channels.values()}
        # results leaving the pipelines / cells entering them
        for name, el in self.elements.items():
            head = el.pipeline[0] if el.latency else None
            can_shift = True
            if el.latency and head is not None:
                can_shift = all(room[id(ch)] for ch in el.successors)
            fire = el.pc < el.total and can_shift and el.ready(el.pc)
            if el.latency == 0:
                if fire and not all(room[id(ch)] for ch in el.successors):
                    fire = False
                if fire:
                    self._retire(el, el.evaluate(el.pc))
            elif can_shift:
                if head is not None:
                    self._retire(el, head)
                    progressed = True
                el.pipeline.pop(0)
                el.pipeline.append(el.evaluate(el.pc) if fire else None)
                if any(v is not None for v in el.pipeline):
                    progressed = True
            if fire:
                if el.first_cycle is None:
                    el.first_cycle = self.cycles
                el.last_cycle = self.cycles
                el.pc += 1
                el.kernel.program_counter = el.pc
                progressed = True
            elif el.pc < el.total:
                el.stall_cycles += 1
        for el in self.elements.values():
            el.release()
        # inputs: one word per cycle to all consumers at once
        for name, data in self.streams.items():
            pc = self.input_pc[name]
            chans = self.input_channels[name]
            if pc < self.total and all(room[id(ch)] for ch in chans):
                for ch in chans:
                    ch.push(data[pc])
                self.input_pc[name] = pc + 1
                self.input_nodes[name].program_counter = pc + 1
                progressed = True
        self.cycles += 1
        return progressed

    def _retire(self, el: _Element, value):
        for ch in el.successors:
            ch.push(value)
        if el.name in self.results:
            n = self.result_count[el.name]
            self.results[el.name][n] = value
            self.result_count[el.name] = n + 1
            self.output_nodes[el.name].program_counter = n + 1
        el.sent += 1

    def all_done(self) -> bool:
        if not self._initialized:
            return False
        if any(pc < self.total for pc in self.input_pc.values()):
            return False
        if any(el.sent < el.total for el in self.elements.values()):
            return False
        return all(n >= self.total for n in self.result_count.values())

    # ---------------------------------------------------------------- driver
    def simulate(self):
        if self.log_level >= LogLevel.MODERATE:
            print("Initialize simulation.")
        self.initialize()
        if self.log_level >= LogLevel.MODERATE:
            print("Running simulation...")
        idle = 0
        while not self.all_done():
            if self.step_execution():
                idle = 0
            else:
                idle += 1
                if idle > 2:
                    raise SimulationDeadlock(self.diagnostics())
        if self.log_level >= LogLevel.MODERATE:
            print("Simulation done after {} cycles.".format(self.cycles))
        self.finalize()

    def finalize(self):
        if self.write_output:
            folder = os.path.join("results", self.program_name, "simulation")
            os.makedirs(folder, exist_ok=True)
            helper.save_output_arrays({k: v.reshape(self.program_description["dimensions"])
                                       for k, v in self.results.items()}, folder)
        if self.log_level >= LogLevel.MODERATE:
            print(self.report())

    def get_result(self) -> Dict[str, np.ndarray]:
        return {k: v.copy() for k, v in self.results.items()}

    def channel_usage(self):
        """{channel name: (capacity analysed for it, maximum occupancy observed)}"""
        return {ch.name: (ch.capacity, ch.max_occupancy)
                for el in self.elements.values() for ch in el.channels.values()}

    def report(self) -> str:
        lines = ["simulated {} cycles for {} cells ({:.1f} % of the cycles stream a word)".format(
            self.cycles, self.total, 100.0 * self.total / max(1, self.cycles))]
        for el in self.elements.values():
            lines.append("  {:<16} latency {:>4}  first result cycle {:>8}  stalled {:>8} cycles".format(
                el.name, el.latency, (el.first_cycle or 0) + el.latency, el.stall_cycles))
        for name, (cap, used) in sorted(self.channel_usage().items()):
            lines.append("  channel {:<24} capacity {:>8}  max. occupancy {:>8}".format(name, cap, used))
        return "\n".join(lines)

    def diagnostics(self, ex_=None) -> str:
        lines = ["dead-lock after {} cycles: no processing element can advance".format(self.cycles)]
        for el in self.elements.values():
            lines.append("  {}: computed {} / {} cells, {} results in flight".format(
                el.name, el.pc, el.total, sum(v is not None for v in el.pipeline)))
            for field, ch in el.channels.items():
                lines.append("    channel {}: {} of {} words held (arrived {}, released {})".format(
                    ch.name, ch.arrived - ch.released, ch.capacity, ch.arrived, ch.released))
        return "\n".join(lines)
