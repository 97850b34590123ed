"""The stencil program as a DAG of Input, Kernel (operator) and Output nodes.

Constructor, attribute names and the analysis pipeline
``import_input -> create_kernels -> compute_kernel_latency -> connect_kernels ->
compute_delay_buffer -> add_channels`` follow reference
``stencilflow/kernel_chain_graph.py:32-102``.  The FPGA-specific results (delay
buffer depths, sliding-window chunks) are still computed because callers and the
report read them; the CUDA planner derives its plane lags from the same
accumulated path lengths, restricted to the streamed dimension.
"""

import ast
import copy
import functools
import operator
import os
from typing import Dict, List

import networkx as nx

from . import helper
from .base_node_class import Input, Output
from .bounded_queue import BoundedQueue
from .kernel import Kernel
from .log_level import LogLevel


class KernelChainGraph:
    def __init__(self, path, plot_graph=False, log_level=LogLevel.NO_LOG, program=None):
        """``path``: JSON program file.  ``program`` (extension): an already parsed
        program dict (as returned by ``helper.parse_json``) to use instead of reading
        ``path``; ``path`` then only provides the name."""
        if isinstance(log_level, int):
            log_level = LogLevel(log_level)
        self.path = os.path.abspath(path)
        self.log_level = log_level
        self.inputs: Dict[str, Dict] = {}
        self.outputs: List[str] = []
        self.dimensions: List[int] = []
        self.program: Dict[str, Dict] = {}
        self.vectorization = 1
        self.kernel_latency = None
        self.channels: Dict[str, Dict] = {}
        self.graph = nx.DiGraph()
        self.input_nodes: Dict[str, Input] = {}
        self.output_nodes: Dict[str, Output] = {}
        self.kernel_nodes: Dict[str, Kernel] = {}
        self.config = helper.parse_json("stencil_chain.config")
        self.name = os.path.splitext(os.path.basename(self.path))[0]
        self.kernel_dimensions = -1
        self.constants = {}
        self._program_override = program
        self._say(LogLevel.MODERATE, "Initialize KernelChainGraph.")
        self._say(LogLevel.MODERATE, "Read input config files.")
        self.import_input()
        self._say(LogLevel.MODERATE, "Create all kernels.")
        self.create_kernels()
        self._say(LogLevel.MODERATE, "Compute kernel latencies.")
        self.compute_kernel_latency()
        self._say(LogLevel.MODERATE, "Connect kernels.")
        self.connect_kernels()
        self._say(LogLevel.MODERATE, "Compute delay buffer sizes.")
        self.compute_delay_buffer()
        if plot_graph:
            self.plot_graph(self.name + ".png")
        self._say(LogLevel.MODERATE, "Add channels to the graph edges.")
        self.add_channels()
        for kernel in self.program:
            text = self.program[kernel]["computation_string"]
            if "sin" in text or "cos" in text or "tan" in text:
                print("Warning: Computation contains sinusoidal functions with experimental latency values.")
        if self.log_level >= LogLevel.MODERATE:
            self.report(self.name)

    def _say(self, level, text):
        if self.log_level >= level:
            print(text)

    # ------------------------------------------------------------------ loading
    def import_input(self):
        """Read the program sections.  Programs with fewer than three dimensions are
        padded in front (``[N, M] -> [1, N, M]``) and use the *last* iterators, i.e. 2-D
        programs are written in ``j, k`` and 1-D programs in ``k``
        (reference kernel_chain_graph.py:364-405)."""
        inp = copy.deepcopy(self._program_override) if self._program_override is not None \
            else helper.parse_json(self.path)
        self.kernel_dimensions = len(inp["dimensions"])
        if not 1 <= self.kernel_dimensions <= 3:
            raise ValueError("Programs must have 1 to 3 dimensions")
        self.constants = copy.copy(inp["constants"]) if "constants" in inp else {}
        self.vectorization = int(inp["vectorization"]) if "vectorization" in inp else 1
        self.program = inp["program"]
        self.inputs = inp["inputs"]
        self.iterators = list(helper.ITERATORS[3 - self.kernel_dimensions:])
        for cfg in self.inputs.values():
            if "input_dims" not in cfg:
                cfg["input_dims"] = cfg["dimensions"] if "dimensions" in cfg else list(self.iterators)
        self.outputs = inp["outputs"]
        self.dimensions = [1] * (3 - self.kernel_dimensions) + list(inp["dimensions"])

    def total_elements(self):
        return functools.reduce(operator.mul, self.dimensions, 1)

    def create_kernels(self):
        self.kernel_nodes = {}
        for name, entry in self.program.items():
            node = Kernel(name=name,
                          kernel_string=str(entry["computation_string"]),
                          dimensions=self.dimensions,
                          data_type=entry["data_type"],
                          boundary_conditions=entry.get("boundary_conditions", {}),
                          raw_inputs=self.inputs,
                          vectorization=self.vectorization,
                          default_dims=self.iterators)
            self.graph.add_node(node)
            self.kernel_nodes[name] = node
        self.input_nodes = {}
        for name, cfg in self.inputs.items():
            node = Input(name=name, data_type=cfg["data_type"],
                         data_queue=BoundedQueue(name=name, maxsize=self.total_elements()))
            self.input_nodes[name] = node
            self.graph.add_node(node)
        self.output_nodes = {}
        for name in self.outputs:
            if name not in self.program:
                raise ValueError("Output {} is not produced by any operator".format(name))
            node = Output(name=name, data_type=self.program[name]["data_type"],
                          dimensions=self.dimensions)
            self.output_nodes[name] = node
            self.graph.add_node(node)

    def compute_kernel_latency(self):
        self.kernel_latency = {n: k.graph.max_latency for n, k in self.kernel_nodes.items()}

    def at_least_one(self, value):
        return value if value > 0 else 1

    # ------------------------------------------------------------------ edges
    def _consumed_names(self, kernel):
        """Names of fields / 0-D inputs an operator reads (anything in its data-flow
        leaves that is a field access or a bare name)."""
        return {n.name for n in kernel.graph.inputs if isinstance(n.name, str)}

    def connect_kernels(self):
        """An operator reading ``x`` depends on input ``x`` or on operator ``x``;
        an output is fed by the operator of the same name
        (reference kernel_chain_graph.py:243-272)."""
        for dest in self.kernel_nodes.values():
            for name in sorted(self._consumed_names(dest)):
                if name == dest.name:
                    continue
                if name in self.kernel_nodes:
                    self.graph.add_edge(self.kernel_nodes[name], dest, channel=None)
                elif name in self.input_nodes:
                    self.graph.add_edge(self.input_nodes[name], dest, channel=None)
        for name, out in self.output_nodes.items():
            self.graph.add_edge(self.kernel_nodes[name], out, channel=None)

    def topological_order(self):
        try:
            return list(nx.topological_sort(self.graph))
        except nx.NetworkXUnfeasible:
            cycle = next(nx.simple_cycles(self.graph))
            raise ValueError("Cycle detected: {}".format([c.name for c in cycle]))

    @staticmethod
    def greater(a, b):
        """Lexicographic ``a > b`` where a ``None`` leading entry never wins."""
        if len(a) == 0 or len(b) == 0:
            return False
        if a[0] is None:
            return False
        if b[0] is None:
            return True
        if a[0] != b[0]:
            return a[0] > b[0]
        return KernelChainGraph.greater(a[1:], b[1:])

    def compute_delay_buffer(self):
        """Size the FIFO on every edge so that fork/join paths meet in step.

        Every node carries, per *program input*, the accumulated stream distance
        ``[di, dj, dk + latency, via]`` of every path from that input.  The longest
        path (plus one cycle) sets the pace; each other path gets a buffer of the
        flattened difference (reference kernel_chain_graph.py:476-559)."""
        for node in self.topological_order():
            for src in node.input_paths:
                longest = max(node.input_paths[src])
                longest[2] += 1
                for entry in node.input_paths[src]:
                    via = entry[-1]
                    depth = helper.convert_3d_to_1d(
                        dimensions=self.dimensions,
                        index=helper.list_subtract_cwise(longest[:-1], entry[:-1]))
                    queue = BoundedQueue(name=via, maxsize=depth)
                    queue.import_data([None] * queue.maxsize)
                    node.delay_buffer[via] = queue
            if isinstance(node, Input):
                node.delay_buffer = BoundedQueue(name=node.name, maxsize=1, collection=[None])
            for succ in self.graph.successors(node):
                if isinstance(node, Input):
                    succ.input_paths.setdefault(node.name, []).append(
                        [0] * len(self.dimensions) + [node.name])
                elif isinstance(node, Kernel):
                    ahead = [0, 0, 0]
                    for field, acc in node.graph.accesses.items():
                        far = max(acc, key=lambda idx: [x if x is not None else -10**9 for x in idx])
                        if KernelChainGraph.greater(far, ahead):
                            ahead = far
                    latency = node.graph.max_latency
                    for src, paths in node.input_paths.items():
                        base = max(paths)
                        total = [a + d if a is not None else d for a, d in zip(ahead, base)]
                        total[-1] += latency
                        total.append(node.name)
                        succ.input_paths.setdefault(src, []).append(total)

    def add_channels(self):
        """One channel record per edge: delay FIFO, sliding-window chunks and type
        (reference kernel_chain_graph.py:274-362)."""
        self.channels = {}
        for src, dest in self.graph.edges:
            name = src.name + "_" + dest.name
            if isinstance(dest, Kernel):
                channel = {
                    "name": name,
                    "delay_buffer": dest.delay_buffer.get(src.name, BoundedQueue(src.name, 1)),
                    "internal_buffer": dest.internal_buffer.get(src.name, []),
                    "data_type": src.data_type,
                }
                if isinstance(src, Input):
                    channel["input_dims"] = self.inputs[src.name].get("input_dims")
            else:
                channel = {
                    "name": name,
                    "delay_buffer": dest.delay_buffer.get(src.name, BoundedQueue(src.name, 1)),
                    "internal_buffer": {},
                    "data_type": src.data_type,
                }
            self.channels[name] = channel
            src.outputs[dest.name] = channel
            dest.inputs[src.name] = channel
            self.graph[src][dest]["channel"] = channel

    # ------------------------------------------------------------------ models
    def compute_critical_path_dim(self):
        crit = [0] * len(self.dimensions)
        for output in self.outputs:
            node = self.kernel_nodes[output]
            if not node.input_paths:
                continue
            lat = node.graph.max_latency
            src = max(node.input_paths)
            path = list(max(node.input_paths[src]))
            path[2] += lat
            crit = path[:-1]
        return crit

    def compute_critical_path(self):
        return helper.convert_3d_to_1d(index=self.compute_critical_path_dim(),
                                       dimensions=self.dimensions)

    def operation_count(self):
        """Per operation type: (ops per cell summed over operators, ops in total)
        (reference kernel_chain_graph.py:721-747)."""
        cells = self.total_elements()
        operations = {}
        for kernel in self.kernel_nodes.values():
            counter = helper.OpCounter()
            counter.visit(ast.parse(kernel.kernel_string.strip()))
            for name, count in counter.operation_count.items():
                per_cell, total = operations.get(name, (0, 0))
                operations[name] = (per_cell + count, total + cells * count)
        return operations

    def minimum_communication_volume(self):
        """Bytes that must cross the off-chip interface when the whole program is
        fused: every input once at its own dimensionality, every output once
        (reference kernel_chain_graph.py:749-768)."""
        volume = 0
        for cfg in self.inputs.values():
            elements = functools.reduce(
                operator.mul,
                [self.dimensions[helper.ITERATORS.index(it)] for it in cfg["input_dims"]], 1)
            volume += cfg["data_type"].bytes * elements
        for name in self.outputs:
            volume += self.program[name]["data_type"].bytes * self.total_elements()
        return volume

    def runtime_lower_bound(self):
        return (self.total_elements() + self.compute_critical_path()) // self.vectorization

    def enumerate_cuts(self):
        """All descendant-closed two-colourings of the operator DAG, i.e. the candidate
        places to cut the operator pipeline in two (reference kernel_chain_graph.py:116-160).
        Returns (list of coloured graphs, node -> index)."""
        kernels = [n for n in self.topological_order() if isinstance(n, Kernel)]
        index = {n: i for i, n in enumerate(kernels)}
        sub = self.graph.subgraph(kernels)
        seen, cuts = set(), []
        frontier = [frozenset()]
        while frontier:
            coloured = frontier.pop()
            for n in kernels:
                if n in coloured:
                    continue
                new = frozenset(coloured | {n} | nx.descendants(sub, n))
                if new in seen or len(new) == len(kernels):
                    continue
                seen.add(new)
                frontier.append(new)
                g = nx.DiGraph(sub)
                for m in g.nodes:
                    g.nodes[m]["color"] = 1 if m in new else 0
                cuts.append(g)
        return cuts, index

    def plot_graph(self, save_path=None):
        print("Plotting is not available in this backend (requested: {}).".format(save_path))

    # ------------------------------------------------------------------ report
    def report(self, name):
        print("Report of {}\n".format(name))
        print("dimensions of data array: {}\n".format(self.dimensions))
        print("channel info:")
        for _, _, channel in self.graph.edges(data="channel"):
            if channel is not None:
                print("internal buffers:\n {}".format(channel["internal_buffer"]))
                print("delay buffers:\n {}".format(channel["delay_buffer"]))
        print()
        sections = [
            ("field access info:", "field accesses", lambda k: k.graph.accesses),
            ("internal buffer size info:", "internal buffer size", lambda k: k.graph.buffer_size),
            ("internal buffer chunks info:", "internal buffer chunks", lambda k: k.internal_buffer),
            ("delay buffer size info:", "delay buffer size", lambda k: k.delay_buffer),
            ("path length info:", "path lengths", lambda k: k.input_paths),
            ("latency info:", "node latency", lambda k: k.graph.max_latency),
        ]
        for title, label, get in sections:
            print(title)
            for n, k in self.kernel_nodes.items():
                print("node name: {}, {}: {}".format(n, label, get(k)))
            print()
        print("critical path info:")
        print("critical path length is {}\n".format(self.compute_critical_path()))
        total = 0
        for _, _, channel in self.graph.edges(data="channel"):
            if channel is not None:
                total += sum(q.maxsize for q in channel["internal_buffer"]) \
                    if isinstance(channel["internal_buffer"], list) else 0
                total += channel["delay_buffer"].maxsize
        print("total buffer info:")
        print("total buffer size: {}\n".format(total))
        print("input kernel string info:")
        for n, k in self.kernel_nodes.items():
            print("input kernel string of {} is: {}".format(n, k.kernel_string))
        print()
        print("relative access kernel string info:")
        for n, k in self.kernel_nodes.items():
            print("relative access kernel string of {} is: {}".format(
                n, k.generate_relative_access_kernel_string()))
        print()
        print("analytical model:")
        print("operation count: {}".format(self.operation_count()))
        print("minimum communication volume: {} bytes".format(self.minimum_communication_volume()))
        print("runtime lower bound: {} cycles".format(self.runtime_lower_bound()))


def main(argv=None):
    """``python -m stencilflow_b200.kernel_chain_graph -stencil_file prog.json [-plot] [-simulate] [-report]
    [-log-level N]`` -- the debugging entry point of reference kernel_chain_graph.py:777-817."""
    import argparse
    import re
    parser = argparse.ArgumentParser()
    parser.add_argument("-stencil_file", required=True)
    parser.add_argument("-plot", action="store_true")
    parser.add_argument("-log-level", default=LogLevel.MODERATE.value, type=int)
    parser.add_argument("-report", action="store_true")
    parser.add_argument("-simulate", action="store_true")
    args = parser.parse_args(argv)
    level = LogLevel(args.log_level)
    description = helper.parse_json(args.stencil_file)
    chain = KernelChainGraph(path=args.stencil_file, plot_graph=args.plot, log_level=level)
    sim = None
    if args.simulate:
        from .simulator import Simulator
        sim = Simulator(program_name=re.match(r"[^\.]+", os.path.basename(args.stencil_file)).group(0),
                        program_description=description, input_nodes=chain.input_nodes,
                        kernel_nodes=chain.kernel_nodes, output_nodes=chain.output_nodes,
                        dimensions=chain.dimensions, write_output=False, log_level=level)
        sim.simulate()
    if args.report:
        if level < LogLevel.MODERATE:          # at MODERATE and above the constructor has printed it
            chain.report(args.stencil_file)
        if sim is not None:
            print(sim.report())
    return chain, sim


if __name__ == "__main__":
    main()
