"""One operator of a stencil program: a named output field computed cell-wise from
neighbouring cells of inputs and of other operators' fields.

Constructor signature and attribute names follow reference ``stencilflow/kernel.py:27-115``.
The cycle-level simulator hooks of the reference (``try_read/try_execute/try_write``)
are not part of this backend; what remains is the static analysis the planner and
the lowering consume.
"""

from typing import Dict, List

from . import expr as ex
from . import helper
from .base_node_class import BaseKernelNodeClass
from .bounded_queue import BoundedQueue
from .compute_graph import ComputeGraph


class Kernel(BaseKernelNodeClass):
    def __init__(self, name, kernel_string, dimensions, data_type, boundary_conditions,
                 raw_inputs, vectorization=1, plot_graph=False, verbose=False,
                 default_dims=None):
        super().__init__(name, BoundedQueue(name="dummy", maxsize=0), data_type)
        self.kernel_string = kernel_string
        self.raw_inputs = raw_inputs
        self.dimensions = dimensions
        self.boundary_conditions = boundary_conditions
        self.verbose = verbose
        self.vectorization = vectorization
        self.config = helper.parse_json("kernel.config")
        self.graph = ComputeGraph(vectorization=vectorization, dimensions=dimensions,
                                  raw_inputs=raw_inputs)
        self.graph.generate_graph(kernel_string, default_dims=default_dims)
        self.graph.calculate_latency()
        self.graph.determine_inputs_outputs()
        self.graph.setup_internal_buffers()
        self.internal_buffer: Dict[str, List[BoundedQueue]] = {}
        self.setup_internal_buffers()
        self.dist_to_center = {}
        self.set_up_dist_to_center()

    # statements of the computation string as typed IR
    @property
    def statements(self):
        return self.graph.statements

    @staticmethod
    def remove_duplicate_accesses(inp):
        out = []
        for row in inp:
            if list(row) not in out:
                out.append(list(row))
        return out

    def setup_internal_buffers(self):
        """Split each field's sliding window at its accesses: sorted from the
        furthest-ahead access backwards, consecutive accesses ``d`` flattened words
        apart give a chunk of ``d`` words (reference kernel.py:388-427).  For 3-D
        Jacobi at 32^3 this yields 992, 31, 2, 31, 992 = two planes."""
        for name in self.graph.accesses:
            self.graph.accesses[name] = self.remove_duplicate_accesses(self.graph.accesses[name])
        for name in self.graph.buffer_size:
            chunks = []
            acc = self.graph.accesses[name]
            acc.sort(key=lambda idx: [x if x is not None else 0 for x in idx], reverse=True)
            if len(acc) == 1:
                chunks.append(BoundedQueue(name=name, maxsize=1, collection=[None]))
            elif len(acc) > 1:
                for pre, cur in zip(acc, acc[1:]):
                    diff = abs(helper.convert_3d_to_1d(
                        index=helper.list_subtract_cwise(pre, cur), dimensions=self.dimensions))
                    if diff:
                        chunks.append(BoundedQueue(name=name, maxsize=diff,
                                                   collection=[None] * diff))
            self.internal_buffer[name] = chunks

    def set_up_dist_to_center(self):
        """Flattened distance from the furthest-ahead access of each field to its
        centre, i.e. how long the FPGA pipeline must fill before the first result."""
        for name in self.graph.accesses:
            furthest = self.graph.max_index[name]
            self.dist_to_center[name] = helper.convert_3d_to_1d(
                dimensions=self.dimensions,
                index=[x if x is None or x > 0 else 0 for x in furthest])

    def generate_relative_access_kernel_string(self, relative_to_center=True,
                                               replace_negative_index=False,
                                               python_syntax=False, flatten_index=True,
                                               output_dimensions=None):
        """Computation with every access rewritten as ``field[<flat offset>]`` (or
        ``field_<offset>`` names), statements joined by ``; ``
        (reference kernel.py:327-368)."""
        dims = output_dimensions if output_dimensions is not None else self.dimensions

        def tap(t):
            off = list(t.offset)
            if not relative_to_center:
                off = helper.list_subtract_cwise(off, self.graph.max_index[t.field])
            if flatten_index:
                val = helper.convert_3d_to_1d(dimensions=dims, index=off)
                text = str(val)
            else:
                text = ", ".join(str(o) for o in off if o is not None)
            if python_syntax:
                return "{}[{}]".format(t.field, text)
            text = text.replace(", ", "_")
            if replace_negative_index:
                text = text.replace("-", "n")
            return "{}_{}".format(t.field, text)

        parts = []
        for s in self.statements[:-1]:
            parts.append("{} = {}".format(s.target, ex.to_source(s.value, tap)))
        parts.append("{} = {}".format(self.name, ex.to_source(self.statements[-1].value, tap)))
        return "; ".join(parts)
