"""Per-operator analysis of a computation string.

Exposes what the rest of the front end (and the reference's callers) read from a
``ComputeGraph``: ``accesses``, ``min_index``/``max_index``, ``buffer_size``,
``max_latency``, ``inputs``/``outputs`` (reference ``stencilflow/compute_graph.py:41-171,461-532``).
The expression itself is held as the typed IR of :mod:`expr` (``statements``)
instead of a networkx graph of AST nodes; the derived quantities are the same.
"""

import math
from typing import Dict, List, Optional

from . import expr as ex
from . import helper


class OutputNode:
    """The assignment target that carries the operator's result."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "Output({})".format(self.name)


class ComputeGraph:
    def __init__(self, verbose=False, dimensions=3, vectorization=1, raw_inputs=None):
        self.verbose = verbose
        self.dimensions = dimensions
        self.vectorization = vectorization
        self.raw_inputs = raw_inputs if raw_inputs is not None else {}
        self.config = helper.parse_json("compute_graph.config")
        self.statements: List[ex.Statement] = []
        self.max_latency = -1
        self.inputs = set()
        self.outputs = set()
        self.min_index: Dict[str, List] = {}
        self.max_index: Dict[str, List] = {}
        self.buffer_size: Dict[str, List] = {}
        self.accesses: Dict[str, List[List]] = {}

    # ------------------------------------------------------------------ parsing
    def _program_iterators(self) -> List[str]:
        dims = self.dimensions
        if isinstance(dims, int):
            return list(helper.ITERATORS[3 - dims:])
        return list(helper.ITERATORS[3 - len(dims):])

    def generate_graph(self, computation_string: str, default_dims: Optional[List[str]] = None):
        field_dims = {
            name: (cfg.get("input_dims") if isinstance(cfg, dict) else None)
            for name, cfg in self.raw_inputs.items()
        }
        self.statements = ex.parse_computation(
            computation_string, field_dims,
            default_dims if default_dims is not None else self._program_iterators())
        # every temporary must feed the result (reference compute_graph.py:237-243)
        targets = [s.target for s in self.statements]
        used = set()
        for s in self.statements:
            used.update(v.name for v in ex.walk(s.value) if isinstance(v, ex.Var))
        for t in targets[:-1]:
            if t not in used and t != targets[-1]:
                raise RuntimeError(
                    "Kernel-internal data flow is not single component (must be connected in the sense "
                    "of a DAG).")
        return self.statements

    # --------------------------------------------------------------- inputs/outputs
    def determine_inputs_outputs(self):
        """Leaves of the data flow: field accesses, free names and literals are inputs,
        the final assignment target is the output."""
        assigned = set()
        leaves = []
        for s in self.statements:
            for node in ex.walk(s.value):
                if isinstance(node, ex.Tap) or isinstance(node, ex.Const):
                    leaves.append(node)
                elif isinstance(node, ex.Var) and node.name not in assigned:
                    leaves.append(node)
            assigned.add(s.target)
        self.inputs = set(leaves)
        self.outputs = {OutputNode(self.statements[-1].target)}

    # ------------------------------------------------------------------- windows
    def setup_internal_buffers(self, relative_to_center=True):
        self.min_index, self.max_index, self.buffer_size, self.accesses = {}, {}, {}, {}
        for node in self.inputs:
            if isinstance(node, ex.Tap):
                idx = node.index
                if node.name in self.min_index:
                    if _lex_less(idx, self.min_index[node.name]):
                        self.min_index[node.name] = idx
                    if not _lex_less(idx, self.max_index[node.name]):
                        self.max_index[node.name] = idx
                else:
                    self.min_index[node.name] = idx
                    self.max_index[node.name] = idx
                self.accesses.setdefault(node.name, []).append(idx)
            elif isinstance(node, ex.Var) and node.name in self.raw_inputs:
                self.min_index[node.name] = [0, 0, 0]
                self.max_index[node.name] = [0, 0, 0]
                self.accesses[node.name] = [[0, 0, 0]]
        for name in self.accesses:
            size = [abs(a - b) if a is not None and b is not None else None
                    for a, b in zip(self.max_index[name], self.min_index[name])]
            size[-1] = (size[-1] if size[-1] is not None else 0) + (self.vectorization - 1)
            self.buffer_size[name] = size
        if not relative_to_center:
            for name in self.accesses:
                self.accesses[name] = [
                    helper.list_subtract_cwise(a, self.max_index[name]) for a in self.accesses[name]
                ]

    # ------------------------------------------------------------------- latency
    def calculate_latency(self):
        """Critical path through the expression in FPGA pipeline cycles, using the
        per-operation table of ``compute_graph.config`` (reference compute_graph.py:461-532):
        the result register costs 1, every operation on the way to a leaf adds its
        latency, temporaries chain through their defining statement."""
        table = self.config["op_latency"]
        defs = {}

        def depth(e) -> int:
            if isinstance(e, (ex.Const, ex.Tap)):
                return 0
            if isinstance(e, ex.Var):
                return defs.get(e.name, 0)
            if isinstance(e, ex.Bin):
                cost = table[{"+": "add", "-": "sub", "*": "mult", "/": "div"}[e.op]]
            elif isinstance(e, ex.Neg):
                cost = table["neg"]
            elif isinstance(e, ex.Cmp) or isinstance(e, ex.Logic):
                cost = table["comparison"]
            elif isinstance(e, ex.Select):
                cost = table["conditional"]
            elif isinstance(e, ex.Call):
                cost = table.get(e.fn, table["sqrt"])
            else:
                raise NotImplementedError("Node type {} has not been implemented yet.".format(type(e)))
            return cost + max((depth(c) for c in e.children()), default=0)

        total = 0
        for s in self.statements:
            defs[s.target] = depth(s.value)
            total = defs[s.target]
        self.max_latency = math.ceil((1 + total) / self.vectorization)

    def try_set_max_latency(self, new_val):
        if self.max_latency <= new_val:
            self.max_latency = new_val
            return True
        return False


def _lex_less(a, b):
    """Lexicographic ``a < b`` over index lists whose ``None`` entries line up."""
    for x, y in zip(a, b):
        if x is None or y is None or x == y:
            continue
        return x < y
    return False
