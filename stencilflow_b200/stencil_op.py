"""The operator 5-tuple -- the plugin boundary of the path.

In the reference every operator is distilled to
``Stencil(label, shape, accesses, output_fields, boundary_conditions, code)``
(``stencilflow/stencil/stencil.py:44-59``) by ``_generate_stencil``
(``stencilflow/sdfg_generator.py:68-176``) and a backend is chosen through
``Stencil.implementations`` (``stencil.py:15-19``: "Intel FPGA", "Xilinx", "CPU").
``StencilOp`` is the same record, built from the same front-end objects, and
``IMPLEMENTATIONS`` gains the entry this package exists for: ``"CUDA"``.
"""

import collections
import os
from typing import Dict, List, Optional, Tuple

import networkx as nx

from . import dtypes
from . import expr as ex
from .base_node_class import Input, Output
from .kernel import Kernel

JUNK_VAL = -100000  # value read by out-of-bounds taps under "shrink" (reference stencil/_common.py:8)

IMPLEMENTATIONS = ("CUDA",)

BOUNDARY_TYPES = ("constant", "copy", "shrink")


class Field:
    """An array (or 0-D value) of the program: an input or an operator's result."""

    def __init__(self, name, data_type, dims, extents, kind):
        self.name = name
        self.data_type = data_type
        self.dims = list(dims)                      # iterators this field is indexed by, e.g. ["i", "k"]
        self.shape = tuple(extents[d] for d in dims)
        self.kind = kind                            # "input" | "output" | "intermediate"

    @property
    def is_scalar(self):
        return len(self.dims) == 0

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.size * self.data_type.bytes

    def __repr__(self):
        return "Field({}, {}, {}, {})".format(self.name, self.data_type, self.dims, self.kind)


class StencilOp:
    """One operator ready for lowering.

    ``accesses[field] = (dim_mask, [offset tuples over the field's own dims])`` and
    ``boundary_conditions[field] = {"btype": ..., "value": ...}`` have the reference's
    shape; ``statements`` is the computation as typed IR, ``code`` its original text."""

    implementation = "CUDA"

    def __init__(self, name, shape, iterators, accesses, output_fields, boundary_conditions,
                 code, statements, data_type, scalars):
        self.name = name
        self.label = name
        self.shape = tuple(shape)
        self.iterators = list(iterators)
        self.accesses = accesses
        self.output_fields = output_fields
        self.boundary_conditions = boundary_conditions
        self.code = code
        self.statements = statements
        self.data_type = data_type
        self.scalars = list(scalars)

    def offsets3(self, field) -> List[Tuple[Optional[int], ...]]:
        """Offsets of ``field`` as (di, dj, dk) with None for dimensions it lacks."""
        cache = self.__dict__.setdefault("_offsets3", {})
        if field not in cache:
            mask, offs = self.accesses[field]
            present = [it for it, m in zip(self.iterators, mask) if m]
            out = []
            for off in offs:
                by = dict(zip(present, off))
                out.append(tuple(by.get(it) for it in ("i", "j", "k")))
            cache[field] = out
        return list(cache[field])

    def extent(self, it):
        """(min, max) offset over all array accesses along iterator ``it`` (0,0 if none)."""
        lo = hi = 0
        pos = ex.ITERATORS.index(it)
        for field in self.accesses:
            for off in self.offsets3(field):
                if off[pos] is not None:
                    lo, hi = min(lo, off[pos]), max(hi, off[pos])
        return lo, hi

    def __repr__(self):
        return "StencilOp({}, shape={}, reads={})".format(self.name, self.shape, list(self.accesses))


class StencilProgram:
    """All operators of a chain in execution order plus the table of fields."""

    def __init__(self, name, shape, iterators, fields, ops, outputs, constants, vectorization):
        self.name = name
        self.shape = tuple(shape)            # extents of the program's own iterators
        self.iterators = list(iterators)
        self.fields: Dict[str, Field] = fields
        self.ops: List[StencilOp] = ops
        self.outputs = list(outputs)
        self.constants = constants
        self.vectorization = vectorization

    @property
    def extents(self):
        return dict(zip(self.iterators, self.shape))

    @property
    def shape3(self):
        return (1,) * (3 - len(self.shape)) + self.shape

    @property
    def cells(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def consumers(self, field):
        return [op for op in self.ops if field in op.accesses]


def _normalise_bc(bc, field, op_name):
    if bc is None:
        raise ValueError("Operator {} reads {} out of bounds but gives no boundary condition".format(
            op_name, field))
    kind = bc.get("btype", bc.get("type"))
    if kind not in BOUNDARY_TYPES:
        raise ValueError("Unsupported boundary condition type: {}".format(kind))
    out = {"btype": kind}
    if kind == "constant":
        if "value" not in bc:
            raise ValueError("constant boundary condition of {} in {} lacks a value".format(field, op_name))
        out["value"] = bc["value"]
    return out


def make_program(chain) -> StencilProgram:
    """KernelChainGraph -> operators in topological order (the order ``generate_reference``
    executes them in, ``stencilflow/sdfg_generator.py:638-675``)."""
    iterators = list(chain.iterators)
    shape = list(chain.dimensions[3 - chain.kernel_dimensions:])
    extents = dict(zip(iterators, shape))
    if chain.vectorization > 1 and shape[-1] % chain.vectorization != 0:
        raise ValueError("Shape not divisible by vectorization width")   # sdfg_generator.py:43-45

    fields: Dict[str, Field] = collections.OrderedDict()
    for name, cfg in chain.inputs.items():
        dims = cfg["input_dims"]
        for d in dims:
            if d not in iterators:
                raise ValueError("Input {} uses iterator {} the program does not have".format(name, d))
        fields[name] = Field(name, cfg["data_type"], [it for it in iterators if it in dims], extents, "input")

    order = [n for n in chain.topological_order() if isinstance(n, Kernel)]
    consumed = set()
    for node in order:
        consumed.update(node.graph.accesses.keys())
    for node in order:
        kind = "output" if node.name in chain.outputs else "intermediate"
        if kind == "intermediate" and node.name not in consumed:
            raise ValueError("Orphan operator {}: its result is neither consumed nor an output".format(
                node.name))
        fields[node.name] = Field(node.name, node.data_type, iterators, extents, kind)

    ops = []
    for node in order:
        taps = ex.collect_taps(node.statements)
        free = ex.collect_vars(node.statements)
        accesses = collections.OrderedDict()
        scalars = []
        for name in free:
            if name in fields and fields[name].is_scalar:
                scalars.append(name)
            elif name in chain.constants:
                scalars.append(name)
            elif name in fields:
                raise ValueError("Field {} is used without an index in operator {}".format(name, node.name))
            else:
                raise NameError("Unknown name {} in operator {}".format(name, node.name))
        bcs = {}
        for field, offsets in taps.items():
            if field not in fields:
                raise NameError("Operator {} reads unknown field {}".format(node.name, field))
            f = fields[field]
            if f.is_scalar:
                raise ValueError("0-D input {} cannot be indexed (operator {})".format(field, node.name))
            mask = [it in f.dims for it in iterators]
            offs = []
            for off in offsets:
                offs.append(tuple(off[ex.ITERATORS.index(it)] for it in f.dims))
            accesses[field] = (mask, offs)
            if any(o != 0 for off in offs for o in off):
                bcs[field] = _normalise_bc((node.boundary_conditions or {}).get(field), field, node.name)
            elif field in (node.boundary_conditions or {}):
                try:
                    bcs[field] = _normalise_bc(node.boundary_conditions[field], field, node.name)
                except ValueError:
                    pass
        statements = node.statements
        # SFB200_REASSOCIATE: "auto" (default) = taps that straddle two packed pairs are summed first (the
        # float32 kernels compute on pairs of k-neighbours: one issue slot in eight less) and sums become
        # balanced trees -- a dependent chain of two instead of three additions per cell, which the
        # latency-bound float64 2-D chain (two warps per scheduler) turns into 2.8 % (8.01 -> 7.79 ms with the
        # split loops of round 2; it lost 1.2 % before them) and the float32 Jacobi-3D pass into 0.8 %
        # (3.73 -> 3.70 ms).  "0": the reference's left-to-right order; "1": pairing only, float32 only; "2" /
        # "3": balanced trees for float32 / float64 only.  Both kernel families lower from the rewritten tree,
        # so fused and one-operator results stay bit-identical.
        mode = os.environ.get("SFB200_REASSOCIATE", "auto")
        is32, is64 = node.data_type == dtypes.float32, node.data_type == dtypes.float64
        if (is32 and mode in ("auto", "1", "2")) or (is64 and mode in ("auto", "3")):
            balance = mode != "1"
            statements = [ex.Statement(st.target, ex.pair_odd_taps(st.value, balance)) for st in statements]
        ops.append(StencilOp(
            name=node.name, shape=shape, iterators=iterators, accesses=accesses,
            output_fields={node.name: [0] * len(shape)}, boundary_conditions=bcs,
            code=node.kernel_string, statements=statements, data_type=node.data_type,
            scalars=scalars))
    for out in chain.outputs:
        if out not in fields or fields[out].kind != "output":
            raise ValueError("Output {} is not produced by any operator".format(out))
    return StencilProgram(chain.name, shape, iterators, fields, ops, chain.outputs,
                          dict(chain.constants), chain.vectorization)
