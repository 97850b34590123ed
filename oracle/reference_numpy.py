"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's
CPU program for a StencilFlow JSON stencil program.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  It deliberately shares no code with ``stencilflow_b200`` (it parses the JSON and
the computation strings itself) so that a front-end bug cannot cancel out in a parity test.

What is restated, with the reference lines followed:

* program assembly -- ``stencilflow/sdfg_generator.py:580-677`` (``generate_reference``): operators
  run one at a time in topological order; every operator's field is a full-shape array of the
  operator's ``data_type``; 0-D inputs are scalars; ``vectorization`` is ignored.
* per-cell semantics -- ``stencilflow/stencil/cpu.py:58-115,141-179`` (``ExpandStencilCPU``): a tap
  with offset ``o`` in a present dimension of extent ``N`` is out of bounds iff ``p < -o`` (o<0) or
  ``p >= N - o`` (o>0); out-of-bounds taps read the boundary value: ``constant`` -> the JSON value,
  ``shrink`` -> ``JUNK_VAL = -100000`` (``stencilflow/stencil/_common.py:8``).
* ``copy`` -- the CPU expansion raises ``NameError`` for it (``cpu.py:86-88``), so the Intel FPGA
  expansion is followed: the out-of-bounds tap reads the *centre* tap of the same field at the
  current cell (``stencilflow/stencil/intel_fpga.py:179-185,225-227``).
* dimensions -- ``stencilflow/kernel_chain_graph.py:382-405``: 1-D/2-D programs use the last
  iterators (``k`` / ``j,k``); inputs are indexed by their ``input_dims`` and broadcast along the rest.
* arithmetic types -- DaCe emits tasklet locals as ``auto`` and Python float literals as C++ double
  literals (``dace/dace/codegen/cppunparse.py:351-353,743-753``).  Consequently a float JSON boundary
  value turns a guarded tap into a double, ``0.25 * x`` is evaluated in double, and the result is
  narrowed once when stored.  Mirrored here with numpy scalar types (NEP 50 promotion).
* inputs -- ``stencilflow/helper.py:162-237``: ``constant:<v>`` arrays are allocated at the full
  program shape even for lower-dimensional inputs; embedded lists arrive flat.

PARITY PINNING: the reference stores no output vectors for this path (its program tests compare
FPGA emulation with the CPU SDFG generated from the same JSON, ``test/test_stencilflow.py:188-224``)
and its compiled paths cannot run under Python 3.12 (DaCe 0.10.8 needs <3.10).  This oracle is pinned by
(1) OUTPUTS OF THE REFERENCE ITSELF: its pure-Python dataflow simulator (``stencilflow/simulator.py``,
    ``kernel.py:634-738``, ``calculator.py``) does run here, unmodified, behind a stub for the DaCe type
    names; ``tests/golden/make_reference_sim_golden.py`` ran it on 10 program/input cases inside the
    simulator's envelope (3-D, constant boundaries, full-dimensional inputs) and
    ``tests/test_oracle.py::test_oracles_match_reference_simulator`` requires this module to reproduce
    the stored outputs (max relative error 2e-6 float32 / 1e-13 float64);
(2) hand-derived known answers for every program of ``test/stencils`` (``tests/golden/known_answers.json``),
    which also cover what the simulator cannot run (2-D programs, lower-dimensional and 0-D inputs);
(3) agreement with the independently generated C++/OpenMP restatement in ``reference_cpp.py``.
``shrink`` and ``copy`` boundaries have no executable reference here (the simulator raises
NotImplementedError for them, ``kernel.py:534-540``): for those the pinning is (2)+(3) only.
"""

import ast
import json
import math
import os
import re

import numpy as np

ITERATORS = ("i", "j", "k")
JUNK_VAL = -100000  # stencilflow/stencil/_common.py:8

_NP_TYPES = {
    "float32": np.float32, "float64": np.float64,
    "int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64,
    "uint8": np.uint8, "uint16": np.uint16, "uint32": np.uint32, "uint64": np.uint64,
    "bool": np.bool_,
}


def load_program(path_or_dict):
    if isinstance(path_or_dict, dict):
        prog = json.loads(json.dumps(path_or_dict, default=_json_default))
    else:
        with open(path_or_dict) as f:
            prog = json.load(f)
        prog.setdefault("path", os.path.dirname(os.path.abspath(path_or_dict)))
    return prog


def _json_default(o):
    # tolerate typeclass-like objects (anything with a ``name``) in already-parsed programs
    if hasattr(o, "to_string"):
        return o.to_string()
    if hasattr(o, "name"):
        return o.name
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError(type(o))


class ProgramInfo:
    """Shapes, iterators and evaluation order of a program."""

    def __init__(self, prog):
        dims = list(prog["dimensions"])
        self.ndims = len(dims)
        self.iterators = list(ITERATORS[3 - self.ndims:])
        self.shape = tuple(int(d) for d in dims)                 # JSON shape (1-3 D)
        self.extent = dict(zip(self.iterators, self.shape))
        self.inputs = prog["inputs"]
        self.program = prog["program"]
        self.outputs = list(prog["outputs"])
        self.constants = prog.get("constants", {})
        self.input_dims = {}
        for name, cfg in self.inputs.items():
            d = cfg.get("input_dims", cfg.get("dimensions", None))
            self.input_dims[name] = list(self.iterators) if d is None else list(d)
        self.order = self._topological_order()

    def field_dims(self, name):
        return self.input_dims[name] if name in self.inputs else list(self.iterators)

    def field_shape(self, name):
        return tuple(self.extent[d] for d in self.field_dims(name))

    def field_type(self, name):
        cfg = self.inputs[name] if name in self.inputs else self.program[name]
        return _NP_TYPES[cfg["data_type"]]

    def _reads(self, name):
        tree = ast.parse(self.program[name]["computation_string"].strip())
        return {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)}

    def _topological_order(self):
        deps = {k: sorted(n for n in self._reads(k) if n in self.program and n != k)
                for k in self.program}
        done, order, visiting = set(), [], set()

        def visit(k):
            if k in done:
                return
            if k in visiting:
                raise ValueError("Cycle detected: {}".format(sorted(visiting)))
            visiting.add(k)
            for d in deps[k]:
                visit(d)
            visiting.discard(k)
            done.add(k)
            order.append(k)

        for k in self.program:
            visit(k)
        return order


def materialize_inputs(prog, overrides=None):
    """Inputs as the reference driver would hand them to the program
    (``stencilflow/run_program.py:138-148``), except that arrays come back already
    in the input's own shape.  ``overrides`` replaces selected inputs."""
    info = ProgramInfo(prog)
    out = {}
    for name, cfg in info.inputs.items():
        if overrides and name in overrides:
            val = overrides[name]
        else:
            val = _load_input(cfg, info, prog.get("path"))
        out[name] = _conform(val, name, info)
    return out


def _load_input(cfg, info, prefix):
    data = cfg["data"]
    dtype = _NP_TYPES[cfg["data_type"]]
    dims = cfg.get("input_dims", cfg.get("dimensions", None))
    scalar = dims is not None and len(dims) == 0
    if isinstance(data, str):
        m = re.match(r"([^:]+):(.+)", data)
        if m and m.group(1) == "constant":
            v = float(m.group(2))
            return v if scalar else np.full(info.shape, v, dtype=dtype)
        path = data if os.path.isfile(data) or prefix is None else os.path.join(prefix, data)
        if path.endswith(".csv"):
            return np.genfromtxt(path, dtype, delimiter=",")
        if path.endswith(".dat"):
            return np.fromfile(path, dtype)
        raise ValueError("cannot load input: " + data)
    if scalar:
        return dtype(data)
    return np.array(data, dtype=dtype)


def _conform(val, name, info):
    dtype = info.field_type(name)
    shape = info.field_shape(name)
    if len(shape) == 0:
        return dtype(val)
    arr = np.asarray(val, dtype=dtype)
    n = int(np.prod(shape))
    if arr.shape == shape:
        return arr
    # full-shape or flat buffers: the program reads the first prod(shape) elements
    return np.ascontiguousarray(arr.ravel()[:n]).reshape(shape)


# ----------------------------------------------------------------------------- evaluation


def _offset(node):
    if isinstance(node, ast.Name):
        return node.id, 0
    if isinstance(node, ast.BinOp) and isinstance(node.left, ast.Name):
        r = node.right
        sign = 1
        if isinstance(r, ast.UnaryOp) and isinstance(r.op, ast.USub):
            r, sign = r.operand, -1
        if isinstance(r, ast.Constant) and isinstance(r.value, int):
            v = sign * r.value
            if isinstance(node.op, ast.Add):
                return node.left.id, v
            if isinstance(node.op, ast.Sub):
                return node.left.id, -v
    raise TypeError("Unrecognized offset: " + ast.unparse(node))


def _literal(value):
    """C++ literal typing: floats are doubles; ints stay weakly typed."""
    if isinstance(value, bool):
        return bool(value)
    if isinstance(value, int):
        return int(value)
    return np.float64(value)


_FUNCS = {
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "sinh": np.sinh, "cosh": np.cosh,
    "tanh": np.tanh, "sqrt": np.sqrt, "exp": np.exp, "log": np.log, "fabs": np.abs,
    "abs": np.abs, "floor": np.floor, "ceil": np.ceil, "min": np.minimum, "max": np.maximum,
    "pow": np.power,
}


class _Evaluator:
    def __init__(self, info, opname, fields, scalars):
        self.info = info
        self.op = opname
        self.fields = fields
        self.scalars = scalars
        self.locals = {}
        self.bcs = info.program[opname].get("boundary_conditions", {}) or {}
        self.full = info.shape

    def tap(self, node):
        field = node.value.id
        sl = node.slice
        elts = list(sl.elts) if isinstance(sl, ast.Tuple) else [sl]
        by_name = dict(_offset(e) for e in elts)
        dims = self.info.field_dims(field)
        offs = [by_name[d] for d in dims]
        arr = self.fields[field]
        shape = arr.shape
        # region of the *output* index space whose tap is in bounds
        dst, src, oob = [], [], False
        for n, o in zip(shape, offs):
            lo, hi = max(0, -o), min(n, n - o)
            if lo >= hi:
                oob = True
                lo, hi = 0, 0
            dst.append(slice(lo, hi))
            src.append(slice(lo + o, hi + o))
        guarded = any(o != 0 for o in offs)
        if not guarded:
            val = arr
        else:
            bc = self.bcs.get(field)
            if bc is None:
                raise KeyError("No boundary condition for {} in {}".format(field, self.op))
            kind = bc.get("type", bc.get("btype"))
            if kind == "constant":
                fill = _literal(bc["value"])
            elif kind == "shrink":
                fill = JUNK_VAL
            elif kind == "copy":
                fill = None
            else:
                raise ValueError("Unsupported boundary condition type: {}".format(kind))
            if fill is None:
                val = arr.copy()
            else:
                # (cond ? fill : tap) has the common C++ type of fill and the field
                rtype = np.result_type(fill, arr.dtype) if isinstance(fill, np.generic) else arr.dtype
                val = np.full(shape, fill, dtype=rtype)
            if not oob:
                val[tuple(dst)] = arr[tuple(src)]
        # broadcast along the dimensions the field does not have
        index = tuple(slice(None) if d in dims else None for d in self.info.iterators)
        return val[index]

    def name(self, node):
        n = node.id
        if n in self.locals:
            return self.locals[n]
        if n in self.scalars:
            return self.scalars[n]
        if n in self.fields and self.fields[n].ndim == 0:
            return self.fields[n]
        raise NameError("Unknown name {} in {}".format(n, self.op))

    def ev(self, node):
        if isinstance(node, ast.Constant):
            return _literal(node.value)
        if isinstance(node, ast.Name):
            return self.name(node)
        if isinstance(node, ast.Subscript):
            return self.tap(node)
        if isinstance(node, ast.BinOp):
            a, b = self.ev(node.left), self.ev(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            if isinstance(node.op, ast.Div):
                return a / b
            raise TypeError("Unsupported operator")
        if isinstance(node, ast.UnaryOp):
            v = self.ev(node.operand)
            if isinstance(node.op, ast.USub):
                return -v
            if isinstance(node.op, ast.UAdd):
                return v
            if isinstance(node.op, ast.Not):
                return np.logical_not(v)
            raise TypeError("Unsupported unary operator")
        if isinstance(node, ast.Compare):
            a, b = self.ev(node.left), self.ev(node.comparators[0])
            op = node.ops[0]
            table = {ast.Lt: np.less, ast.LtE: np.less_equal, ast.Gt: np.greater,
                     ast.GtE: np.greater_equal, ast.Eq: np.equal, ast.NotEq: np.not_equal}
            return table[type(op)](a, b)
        if isinstance(node, ast.BoolOp):
            vals = [self.ev(v) for v in node.values]
            fn = np.logical_and if isinstance(node.op, ast.And) else np.logical_or
            out = vals[0]
            for v in vals[1:]:
                out = fn(out, v)
            return out
        if isinstance(node, ast.IfExp):
            c, a, b = self.ev(node.test), self.ev(node.body), self.ev(node.orelse)
            return np.where(c, a, b)
        if isinstance(node, ast.Call):
            return _FUNCS[node.func.id](*[self.ev(a) for a in node.args])
        raise TypeError("Unsupported syntax: " + ast.dump(node))

    def run(self):
        tree = ast.parse(self.info.program[self.op]["computation_string"].strip())
        last = None
        for stmt in tree.body:
            if not isinstance(stmt, ast.Assign):
                continue
            last = stmt.targets[0].id
            self.locals[last] = self.ev(stmt.value)
        result = self.locals[self.op] if self.op in self.locals else self.locals[last]
        out_type = self.info.field_type(self.op)
        return np.ascontiguousarray(np.broadcast_to(np.asarray(result), self.full)).astype(out_type)


def run_reference(prog, inputs=None, keep_intermediates=False):
    """Evaluate the program.  ``inputs``: name -> ndarray / scalar (missing ones are
    materialised from the JSON).  Returns {output name: ndarray of the JSON shape}; with
    ``keep_intermediates`` every operator's field is returned."""
    prog = load_program(prog)
    info = ProgramInfo(prog)
    given = materialize_inputs(prog, inputs or {})
    fields, scalars = {}, {}
    for name, val in given.items():
        if len(info.field_shape(name)) == 0:
            scalars[name] = val
        else:
            fields[name] = val
    for name, c in info.constants.items():
        scalars[name] = _NP_TYPES[c["data_type"]](c["value"])
    for op in info.order:
        fields[op] = _Evaluator(info, op, fields, scalars).run()
    names = list(info.program) if keep_intermediates else info.outputs
    return {n: fields[n] for n in names}


def max_relative_error(reference, result):
    """``max |ref-res| / (max(|ref|,|res|) + eps)`` -- the quantity the parity tests bound."""
    reference = np.asarray(reference)
    result = np.asarray(result)
    if reference.size == 0:
        return 0.0
    eps = np.finfo(reference.dtype).eps
    den = np.maximum(np.abs(reference), np.abs(result)).astype(np.float64) + eps
    return float(np.max(np.abs(reference.astype(np.float64) - result.astype(np.float64)) / den))


def trim_halo(arr, halo):
    """``-halo`` slicing of ``stencilflow/run_program.py:202-209``: every axis loses ``halo`` cells
    on both sides."""
    if halo <= 0:
        return arr
    return arr[tuple(slice(halo, -halo) for _ in arr.shape)]
