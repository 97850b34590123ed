"""Typed expression IR for StencilFlow computation strings.

A program entry's ``computation_string`` is a ``;``-separated list of Python
assignments whose right-hand sides use ``+ - * /``, unary minus, comparisons,
``and``/``or``, ternaries, calls of a few math functions, numeric literals,
bare names (0-D inputs, program constants, earlier temporaries) and field
accesses ``f[i-1, j, k+2]`` whose indices are ``<iterator>`` or
``<iterator> +/- <int>`` (reference ``stencilflow/compute_graph.py:81-110,203-326``,
``stencilflow/compute_graph_nodes.py:189-238``).

The reference walks the pre-3.9 ``ast`` shape (``node.slice.value``) and rewrites
the strings textually for 1-D/2-D programs; this module parses with the current
``ast`` and resolves every index by the *name* of its iterator, which yields the
same ``[di, dj, dk]`` offsets (``None`` for dimensions a field does not have).
"""

import ast
from typing import Callable, Dict, List, Optional, Sequence, Tuple

ITERATORS = ("i", "j", "k")

# name in a computation string -> (C/CUDA spelling, arity).  Latency classes for
# the FPGA model live in compute_graph.config.
FUNCTIONS = {
    "sin": ("sin", 1), "cos": ("cos", 1), "tan": ("tan", 1),
    "sinh": ("sinh", 1), "cosh": ("cosh", 1), "tanh": ("tanh", 1),
    "sqrt": ("sqrt", 1), "exp": ("exp", 1), "log": ("log", 1),
    "fabs": ("fabs", 1), "abs": ("fabs", 1), "floor": ("floor", 1), "ceil": ("ceil", 1),
    "min": ("fmin", 2), "max": ("fmax", 2), "pow": ("pow", 2),
}

_BINOPS = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
_CMPOPS = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==", ast.NotEq: "!="}


class Expr:
    __slots__ = ()

    def children(self) -> Sequence["Expr"]:
        return ()


class Const(Expr):
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value

    @property
    def name(self):
        return self.value

    def __repr__(self):
        return repr(self.value)


class Tap(Expr):
    """Field access; ``offset`` is a 3-tuple over (i, j, k), ``None`` = dimension absent."""
    __slots__ = ("field", "offset")

    def __init__(self, field: str, offset: Tuple[Optional[int], ...]):
        self.field = field
        self.offset = tuple(offset)

    @property
    def name(self):
        return self.field

    @property
    def index(self):
        return list(self.offset)

    def __repr__(self):
        return "{}{}".format(self.field, list(self.offset))


class Var(Expr):
    """Bare name: a 0-D input, a program constant or a cell-local temporary."""
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


class Bin(Expr):
    __slots__ = ("op", "a", "b")

    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def children(self):
        return (self.a, self.b)

    def __repr__(self):
        return "({} {} {})".format(self.a, self.op, self.b)


class Neg(Expr):
    __slots__ = ("a",)

    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)

    def __repr__(self):
        return "(-{})".format(self.a)


class Cmp(Expr):
    __slots__ = ("op", "a", "b")

    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def children(self):
        return (self.a, self.b)

    def __repr__(self):
        return "({} {} {})".format(self.a, self.op, self.b)


class Logic(Expr):
    """``and`` / ``or`` / ``not`` over boolean sub-expressions."""
    __slots__ = ("op", "args")

    def __init__(self, op, args):
        self.op, self.args = op, list(args)

    def children(self):
        return self.args

    def __repr__(self):
        if self.op == "not":
            return "(not {})".format(self.args[0])
        return "(" + " {} ".format(self.op).join(map(repr, self.args)) + ")"


class Select(Expr):
    __slots__ = ("cond", "a", "b")

    def __init__(self, cond, a, b):
        self.cond, self.a, self.b = cond, a, b

    def children(self):
        return (self.cond, self.a, self.b)

    def __repr__(self):
        return "({} if {} else {})".format(self.a, self.cond, self.b)


class Call(Expr):
    __slots__ = ("fn", "args")

    def __init__(self, fn, args):
        self.fn, self.args = fn, list(args)

    def children(self):
        return self.args

    def __repr__(self):
        return "{}({})".format(self.fn, ", ".join(map(repr, self.args)))


class Statement:
    __slots__ = ("target", "value")

    def __init__(self, target: str, value: Expr):
        self.target, self.value = target, value

    def __repr__(self):
        return "{} = {}".format(self.target, self.value)


def walk(e: Expr):
    yield e
    for c in e.children():
        yield from walk(c)


def _index_offset(node: ast.AST) -> Tuple[str, int]:
    """``k`` -> ("k", 0); ``j - 3`` -> ("j", -3)."""
    if isinstance(node, ast.Name):
        return node.id, 0
    if isinstance(node, ast.BinOp) and isinstance(node.left, ast.Name) \
            and isinstance(node.op, (ast.Add, ast.Sub)):
        rhs = node.right
        sign = 1
        if isinstance(rhs, ast.UnaryOp) and isinstance(rhs.op, ast.USub):
            rhs, sign = rhs.operand, -1
        if isinstance(rhs, ast.Constant) and isinstance(rhs.value, int):
            val = sign * int(rhs.value)
            return node.left.id, val if isinstance(node.op, ast.Add) else -val
    raise TypeError("Unrecognized offset: {}".format(ast.unparse(node)))


class _Builder:
    def __init__(self, field_dims: Dict[str, Optional[List[str]]], default_dims: List[str]):
        self.field_dims = field_dims
        self.default_dims = list(default_dims)

    def subscript(self, node: ast.Subscript) -> Tap:
        if not isinstance(node.value, ast.Name):
            raise TypeError("Only subscripts of variables are supported")
        field = node.value.id
        sl = node.slice
        elts = list(sl.elts) if isinstance(sl, ast.Tuple) else [sl]
        by_name = {}
        for e in elts:
            it, off = _index_offset(e)
            if it not in ITERATORS:
                raise TypeError("Unknown iterator '{}' in access to {}".format(it, field))
            by_name[it] = off
        dims = self.field_dims.get(field)
        if dims is None:
            dims = self.default_dims
        # A leading iterator the program does not iterate over (size-1 padding,
        # e.g. "a[i, j, k]" in a program declared with dimensions [1, N, M]) is dropped,
        # like the reference prunes indices (compute_graph_nodes.py:223-226).
        missing = [d for d in dims if d not in by_name]
        if missing:
            raise KeyError("Access {} lacks iterator(s) {} of field '{}'".format(
                ast.unparse(node), missing, field))
        return Tap(field, tuple(by_name[d] if d in dims else None for d in ITERATORS))

    def expr(self, node: ast.AST) -> Expr:
        if isinstance(node, ast.Constant):
            if isinstance(node.value, (int, float, bool)):
                return Const(node.value)
            raise TypeError("Unsupported literal: {!r}".format(node.value))
        if isinstance(node, ast.Name):
            return Var(node.id)
        if isinstance(node, ast.Subscript):
            return self.subscript(node)
        if isinstance(node, ast.BinOp):
            if type(node.op) not in _BINOPS:
                raise TypeError("Unsupported operator: {}".format(type(node.op).__name__))
            return Bin(_BINOPS[type(node.op)], self.expr(node.left), self.expr(node.right))
        if isinstance(node, ast.UnaryOp):
            if isinstance(node.op, ast.USub):
                return Neg(self.expr(node.operand))
            if isinstance(node.op, ast.UAdd):
                return self.expr(node.operand)
            if isinstance(node.op, ast.Not):
                return Logic("not", [self.expr(node.operand)])
            raise TypeError("Unsupported unary operator: {}".format(type(node.op).__name__))
        if isinstance(node, ast.Compare):
            if len(node.ops) != 1 or type(node.ops[0]) not in _CMPOPS:
                raise TypeError("Unsupported comparison: {}".format(ast.unparse(node)))
            return Cmp(_CMPOPS[type(node.ops[0])], self.expr(node.left),
                       self.expr(node.comparators[0]))
        if isinstance(node, ast.BoolOp):
            op = "and" if isinstance(node.op, ast.And) else "or"
            return Logic(op, [self.expr(v) for v in node.values])
        if isinstance(node, ast.IfExp):
            return Select(self.expr(node.test), self.expr(node.body), self.expr(node.orelse))
        if isinstance(node, ast.Call):
            if not isinstance(node.func, ast.Name) or node.func.id not in FUNCTIONS:
                raise TypeError("Unsupported function: {}".format(ast.unparse(node.func)))
            fn = node.func.id
            if len(node.args) != FUNCTIONS[fn][1]:
                raise TypeError("{} expects {} argument(s)".format(fn, FUNCTIONS[fn][1]))
            return Call(fn, [self.expr(a) for a in node.args])
        raise Exception("Unknown AST type {}".format(type(node)))


def parse_computation(computation_string: str,
                      field_dims: Dict[str, Optional[List[str]]],
                      default_dims: List[str]) -> List[Statement]:
    """Parse a computation string into assignments.

    ``field_dims[name]`` lists the iterators a program *input* is indexed by
    (``[]`` for 0-D); every other field has ``default_dims`` (the program's own
    iterators).  Iterators appearing in a subscript but not in the field's
    dims are ignored, so 3-D style strings work in padded ``[1, N, M]`` programs.
    """
    tree = ast.parse(computation_string.strip())
    builder = _Builder(field_dims, default_dims)
    stmts = []
    for node in tree.body:
        if isinstance(node, ast.Assign):
            if len(node.targets) != 1 or not isinstance(node.targets[0], ast.Name):
                raise TypeError("Only simple assignments are supported")
            stmts.append(Statement(node.targets[0].id, builder.expr(node.value)))
        elif isinstance(node, ast.Expr):
            continue  # expression statements carry no data flow (compute_graph.py:210-219)
        else:
            raise Exception("Unknown AST type {}".format(type(node)))
    if not stmts:
        raise ValueError("Computation string has no assignment: " + computation_string)
    return stmts


def collect_taps(stmts: Sequence[Statement]) -> Dict[str, List[Tuple[Optional[int], ...]]]:
    """field -> distinct offsets in order of first appearance."""
    taps: Dict[str, List[Tuple[Optional[int], ...]]] = {}
    for s in stmts:
        for e in walk(s.value):
            if isinstance(e, Tap):
                lst = taps.setdefault(e.field, [])
                if e.offset not in lst:
                    lst.append(e.offset)
    return taps


def collect_vars(stmts: Sequence[Statement]) -> List[str]:
    """Bare names read before (or without) being assigned in the same string."""
    assigned, free = set(), []
    for s in stmts:
        for e in walk(s.value):
            if isinstance(e, Var) and e.name not in assigned and e.name not in free:
                free.append(e.name)
        assigned.add(s.target)
    return free


_PREC = {"or": 1, "and": 2, "not": 3, "cmp": 4, "+": 5, "-": 5, "*": 6, "/": 6, "neg": 7}


def emit_c(e: Expr,
           tap: Callable[[Tap], str],
           var: Callable[[str], str],
           literal: Callable[[object], str],
           call: Callable[[str, List[str]], str]) -> str:
    """Render ``e`` as a fully parenthesised C expression.  The four callbacks decide
    how taps, names, literals and calls are spelled, which is where the CUDA lowering
    and any other C-family back end differ."""
    def go(x: Expr) -> str:
        if isinstance(x, Const):
            return literal(x.value)
        if isinstance(x, Tap):
            return tap(x)
        if isinstance(x, Var):
            return var(x.name)
        if isinstance(x, Bin):
            return "({} {} {})".format(go(x.a), x.op, go(x.b))
        if isinstance(x, Neg):
            return "(-{})".format(go(x.a))
        if isinstance(x, Cmp):
            return "({} {} {})".format(go(x.a), x.op, go(x.b))
        if isinstance(x, Logic):
            if x.op == "not":
                return "(!{})".format(go(x.args[0]))
            return "(" + (" && " if x.op == "and" else " || ").join(go(a) for a in x.args) + ")"
        if isinstance(x, Select):
            return "({} ? {} : {})".format(go(x.cond), go(x.a), go(x.b))
        if isinstance(x, Call):
            return call(x.fn, [go(a) for a in x.args])
        raise TypeError(type(x))
    return go(e)


def to_source(e: Expr, tap: Optional[Callable[[Tap], str]] = None) -> str:
    """Python-syntax rendering (used for reports and relative-access strings)."""
    tap = tap or (lambda t: "{}[{}]".format(
        t.field, ", ".join("{}{}".format(it, "" if o == 0 else "{:+d}".format(o))
                           for it, o in zip(ITERATORS, t.offset) if o is not None)))

    def go(x: Expr) -> str:
        if isinstance(x, Const):
            return repr(x.value)
        if isinstance(x, Tap):
            return tap(x)
        if isinstance(x, Var):
            return x.name
        if isinstance(x, Bin):
            return "({} {} {})".format(go(x.a), x.op, go(x.b))
        if isinstance(x, Neg):
            return "(-{})".format(go(x.a))
        if isinstance(x, Cmp):
            return "({} {} {})".format(go(x.a), x.op, go(x.b))
        if isinstance(x, Logic):
            if x.op == "not":
                return "(not {})".format(go(x.args[0]))
            return "(" + " {} ".format(x.op).join(go(a) for a in x.args) + ")"
        if isinstance(x, Select):
            return "({} if {} else {})".format(go(x.a), go(x.cond), go(x.b))
        if isinstance(x, Call):
            return "{}({})".format(x.fn, ", ".join(go(a) for a in x.args))
        raise TypeError(type(x))
    return go(e)


def pair_odd_taps(e: Expr, balance: bool = False) -> Expr:
    """Re-associates sums so that the taps with an odd innermost offset are added to each other first:
    ``((((a[i-1] + a[i+1]) + a[j-1]) + a[j+1]) + a[k-1]) + a[k+1]`` becomes
    ``(((a[i-1] + a[i+1]) + a[j-1]) + a[j+1]) + (a[k-1] + a[k+1])``.

    Why: the float32 kernels compute on aligned *pairs* of neighbouring k-cells (``add.f32x2``); a tap with
    an odd k-offset straddles two pairs, so adding it to a packed running sum costs two scalar additions.
    Adding the straddling taps to each other first costs two scalar additions per tap but one, and a single
    packed addition for the lot -- one issue slot in eight less for a 7-point stencil.  Both kernel
    families are lowered from the rewritten tree, so fused and one-operator results stay bit-identical;
    against the reference the change is an admissible re-association (its CPU program is built with
    ``-ffast-math``, ``dace/dace/config_schema.yml:272``) well inside the 1e-5 tolerance.

    ``balance``: the sums are built as balanced trees instead of left-to-right chains -- a shorter dependent
    chain per cell (``(a + b) + (c + d)``: two additions deep instead of three), which is what a
    latency-bound float64 kernel with two warps per scheduler needs."""
    if isinstance(e, Bin):
        if e.op == "+":
            terms = []
            node = e
            while isinstance(node, Bin) and node.op == "+":
                terms.append(node.b)
                node = node.a
            terms.append(node)
            terms = [pair_odd_taps(t, balance) for t in reversed(terms)]
            odd = [t for t in terms if isinstance(t, Tap) and t.offset[2] is not None and t.offset[2] % 2 != 0]

            def chain(ts):
                if balance and len(ts) > 2:
                    half = (len(ts) + 1) // 2
                    return Bin("+", chain(ts[:half]), chain(ts[half:]))
                acc = ts[0]
                for t in ts[1:]:
                    acc = Bin("+", acc, t)
                return acc

            if len(odd) >= 2 and len(odd) < len(terms):
                rest = [t for t in terms if not any(t is o for o in odd)]
                return Bin("+", chain(rest), chain(odd))
            return chain(terms)
        return Bin(e.op, pair_odd_taps(e.a, balance), pair_odd_taps(e.b, balance))
    if isinstance(e, Neg):
        return Neg(pair_odd_taps(e.a, balance))
    if isinstance(e, Cmp):
        return Cmp(e.op, pair_odd_taps(e.a, balance), pair_odd_taps(e.b, balance))
    if isinstance(e, Logic):
        return Logic(e.op, [pair_odd_taps(a, balance) for a in e.args])
    if isinstance(e, Select):
        return Select(pair_odd_taps(e.cond, balance), pair_odd_taps(e.a, balance), pair_odd_taps(e.b, balance))
    if isinstance(e, Call):
        return Call(e.fn, [pair_odd_taps(a, balance) for a in e.args])
    return e
