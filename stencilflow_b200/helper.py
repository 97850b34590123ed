"""Front-end utilities shared by every stage of the path: JSON loading, array I/O,
index arithmetic and the result comparison.

Public names and argument meaning follow reference ``stencilflow/helper.py`` so that
callers (and the reference's own unit tests, ``test/test_stencilflow.py:114-162``)
read the same.  Differences are deliberate and listed where they occur.
"""

import ast
import collections
import functools
import json
import operator
import os
import re

import numpy as np

from . import dtypes

ITERATORS = ["i", "j", "k"]

_PACKAGE_DIR = os.path.dirname(os.path.realpath(__file__))


def str_to_dtype(dtype_str):
    """JSON ``data_type`` string -> typeclass (reference helper.py:47-59)."""
    if not isinstance(dtype_str, str):
        raise TypeError("Expected string, got: " + type(dtype_str).__name__)
    return dtypes.from_string(dtype_str)


def _convert_data_types(node):
    for key, val in node.items():
        if isinstance(val, dict):
            _convert_data_types(val)
        elif key == "data_type" and isinstance(val, str):
            node[key] = str_to_dtype(val)


def parse_json(config_path):
    """Load a program (or ``*.config``) file.  Adds ``"path"`` (directory of the
    file) and turns every ``"data_type"`` string into a typeclass
    (reference helper.py:62-92).  Paths that do not exist are retried relative
    to the package directory, which is how the ``*.config`` files are found."""
    if not os.path.isfile(config_path):
        packaged = os.path.join(_PACKAGE_DIR, config_path)
        if not os.path.isfile(packaged):
            raise RuntimeError("file {} does not exists.".format(config_path))
        config_path = packaged
    with open(config_path, "r") as handle:
        config = json.load(handle)
    config["path"] = os.path.dirname(os.path.abspath(config_path))
    _convert_data_types(config)
    return config


def max_dict_entry_key(dict1):
    if not isinstance(dict1, dict):
        raise Exception("dict1 should be of type {}, but is of type {}".format(dict, type(dict1)))
    return max(dict1, key=dict1.get)


def _check_list(name, value):
    if not isinstance(value, list):
        raise Exception("{} should be of type {}, but is of type {}".format(name, list, type(value)))


def list_add_cwise(list1, list2):
    _check_list("list1", list1)
    _check_list("list2", list2)
    return [a + b for a, b in zip(list1, list2)]


def list_subtract_cwise(list1, list2):
    """Component-wise difference; a ``None`` (absent dimension) on either side
    stays ``None`` (reference helper.py:130-144)."""
    _check_list("list1", list1)
    _check_list("list2", list2)
    return [a - b if a is not None and b is not None else None for a, b in zip(list1, list2)]


def dim_to_abs_val(input, dimensions):
    """Row-major flattening of an index/extent vector (reference helper.py:147-159)."""
    strides = [functools.reduce(operator.mul, dimensions[d + 1:], 1) for d in range(len(dimensions))]
    return sum(x * s for x, s in zip(input, strides))


def num_dims(index):
    return sum(1 for x in index if x is not None)


def convert_3d_to_1d(dimensions, index):
    """Flatten a 3-entry extent that may contain ``None`` for absent dimensions
    (reference helper.py:293-314)."""
    if not index:
        return 0
    present = num_dims(index)
    if present == 3:
        return dim_to_abs_val(index, dimensions)
    if present == 2:
        if index[0] is None:
            return index[1] * dimensions[2] + index[2]
        if index[1] is None:
            return index[0] * dimensions[2] + index[2]
        return index[0] * dimensions[1] + index[1]
    if present == 1:
        return [x for x in index if x is not None][0]
    return 0


def aligned(a, alignment=16):
    """Return ``a`` or an equal copy whose data pointer is ``alignment``-byte aligned."""
    if a.ctypes.data % alignment == 0:
        return a
    extra = alignment // a.itemsize + 1
    buf = np.empty(a.size + extra, dtype=a.dtype)
    ofs = (-buf.ctypes.data % alignment) // a.itemsize
    view = buf[ofs:ofs + a.size].reshape(a.shape)
    np.copyto(view, a)
    assert view.ctypes.data % alignment == 0
    return view


def _is_scalar_input(input_config):
    dims = input_config.get("input_dims", None)
    return dims is not None and len(dims) == 0


def load_array(input_config, prefix=None, shape=None):
    """Materialise one program input (reference helper.py:162-217).

    ``data`` may be ``"constant:<v>"`` (filled at the full program ``shape``; a
    0-D input yields the Python float), a ``.csv``/``.dat`` path (``.dat`` is raw
    ``ndarray.tofile`` content, returned flat), or an embedded list (returned
    flat, not reshaped).  The reference's ``"random:"`` branch cannot execute
    (it references undefined names, helper.py:189-196); here ``"random:<lo>,<hi>"``
    is implemented with a fixed seed so that runs are reproducible.
    """
    data = input_config["data"]
    dtype = input_config["data_type"].type
    if isinstance(data, str):
        m = re.match(r"([^:]+):(.+)", data)
        if m and not os.path.isfile(data):
            kind, arg = m.group(1), m.group(2)
            scalar = _is_scalar_input(input_config)
            if shape is None and not scalar:
                raise ValueError("Must provide shape when using generated inputs")
            if kind == "constant":
                val = float(arg)
                if scalar:
                    return val
                arr = np.empty(shape, dtype=dtype)
                arr[:] = val
                return arr
            if kind == "random":
                lo, hi = (float(x) for x in re.split(r"[,;:\s]+", arg.strip())[:2])
                rng = np.random.default_rng(1234)
                if scalar:
                    return float(lo + (hi - lo) * rng.random())
                return (lo + (hi - lo) * rng.random(size=tuple(shape))).astype(dtype)
            raise ValueError("Unknown generation: " + kind)
        path = data
        if not os.path.isfile(path) and prefix is not None:
            path = os.path.join(prefix, data)
        if not os.path.isfile(path):
            raise FileNotFoundError("File {} does not exists.".format(data))
        if path.endswith(".csv"):
            return np.genfromtxt(path, dtype, delimiter=",")
        if path.endswith(".dat"):
            return np.fromfile(path, dtype)
        raise ValueError("Invalid file type: " + path)
    if _is_scalar_input(input_config) or (shape is not None and len(shape) == 0):
        return dtype(data)
    if isinstance(data, np.ndarray):
        return data
    return np.array(data, dtype=dtype)


def load_input_arrays(input_configs, prefix=None, shape=None):
    """Load all inputs of a program; arrays come back 64-byte aligned
    (reference helper.py:220-237)."""
    arrays = {}
    for name, source in input_configs.items():
        arr = load_array(source, prefix, shape)
        if isinstance(arr, np.ndarray) and arr.ndim > 0:
            arr = aligned(arr, 64)
        arrays[name] = arr
    return arrays


def save_array(array, path):
    array.tofile(path)


def save_output_arrays(outputs, output_dir=str()):
    """``<output_dir>/<name>.dat`` as raw bytes (reference helper.py:249-258)."""
    for name, data in outputs.items():
        save_array(data, os.path.join(output_dir, name + ".dat"))


def relative_difference(reference, result):
    """Element-wise ``|ref-res| / (max(|ref|,|res|) + eps)``."""
    reference = np.asarray(reference)
    result = np.asarray(result)
    eps = np.finfo(reference.dtype).eps if np.issubdtype(reference.dtype, np.floating) else 0
    scale = np.maximum(np.abs(reference), np.abs(result)) + eps
    return np.abs(reference - result) / scale


def arrays_are_equal(reference, result, tolerance=1e-5):
    """True iff the maximum relative difference is within ``tolerance``.

    Same signature and default as reference helper.py:261-276.  The reference
    divides by the *signed* element-wise maximum, which lets any pair of
    negative numbers pass; here the divisor uses magnitudes, i.e. the check is
    strictly tighter.  NaNs never compare equal.
    """
    if not isinstance(reference, np.ndarray):
        reference = load_array(reference)
    if not isinstance(result, np.ndarray):
        result = load_array(result)
    if reference.shape != result.shape:
        return False
    if reference.size == 0:
        return True
    return bool(np.all(relative_difference(reference, result) <= tolerance))


def unique(iterable):
    """Drop duplicates, keep first occurrences, keep the container type."""
    try:
        seen = []
        for x in iterable:
            if x not in seen:
                seen.append(x)
        return type(iterable)(seen)
    except TypeError:
        return type(iterable)(collections.OrderedDict(zip(map(str, iterable), iterable)).values())


class OpCounter(ast.NodeVisitor):
    """Counts arithmetic operations of a computation string the way reference
    helper.py:341-365 does: a binary operator counts when at least one operand
    is a field access or another binary operation; every call counts."""

    def __init__(self):
        self._operation_count = {}

    @property
    def operation_count(self):
        return self._operation_count

    def _bump(self, name):
        self._operation_count[name] = self._operation_count.get(name, 0) + 1

    def visit_BinOp(self, node):
        if any(isinstance(side, (ast.Subscript, ast.BinOp)) for side in (node.left, node.right)):
            self._bump(type(node.op).__name__)
        self.generic_visit(node)

    def visit_Call(self, node):
        self._bump(node.func.id)
        self.generic_visit(node)
