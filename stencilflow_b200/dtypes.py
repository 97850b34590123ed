"""Minimal stand-in for the slice of ``dace.dtypes`` the StencilFlow front end touches.

The reference converts every ``"data_type"`` string of a program into a DaCe
typeclass (reference ``stencilflow/helper.py:47-59``) and afterwards only uses
``.type`` (numpy scalar type), ``.bytes``, ``.ctype``, ``__call__`` and an
``isinstance(..., typeclass)`` check (reference ``stencilflow/base_node_class.py:69-72``,
``stencilflow/kernel_chain_graph.py:761-767``).  DaCe itself is not a dependency of
this package, so those few members are provided here.
"""

import numpy as np


class typeclass:
    """A named scalar type: numpy type + C spelling + size in bytes."""

    def __init__(self, name: str, np_type, ctype: str):
        self.name = name
        self.type = np_type
        self.ctype = ctype
        self.bytes = np.dtype(np_type).itemsize
        self.dtype = self  # DaCe typeclasses are their own ``dtype``

    def __call__(self, value):
        return self.type(value)

    def to_string(self) -> str:
        return self.name

    def as_numpy_dtype(self):
        return np.dtype(self.type)

    @property
    def is_float(self) -> bool:
        return np.issubdtype(self.type, np.floating)

    def __repr__(self):
        return self.name

    def __eq__(self, other):
        return isinstance(other, typeclass) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


bool_ = typeclass("bool", np.bool_, "bool")
int8 = typeclass("int8", np.int8, "signed char")
int16 = typeclass("int16", np.int16, "short")
int32 = typeclass("int32", np.int32, "int")
int64 = typeclass("int64", np.int64, "long long")
uint8 = typeclass("uint8", np.uint8, "unsigned char")
uint16 = typeclass("uint16", np.uint16, "unsigned short")
uint32 = typeclass("uint32", np.uint32, "unsigned int")
uint64 = typeclass("uint64", np.uint64, "unsigned long long")
float32 = typeclass("float32", np.float32, "float")
float64 = typeclass("float64", np.float64, "double")

_BY_NAME = {
    t.name: t
    for t in (bool_, int8, int16, int32, int64, uint8, uint16, uint32, uint64,
              float32, float64)
}


def from_string(name: str) -> typeclass:
    """Resolve a JSON ``data_type`` string; unknown names raise AttributeError
    exactly like ``getattr(dace.dtypes, name)`` would (reference helper.py:54-59)."""
    try:
        return _BY_NAME[name]
    except KeyError:
        raise AttributeError("Unsupported data type: " + str(name))


def from_numpy(np_type) -> typeclass:
    dt = np.dtype(np_type)
    for t in _BY_NAME.values():
        if np.dtype(t.type) == dt:
            return t
    raise AttributeError("Unsupported data type: " + str(np_type))
