"""Synthetic field generator: host mirror of ``sfb_fill_hash`` (csrc/sfb200_runtime.cu).

Large benchmark fields are generated directly in HBM; the same counter-based hash evaluated with
numpy gives bit-identical values on the host, so reduced-size parity runs and full-size device runs
use one generator (U[0,1) from a fixed-seed counter hash).
"""

import numpy as np


def _mix32(x):
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def hash_unit(idx, seed):
    """u in [0,1) with 24 random bits for uint64 indices ``idx``."""
    idx = np.asarray(idx, dtype=np.uint64)
    seed = np.uint64(seed)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        s = _mix32(np.uint32(seed & np.uint64(0xFFFFFFFF)) + np.uint32(0x9e3779b9) * hi +
                   np.uint32(0x85ebca6b) * np.uint32(seed >> np.uint64(32)))
        h = _mix32(lo ^ s)
    return (h >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def fill_hash(shape, dtype, seed, lo=0.0, hi=1.0, index_offset=0):
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64) + np.uint64(index_offset)
    vals = np.float64(lo) + np.float64(hi - lo) * hash_unit(idx, seed)
    return vals.astype(dtype).reshape(shape)
