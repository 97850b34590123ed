"""Console verbosity levels (mirrors reference ``stencilflow/log_level.py:15-24``)."""

import enum
import functools


@functools.total_ordering
class LogLevel(enum.Enum):
    NO_LOG = 0
    BASIC = 1
    MODERATE = 2
    FULL = 3

    def __lt__(self, other):
        if isinstance(other, LogLevel):
            return self.value < other.value
        if isinstance(other, int):
            return self.value < other
        return NotImplemented
