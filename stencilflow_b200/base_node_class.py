"""Common state of the three node kinds of the program DAG (Input, Kernel, Output).

Field names follow reference ``stencilflow/base_node_class.py:45-90`` because
``KernelChainGraph.compute_delay_buffer``/``add_channels`` and external callers
address them by name.
"""

from . import dtypes
from .bounded_queue import BoundedQueue


class BaseKernelNodeClass:
    def __init__(self, name, data_queue, data_type, verbose=False):
        if not isinstance(data_type, dtypes.typeclass):
            raise TypeError("Expected dtypes.typeclass, got: " + type(data_type).__name__)
        self.name = name
        self.data_queue = data_queue
        self.data_type = data_type
        self.verbose = verbose
        self.input_paths = {}     # program input -> [[di, dj, dk, predecessor name], ...]
        self.inputs = {}          # predecessor name -> channel
        self.outputs = {}         # successor name -> channel
        self.delay_buffer = {}    # predecessor name -> BoundedQueue
        self.program_counter = 0

    def generate_label(self):
        return self.name

    def __repr__(self):
        return "{}({})".format(type(self).__name__, self.name)


class Input(BaseKernelNodeClass):
    """A program input array (or 0-D value)."""

    def __init__(self, name, data_type, data_queue=None):
        super().__init__(name, data_queue if data_queue is not None else BoundedQueue(name, 1),
                         data_type)
        self.queues = {}
        self.dimension_size = self.data_queue.maxsize


class Output(BaseKernelNodeClass):
    """A program output: the sink an operator of the same name writes to."""

    def __init__(self, name, data_type, dimensions, data_queue=None):
        super().__init__(name, data_queue if data_queue is not None else BoundedQueue("dummy", 0),
                         data_type)
        self.dimensions = dimensions
