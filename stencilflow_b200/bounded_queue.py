"""Fixed-capacity FIFO used to describe delay and sliding-window buffers.

API and error behaviour follow reference ``stencilflow/bounded_queue.py`` as pinned
by its unit tests (reference ``test/test_stencilflow.py:17-84``): a queue has at
least capacity 1, over/underflow raise ``RuntimeError``, the ``try_*`` variants
return ``False`` instead.  In this backend the queues are analysis artefacts only
(their ``maxsize`` drives the shared-memory ring planner); no data flows through
them at run time.
"""

import collections

import numpy as np


class BoundedQueue:
    def __init__(self, name, maxsize, swap_out=False, collection=(), verbose=False):
        self.name = name
        self.maxsize = maxsize if maxsize > 0 else 1
        self.swap_out = swap_out
        self.verbose = verbose
        # newest element sits at index 0, the next one to leave at index size-1
        self.queue = collections.deque(collection, self.maxsize)
        self.current_size = len(self.queue)

    def __repr__(self):
        return "BoundedQueue: {}, current size: {}, max size: {}".format(
            self.name, self.current_size, self.maxsize)

    __str__ = __repr__

    def import_data(self, data):
        if len(data) > self.maxsize:
            raise RuntimeError(
                "max size of queue ({}) is smaller than the data collection size ({})".format(
                    self.maxsize, len(data)))
        self.queue = collections.deque(data, self.maxsize)
        self.current_size = len(data)

    def export_data(self):
        return np.array(self.queue)[::-1]

    def size(self):
        return self.current_size

    def is_empty(self):
        return self.current_size == 0

    def is_full(self):
        return self.current_size == self.maxsize

    def enqueue(self, item):
        if not self.try_enqueue(item):
            raise RuntimeError("buffer {} overflow occurred".format(self.name))

    def try_enqueue(self, item):
        if self.current_size >= self.maxsize:
            return False
        self.queue.appendleft(item)
        self.current_size += 1
        return True

    def dequeue(self):
        if self.current_size == 0:
            raise RuntimeError("buffer {} underflow occurred".format(self.name))
        self.current_size -= 1
        return self.queue.pop()

    def try_dequeue(self):
        if self.current_size == 0:
            return False
        self.current_size -= 1
        return self.queue.pop()

    def peek(self, index):
        if index >= self.current_size:
            raise RuntimeError(
                "buffer {} index out of bound access occurred".format(self.name))
        return self.queue[index]

    def try_peek_last(self):
        if self.current_size == 0:
            return False
        return self.queue[self.current_size - 1]
