"""Pinned (page-locked) host buffers for result delivery.

Results are copied device->host straight into page-locked memory (full PCIe rate, no driver-side
staging) and handed to the caller as ordinary numpy arrays that own that memory: when the last
view of an array is garbage-collected the buffer goes back to a free list.  The amount of pinned
memory held by live results is capped; beyond the cap plain pageable arrays are returned.
"""

from __future__ import annotations

import ctypes as C
import math
import os
import threading
import weakref
from collections import defaultdict

import numpy as np

from . import _lib

MAX_OUTSTANDING = int(os.environ.get("BLDFM_B200_PINNED_MAX", str(4 << 30)))
_GRAIN = 1 << 16


class _Pool:
    def __init__(self):
        self.free = defaultdict(list)
        self.outstanding = 0
        self.lock = threading.Lock()
        self.pid = os.getpid()
        self._last_ptr = None

    def _give_back(self, nbytes, ptr, pid):
        if pid != os.getpid():
            return
        with self.lock:
            self.outstanding -= nbytes
            self.free[nbytes].append(ptr)

    def empty2(self, shape, dtype):
        """(array, pinned): like ``empty`` but also tells whether the array really is page-locked."""
        a = self.empty(shape, dtype)
        return a, a.base is not None

    def empty3(self, shape, dtype):
        """(array, pinned, address): the address spares the caller ``array.ctypes.data`` (microseconds)."""
        self._last_ptr = None
        a = self.empty(shape, dtype)
        ptr = self._last_ptr
        if ptr is None:
            return a, False, a.ctypes.data
        return a, True, ptr

    def empty(self, shape, dtype):
        """np.empty(shape, dtype) on pinned memory when possible."""
        dtype = np.dtype(dtype)
        count = math.prod(shape)
        nbytes = max(_GRAIN, -(-count * dtype.itemsize // _GRAIN) * _GRAIN)
        with self.lock:
            if os.getpid() != self.pid:          # forked child: the parent's pins are not ours
                self.free.clear()
                self.outstanding = 0
                self.pid = os.getpid()
            if self.outstanding + nbytes > MAX_OUTSTANDING:
                return np.empty(shape, dtype)
            lst = self.free[nbytes]
            ptr = lst.pop() if lst else None
        if ptr is None:
            p = C.c_void_p()
            if _lib.lib().bldfm_host_alloc(nbytes, C.byref(p)) != _lib.OK:
                return np.empty(shape, dtype)
            ptr = p.value
        with self.lock:
            self.outstanding += nbytes
        raw = (C.c_char * nbytes).from_address(ptr)
        weakref.finalize(raw, self._give_back, nbytes, ptr, os.getpid())
        self._last_ptr = ptr
        return np.ndarray(shape, dtype, raw)

    def trim(self):
        """Release every free pinned buffer back to the driver."""
        with self.lock:
            items = [(n, p) for n, lst in self.free.items() for p in lst]
            self.free.clear()
        for _, p in items:
            _lib.lib().bldfm_host_free(p)


pool = _Pool()
