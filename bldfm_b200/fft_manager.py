"""Plan cache -- the CUDA counterpart of the reference's FFTManager singleton
(src/bldfm/fft_manager.py:12-145).

The reference keeps one process-wide pyFFTW manager (thread count, plan cache, wisdom file).
Here the reusable state is a ``bldfm_plan`` per (device, geometry): cuFFT plans, workspaces and
staged tables.  ``get_fft_manager`` / ``reset_fft_manager`` keep their names and call shapes so
that ``interface._worker_*``-style code (interface.py:216-223) keeps working.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

from . import _lib
from . import config


class PlanManager:
    """Holds the ``bldfm_plan`` handles of this process."""

    def __init__(self, num_threads=1, cache_keepalive=30):
        # accepted for signature compatibility (fft_manager.py:22-24); unused on the GPU
        self.num_threads = num_threads
        self.cache_keepalive = cache_keepalive
        self._plans = {}
        self._memo_owners = []
        self._pid = os.getpid()

    def _forget_memos(self):
        for g in self._memo_owners:
            g.__dict__.pop("_plans", None)
        self._memo_owners = []

    def plan(self, geom: _lib.Geometry, device: int | None = None):
        if os.getpid() != self._pid:
            # forked child: the parent's CUDA context is unusable here -- forget, never destroy
            self._plans = {}
            self._forget_memos()
            self._pid = os.getpid()
        dev = config.DEVICE if device is None else int(device)
        # a bldfm_plan is not thread-safe (include/bldfm_b200.h): every host thread gets its own
        tid = threading.get_ident()
        # fast path: the handle is remembered on the (cached) Geometry object itself
        memo = geom.__dict__.get("_plans")
        if memo is not None:
            hit = memo.get((id(self), dev, tid))
            if hit is not None:
                return hit
        key = (dev, tid) + geom.key()
        h = self._plans.get(key)
        if h is not None:
            self._plans[key] = self._plans.pop(key)        # most recently used goes last
        if h is None:
            self._evict_if_needed()
            h = C.c_void_p()
            _lib.check(_lib.lib().bldfm_plan_create(C.byref(geom), dev, C.byref(h)))
            self._plans[key] = h
        geom.__dict__.setdefault("_plans", {})[(id(self), dev, tid)] = h
        self._memo_owners.append(geom)
        return h

    def _evict_if_needed(self):
        """Destroy least-recently-used plans while the cached workspaces exceed the budget."""
        L = _lib.lib()
        while len(self._plans) > 0 and self.workspace_bytes() > config.MAX_WORKSPACE_BYTES:
            key = next(iter(self._plans))
            L.bldfm_plan_destroy(self._plans.pop(key))
            self._forget_memos()

    def clear_cache(self):
        if os.getpid() == self._pid:
            for h in self._plans.values():
                _lib.lib().bldfm_plan_destroy(h)
        self._plans = {}
        self._forget_memos()

    def launch_count(self) -> int:
        return sum(int(_lib.lib().bldfm_plan_launch_count(h)) for h in self._plans.values())

    def workspace_bytes(self) -> int:
        return sum(int(_lib.lib().bldfm_plan_workspace_bytes(h)) for h in self._plans.values())


_fft_manager = None


def get_fft_manager(num_threads=1, cache_keepalive=30):
    """Get or create the process-wide plan manager (fft_manager.py:121-139)."""
    global _fft_manager
    if _fft_manager is None:
        _fft_manager = PlanManager(num_threads=num_threads, cache_keepalive=cache_keepalive)
    else:
        _fft_manager.num_threads = num_threads
    return _fft_manager


def reset_fft_manager():
    """Drop all plans (fft_manager.py:142-145); safe to call in a forked worker."""
    global _fft_manager
    if _fft_manager is not None:
        _fft_manager.clear_cache()
    _fft_manager = None
