"""Disk cache of Green's-function results with the reference's key bytes and file layout
(src/bldfm/cache.py:19-83): SHA-256 over the raw bytes of z, the five profiles, domain, modes,
meas_pt, str(halo) and precision; one ``<key>.npz`` with X, Y, Z, conc, flx per entry.  Caches
written by the reference are readable here and vice versa.

``background=True`` moves the ``np.savez`` of ``put`` (about 10 ms for a 512x512 entry -- a hundred times the
GPU solve it memoises) onto a writer thread: ``put`` snapshots the arrays and returns, entries that are still
queued are served from memory by ``get``, files appear atomically (temporary name + rename) and
``flush()`` / interpreter exit wait for the queue.  The files hold the same members, byte for byte, as the
synchronous path writes.
"""

from __future__ import annotations

import atexit
import hashlib
import logging
import os
import queue
import threading
from pathlib import Path

import numpy as np

logger = logging.getLogger("bldfm.cache")


def cache_key(z, profiles, domain, modes, meas_pt, halo, precision) -> str:
    """Hex digest identical to GreensFunctionCache._compute_key (cache.py:36-47)."""
    h = hashlib.sha256()
    h.update(np.asarray(z).tobytes())
    for arr in profiles:
        h.update(np.asarray(arr).tobytes())
    for item in (domain, modes, meas_pt):
        h.update(np.asarray(item).tobytes())
    h.update(str(halo).encode())
    h.update(precision.encode())
    return h.hexdigest()


def _snapshot(a):
    """An array the writer thread can rely on: read-only views (the zero-copy grids) are immutable already,
    anything the caller could still modify is copied."""
    a = np.asarray(a)
    return a if not a.flags.writeable else a.copy()


class GreensFunctionCache:
    """Drop-in for bldfm.cache.GreensFunctionCache (cache.py:19-83)."""

    def __init__(self, cache_dir=".bldfm_cache", background=False):
        self.cache_dir = Path(cache_dir)
        self.cache_dir.mkdir(parents=True, exist_ok=True)
        self.background = bool(background)
        self._pending = {}
        self._lock = threading.Lock()
        self._queue = None
        self._thread = None
        self._error = None

    def _compute_key(self, z, profiles, domain, modes, meas_pt, halo, precision):
        return cache_key(z, profiles, domain, modes, meas_pt, halo, precision)

    def _path(self, *key_args):
        return self.cache_dir / f"{cache_key(*key_args)}.npz"

    def get(self, z, profiles, domain, modes, meas_pt, halo, precision):
        path = self._path(z, profiles, domain, modes, meas_pt, halo, precision)
        with self._lock:
            queued = self._pending.get(path.stem)
        if queued is not None:
            logger.debug("Cache hit (queued for writing): %s", path.stem[:12])
            X, Y, Z, conc, flx = queued
            return (X, Y, Z), conc, flx
        if not path.exists():
            logger.debug("Cache miss: %s", path.stem[:12])
            return None
        logger.debug("Cache hit: %s", path.stem[:12])
        data = np.load(path)
        return (data["X"], data["Y"], data["Z"]), data["conc"], data["flx"]

    @staticmethod
    def _write(path, X, Y, Z, conc, flx):
        tmp = path.with_name(f".{path.stem}.{os.getpid()}.tmp.npz")
        np.savez(tmp, X=X, Y=Y, Z=Z, conc=conc, flx=flx)
        os.replace(tmp, path)

    def put(self, z, profiles, domain, modes, meas_pt, halo, precision, grid, conc, flx):
        path = self._path(z, profiles, domain, modes, meas_pt, halo, precision)
        X, Y, Z = grid
        if not self.background:
            self._write(path, X, Y, Z, conc, flx)
            logger.debug("Cached: %s", path.stem[:12])
            return
        if self._error is not None:
            err, self._error = self._error, None
            raise err
        item = tuple(_snapshot(a) for a in (X, Y, Z, conc, flx))
        with self._lock:
            self._pending[path.stem] = item
        self._start()
        self._queue.put((path, item))

    # ---- writer thread ----------------------------------------------------------------------------
    def _start(self):
        if self._thread is not None and self._thread.is_alive():
            return
        self._queue = queue.Queue()
        self._thread = threading.Thread(target=self._run, name="bldfm-cache-writer", daemon=True)
        self._thread.start()
        atexit.register(self.flush)

    def _run(self):
        while True:
            path, item = self._queue.get()
            try:
                self._write(path, *item)
                logger.debug("Cached: %s", path.stem[:12])
            except Exception as e:           # surfaced by the next put() / flush()
                self._error = e
            finally:
                with self._lock:
                    if self._pending.get(path.stem) is item:
                        del self._pending[path.stem]
                self._queue.task_done()

    def flush(self):
        """Wait until every queued entry is on disk (no-op for the synchronous cache)."""
        if self._queue is not None:
            self._queue.join()
        if self._error is not None:
            err, self._error = self._error, None
            raise err

    def clear(self):
        self.flush()
        count = 0
        for f in self.cache_dir.glob("*.npz"):
            f.unlink()
            count += 1
        logger.info("Cleared %d cache entries", count)
