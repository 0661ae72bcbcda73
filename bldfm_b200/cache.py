"""Disk cache of Green's-function results with the reference's key bytes and file layout
(src/bldfm/cache.py:19-83): SHA-256 over the raw bytes of z, the five profiles, domain, modes,
meas_pt, str(halo) and precision; one ``<key>.npz`` with X, Y, Z, conc, flx per entry.  Caches
written by the reference are readable here and vice versa.
"""

from __future__ import annotations

import hashlib
import logging
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


class GreensFunctionCache:
    """Drop-in for bldfm.cache.GreensFunctionCache (cache.py:19-83)."""

    def __init__(self, cache_dir=".bldfm_cache"):
        self.cache_dir = Path(cache_dir)
        self.cache_dir.mkdir(parents=True, exist_ok=True)

    def _compute_key(self, z, profiles, domain, modes, meas_pt, halo, precision):
        return cache_key(z, profiles, domain, modes, meas_pt, halo, precision)

    def _path(self, *key_args):
        return self.cache_dir / f"{cache_key(*key_args)}.npz"

    def get(self, z, profiles, domain, modes, meas_pt, halo, precision):
        path = self._path(z, profiles, domain, modes, meas_pt, halo, precision)
        if not path.exists():
            logger.debug("Cache miss: %s", path.stem[:12])
            return None
        logger.debug("Cache hit: %s", path.stem[:12])
        data = np.load(path)
        return (data["X"], data["Y"], data["Z"]), data["conc"], data["flx"]

    def put(self, z, profiles, domain, modes, meas_pt, halo, precision, grid, conc, flx):
        path = self._path(z, profiles, domain, modes, meas_pt, halo, precision)
        X, Y, Z = grid
        np.savez(path, X=X, Y=Y, Z=Z, conc=conc, flx=flx)
        logger.debug("Cached: %s", path.stem[:12])

    def clear(self):
        count = 0
        for f in self.cache_dir.glob("*.npz"):
            f.unlink()
            count += 1
        logger.info("Cleared %d cache entries", count)
