"""GreensFunctionCache: byte-identical keys are pinned in test_oracle.py; here the write-behind mode
(reference: src/bldfm/cache.py:67-75 writes synchronously with np.savez)."""
import zipfile

import numpy as np

from bldfm_b200.cache import GreensFunctionCache


def _entry(seed, n=64):
    rng = np.random.default_rng(seed)
    z = np.linspace(0.1, 20.0, 9)
    profiles = tuple(rng.random(9) for _ in range(5))
    x = np.linspace(0, 100.0, n, endpoint=False)
    X = np.broadcast_to(x[None, :], (n, n))
    Y = np.broadcast_to(x[:, None], (n, n))
    Z = np.full((n, n), 10.0)
    conc, flx = rng.random((n, n)), rng.random((n, n))
    key = (z, profiles, (100.0, 100.0), (n, n), (float(seed), 2.0), None, "double")
    return key, (X, Y, Z), conc, flx


def _members(path):
    with zipfile.ZipFile(path) as zf:
        return [(i.filename, zf.read(i.filename)) for i in zf.infolist()]


def test_background_writer_files_equal_synchronous_ones(tmp_path):
    sync = GreensFunctionCache(tmp_path / "sync")
    back = GreensFunctionCache(tmp_path / "back", background=True)
    entries = [_entry(s) for s in range(12)]
    originals = [e[2].copy() for e in entries]
    for key, grid, conc, flx in entries:
        sync.put(*key, grid, conc, flx)
        back.put(*key, grid, conc, flx)
        # the caller may reuse its buffers right after put(): the queued snapshot must not change
        conc += 1.0
    # queued entries are served from memory, with the values at the time of put()
    key, grid, conc, flx = entries[-1]
    hit = back.get(*key)
    assert hit is not None and np.array_equal(hit[1], originals[-1]) and np.array_equal(hit[2], flx)
    back.flush()
    files_s = sorted(p.name for p in (tmp_path / "sync").glob("*.npz"))
    files_b = sorted(p.name for p in (tmp_path / "back").glob("*.npz"))
    assert files_s == files_b and len(files_b) == 12
    assert not list((tmp_path / "back").glob(".*tmp*"))
    for name in files_s:
        # same members in the same order, every .npy member byte for byte (the zip headers carry a timestamp)
        assert _members(tmp_path / "sync" / name) == _members(tmp_path / "back" / name)
    # after the flush the entry comes from disk and is still the same
    hit = back.get(*key)
    assert np.array_equal(hit[1], originals[-1]) and np.array_equal(hit[0][0], grid[0])
    assert back.get(*_entry(99)[0]) is None
    back.clear()
    assert not list((tmp_path / "back").glob("*.npz"))


def test_background_writer_reports_errors(tmp_path):
    import pytest
    back = GreensFunctionCache(tmp_path / "gone", background=True)
    key, grid, conc, flx = _entry(1)
    (tmp_path / "gone").rmdir()
    back.put(*key, grid, conc, flx)
    with pytest.raises(OSError):
        back.flush()
