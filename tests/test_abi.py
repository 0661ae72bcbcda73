"""CPU tests of the boundary: the C-ABI library loads, exports everything include/*.h declares,
its pure-host helpers match numpy bit for bit, and compute calls fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def L():
    from bldfm_b200 import build, _lib
    build.build()
    return _lib


def test_header_and_library_export_the_same_symbols(L):
    header = (ROOT / "include" / "bldfm_b200.h").read_text()
    declared = set(re.findall(r"\b(bldfm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS)
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bldfm_version()


def test_geometry_matches_reference_bookkeeping(L, oracle):
    cases = [((512, 512), (4000.0, 4000.0), (512, 512), None),
             ((128, 64), (100.0, 700.0), (64, 128), None),
             ((48, 64), (500.0, 375.0), (48, 32), 200.0),
             ((9, 11), (220.0, 180.0), (512, 512), 45.0),
             ((256, 512), (2000.0, 1000.0), (512, 512), None),
             ((40, 48), (960.0, 800.0), (24, 16), 0.0)]
    for shape, domain, modes, halo in cases:
        g = L.geometry(shape, domain, modes, halo)
        o = oracle.geometry(shape, domain, modes, halo)
        for k in ("nx", "ny", "px", "py", "nxe", "nye", "nlx", "nly", "nfx", "nfy"):
            assert getattr(g, k) == o[k], (k, shape, modes)
        assert g.dx == o["dx"] and g.dy == o["dy"] and g.halo == o["halo"]
        lx = np.empty(g.nlx)
        ly = np.empty(g.nly)
        assert L.lib().bldfm_wavenumbers(C.byref(g), L.dptr(lx), L.dptr(ly)) == 0
        olx, oly = oracle.wavenumbers(o)
        assert np.array_equal(lx, olx) and np.array_equal(ly, oly)


def test_error_mapping_without_gpu(L):
    from bldfm_b200 import steady_state_transport_solver
    z = np.linspace(0.1, 20.0, 10)
    prof = (z, z, z, z, z)
    with pytest.raises(ValueError, match="modes must consist of even numbers."):
        steady_state_transport_solver(np.zeros((8, 8)), z, prof, (10.0, 10.0), 3, modes=(7, 8))
    with pytest.raises(ValueError, match="precision must be single \\(default\\) or double."):
        steady_state_transport_solver(np.zeros((8, 8)), z, prof, (10.0, 10.0), 3, modes=(8, 8),
                                      precision="half")
    assert L.lib().bldfm_output_is_f32(0, 0.0, 0.0) == 1
    assert L.lib().bldfm_output_is_f32(0, 1.0, 0.0) == 0
    assert L.lib().bldfm_output_is_f32(L.DOUBLE, 0.0, 0.0) == 0
    assert L.lib().bldfm_output_is_f32(L.FOOTPRINT, 0.0, 0.0) == 0


def test_no_cpu_fallback(L):
    """Without a CUDA device the product must raise, never compute on the host."""
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    from bldfm_b200 import steady_state_transport_solver, ivp_solver
    z = np.linspace(0.1, 20.0, 10)
    prof = (z, z, z, z, z)
    with pytest.raises(L.BldfmError, match="no CUDA device"):
        steady_state_transport_solver(np.zeros((8, 8)), z, prof, (10.0, 10.0), 3, modes=(8, 8),
                                      footprint=True)
    with pytest.raises(L.BldfmError, match="no CUDA device"):
        ivp_solver((np.ones(4, complex), np.zeros(4, complex)), prof, z, [3], np.ones(4), np.ones(4))


def test_product_never_imports_oracle():
    pkg = ROOT / "bldfm_b200"
    for f in pkg.rglob("*.py"):
        assert "oracle" not in f.read_text(), f
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle" not in f.read_text(), f
    # the measurement scripts time the product only; parity tools live under tests/
    for f in (ROOT / "scripts").glob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle", f.read_text(), re.M), f


@pytest.mark.parametrize("shape,modes,halo", [
    ((512, 512), (512, 512), None), ((32, 48), (48, 32), 0.0), ((24, 40), (16, 8), 35.0), ((2, 2), (2, 2), None),
    ((15, 45), (512, 512), 0.0), ((33, 20), (512, 512), 0.0), ((20, 33), (512, 512), 0.0), ((100, 100), (50, 66), None),
])
def test_march_thread_map_covers_every_mode_once(L, shape, modes, halo):
    """Half-plane march (csrc/march.cuh): rows ky <= nly/2 plus the mirror stores plus the extra Nyquist-column
    threads must write each retained mode exactly once -- in one launch, and split in row blocks like the
    sharded solve does; the full-plane map is the identity."""
    g = L.geometry(shape, (shape[1] * 7.0, shape[0] * 7.0), modes, halo)
    lib = L.lib()
    nth = C.c_int64(0)

    def cover(row0, rows, half, count):
        L.check(lib.bldfm_march_coverage(C.byref(g), row0, rows, half, count.ctypes.data_as(C.c_void_p), C.byref(nth)))
        return nth.value

    nrow = g.nly // 2 + 1
    count = np.zeros((g.nly, g.nlx), np.int32)
    n = cover(0, nrow, 1, count)
    assert (count == 1).all()
    # about half the modes are marched (the reference marches all of them)
    assert n == g.nlx * nrow + ((g.nly - 1) // 2 if g.nlx % 2 == 0 else 0)
    for G in (2, 3, 8):
        rp = -(-nrow // G)
        if (G - 1) * rp >= nrow:
            continue
        count = np.zeros((g.nly, g.nlx), np.int32)
        total = 0
        for r in range(G):
            total += cover(r * rp, min(rp, nrow - r * rp), 1, count)
        assert (count == 1).all(), G
        assert total == n
    count = np.zeros((g.nly, g.nlx), np.int32)
    assert cover(0, g.nly, 0, count) == g.nlx * g.nly and (count == 1).all()
    with pytest.raises(ValueError):
        cover(0, nrow + 1, 1, count)


def test_kappa_matches_the_oracle_formula(L, oracle):
    """bldfm_kappa (host arithmetic behind BLDFM_MARCH_AUTO) == SURVEY.md Appendix C's conditioning number as
    the oracle computes it, on every golden case; the default gate is 8.5."""
    import ctypes as C
    from bldfm_b200 import _lib
    from conftest import SOLVE_CASES, load_case
    lib = L.lib()
    assert abs(lib.bldfm_auto_kappa_limit() - 8.5) < 1e-12 or "BLDFM_B200_AUTO_KAPPA" in __import__("os").environ
    for name in SOLVE_CASES:
        kw, _ = load_case(name)
        shape = np.asarray(kw["srf_flx"]).shape
        g = oracle.geometry(shape, kw["domain"], kw["modes"], kw.get("halo"))
        geom = _lib.geometry(shape, kw["domain"], kw["modes"], kw.get("halo"))
        z = np.asarray(kw["z"], dtype=np.float64)
        lvl = int(np.max(np.atleast_1d(kw["levels"])))
        prob, keep = _lib.make_problem(z, kw["profiles"], (0.0, 0.0), 0.0)
        kap = C.c_double(0.0)
        assert lib.bldfm_kappa(C.byref(geom), C.byref(prob), lvl, C.byref(kap)) == 0
        want = oracle.kappa(z, [np.asarray(p, dtype=np.float64) for p in kw["profiles"]], g, float(z[lvl]))
        assert abs(kap.value - want) <= 1e-9 * max(1.0, abs(want)), name


def test_make_grid_modes_match_the_reference_meshgrid():
    """solver.make_grid: the three ways of producing (X, Y, Z) (config.GRID_COPY) give the values and shapes of
    the reference's np.meshgrid + squeeze (src/bldfm/solver.py:293-298); the default copy-on-write arrays are
    writable and independent between calls, the zero-copy views are read-only."""
    from bldfm_b200.solver import make_grid
    z = np.linspace(0.1, 20.0, 9)
    for lv in (np.array([5]), np.array([0, 3, 8])):
        for (nx, ny) in ((16, 12), (7, 5)):
            x = np.linspace(0, 100.0, nx, endpoint=False)
            y = np.linspace(0, 60.0, ny, endpoint=False)
            Z, Y, X = np.meshgrid(z[lv], y, x, indexing="ij")
            want = (np.squeeze(X), np.squeeze(Y), np.squeeze(Z))
            for mode in ("cow", "1", "0"):
                got = make_grid(z, lv, (100.0, 60.0), nx, ny, mode=mode)
                for a, b in zip(got, want):
                    assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), mode
            g1 = make_grid(z, lv, (100.0, 60.0), nx, ny, mode="cow")
            g2 = make_grid(z, lv, (100.0, 60.0), nx, ny, mode="cow")
            g1[0][...] -= 50.0                      # what a drop-in caller may do: X -= x0
            g1[2][...] = 0.0
            assert np.array_equal(g2[0], want[0]) and np.array_equal(g2[2], want[2])
            assert np.array_equal(g1[0], want[0] - 50.0)
            ro = make_grid(z, lv, (100.0, 60.0), nx, ny, mode="0")
            with pytest.raises(ValueError):
                ro[0][...] = 1.0


def test_sweep_admissibility_bounds():
    """bldfm_sweep_admissible (host arithmetic in front of every BLDFM_MARCH_SWEEP / AUTO launch): BASELINE
    configs 2 and 5 (same dx and wavenumber range) are admissible at every level; a column whose swept vector
    would grow beyond 2^512 (centimetre cells), a step with |T| h^2 / Kz > 1.5, and degenerate profiles are not."""
    from bldfm_b200 import _lib
    from bldfm_b200.pbl_model import vertical_profiles
    lib = _lib.lib()

    def ok(shape, domain, modes, z, prof, lvl):
        geom = _lib.geometry(shape, domain, modes, None)
        prob, keep = _lib.make_problem(z, prof, (0.0, 0.0), 0.0)
        out = C.c_int32(-1)
        assert lib.bldfm_sweep_admissible(C.byref(geom), C.byref(prob), lvl, C.byref(out)) == 0
        return out.value

    z, prof = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    for lvl in (0, 1, 64, len(z) - 1):
        assert ok((512, 512), (4000.0, 4000.0), (512, 512), z, prof, lvl) == 1
    z5, prof5 = vertical_profiles(256, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    assert ok((4096, 4096), (32000.0, 32000.0), (4096, 4096), z5, prof5, 256) == 1
    # 4 cm cells: growth exponent of the column far beyond 355
    assert ok((512, 512), (20.0, 20.0), (512, 512), z, prof, 64) == 0
    # one coarse step: |T| h^2 / Kz > 1.5 at the largest wavenumber
    zc = np.array([0.0, 1.0, 2.0, 12.0])
    one = np.ones(4)
    assert ok((64, 64), (640.0, 640.0), (64, 64), zc, (2 * one, one, one, one, 0.5 * one), 2) == 0
    assert ok((64, 64), (64000.0, 64000.0), (64, 64), zc, (2 * one, one, one, one, 0.5 * one), 2) == 1
    # non-increasing heights are refused rather than swept
    assert ok((64, 64), (64000.0, 64000.0), (64, 64), np.array([0.0, 1.0, 1.0, 3.0]), (one, one, one, one, one), 2) == 0


def test_flag_and_status_constants_match_the_header():
    """Every `#define BLDFM_<NAME> <value>` of include/bldfm_b200.h has the same value in the ctypes binding
    (flags as `_lib.<NAME>`, status codes as `_lib.<NAME>`), flags are distinct bits, and nothing is missing."""
    import re
    from bldfm_b200 import _lib
    text = (ROOT / "include" / "bldfm_b200.h").read_text()
    defs = {m.group(1): int(m.group(2), 0)
            for m in re.finditer(r"^#define\s+BLDFM_([A-Z0-9_]+)\s+(-?(?:0x[0-9a-fA-F]+|\d+))\b", text, re.M)}
    flags = {k: v for k, v in defs.items() if not k.startswith("ERR_") and k != "OK" and v > 0}
    assert len(flags) >= 14 and "MARCH_SWEEP" in flags and "OUT_MAPPED" in flags
    for name, value in defs.items():
        assert hasattr(_lib, name), f"_lib.{name} is missing"
        assert getattr(_lib, name) == value, name
    bits = sorted(flags.values())
    assert all(b & (b - 1) == 0 for b in bits) and len(set(bits)) == len(bits)
