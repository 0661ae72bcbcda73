"""CPU tests: the oracle (oracle/) against golden vectors generated from the unmodified reference."""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, SOLVE_CASES, load_case, rel_l2


def test_ivp_bitwise_against_numba_golden(oracle):
    d = np.load(GOLDEN / "ivp.npz")
    profs = (d["u"], d["v"], d["Kx"], d["Ky"], d["Kz"])
    for tag in ("a", "b"):
        for nthreads in (1, 3):
            pt, qt, P, Q = oracle.ivp((d[f"{tag}_p0"], d[f"{tag}_q0"]), profs, d["z"], d["levels"],
                                      d["Lx"], d["Ly"], nthreads=nthreads)
            assert np.array_equal(pt, d[f"{tag}_ptop"])
            assert np.array_equal(qt, d[f"{tag}_qtop"])
            assert np.array_equal(P, d[f"{tag}_P"])
            assert np.array_equal(Q, d[f"{tag}_Q"])


@pytest.mark.parametrize("name", SOLVE_CASES)
@pytest.mark.parametrize("precision", ["single", "double"])
def test_solve_matches_reference_golden(oracle, name, precision):
    kw, d = load_case(name)
    grid, conc, flx = oracle.solve(precision=precision, **kw)
    ref_c, ref_f = d[f"conc_{precision}"], d[f"flx_{precision}"]
    assert conc.shape == ref_c.shape and flx.shape == ref_f.shape
    assert conc.dtype == ref_c.dtype and flx.dtype == ref_f.dtype
    # same numpy/scipy build as the generator -> identical; allow last-digit FFT differences
    tol = 1e-6 if ref_c.dtype == np.float32 else 1e-13
    assert rel_l2(conc, ref_c) <= tol
    assert rel_l2(flx, ref_f) <= tol
    if precision == "double":
        for got, key in zip(grid, ("X", "Y", "Z")):
            assert np.array_equal(got, d[key])


def test_reference_regression_goldens(oracle):
    """The reference's own tests/references/{source_area,plume_3d}.npz (atol 1e-6, rtol 1e-5)."""
    ref = np.load(GOLDEN / "refgold.npz")
    for name in ("source_area", "plume_3d"):
        kw, _ = load_case(name)
        _, conc, flx = oracle.solve(precision="single", **kw)
        np.testing.assert_allclose(conc, ref[f"{name}_conc"], atol=1e-6, rtol=1e-5)
        np.testing.assert_allclose(flx, ref[f"{name}_flx"], atol=1e-6, rtol=1e-5)
        assert rel_l2(flx, ref[f"{name}_flx"]) < 1e-12


def test_reference_goldens_in_place(oracle):
    """Same check straight from /root/reference when it is mounted (build container only)."""
    refdir = Path("/root/reference/tests/references")
    if not refdir.exists():
        pytest.skip("reference tree not mounted")
    for name in ("source_area", "plume_3d"):
        ref = np.load(refdir / f"{name}.npz")
        kw, _ = load_case(name)
        _, conc, flx = oracle.solve(precision="single", **kw)
        np.testing.assert_allclose(conc, ref["conc"], atol=1e-6, rtol=1e-5)
        np.testing.assert_allclose(flx, ref["flx"], atol=1e-6, rtol=1e-5)


def test_error_messages(oracle):
    kw, _ = load_case("source_area")
    with pytest.raises(ValueError, match="modes must consist of even numbers."):
        oracle.solve(**{**kw, "modes": (63, 128)})
    with pytest.raises(ValueError, match="precision must be single"):
        oracle.solve(precision="half", **kw)


def test_footprint_sums_to_one(oracle):
    """Sum of the flux footprint over the padded domain is 1 by construction;
    over the cropped domain it lies in (0.25, 1.05] (reference tests/test_integration.py:295-399)."""
    kw, _ = load_case("source_area")
    _, conc, flx = oracle.solve(precision="double", **kw)
    assert 0.25 < flx.sum() <= 1.05


def test_profiles_bitwise():
    from bldfm_b200.pbl_model import vertical_profiles
    d = np.load(GOLDEN / "profiles.npz")
    cases = json.loads(str(d["cases"]))
    for i, c in enumerate(cases):
        c["wind"] = tuple(c["wind"])
        z, p = vertical_profiles(**c)
        assert np.array_equal(z, d[f"z{i}"])
        for name, a in zip(("u", "v", "Kx", "Ky", "Kz"), p):
            assert np.array_equal(np.asarray(a, dtype=np.float64).reshape(-1), d[f"{name}{i}"])


def test_synthetic_generators_bitwise():
    """bldfm_b200.synthetic == the reference's generators (src/bldfm/synthetic.py:12-185) on BASELINE config 4's
    inputs: 1440 half-hourly steps with seed 0, 8 towers on a 500 m grid."""
    from bldfm_b200.synthetic import generate_synthetic_timeseries, generate_towers_grid
    d = np.load(GOLDEN / "synthetic.npz")
    a = generate_synthetic_timeseries(n_timesteps=1440, seed=0)
    for k in ("ustar", "mol", "wind_speed", "wind_dir"):
        assert np.array_equal(np.array(a[k]), d[k]), k
    assert a["timestamps"][0] == str(d["t_first"]) and a["timestamps"][-1] == str(d["t_last"])
    assert generate_towers_grid(n_towers=8, layout="grid", spacing_m=500, z_m=10.0, seed=0) == json.loads(str(d["towers"]))
    assert generate_towers_grid(n_towers=5, layout="random", seed=2) == json.loads(str(d["towers_random"]))
    with pytest.raises(ValueError, match="Unknown layout"):
        generate_towers_grid(layout="ring")


def test_profile_errors():
    from bldfm_b200.pbl_model import vertical_profiles
    with pytest.raises(ValueError, match="Either z0 or ustar"):
        vertical_profiles(8, 10.0, (1.0, 0.0), ustar=0.3, z0=0.1)
    with pytest.raises(ValueError, match="Invalid closure type"):
        vertical_profiles(8, 10.0, (1.0, 0.0), ustar=0.3, closure="NOPE")


def test_cache_keys_identical_to_reference():
    """GreensFunctionCache keys must stay byte-identical to the reference's (cache.py:36-47)."""
    from bldfm_b200.cache import GreensFunctionCache, cache_key
    from bldfm_b200.pbl_model import vertical_profiles
    keys = json.loads((GOLDEN / "cache_keys.json").read_text())
    z, profs = vertical_profiles(16, 10.0, (0.0, -6.0), 0.5)
    for rec in keys.values():
        args = (z, profs, tuple(rec["domain"]), tuple(rec["modes"]), tuple(rec["meas_pt"]), rec["halo"], rec["precision"])
        assert cache_key(*args) == rec["key"]


def test_cache_put_get_clear(tmp_path):
    from bldfm_b200.cache import GreensFunctionCache
    c = GreensFunctionCache(cache_dir=tmp_path / "cc")
    z = np.linspace(0.1, 20, 9)
    profs = tuple(z * (i + 1) for i in range(5))
    args = (z, profs, (10.0, 20.0), (8, 8), (1.0, 2.0), None, "single")
    assert c.get(*args) is None
    grid = (np.ones((4, 4)), np.zeros((4, 4)), np.full((4, 4), 2.0))
    c.put(*args, grid, np.arange(16.0).reshape(4, 4), -np.arange(16.0).reshape(4, 4))
    g, conc, flx = c.get(*args)
    assert np.array_equal(conc, np.arange(16.0).reshape(4, 4)) and np.array_equal(g[2], grid[2])
    assert c.get(z, profs, (10.0, 20.0), (8, 8), (1.0, 2.5), None, "single") is None
    c.clear()
    assert c.get(*args) is None


@pytest.mark.parametrize("shape,modes", [((48, 64), (64, 48)), ((24, 40), (16, 8)), ((15, 45), (512, 512))])
def test_reference_spectra_are_conjugate_symmetric(oracle, shape, modes):
    """The property the half-plane march of the CUDA path rests on (bldfm_b200/csrc/march.cuh), checked on
    the CPU restatement of the reference (itself bit-identical to the reference's numba ivp_solver): in
    footprint mode the combined, shifted spectra satisfy S[-ky][-kx] == conj(S[ky][kx]) BIT FOR BIT for every
    retained mode whose partner is retained too (everything but the Nyquist row/column of even sizes)."""
    from bldfm_b200.pbl_model import vertical_profiles
    ny, nx = shape
    dom = (nx * 9.0, ny * 11.0)
    z, prof = vertical_profiles(12, 8.0, (2.5, -3.5), ustar=0.35, mol=-80.0)
    halo = 0.0 if modes[0] > 100 else None
    tp, tq = oracle.solve(np.zeros((ny, nx)), z, prof, dom, [0, 7, 12], modes=modes, meas_pt=(dom[0] * 0.4, dom[1] * 0.6),
                          srf_bg_conc=0.2, footprint=True, halo=halo, precision="double", return_spectral=True)
    nlv, nly, nlx = tp.shape
    fx = np.fft.fftfreq(nlx, 1.0 / nlx).astype(int)
    fy = np.fft.fftfreq(nly, 1.0 / nly).astype(int)
    ix = {f: i for i, f in enumerate(fx)}
    iy = {f: i for i, f in enumerate(fy)}
    checked = 0
    for ky, f_y in enumerate(fy):
        for kx, f_x in enumerate(fx):
            if -f_y in iy and -f_x in ix:
                my, mx = iy[-f_y], ix[-f_x]
                assert np.array_equal(tp[:, my, mx], np.conj(tp[:, ky, kx]))
                assert np.array_equal(tq[:, my, mx], np.conj(tq[:, ky, kx]))
                checked += 1
    assert checked >= (nlx - 1) * (nly - 1)


def test_downward_sweep_is_the_same_discrete_solution_as_linear_shooting(oracle):
    """The identity behind MARCH_MODE="sweep" (bldfm_b200/csrc/march.cuh::sweep_body), on the CPU: one vector swept
    downward from the radiation condition with adj(M_i), scaled by q0 * prod_{i<L} det(M_i) / (w_0).q, is the
    solution the reference's two upward IVPs + alpha produce (solver.py:220-235).  In extended precision the two
    agree to round-off; in binary64 the sweep stays at 1e-15 while shooting carries the e^{2 kappa} round-off of
    SURVEY.md Appendix C (1e-12 at kappa = 7.1, 1e-10 at kappa = 10.3)."""
    import importlib.util
    from conftest import ROOT
    O = oracle
    spec = importlib.util.spec_from_file_location("sweep_accuracy", ROOT / "tests" / "tools" / "sweep_accuracy.py")
    sa = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sa)
    from bldfm_b200.pbl_model import vertical_profiles
    rng = np.random.default_rng(3)
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a - b, np.clongdouble)) / np.linalg.norm(b))  # noqa: E731
    for dom, shoot_lo, shoot_hi in ((4000.0, 1e-13, 2e-11), (2000.0, 5e-12, 2e-9)):
        z, prof = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
        g = O.geometry((512, 512), (dom, dom), (512, 512), None)
        lx, ly = O.wavenumbers(g)
        ix = rng.integers(1, 512, 600)
        iy = rng.integers(0, 512, 600)
        a = (z, prof, g, lx, ly, ix, iy, 64)
        tp, tq = sa.run(np.longdouble, np.clongdouble, True, *a)
        xp, xq = sa.run(np.longdouble, np.clongdouble, False, *a)
        fp, fq = sa.run(np.float64, np.complex128, False, *a)
        sp, sq = sa.run(np.float64, np.complex128, True, *a)
        if np.finfo(np.longdouble).nmant > 52:                       # x87 extended precision available
            assert rel(xp, tp) <= 1e-12 and rel(xq, tq) <= 1e-12     # same discrete solution
            assert rel(sp, tp) <= 2e-14 and rel(sq, tq) <= 2e-14     # the sweep does not amplify round-off
            assert shoot_lo <= rel(fq, tq) <= shoot_hi               # shooting does, by e^{2 kappa}
        assert rel(sp, fp) <= shoot_hi and rel(sq, fq) <= shoot_hi   # binary64 against binary64
