"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (north_star): rel-L2 <= 1e-10 in FP64, <= 1e-5 in the FP32-storage mode; the march
itself (ivp_solver) is compared BITWISE.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, SOLVE_CASES, load_case, rel_l2, rel_l2c

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-10
TOL_F32 = 1e-5


@pytest.fixture(scope="module")
def B(gpu_lib):
    import bldfm_b200
    return bldfm_b200


def test_ivp_bitwise_against_reference_golden(B):
    d = np.load(GOLDEN / "ivp.npz")
    profs = (d["u"], d["v"], d["Kx"], d["Ky"], d["Kz"])
    for tag in ("a", "b"):
        pt, qt, P, Q = B.ivp_solver((d[f"{tag}_p0"], d[f"{tag}_q0"]), profs, d["z"], d["levels"],
                                    d["Lx"], d["Ly"])
        assert np.array_equal(pt, d[f"{tag}_ptop"])
        assert np.array_equal(qt, d[f"{tag}_qtop"])
        assert np.array_equal(P, d[f"{tag}_P"])
        assert np.array_equal(Q, d[f"{tag}_Q"])


def test_ivp_bitwise_against_oracle_random(B, oracle):
    from bldfm_b200.pbl_model import vertical_profiles
    rng = np.random.default_rng(7)
    z, profs = vertical_profiles(128, 10.0, (2.0, -5.0), ustar=0.35, mol=-30.0)
    M = 5000
    Lx = rng.uniform(-0.2, 0.2, M)
    Ly = rng.uniform(-0.2, 0.2, M)
    p0 = rng.normal(size=M) + 1j * rng.normal(size=M)
    q0 = rng.normal(size=M) + 1j * rng.normal(size=M)
    lv = np.array([0, 3, 128, len(z) - 1])
    got = B.ivp_solver((p0, q0), profs, z, lv, Lx, Ly)
    ref = oracle.ivp((p0, q0), profs, z, lv, Lx, Ly, nthreads=4)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)


def test_ivp_fma_mode_is_close_but_optional(B, oracle):
    d = np.load(GOLDEN / "ivp.npz")
    profs = (d["u"], d["v"], d["Kx"], d["Ky"], d["Kz"])
    prev, B.config.MARCH_MODE = B.config.MARCH_MODE, "fma"
    try:
        pt, qt, P, Q = B.ivp_solver((d["a_p0"], d["a_q0"]), profs, d["z"], d["levels"], d["Lx"], d["Ly"])
    finally:
        B.config.MARCH_MODE = prev
    assert rel_l2c(pt, d["a_ptop"]) < 1e-12
    assert rel_l2c(P, d["a_P"]) < 1e-12


@pytest.mark.parametrize("name", SOLVE_CASES)
@pytest.mark.parametrize("precision", ["single", "double"])
def test_spectral_stage_against_oracle(B, oracle, name, precision):
    from bldfm_b200.solver import spectral_fields
    kw, _ = load_case(name)
    tp, tq = spectral_fields(precision=precision, **kw)
    otp, otq = oracle.solve(precision=precision, return_spectral=True, **kw)
    # the combine amplifies ulp-level differences in alpha by e^{2 kappa};
    # the ill-conditioned goldens (dx ~ 1.5 m) sit near 1e-12, the well-conditioned ones at 1e-15
    tol = TOL_F64 if precision == "double" else 2e-7
    assert tp.shape == otp.shape
    assert rel_l2c(tp, otp) <= tol, rel_l2c(tp, otp)
    assert rel_l2c(tq, otq) <= tol, rel_l2c(tq, otq)


@pytest.mark.parametrize("fft", ["hermitian", "full", "library"])
@pytest.mark.parametrize("name", SOLVE_CASES)
@pytest.mark.parametrize("precision", ["single", "double"])
def test_solve_against_reference_golden(B, name, precision, fft):
    """All three back-transform paths: real-output half-work passes (default), full complex pruned
    passes, and the cuFFT library path."""
    kw, d = load_case(name)
    B.config.FFT_LIBRARY = fft == "library"
    B.config.FFT_FULL = fft == "full"
    try:
        grid, conc, flx = B.steady_state_transport_solver(precision=precision, **kw)
    finally:
        B.config.FFT_LIBRARY = False
        B.config.FFT_FULL = False
    ref_c, ref_f = d[f"conc_{precision}"], d[f"flx_{precision}"]
    assert conc.shape == ref_c.shape and flx.shape == ref_f.shape
    assert conc.dtype == ref_c.dtype and flx.dtype == ref_f.dtype
    if ref_c.dtype == np.float32 or precision == "single":
        tol = TOL_F32
    else:
        tol = TOL_F64
    assert rel_l2(conc, ref_c) <= tol, rel_l2(conc, ref_c)
    assert rel_l2(flx, ref_f) <= tol, rel_l2(flx, ref_f)
    for got, key in zip(grid, ("X", "Y", "Z")):
        assert got.shape == d[key].shape
        assert np.array_equal(got, d[key])


def test_reference_regression_goldens(B):
    """The reference's own tests/references goldens with its own tolerances (test_regression.py:18-24)."""
    ref = np.load(GOLDEN / "refgold.npz")
    for name in ("source_area", "plume_3d"):
        kw, _ = load_case(name)
        _, conc, flx = B.steady_state_transport_solver(precision="single", **kw)
        np.testing.assert_allclose(conc, ref[f"{name}_conc"], atol=1e-6, rtol=1e-5)
        np.testing.assert_allclose(flx, ref[f"{name}_flx"], atol=1e-6, rtol=1e-5)


def _config2(n=512):
    from bldfm_b200.pbl_model import vertical_profiles
    z, profs = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    return dict(srf_flx=np.zeros((n, n)), z=z, profiles=profs, domain=(4000.0, 4000.0), levels=64,
                modes=(n, n), meas_pt=(2000.0, 2000.0), footprint=True, precision="double")


def test_baseline_config2_full_size_against_oracle(B, oracle):
    """BASELINE config 2 (512x512x64 FP64 unstable footprint) at full size: rel-L2 <= 1e-10."""
    kw = _config2()
    _, conc, flx = B.steady_state_transport_solver(**kw)
    _, oc, of = oracle.solve(nthreads=oracle.max_threads(), **kw)
    assert rel_l2(conc, oc) <= TOL_F64, rel_l2(conc, oc)
    assert rel_l2(flx, of) <= TOL_F64, rel_l2(flx, of)
    # size-independent properties: footprint weights sum to ~1 over the domain, finite
    assert np.isfinite(flx).all() and np.isfinite(conc).all()
    assert 0.25 < flx.sum() <= 1.05


def _last_march_mode(B, kw):
    from bldfm_b200 import _lib
    geom = _lib.geometry(np.asarray(kw["srf_flx"]).shape, kw["domain"], kw["modes"], kw.get("halo"))
    return int(_lib.lib().bldfm_plan_last_march_mode(B.get_fft_manager().plan(geom)))


def test_config2_full_size_every_march_mode(B, oracle):
    """BASELINE config 2 at FULL size in every arithmetic mode: each within 1e-10 of the oracle; "auto" (the
    default) picks the downward sweep here (kappa = 7.1 <= 8.5, one output level) and is bit-identical to "sweep"."""
    kw = _config2()
    _, oc, of = oracle.solve(nthreads=oracle.max_threads(), **kw)
    prev = B.config.MARCH_MODE
    got = {}
    try:
        for mode in ("exact", "fma", "sweep", "auto"):
            B.config.MARCH_MODE = mode
            _, conc, flx = B.steady_state_transport_solver(**kw)
            got[mode] = (conc, flx, _last_march_mode(B, kw))
            assert rel_l2(conc, oc) <= TOL_F64 and rel_l2(flx, of) <= TOL_F64, mode
    finally:
        B.config.MARCH_MODE = prev
    assert got["exact"][2] == 0 and got["fma"][2] == 1 and got["sweep"][2] == 2 and got["auto"][2] == 2
    assert np.array_equal(got["auto"][0], got["sweep"][0]) and np.array_equal(got["auto"][1], got["sweep"][1])
    # the bit-mirrored march sits two orders of magnitude closer (1e-14 vs 1e-12): the fast modes differ from
    # the reference by the reference's own round-off (3.7e-13 / 1.2e-12 here)
    assert rel_l2(got["exact"][1], of) <= 1e-13
    assert rel_l2(got["fma"][1], of) <= 1e-11
    assert rel_l2(got["sweep"][1], of) <= 1e-11


def test_auto_mode_falls_back_to_the_exact_march_when_ill_conditioned(B, oracle):
    """kappa gate of BLDFM_MARCH_AUTO: on a 1000 m domain (kappa = 15.3) the FMA march
    would deviate by ~1e-8, so auto must take the bit-mirrored march; in between (2000 m, kappa = 10.3) too."""
    from bldfm_b200 import _lib
    prev = B.config.MARCH_MODE
    try:
        for dom, kap_lo in ((1000.0, 15.0), (2000.0, 10.0)):
            kw = _config2(256)
            kw["domain"] = (dom, dom)
            kw["modes"] = (256, 256)
            kw["meas_pt"] = (dom / 2, dom / 2)
            # same grid spacing / wavenumber range as the 512^2 case of Appendix C
            kw["domain"] = (dom / 2, dom / 2)
            kw["meas_pt"] = (dom / 4, dom / 4)
            geom = _lib.geometry((256, 256), kw["domain"], kw["modes"], None)
            prob, keep = _lib.make_problem(kw["z"], kw["profiles"], kw["meas_pt"], 0.0)
            kap = C.c_double(0.0)
            _lib.check(_lib.lib().bldfm_kappa(C.byref(geom), C.byref(prob), 64, C.byref(kap)))
            assert kap.value > kap_lo > _lib.lib().bldfm_auto_kappa_limit()
            B.config.MARCH_MODE = "auto"
            _, ca, fa = B.steady_state_transport_solver(**kw)
            assert _last_march_mode(B, kw) == 0
            B.config.MARCH_MODE = "exact"
            _, ce, fe = B.steady_state_transport_solver(**kw)
            assert np.array_equal(ca, ce) and np.array_equal(fa, fe)
    finally:
        B.config.MARCH_MODE = prev


def test_fma_mode_within_tolerance_on_a_moderately_conditioned_case(B, oracle):
    kw = _config2(256)
    kw["domain"] = (2000.0, 2000.0)
    kw["meas_pt"] = (1000.0, 1000.0)
    prev, B.config.MARCH_MODE = B.config.MARCH_MODE, "fma"
    try:
        _, conc, flx = B.steady_state_transport_solver(**kw)
    finally:
        B.config.MARCH_MODE = prev
    _, oc, of = oracle.solve(nthreads=oracle.max_threads(), **kw)
    assert rel_l2(conc, oc) <= TOL_F64
    assert rel_l2(flx, of) <= TOL_F64


def test_linearity_in_source(B):
    """Non-footprint solves are linear in srf_flx (size-independent property)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    z, profs = vertical_profiles(16, 10.0, (5.0, 1.0), ustar=0.4)
    dom = (1600.0, 800.0)
    a = ideal_source((128, 64), dom, shape="diamond")
    b = ideal_source((128, 64), dom, src_loc=(300.0, 500.0), shape="point")
    kw = dict(z=z, profiles=profs, domain=dom, levels=[4, 16], modes=(128, 64), meas_pt=(800.0, 400.0),
              precision="double")
    _, ca, fa = B.steady_state_transport_solver(a, **kw)
    _, cb, fb = B.steady_state_transport_solver(b, **kw)
    _, cs, fs = B.steady_state_transport_solver(2.0 * a - 3.0 * b, **kw)
    assert rel_l2(cs, 2.0 * ca - 3.0 * cb) < 1e-12
    assert rel_l2(fs, 2.0 * fa - 3.0 * fb) < 1e-12


def test_footprint_convolution_equals_dispersion(B):
    """Green's-function identity: flux at the tower from a dispersion solve equals
    sum(footprint * source) (what point_measurement computes, utils.py:80-92)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source, point_measurement
    z, profs = vertical_profiles(16, 10.0, (4.0, 2.0), ustar=0.4, mol=-100.0)
    dom = (1280.0, 1280.0)
    n = 128
    src = ideal_source((n, n), dom, src_loc=(400.0, 500.0), shape="circle")
    tower = (800.0, 800.0)   # on a grid node: 800/10
    kw = dict(z=z, profiles=profs, domain=dom, levels=16, modes=(n, n), precision="double")
    _, _, fp = B.steady_state_transport_solver(np.zeros((n, n)), meas_pt=tower, footprint=True, **kw)
    _, _, fl = B.steady_state_transport_solver(src, meas_pt=(0.0, 0.0), footprint=False, **kw)
    ix, iy = int(round(tower[0] / (dom[0] / n))), int(round(tower[1] / (dom[1] / n)))
    assert abs(point_measurement(fp, src) - fl[iy, ix]) <= 1e-9 * abs(fl[iy, ix])


def test_scalar_vs_array_levels_and_dtypes(B):
    kw, d = load_case("noshift")
    g1, c1, f1 = B.steady_state_transport_solver(precision="single", **kw)
    assert c1.dtype == np.float32 and c1.ndim == 2
    kw2 = dict(kw)
    kw2["levels"] = [int(kw["levels"])]
    g2, c2, f2 = B.steady_state_transport_solver(precision="single", **kw2)
    assert np.array_equal(c1, c2) and c2.ndim == 2
    kw3 = dict(kw)
    kw3["levels"] = [2, 16]
    g3, c3, f3 = B.steady_state_transport_solver(precision="double", **kw3)
    assert c3.shape == (2,) + c1.shape and c3.dtype == np.float64
    assert g3[2].shape == c3.shape


def test_errors_and_clamp(B, caplog):
    import logging
    kw, d = load_case("clamp")
    with caplog.at_level(logging.INFO, logger="bldfm.solver"):
        B.steady_state_transport_solver(precision="double", **kw)
    assert any("Setting both equal." in r.message for r in caplog.records)
    with pytest.raises(IndexError):
        B.steady_state_transport_solver(precision="double", **{**kw, "levels": 999})
    kw2, _ = load_case("analytic")
    with pytest.raises(ValueError):
        B.steady_state_transport_solver(**{**kw2, "levels": [3, 12]})


def test_cache_roundtrip(B, tmp_path):
    from bldfm_b200.cache import GreensFunctionCache
    kw, d = load_case("source_area")
    cache = GreensFunctionCache(cache_dir=tmp_path / "c")
    r1 = B.steady_state_transport_solver(precision="double", cache=cache, **kw)
    assert len(list((tmp_path / "c").glob("*.npz"))) == 1
    r2 = B.steady_state_transport_solver(precision="double", cache=cache, **kw)
    assert np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
    for a, b in zip(r1[0], r2[0]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("fft", ["hermitian", "full", "library"])
@pytest.mark.parametrize("shape,modes,halo", [
    ((32, 48), (48, 32), 0.0),        # no halo: transform length == retained modes (Nyquist aliases)
    ((30, 20), (20, 30), 0.0),        # sizes with factors 3 and 5
    ((24, 40), (16, 8), 35.0),        # truncated modes, small halo
    ((2, 2), (2, 2), None),           # smallest grid
    ((15, 45), (512, 512), 0.0),      # odd sizes 3^2*5 x 3*5, modes clamped to odd counts, in-house path
    ((88, 56), (56, 88), 0.0),        # generic-radix stages: 56 = 8*7, 88 = 8*11, no truncation
    ((44, 28), (16, 20), 126.0),      # padded 56 x 66 = (8*7) x (2*3*11), truncated modes
    ((208, 208), (64, 48), 0.0),      # 208 = 16*13
    ((33, 21), (512, 512), 50.0),     # odd sizes, modes clamped; 7 and 11 as factors -> library fallback
])
def test_edge_geometries_against_oracle(B, oracle, shape, modes, halo, fft):
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    ny, nx = shape
    dom = (nx * 9.0, ny * 11.0)
    z, prof = vertical_profiles(12, 8.0, (2.5, -3.5), ustar=0.35, mol=-80.0)
    rng = np.random.default_rng(nx * 100 + ny)
    src = rng.random((ny, nx))
    B.config.FFT_LIBRARY = fft == "library"
    B.config.FFT_FULL = fft == "full"
    try:
        for footprint in (True, False):
            kw = dict(srf_flx=src, z=z, profiles=prof, domain=dom, levels=[0, 7, 12], modes=modes,
                      meas_pt=(dom[0] * 0.4, dom[1] * 0.6), footprint=footprint, halo=halo, precision="double")
            _, c, f = B.steady_state_transport_solver(**kw)
            _, oc, of = oracle.solve(**kw)
            assert c.shape == oc.shape
            assert rel_l2(c, oc) <= TOL_F64, (footprint, rel_l2(c, oc))
            assert rel_l2(f, of) <= TOL_F64, (footprint, rel_l2(f, of))
    finally:
        B.config.FFT_LIBRARY = False
        B.config.FFT_FULL = False


@pytest.mark.parametrize("shape,modes,halo", [
    ((32, 48), (48, 32), 0.0),        # even x even, no truncation
    ((24, 40), (16, 8), 35.0),        # truncated modes
    ((2, 2), (2, 2), None),           # only row 0 and the Nyquist row/column
    ((15, 45), (512, 512), 0.0),      # clamped to odd x odd: every non-zero mode has a partner
    ((33, 20), (512, 512), 0.0),      # odd rows, even columns
    ((64, 64), (64, 64), None),
])
def test_half_plane_march_equals_full_march(B, shape, modes, halo):
    """The default march covers rows ky <= nly/2 and stores conjugates for the rest (march.cuh);
    BLDFM_MARCH_FULL marches every retained mode like the reference.  Footprint spectra must agree
    BITWISE (the source spectrum is an exact constant); with a real source field the two differ only
    by the Hermitian asymmetry of the computed source spectrum: ulp-level, amplified by the shooting
    combine like any other round-off (5e-12 on the worst of these small grids)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.solver import spectral_fields
    ny, nx = shape
    dom = (nx * 9.25, ny * 11.5)      # geometry unique to this test -> a fresh plan, no stale spectra
    z, prof = vertical_profiles(12, 8.0, (2.5, -3.5), ustar=0.35, mol=-80.0)
    rng = np.random.default_rng(nx * 100 + ny)
    src = rng.random((ny, nx))
    for footprint in (True, False):
        for levels in ([12], [0, 7, 12]):
            kw = dict(srf_flx=src, z=z, profiles=prof, domain=dom, levels=levels, modes=modes,
                      meas_pt=(dom[0] * 0.4, dom[1] * 0.6), footprint=footprint, halo=halo,
                      srf_bg_conc=0.3, precision="double")
            hp, hq = spectral_fields(**kw)
            B.config.MARCH_FULL = True
            try:
                fp, fq = spectral_fields(**kw)
            finally:
                B.config.MARCH_FULL = False
            if footprint:
                assert np.array_equal(hp, fp) and np.array_equal(hq, fq)
            else:
                assert rel_l2c(hp, fp) <= TOL_F64 and rel_l2c(hq, fq) <= TOL_F64


@pytest.mark.parametrize("shape", [(256, 256), (512, 512), (256, 512), (1024, 1024), (2048, 1024), (4096, 4096)])
def test_default_halo_passes_equal_generic_passes(B, shape):
    """Default-halo geometry (modes == grid, padded = 3 x grid): the back-transform runs the specialised
    sparse radix-24 passes (fft24.cuh, Q = 32 ... 512) or, for large launches, the sparse radix-48 passes
    (fft48.cuh); BLDFM_B200_FFT24=0 forces the generic real-output passes.  Same operands, different
    factorisations: agreement to round-off."""
    import os
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    ny, nx = shape
    dom = (3000.0, 3000.0)
    z, prof = vertical_profiles(8, 6.0, (3.0, -2.0), ustar=0.4, mol=-200.0)
    cases = [dict(footprint=True, meas_pt=(1300.0, 1700.0), precision="double", levels=[3, 8])]
    if nx <= 1024:
        src = ideal_source((nx, ny), dom, src_loc=(900.0, 2000.0), shape="circle")
        cases.append(dict(footprint=False, meas_pt=(1500.0, 1200.0), precision="double", levels=8, srf_flx=src))
        cases.append(dict(footprint=False, meas_pt=(0.0, 0.0), precision="single", levels=[2, 8], srf_flx=src))
    for case in cases:
        kw = dict(srf_flx=np.zeros((ny, nx)), z=z, profiles=prof, domain=dom, modes=(nx, ny))
        kw.update(case)
        _, c1, f1 = B.steady_state_transport_solver(**kw)
        c1, f1 = c1.copy(), f1.copy()
        # the two-stage variant (fft48.cuh, P = 256 / 512) is picked for large launches only: force it
        from bldfm_b200 import _lib
        _lib.set_option("BLDFM_B200_FFT48", 2)
        try:
            _, c2, f2 = B.steady_state_transport_solver(**kw)
            c2, f2 = c2.copy(), f2.copy()
        finally:
            _lib.set_option("BLDFM_B200_FFT48", None)
        # the persistent bulk-copy pipelined variant (fft24p.cuh) is picked for large launches only: force it;
        # it performs the same operations per element as k_fft24 -> bit-identical
        _lib.set_option("BLDFM_B200_FFT24P", 2)
        try:
            _, c3, f3 = B.steady_state_transport_solver(**kw)
            c3, f3 = c3.copy(), f3.copy()
        finally:
            _lib.set_option("BLDFM_B200_FFT24P", None)
        assert np.array_equal(c3, c1) and np.array_equal(f3, f1), case["footprint"]
        _lib.set_option("BLDFM_B200_FFT24", 0)
        try:
            _, c0, f0 = B.steady_state_transport_solver(**kw)
        finally:
            _lib.set_option("BLDFM_B200_FFT24", None)
        tol = 1e-13 if c0.dtype == np.float64 else 2e-6
        assert c1.dtype == c0.dtype
        assert rel_l2(c1, c0) <= tol, (case["footprint"], rel_l2(c1, c0))
        assert rel_l2(f1, f0) <= tol, (case["footprint"], rel_l2(f1, f0))
        assert rel_l2(c2, c0) <= tol, (case["footprint"], rel_l2(c2, c0))
        assert rel_l2(f2, f0) <= tol, (case["footprint"], rel_l2(f2, f0))


def test_baseline_config1_minimal_yaml_against_oracle(B, oracle):
    """BASELINE config 1 = examples/configs/minimal.yaml as shipped (512x256, n=16, domain 2000x1000 m, modes
    (512,512), default halo -> padded 1536x1280, diamond source, non-footprint, tower at (0,0)): default
    precision="single" (float32 fields, tolerance 1e-5) and precision="double" (1e-10)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import compute_wind_fields, ideal_source
    u, v = compute_wind_fields(4.123, 256.0)
    z, prof = vertical_profiles(16, 10.0, (u, v), ustar=0.4, mol=1e9)
    src = ideal_source((512, 256), (2000.0, 1000.0))
    for precision, tol in (("single", TOL_F32), ("double", TOL_F64)):
        kw = dict(srf_flx=src, z=z, profiles=prof, domain=(2000.0, 1000.0), levels=16, modes=(512, 512),
                  meas_pt=(0.0, 0.0), footprint=False, precision=precision)
        _, c, f = B.steady_state_transport_solver(**kw)
        _, oc, of = oracle.solve(nthreads=oracle.max_threads(), **kw)
        assert c.dtype == oc.dtype and c.shape == oc.shape == (256, 512)
        assert rel_l2(c, oc) <= tol, (precision, rel_l2(c, oc))
        assert rel_l2(f, of) <= tol, (precision, rel_l2(f, of))


def test_baseline_config3_replica_against_oracle(B, oracle):
    """BASELINE config 3 (3-D plume, all levels to z_m written out) at a size the oracle finishes in seconds:
    256x256, n=32 -> 33 output levels, neutral MOST, point source, default halo (the fft24 plan with Q = 32)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    n, nz = 256, 32
    z, prof = vertical_profiles(nz, 10.0, (6.0, 0.0), ustar=0.4)
    dom = (8000.0 * n / 1024, 8000.0 * n / 1024)
    src = ideal_source((n, n), dom, src_loc=(dom[0] / 4, dom[1] / 2), shape="point")
    kw = dict(srf_flx=src, z=z, profiles=prof, domain=dom, levels=np.arange(0, nz + 1), modes=(n, n),
              meas_pt=(0.0, 0.0), footprint=False, precision="double")
    _, c, f = B.steady_state_transport_solver(**kw)
    _, oc, of = oracle.solve(nthreads=oracle.max_threads(), **kw)
    assert c.shape == oc.shape == (nz + 1, n, n)
    assert rel_l2(c, oc) <= TOL_F64, rel_l2(c, oc)
    assert rel_l2(f, of) <= TOL_F64, rel_l2(f, of)
    # level by level as well (the upper levels carry little mass and would hide in the global norm)
    for l in range(nz + 1):
        assert rel_l2(f[l], of[l]) <= 1e-9, (l, rel_l2(f[l], of[l]))


@pytest.mark.gpu
@pytest.mark.parametrize("deliver32", [False, True])
def test_direct_host_stores_equal_the_copy_path(deliver32):
    """BLDFM_OUT_MAPPED: the last transform pass writing the page-locked result itself gives the same bits as
    device buffers + D2H copies (footprint and source mode, one and several levels)."""
    import bldfm_b200
    from bldfm_b200 import _lib
    rng = np.random.default_rng(5)
    z = np.linspace(0.5, 12.0, 24)
    prof = (2.0 + 0.1 * z, 0.5 + 0.02 * z, 0.3 + 0.05 * z, 0.3 + 0.05 * z, 0.2 + 0.04 * z)
    old = bldfm_b200.config.DELIVER_FLOAT32
    bldfm_b200.config.DELIVER_FLOAT32 = deliver32
    try:
        for footprint, levels, q0 in ((True, 23, np.zeros((96, 128))), (False, [5, 11, 23], rng.random((64, 80)))):
            got = []
            for direct in (0, 64 << 20):
                _lib.set_option("BLDFM_B200_DIRECT_HOST", direct)
                _, c, f = bldfm_b200.steady_state_transport_solver(q0, z, prof, (300.0, 240.0), levels, modes=(64, 64),
                                                                   meas_pt=(40.0, 30.0), footprint=footprint,
                                                                   precision="double")
                got.append((np.array(c), np.array(f)))
            assert got[0][0].dtype == (np.float32 if deliver32 else np.float64)
            assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
            assert np.isfinite(got[1][0]).all() and np.abs(got[1][1]).max() > 0
    finally:
        _lib.set_option("BLDFM_B200_DIRECT_HOST", None)
        bldfm_b200.config.DELIVER_FLOAT32 = old


def _sweep_cases():
    rng = np.random.default_rng(11)
    z = np.concatenate([np.linspace(0.2, 10.0, 33), 10.0 + np.cumsum(np.linspace(0.5, 6.0, 20))])
    u = 1.5 + 0.4 * np.log(z / 0.1)
    prof = (u, 0.3 * u, 0.5 + 0.3 * z, 0.4 + 0.25 * z, 0.1 + 0.35 * z)
    q0 = rng.random((96, 128))
    return z, prof, q0


@pytest.mark.gpu
@pytest.mark.parametrize("footprint", [True, False])
@pytest.mark.parametrize("level", [0, 17, 32, 52])
@pytest.mark.parametrize("precision", ["double", "single"])
def test_downward_sweep_equals_the_shooting_march(footprint, level, precision):
    """MARCH_MODE="sweep" (one downward sweep from the radiation condition, march.cuh::sweep_body) against the
    bit-mirrored shooting march: footprint and source mode (complex source spectrum, background concentration in
    mode (0,0)), output level at the ground, inside, at the measurement height and at the top of the column,
    shifted towers, both precisions.  The two differ by round-off only -- and that round-off is the shooting
    march's (it grows with e^{2 kappa(level)}): the yardstick is therefore the difference
    between the two SHOOTING marches (bit-mirrored and FMA-contracted), which the sweep must not exceed."""
    import bldfm_b200 as B
    z, prof, q0 = _sweep_cases()
    kw = dict(srf_flx=q0, z=z, profiles=prof, domain=(1200.0, 900.0), levels=level, modes=(96, 64),
              meas_pt=(310.0, 405.0), srf_bg_conc=0.0 if footprint else 0.7, footprint=footprint,
              precision=precision)
    prev = B.config.MARCH_MODE
    got = {}
    try:
        for mode, code in (("exact", 0), ("fma", 1), ("sweep", 2)):
            B.config.MARCH_MODE = mode
            _, c, f = B.steady_state_transport_solver(**kw)
            got[mode] = (np.array(c), np.array(f))
            assert _last_march_mode(B, kw) == code
    finally:
        B.config.MARCH_MODE = prev
    (c0, f0), (c1, f1), (c2, f2) = got["exact"], got["fma"], got["sweep"]
    assert c2.dtype == c0.dtype and c2.shape == c0.shape
    floor = 1e-13 if precision == "double" else 2e-7
    assert rel_l2(c2, c0) <= max(floor, 4.0 * rel_l2(c1, c0)), (rel_l2(c2, c0), rel_l2(c1, c0))
    assert rel_l2(f2, f0) <= max(floor, 4.0 * rel_l2(f1, f0)), (rel_l2(f2, f0), rel_l2(f1, f0))
    assert rel_l2(c2, c0) <= 1e-9 and rel_l2(f2, f0) <= (1e-9 if precision == "double" else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("footprint", [True, False])
@pytest.mark.parametrize("precision", ["double", "single"])
def test_sweep_for_several_levels_equals_the_shooting_march(footprint, precision):
    """Several output levels in sweep mode (march.cuh::sweep_multi_body: sweep down for alpha, then ONE vector
    upward) against the bit-mirrored shooting march, with the FMA shooting march as the yardstick; levels given
    unsorted and with the top of the column among them."""
    import bldfm_b200 as B
    z, prof, q0 = _sweep_cases()
    kw = dict(srf_flx=q0, z=z, profiles=prof, domain=(1200.0, 900.0), levels=[17, 3, 52, 32, 0], modes=(96, 64),
              meas_pt=(310.0, 405.0), srf_bg_conc=0.0 if footprint else 0.7, footprint=footprint,
              precision=precision)
    prev = B.config.MARCH_MODE
    got = {}
    try:
        for mode, code in (("exact", 0), ("fma", 1), ("sweep", 2)):
            B.config.MARCH_MODE = mode
            _, c, f = B.steady_state_transport_solver(**kw)
            got[mode] = (np.array(c), np.array(f))
            assert _last_march_mode(B, kw) == code
    finally:
        B.config.MARCH_MODE = prev
    (c0, f0), (c1, f1), (c2, f2) = got["exact"], got["fma"], got["sweep"]
    assert c2.shape == c0.shape == (5, 96, 128) and c2.dtype == c0.dtype
    floor = 1e-13 if precision == "double" else 2e-7
    for lv in range(5):
        assert rel_l2(c2[lv], c0[lv]) <= max(floor, 4.0 * rel_l2(c1[lv], c0[lv])), lv
        assert rel_l2(f2[lv], f0[lv]) <= max(floor, 4.0 * rel_l2(f1[lv], f0[lv])), lv


@pytest.mark.gpu
def test_sweep_never_overflows():
    """A column on which the swept vector could leave the binary64 range (here: 0.06 m cells, growth bound far
    beyond 2^512) is marched upward (FMA-contracted) even when the sweep is asked for -- and "auto" takes the
    bit-mirrored march there because the reference itself is round-off dominated."""
    import bldfm_b200 as B
    z, prof, q0 = _sweep_cases()
    prev = B.config.MARCH_MODE
    try:
        B.config.MARCH_MODE = "sweep"
        for levels in (32, [3, 17, 32]):
            kw = dict(srf_flx=q0, z=z, profiles=prof, domain=(8.0, 6.0), levels=levels, modes=(96, 64),
                      footprint=True, precision="double")
            B.steady_state_transport_solver(**kw)
            assert _last_march_mode(B, kw) == 1
        B.config.MARCH_MODE = "auto"
        B.steady_state_transport_solver(**kw)
        assert _last_march_mode(B, kw) == 0
    finally:
        B.config.MARCH_MODE = prev
