"""GPU tests of the batched drivers (interface.py mirror) against the oracle."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _config(footprint=True, precision="double", full_output=False):
    from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
    towers = [Tower("A", 10.0, 400.0, 400.0), Tower("B", 10.0, 560.0, 320.0), Tower("C", 6.0, 240.0, 480.0)]
    met = Met(ustar=[0.4, 0.5, 0.3, 0.45], mol=[-50.0, -80.0, 100.0, 1e9],
              wind_speed=[4.0, 5.0, 3.0, 6.0], wind_dir=[270.0, 250.0, 200.0, 10.0])
    dom = Domain(nx=64, ny=48, xmax=960.0, ymax=720.0, nz=16, modes=(64, 48), full_output=full_output)
    return Config(dom, towers, met, SolverOptions(footprint=footprint, precision=precision), Parallel())


def _oracle_result(oracle, cfg, tower, mi):
    from bldfm_b200 import interface
    step = cfg.met.get_step(mi)
    z, prof = interface._profiles_for(cfg, tower.z_m, step)
    dom, sol = cfg.domain, cfg.solver
    return oracle.solve(interface._surface_flux(cfg, None), z, prof, (dom.xmax, dom.ymax), interface._levels(cfg),
                        modes=dom.modes, meas_pt=(tower.x, tower.y), footprint=sol.footprint, halo=dom.halo,
                        precision=sol.precision)


@pytest.mark.parametrize("footprint", [True, False])
def test_multitower_matches_oracle_and_single(gpu_lib, oracle, footprint):
    import bldfm_b200
    cfg = _config(footprint=footprint)
    res = bldfm_b200.run_bldfm_multitower(cfg)
    assert set(res) == {"A", "B", "C"}
    for tower in cfg.towers:
        assert len(res[tower.name]) == 4
        for mi, r in enumerate(res[tower.name]):
            grid, oc, of = _oracle_result(oracle, cfg, tower, mi)
            assert r["conc"].shape == oc.shape
            assert rel_l2(r["conc"], oc) <= 1e-10 and rel_l2(r["flx"], of) <= 1e-10
            assert r["tower_name"] == tower.name and r["timestamp"] == mi
            assert r["tower_xy"] == (tower.x, tower.y)
            for a, b in zip(r["grid"], grid):
                assert np.array_equal(a, b)
            single = bldfm_b200.run_bldfm_single(cfg, tower, met_index=mi)
            assert np.array_equal(single["conc"], r["conc"]) and np.array_equal(single["flx"], r["flx"])


def test_timeseries_full_output_and_cache(gpu_lib, oracle, tmp_path, monkeypatch):
    import bldfm_b200
    monkeypatch.chdir(tmp_path)
    cfg = _config(footprint=True, precision="single", full_output=True)
    cfg.parallel.use_cache = True
    r1 = bldfm_b200.run_bldfm_timeseries(cfg, cfg.towers[1])
    assert len(list((tmp_path / ".bldfm_cache").glob("*.npz"))) == 4
    r2 = bldfm_b200.run_bldfm_timeseries(cfg, cfg.towers[1])
    for a, b in zip(r1, r2):
        assert a["conc"].shape == (17, 48, 64)
        assert np.array_equal(a["conc"], b["conc"]) and np.array_equal(a["flx"], b["flx"])
    _, oc, of = _oracle_result(oracle, cfg, cfg.towers[1], 2)
    assert rel_l2(r1[2]["conc"], oc) <= 1e-5 and rel_l2(r1[2]["flx"], of) <= 1e-5


def test_parallel_without_process_group_equals_multitower(gpu_lib):
    import bldfm_b200
    cfg = _config()
    a = bldfm_b200.run_bldfm_parallel(cfg, max_workers=4, parallel_over="both")
    b = bldfm_b200.run_bldfm_multitower(cfg)
    for name in a:
        for ra, rb in zip(a[name], b[name]):
            assert np.array_equal(ra["flx"], rb["flx"])
    with pytest.raises(ValueError, match="Unknown parallel_over"):
        bldfm_b200.run_bldfm_parallel(cfg, parallel_over="nope")


def test_batch_shares_marches_between_towers(gpu_lib):
    """8 towers at one height and one met step: one march, eight phase shifts."""
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    z, prof = vertical_profiles(16, 10.0, (3.0, -2.0), ustar=0.4, mol=-60.0)
    pts = [(100.0 + 90.0 * i, 700.0 - 70.0 * i) for i in range(8)]
    kw = dict(domain=(960.0, 960.0), levels=16, modes=(64, 64), footprint=True, precision="double")
    conc, flx = bldfm_b200.solve_batched(np.zeros((64, 64)), [z] * 8, [prof] * 8, meas_pts=pts, **kw)
    for i, pt in enumerate(pts):
        _, c1, f1 = bldfm_b200.steady_state_transport_solver(np.zeros((64, 64)), z, prof, meas_pt=pt, **kw)
        assert np.array_equal(conc[i, 0], c1) and np.array_equal(flx[i, 0], f1)


def test_measure_batched_equals_point_measurement(gpu_lib):
    """f-4: footprint x flux-map sums on the device == point_measurement on the host fields."""
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source, point_measurement
    dom = (960.0, 720.0)
    flux_map = ideal_source((64, 48), dom, shape="circle") + 0.1
    zs, profs, pts = [], [], []
    for i in range(5):
        z, p = vertical_profiles(16, 10.0, (3.0 + 0.2 * i, -2.0), ustar=0.4, mol=-60.0 - 10 * i)
        zs.append(z); profs.append(p); pts.append((300.0 + 40 * i, 350.0))
    kw = dict(domain=dom, levels=[8, 16], modes=(64, 48), footprint=True, precision="double")
    conc, flx = bldfm_b200.solve_batched(np.zeros((48, 64)), zs, profs, meas_pts=pts, **kw)
    cw, fw = bldfm_b200.measure_batched(flux_map, np.zeros((48, 64)), zs, profs, meas_pts=pts, **kw)
    assert cw.shape == (5, 2) and fw.shape == (5, 2)
    for b in range(5):
        for l in range(2):
            assert abs(fw[b, l] - point_measurement(flx[b, l], flux_map)) <= 1e-12 * abs(fw[b, l])
            assert abs(cw[b, l] - point_measurement(conc[b, l], flux_map)) <= 1e-12 * abs(cw[b, l])
    # enqueue-only variant: results land in pinned arrays, valid after synchronize()
    cw2, fw2 = bldfm_b200.measure_batched(flux_map, np.zeros((48, 64)), zs, profs, meas_pts=pts, wait=False, **kw)
    cw3, fw3 = bldfm_b200.measure_batched(2.0 * flux_map, np.zeros((48, 64)), zs, profs, meas_pts=pts, wait=False, **kw)
    bldfm_b200.solver.synchronize()
    assert np.array_equal(cw2, cw) and np.array_equal(fw2, fw)
    assert np.allclose(fw3, 2.0 * fw, rtol=1e-13, atol=0.0)


def test_single_precision_batch_mixes_origin_and_shifted_towers(gpu_lib):
    """precision="single", footprint=False: a tower at exactly (0,0) gets float32 fields, shifted towers
    float64 (solver.py:177-185,254-262); the batched drivers must accept the mix and keep each task's dtype."""
    import bldfm_b200
    from bldfm_b200.schema import Tower
    cfg = _config(footprint=False, precision="single")
    cfg.towers = [Tower("O", 10.0, 0.0, 0.0), cfg.towers[0], Tower("O2", 10.0, 0.0, 0.0), cfg.towers[1]]
    res = bldfm_b200.run_bldfm_multitower(cfg)
    for tower in cfg.towers:
        want = np.float32 if (tower.x, tower.y) == (0.0, 0.0) else np.float64
        for mi, r in enumerate(res[tower.name]):
            single = bldfm_b200.run_bldfm_single(cfg, tower, met_index=mi)
            assert r["conc"].dtype == want == single["conc"].dtype and r["flx"].dtype == want
            assert np.array_equal(r["conc"], single["conc"]) and np.array_equal(r["flx"], single["flx"])
    # the low-level entry point merges the two sub-batches into one float64 array
    from bldfm_b200.pbl_model import vertical_profiles
    z, prof = vertical_profiles(16, 10.0, (3.0, -2.0), ustar=0.4, mol=-60.0)
    src = bldfm_b200.ideal_source((64, 48), (960.0, 720.0))
    kw = dict(domain=(960.0, 720.0), levels=16, modes=(64, 48), footprint=False, precision="single")
    conc, flx = bldfm_b200.solve_batched(src, [z] * 3, [prof] * 3, meas_pts=[(0.0, 0.0), (300.0, 200.0), (0.0, 0.0)], **kw)
    assert conc.dtype == np.float64
    _, c0, f0 = bldfm_b200.steady_state_transport_solver(src, z, prof, meas_pt=(0.0, 0.0), **kw)
    _, c1, f1 = bldfm_b200.steady_state_transport_solver(src, z, prof, meas_pt=(300.0, 200.0), **kw)
    assert c0.dtype == np.float32 and c1.dtype == np.float64
    assert np.array_equal(conc[0, 0], c0) and np.array_equal(conc[2, 0], c0) and np.array_equal(flx[1, 0], f1)


def test_timeseries_accepts_a_tower_outside_the_config(gpu_lib):
    import bldfm_b200
    from bldfm_b200.schema import Tower
    cfg = _config()
    other = Tower("X", 8.0, 333.0, 222.0)
    res = bldfm_b200.run_bldfm_timeseries(cfg, other)
    assert len(res) == 4 and res[0]["tower_name"] == "X"
    single = bldfm_b200.run_bldfm_single(cfg, other, met_index=3)
    assert np.array_equal(res[3]["flx"], single["flx"])


def test_measure_sees_in_place_edits_of_the_weight_map(gpu_lib):
    """The weight map is uploaded on every call: a localised in-place edit of the same buffer must show."""
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source, point_measurement
    dom = (960.0, 720.0)
    w = ideal_source((64, 48), dom, shape="circle") + 0.1
    z, p = vertical_profiles(16, 10.0, (3.0, -2.0), ustar=0.4, mol=-60.0)
    kw = dict(domain=dom, levels=16, modes=(64, 48), footprint=True, precision="double")
    _, flx = bldfm_b200.solve_batched(np.zeros((48, 64)), [z], [p], meas_pts=[(300.0, 350.0)], **kw)
    _, f1 = bldfm_b200.measure_batched(w, np.zeros((48, 64)), [z], [p], meas_pts=[(300.0, 350.0)], **kw)
    w[17, 23] += 5.0                      # one cell, same buffer, same address
    _, f2 = bldfm_b200.measure_batched(w, np.zeros((48, 64)), [z], [p], meas_pts=[(300.0, 350.0)], **kw)
    assert abs(f2[0, 0] - point_measurement(flx[0, 0], w)) <= 1e-12 * abs(f2[0, 0])
    assert abs((f2[0, 0] - f1[0, 0]) - 5.0 * flx[0, 0, 17, 23]) <= 1e-9 * abs(5.0 * flx[0, 0, 17, 23])


def test_measure_and_aggregate_drivers_match_host_reductions(gpu_lib, oracle):
    """f-4: run_bldfm_measure == point_measurement of each footprint; run_bldfm_aggregate == np.mean over the
    timesteps (examples/timeseries_example.py:46) of the delivered fields AND of the oracle's fields."""
    import bldfm_b200
    from bldfm_b200 import interface
    from bldfm_b200.utils import ideal_source, point_measurement
    cfg = _config(footprint=True)
    flux_map = ideal_source((64, 48), (960.0, 720.0), shape="circle") + 0.05
    full = bldfm_b200.run_bldfm_multitower(cfg)
    meas = interface.run_bldfm_measure(cfg, flux_map, chunk_groups=2)
    agg = interface.run_bldfm_aggregate(cfg, chunk_groups=2)
    for tower in cfg.towers:
        for mi in range(4):
            want = point_measurement(full[tower.name][mi]["flx"], flux_map)
            assert abs(meas[tower.name]["flx"][mi] - want) <= 1e-12 * abs(want)
        mean_flx = np.mean([r["flx"] for r in full[tower.name]], axis=0)
        mean_conc = np.mean([r["conc"] for r in full[tower.name]], axis=0)
        assert np.abs(agg[tower.name]["flx"] - mean_flx).max() <= 1e-15 * np.abs(mean_flx).max()
        assert np.abs(agg[tower.name]["conc"] - mean_conc).max() <= 1e-15 * np.abs(mean_conc).max()
        omean = np.mean([_oracle_result(oracle, cfg, tower, mi)[2] for mi in range(4)], axis=0)
        assert rel_l2(agg[tower.name]["flx"], omean) <= 1e-10
        assert agg[tower.name]["n"] == 4


def test_float32_delivery_is_opt_in_and_rounds_the_float64_result(gpu_lib):
    """config.DELIVER_FLOAT32: same solve, fields rounded to float32 on the device before the copy."""
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    z, prof = vertical_profiles(16, 10.0, (3.0, -2.0), ustar=0.4, mol=-60.0)
    kw = dict(srf_flx=np.zeros((48, 64)), z=z, profiles=prof, domain=(960.0, 720.0), levels=[8, 16], modes=(64, 48),
              meas_pt=(300.0, 350.0), footprint=True, precision="double")
    _, c64, f64 = bldfm_b200.steady_state_transport_solver(**kw)
    assert c64.dtype == np.float64
    bldfm_b200.config.DELIVER_FLOAT32 = True
    try:
        _, c32, f32 = bldfm_b200.steady_state_transport_solver(**kw)
        cb, fb = bldfm_b200.solve_batched(kw["srf_flx"], [z, z], [prof, prof], domain=kw["domain"], levels=kw["levels"],
                                          modes=kw["modes"], meas_pts=[kw["meas_pt"]] * 2, footprint=True,
                                          precision="double")
    finally:
        bldfm_b200.config.DELIVER_FLOAT32 = False
    assert c32.dtype == np.float32 and f32.dtype == np.float32 and cb.dtype == np.float32
    assert np.array_equal(c32, c64.astype(np.float32)) and np.array_equal(f32, f64.astype(np.float32))
    assert np.array_equal(fb[1], f64.astype(np.float32))
