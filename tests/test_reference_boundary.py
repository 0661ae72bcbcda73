"""The drop-in boundary seen from the REFERENCE's own caller: bldfm.interface.run_bldfm_single
(src/bldfm/interface.py:115-128) calls steady_state_transport_solver by keyword; here the reference's
unmodified interface module (staged copy under oracle/_ref, see oracle/stage_ref.py) is run with that one
name rebound to bldfm_b200's solver -- the patch INTEGRATION.md section 2 describes.

  CPU (no GPU): the call is recorded and bound against the drop-in's signature; bldfm_b200's own
                run_bldfm_single must pass the same arguments, value for value.
  GPU         : the reference's run_bldfm_single over the CUDA solver == the reference over its own numba/FFT
                solver (<= 1e-10 rel-L2 in FP64, grids equal, dict layout equal).
"""
import importlib
import inspect
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_l2

REF = ROOT / "oracle" / "_ref"


@pytest.fixture(scope="module")
def ref_interface():
    if not (REF / "src" / "bldfm" / "interface.py").exists():
        pytest.skip("oracle/_ref is not staged (python oracle/stage_ref.py where /root/reference exists)")
    os.environ.setdefault("NUMBA_CACHE_DIR", str(REF / "numba_cache"))
    added = [str(REF / "shims"), str(REF / "src")]
    sys.path[:0] = added
    try:
        mod = importlib.import_module("bldfm.interface")
        cfgmod = importlib.import_module("bldfm.config_parser")
        yield mod, cfgmod
    finally:
        for p in added:
            sys.path.remove(p)


def _ref_config(cfgmod, footprint=True, precision="double", nt=2):
    return cfgmod.parse_config_dict({
        "domain": {"nx": 64, "ny": 48, "xmax": 960.0, "ymax": 720.0, "nz": 16, "modes": [64, 48]},
        "towers": [{"name": "A", "lat": 0.0, "lon": 0.0, "z_m": 10.0, "x": 400.0, "y": 400.0},
                   {"name": "B", "lat": 0.0, "lon": 0.0, "z_m": 6.0, "x": 560.0, "y": 320.0}],
        "met": {"ustar": [0.4, 0.5][:nt], "mol": [-50.0, 100.0][:nt], "wind_speed": [4.0, 5.0][:nt],
                "wind_dir": [270.0, 200.0][:nt]},
        "solver": {"closure": "MOST", "footprint": footprint, "precision": precision},
    })


def test_reference_caller_binds_to_the_dropin_signature(ref_interface, monkeypatch):
    import bldfm_b200
    from bldfm_b200 import interface as ours
    mod, cfgmod = ref_interface
    cfg = _ref_config(cfgmod)
    for t, (x, y) in zip(cfg.towers, [(400.0, 400.0), (560.0, 320.0)]):
        t.x, t.y = x, y
    sig = inspect.signature(bldfm_b200.steady_state_transport_solver)
    assert list(sig.parameters) == list(inspect.signature(
        importlib.import_module("bldfm.solver").steady_state_transport_solver).parameters)
    calls = {}

    def recorder(tag):
        def fake(*args, **kw):
            bound = sig.bind(*args, **kw)         # raises TypeError if the caller does not fit the drop-in
            bound.apply_defaults()
            calls.setdefault(tag, []).append(bound.arguments)
            ny, nx = np.asarray(bound.arguments["srf_flx"]).shape
            return (None, None, None), np.zeros((ny, nx)), np.zeros((ny, nx))
        return fake

    monkeypatch.setattr(mod, "steady_state_transport_solver", recorder("ref"))
    monkeypatch.setattr(ours, "steady_state_transport_solver", recorder("ours"))
    for tower in cfg.towers:
        for mi in range(2):
            r = mod.run_bldfm_single(cfg, tower, met_index=mi)
            o = ours.run_bldfm_single(cfg, tower, met_index=mi)
            assert set(r) == set(o) == {"grid", "conc", "flx", "tower_name", "tower_xy", "timestamp", "params"}
            assert r["tower_name"] == o["tower_name"] and r["timestamp"] == o["timestamp"] and r["params"] == o["params"]
    assert len(calls["ref"]) == len(calls["ours"]) == 4
    for a, b in zip(calls["ref"], calls["ours"]):
        assert set(a) == set(b)
        for k in a:
            if k == "profiles":
                assert all(np.array_equal(p, q) for p, q in zip(a[k], b[k]))       # bitwise the same inputs
            elif isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], k


@pytest.mark.gpu
@pytest.mark.parametrize("footprint,precision", [(True, "double"), (False, "double"), (False, "single")])
def test_reference_run_bldfm_single_over_the_cuda_solver(ref_interface, monkeypatch, gpu_lib, footprint, precision):
    import bldfm_b200
    mod, cfgmod = ref_interface
    cfg = _ref_config(cfgmod, footprint=footprint, precision=precision)
    for t, (x, y) in zip(cfg.towers, [(400.0, 400.0), (560.0, 320.0)]):
        t.x, t.y = x, y
    stock = [mod.run_bldfm_single(cfg, tower, met_index=mi) for tower in cfg.towers for mi in range(2)]
    monkeypatch.setattr(mod, "steady_state_transport_solver", bldfm_b200.steady_state_transport_solver)
    patched = [mod.run_bldfm_single(cfg, tower, met_index=mi) for tower in cfg.towers for mi in range(2)]
    tol = 1e-10 if precision == "double" else 1e-5
    for a, b in zip(stock, patched):
        assert a["conc"].shape == b["conc"].shape and a["conc"].dtype == b["conc"].dtype
        assert rel_l2(b["conc"], a["conc"]) <= tol and rel_l2(b["flx"], a["flx"]) <= tol
        for ga, gb in zip(a["grid"], b["grid"]):
            assert np.array_equal(ga, gb)
        assert a["tower_name"] == b["tower_name"] and a["params"] == b["params"]
        b["grid"][0][...] -= 1.0            # drop-in callers may edit the grids in place
