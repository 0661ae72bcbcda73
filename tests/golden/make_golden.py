"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (the reference is not available on the GPU box):

    PYTHONPATH=tests/golden/shims:/root/reference/src NUMBA_CACHE_DIR=/tmp/numba_cache \
        python tests/golden/make_golden.py

The shims (tests/golden/shims: `abltk` logger/paths stub, `pyfftw` -> scipy.fft) only satisfy the
reference's imports; no reference file is modified or copied.  Outputs:

  ivp.npz        inputs + outputs of bldfm.solver.ivp_solver (numba) on random modes  -> BITWISE pin
  profiles.npz   bldfm.pbl_model.vertical_profiles outputs for several closures       -> bitwise pin
  solve_*.npz    inputs + outputs of bldfm.solver.steady_state_transport_solver for small cases
                 (incl. the scenarios behind the reference's own tests/references/*.npz)
  refgold.npz    the reference's regression goldens source_area / plume_3d (conc, flx only)
  cache_keys.json  GreensFunctionCache keys of the reference for fixed inputs (byte-identity pin)
  synthetic.npz  bldfm.synthetic generators for BASELINE config 4 (1440 met steps, 8 towers)      -> bitwise pin
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

from bldfm.pbl_model import vertical_profiles  # noqa: E402  (reference)
from bldfm.solver import ivp_solver, steady_state_transport_solver  # noqa: E402  (reference)
from bldfm.utils import ideal_source  # noqa: E402  (reference)

REF_TESTS = Path("/root/reference/tests/references")


def save(name, **arrays):
    path = HERE / name
    np.savez_compressed(path, **arrays)
    print(f"{name}: {path.stat().st_size / 1024:.0f} KiB")


def gen_ivp():
    rng = np.random.default_rng(20240917)
    z, profs = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    M = 768
    Lx = rng.uniform(-0.134, 0.134, M)
    Ly = rng.uniform(-0.134, 0.134, M)
    Lx[:4] = [0.0, 0.05, 0.0, -0.134]
    Ly[:4] = [0.03, 0.0, -0.134, 0.0]
    one = np.ones(M, complex)
    zero = np.zeros(M, complex)
    q0 = (rng.normal(size=M) + 1j * rng.normal(size=M)) * 4.2e-7
    levels = np.array([0, 1, 17, 64, 104])
    out = {}
    for tag, pq in (("a", (one, zero)), ("b", (zero, q0))):
        pt, qt, P, Q = ivp_solver(pq, profs, z, levels, Lx, Ly)
        out.update({f"{tag}_p0": pq[0], f"{tag}_q0": pq[1], f"{tag}_ptop": pt, f"{tag}_qtop": qt,
                    f"{tag}_P": P, f"{tag}_Q": Q})
    save("ivp.npz", z=z, u=profs[0], v=profs[1], Kx=profs[2], Ky=profs[3], Kz=profs[4],
         levels=levels, Lx=Lx, Ly=Ly, **out)


PROFILE_CASES = [
    dict(n=64, meas_height=10.0, wind=(-3.0, -4.0), ustar=0.4, mol=-50.0),
    dict(n=16, meas_height=10.0, wind=(0.0, -6.0), ustar=0.5),
    dict(n=32, meas_height=5.0, wind=(2.0, 1.0), z0=0.1, mol=100.0),
    dict(n=16, meas_height=10.0, wind=(4.0, 0.0), ustar=0.3, closure="CONSTANT"),
    dict(n=16, meas_height=10.0, wind=(4.0, 1.0), ustar=0.3, mol=-20.0, closure="MOSTM"),
    dict(n=16, meas_height=10.0, wind=(4.0, 1.0), ustar=0.3, closure="OAAHOC", tke=0.8),
    dict(n=20, meas_height=12.0, wind=(4.0, 1.0), ustar=0.3, domain_height=40.0, stretch=15.0, prsc=0.8),
]


def gen_profiles():
    out = {"cases": np.array(json.dumps(PROFILE_CASES))}
    for i, c in enumerate(PROFILE_CASES):
        z, p = vertical_profiles(**c)
        out[f"z{i}"] = z
        for name, a in zip(("u", "v", "Kx", "Ky", "Kz"), p):
            out[f"{name}{i}"] = np.asarray(a, dtype=np.float64).reshape(-1)
    save("profiles.npz", **out)


def solve_cases():
    """name -> (kwargs for the solver, profile kwargs)."""
    cases = {}
    # scenario of tests/conftest.py:196-230 (source_area.npz)
    cases["source_area"] = (
        dict(srf_flx=np.zeros((128, 64)), domain=(100.0, 700.0), levels=16, modes=(64, 128),
             meas_pt=(50.0, 0.0), footprint=True),
        dict(n=16, meas_height=10.0, wind=(0.0, -6.0), ustar=0.5))
    # scenario of tests/conftest.py:233-262 (plume_3d.npz)
    cases["plume_3d"] = (
        dict(srf_flx=ideal_source((64, 32), (800.0, 100.0)), domain=(800.0, 100.0),
             levels=np.arange(0, 17, 2), modes=(64, 32), meas_pt=(400.0, 50.0), footprint=False),
        dict(n=16, meas_height=10.0, wind=(6.0, 0.0), ustar=0.4))
    # no phase shift: float32 outputs in single precision
    cases["noshift"] = (
        dict(srf_flx=ideal_source((64, 32), (800.0, 100.0)), domain=(800.0, 100.0), levels=16,
             modes=(64, 32), meas_pt=(0.0, 0.0), footprint=False, srf_bg_conc=0.25),
        dict(n=16, meas_height=10.0, wind=(6.0, 0.0), ustar=0.4))
    # well-conditioned unstable footprint (dx = 7.8 m like BASELINE config 2), custom halo, top level
    cases["fp_unstable"] = (
        dict(srf_flx=np.zeros((48, 64)), domain=(500.0, 375.0), levels=[0, 5, 32, 52], modes=(48, 32),
             meas_pt=(250.0, 180.0), footprint=True, halo=200.0),
        dict(n=32, meas_height=10.0, wind=(-3.0, -4.0), ustar=0.4, mol=-50.0))
    # stable stratification, circle source, truncated modes, levels unsorted / duplicated
    cases["stable_trunc"] = (
        dict(srf_flx=ideal_source((48, 40), (960.0, 800.0), shape="circle"), domain=(960.0, 800.0),
             levels=[8, 2, 8, 20], modes=(24, 16), meas_pt=(100.0, 300.0), footprint=False),
        dict(n=20, meas_height=8.0, wind=(2.0, 3.0), z0=0.05, mol=80.0))
    # modes larger than the padded grid: clamp to (nxe, nye) (solver.py:122-127), odd padded size
    cases["clamp"] = (
        dict(srf_flx=np.zeros((9, 11)), domain=(220.0, 180.0), levels=10, modes=(512, 512),
             meas_pt=(110.0, 90.0), footprint=True, halo=45.0),
        dict(n=10, meas_height=6.0, wind=(3.0, -1.0), ustar=0.35, mol=-200.0))
    # analytic branch, constant profiles (solver.py:193-202)
    cases["analytic"] = (
        dict(srf_flx=ideal_source((32, 24), (640.0, 480.0)), domain=(640.0, 480.0), levels=12,
             modes=(32, 24), meas_pt=(320.0, 240.0), footprint=False, analytic=True, halo=300.0,
             srf_bg_conc=1.5),
        dict(n=12, meas_height=10.0, wind=(4.0, 1.0), ustar=0.3, closure="CONSTANT"))
    cases["analytic_fp"] = (
        dict(srf_flx=np.zeros((24, 32)), domain=(640.0, 480.0), levels=12, modes=(32, 24),
             meas_pt=(320.0, 240.0), footprint=True, analytic=True, halo=300.0),
        dict(n=12, meas_height=10.0, wind=(4.0, 1.0), ustar=0.3, closure="CONSTANT"))
    # MOSTM closure (Kx != Ky), numeric
    cases["mostm"] = (
        dict(srf_flx=np.zeros((32, 32)), domain=(400.0, 400.0), levels=[10, 16], modes=(32, 32),
             meas_pt=(200.0, 200.0), footprint=True),
        dict(n=16, meas_height=10.0, wind=(4.0, 1.0), ustar=0.3, mol=-20.0, closure="MOSTM"))
    return cases


def gen_solves():
    index = {}
    for name, (kw, pkw) in solve_cases().items():
        z, profs = vertical_profiles(**pkw)
        arrays = dict(z=z, u=profs[0], v=profs[1], Kx=profs[2], Ky=profs[3], Kz=profs[4],
                      srf_flx=np.asarray(kw["srf_flx"], dtype=np.float64),
                      levels=np.asarray(kw["levels"]))
        meta = {k: v for k, v in kw.items() if k not in ("srf_flx", "levels")}
        meta["levels_scalar"] = bool(np.ndim(kw["levels"]) == 0)
        for prec in ("single", "double"):
            grid, conc, flx = steady_state_transport_solver(z=z, profiles=profs, precision=prec, **kw)
            arrays[f"conc_{prec}"] = conc
            arrays[f"flx_{prec}"] = flx
            if prec == "double":
                arrays["X"], arrays["Y"], arrays["Z"] = grid
        arrays["meta"] = np.array(json.dumps(meta))
        save(f"solve_{name}.npz", **arrays)
        index[name] = meta
    (HERE / "index.json").write_text(json.dumps(index, indent=1, sort_keys=True))


def gen_cache_keys():
    """Reference cache keys (bldfm.cache.GreensFunctionCache._compute_key) for fixed inputs."""
    import tempfile

    from bldfm.cache import GreensFunctionCache

    cache = GreensFunctionCache(cache_dir=tempfile.mkdtemp())
    z, profs = vertical_profiles(16, 10.0, (0.0, -6.0), 0.5)
    keys = {}
    for i, (domain, modes, meas_pt, halo, prec) in enumerate([
            ((100.0, 700.0), (64, 128), (50.0, 0.0), None, "single"),
            ((100.0, 700.0), (64, 128), (50.0, 0.0), 200.0, "double"),
            ((4000.0, 4000.0), (512, 512), (2000, 2000), None, "double")]):
        keys[str(i)] = dict(domain=domain, modes=modes, meas_pt=meas_pt, halo=halo, precision=prec,
                            key=cache._compute_key(z, profs, domain, modes, meas_pt, halo, prec))
    (HERE / "cache_keys.json").write_text(json.dumps(keys, indent=1))


def gen_refgold():
    out = {}
    for name in ("source_area", "plume_3d"):
        d = np.load(REF_TESTS / f"{name}.npz")
        out[f"{name}_conc"] = d["conc"]
        out[f"{name}_flx"] = d["flx"]
    save("refgold.npz", **out)


def gen_synthetic():
    """BASELINE config 4's inputs: the reference's synthetic met series and tower grid."""
    from bldfm.synthetic import generate_synthetic_timeseries, generate_towers_grid

    a = generate_synthetic_timeseries(n_timesteps=1440, seed=0)
    save("synthetic.npz", ustar=a["ustar"], mol=a["mol"], wind_speed=a["wind_speed"], wind_dir=a["wind_dir"],
         t_first=a["timestamps"][0], t_last=a["timestamps"][-1],
         towers=json.dumps(generate_towers_grid(n_towers=8, layout="grid", spacing_m=500, z_m=10.0, seed=0)),
         towers_random=json.dumps(generate_towers_grid(n_towers=5, layout="random", seed=2)))


if __name__ == "__main__":
    os.chdir("/tmp")
    gen_synthetic()
    gen_ivp()
    gen_profiles()
    gen_solves()
    gen_refgold()
    gen_cache_keys()
