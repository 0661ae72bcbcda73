"""Time the other BASELINE configs (1, 3, 4-subsample) on one GPU.  Prints one JSON line per config.
(Parity of these configs against the oracle lives in tests/test_gpu_parity.py.)

    python scripts/bench_configs.py [--configs 1,3,4] [--reps 5]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import bldfm_b200
from bldfm_b200.pbl_model import vertical_profiles
from bldfm_b200.utils import compute_wind_fields, ideal_source


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def timeit(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)), float(np.min(t))


def config1(precision):
    """examples/configs/minimal.yaml as shipped: 512x256, n=16, domain 2000x1000, non-footprint."""
    u, v = compute_wind_fields(4.123, 256.0)
    z, prof = vertical_profiles(16, 10.0, (u, v), ustar=0.4, mol=1e9)
    src = ideal_source((512, 256), (2000.0, 1000.0))
    return dict(srf_flx=src, z=z, profiles=prof, domain=(2000.0, 1000.0), levels=16, modes=(512, 512),
                meas_pt=(0.0, 0.0), footprint=False, precision=precision)


def config3(n=1024, nz=128):
    z, prof = vertical_profiles(nz, 10.0, (6.0, 0.0), ustar=0.4)
    dom = (8000.0 * n / 1024, 8000.0 * n / 1024)
    src = ideal_source((n, n), dom, src_loc=(dom[0] / 4, dom[1] / 2), shape="point")
    return dict(srf_flx=src, z=z, profiles=prof, domain=dom, levels=np.arange(0, nz + 1), modes=(n, n),
                meas_pt=(0.0, 0.0), footprint=False, precision="double")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    which = set(args.configs.split(","))

    if "1" in which:
        for prec in ("single", "double"):
            kw = config1(prec)
            med, best = timeit(lambda: bldfm_b200.steady_state_transport_solver(**kw), args.reps)
            out = {"config": 1, "precision": prec, "e2e_ms_median": med * 1e3, "e2e_ms_min": best * 1e3}
            print(json.dumps(out), flush=True)

    if "3" in which:
        kw = config3()
        nlv = len(kw["levels"])
        med, best = timeit(lambda: bldfm_b200.steady_state_transport_solver(**kw), max(2, args.reps // 2))
        from bldfm_b200 import _lib
        import ctypes as C
        geom = _lib.geometry(kw["srf_flx"].shape, kw["domain"], kw["modes"], None)
        plan = bldfm_b200.get_fft_manager().plan(geom)
        L = _lib.lib()
        L.bldfm_plan_set_profiling(plan, 1)
        bldfm_b200.steady_state_transport_solver(**kw)
        tm = _lib.Timings()
        L.bldfm_plan_last_timings(plan, C.byref(tm))
        L.bldfm_plan_set_profiling(plan, 0)
        gpu_ms = tm.forward_ms + tm.march_ms + tm.inverse_ms
        out = {"config": 3, "shape": "1024x1024x129 levels (208 z-levels), FP64, all levels out",
               "e2e_ms_median": med * 1e3, "e2e_ms_min": best * 1e3, "forward_ms": tm.forward_ms,
               "march_ms": tm.march_ms, "inverse_ms": tm.inverse_ms, "device_ms": gpu_ms,
               "out_bytes": 2 * nlv * 1024 * 1024 * 8, "out_gbs_device": 2 * nlv * 1024 * 1024 * 8 / gpu_ms / 1e6,
               "workspace_gb": bldfm_b200.get_fft_manager().workspace_bytes() / 1e9}
        print(json.dumps(out), flush=True)
        bldfm_b200.reset_fft_manager()

    if "4" in which:
        # 8 towers x T half-hourly met steps, per-solve grid as config 2 (footprints, FP64)
        from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
        T = 48
        rng = np.random.default_rng(0)
        hours = np.arange(T) * 0.5
        ustar = (0.45 + 0.25 * np.sin(2 * np.pi * (hours - 6) / 24) + 0.02 * rng.normal(size=T)).clip(0.1, 0.8)
        mol = np.where(np.sin(2 * np.pi * (hours - 6) / 24) > 0, -1.0, 1.0) * (50.0 + 400.0 * rng.random(T))
        ws = (4.5 + 3.0 * np.sin(2 * np.pi * (hours - 8) / 24) + 0.3 * rng.normal(size=T)).clip(1.0, 8.0)
        wd = (270.0 + 30.0 * rng.normal(size=T)) % 360.0
        towers = [Tower(f"T{i}", 10.0, 1000.0 + 500.0 * (i % 4), 1500.0 + 500.0 * (i // 4)) for i in range(8)]
        cfg = Config(Domain(nx=512, ny=512, xmax=4000.0, ymax=4000.0, nz=64, modes=(512, 512)), towers,
                     Met(ustar=list(ustar), mol=list(mol), wind_speed=list(ws), wind_dir=list(wd)),
                     SolverOptions(footprint=True, precision="double"), Parallel())
        bldfm_b200.run_bldfm_multitower(cfg)
        t0 = time.perf_counter()
        res = bldfm_b200.run_bldfm_multitower(cfg)
        dt = time.perf_counter() - t0
        nsolve = 8 * T
        print(json.dumps({"config": 4, "shape": f"8 towers x {T} met steps of config-2 size ({nsolve} footprints, {T} marches)",
                          "wall_s": dt, "footprints_per_s": nsolve / dt,
                          "note": "host numpy results, 4 MB per footprint"}), flush=True)


if __name__ == "__main__":
    main()
