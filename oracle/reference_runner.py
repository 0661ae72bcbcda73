"""TEST INFRASTRUCTURE -- times the reference's OWN CPU implementation (numba ivp_solver + FFT + process pool)
from the staged copy under oracle/_ref (see stage_ref.py).  Run as a separate process:

    python oracle/reference_runner.py --mode throughput|latency [--cores C] [--reps K] [--tasks T]

and prints one JSON line.  Workload = BASELINE config 2 (512x512, n=64 -> 105 levels, modes 512x512,
domain 4000 m, default halo, unstable MOST, footprint, FP64):

  latency     bldfm.solver.steady_state_transport_solver, config.NUM_THREADS = cores (numba parallel=True,
              SURVEY.md 8d (i)), warm second call onwards
  throughput  bldfm.interface.run_bldfm_parallel(cfg, max_workers=cores, parallel_over="both") over T solves
              (default 2*cores), wall clock incl. pool start-up and result return

The FFT backend is scipy.fft standing in for pyFFTW (absent from the image; <= 9 % of the CPU time).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"


def _setup(cores):
    if not (REF / "src" / "bldfm" / "solver.py").exists():
        print(json.dumps({"unavailable": "oracle/_ref is not staged (python oracle/stage_ref.py in the build container)"}))
        sys.exit(0)
    sys.path[:0] = [str(REF / "shims"), str(REF / "src")]
    os.environ.setdefault("NUMBA_CACHE_DIR", str(REF / "numba_cache"))
    os.environ["NUMBA_NUM_THREADS"] = str(cores)
    os.environ.setdefault("PYTHONPATH", "")
    # pool workers are fresh interpreters only under spawn/forkserver; under fork they inherit sys.path
    os.environ["PYTHONPATH"] = os.pathsep.join([str(REF / "shims"), str(REF / "src"), os.environ["PYTHONPATH"]])


def _config2_call():
    import numpy as np
    from bldfm.pbl_model import vertical_profiles
    z, profs = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    return dict(srf_flx=np.zeros((512, 512)), z=z, profiles=profs, domain=(4000.0, 4000.0), levels=64,
                modes=(512, 512), meas_pt=(2000.0, 2000.0), footprint=True, precision="double")


def _config2_cfg(ntasks):
    """BASELINE config 2 as a BLDFMConfig with `ntasks` met rows (wind (-3,-4) <-> speed 5, direction from
    atan2: compute_wind_fields gives u = -ws*sin(wd), v = -ws*cos(wd))."""
    import numpy as np
    from bldfm.config_parser import parse_config_dict
    wd = float(np.degrees(np.arctan2(3.0, 4.0)))
    return parse_config_dict({
        "domain": {"nx": 512, "ny": 512, "xmax": 4000.0, "ymax": 4000.0, "nz": 64, "modes": [512, 512]},
        "towers": [{"name": "T", "lat": 0.0, "lon": 0.0, "z_m": 10.0, "x": 2000.0, "y": 2000.0}],
        "met": {"ustar": [0.4] * ntasks, "mol": [-50.0] * ntasks, "wind_speed": [5.0] * ntasks,
                "wind_dir": [wd] * ntasks},
        "solver": {"closure": "MOST", "footprint": True, "precision": "double"},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"])
    ap.add_argument("--cores", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--tasks", type=int, default=0)
    a = ap.parse_args()
    _setup(a.cores)
    import numpy as np  # noqa: F401
    import bldfm.config as rcfg
    out = {"mode": a.mode, "cores": a.cores, "kind": "reference",
           "fft": "scipy.fft (pyfftw shim)", "numba_threads": a.cores}
    if a.mode == "latency":
        from bldfm.solver import steady_state_transport_solver
        rcfg.NUM_THREADS = a.cores
        kw = _config2_call()
        t_first = time.perf_counter()
        steady_state_transport_solver(**kw)                 # JIT / cache load
        out["first_call_s"] = time.perf_counter() - t_first
        for _ in range(max(0, a.warmup - 1)):
            steady_state_transport_solver(**kw)
        times = []
        t_all = time.perf_counter()
        for _ in range(a.reps):
            t0 = time.perf_counter()
            steady_state_transport_solver(**kw)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > 120.0:          # bounded: the run must end within minutes
                break
        out.update(reps=len(times), s_per_solve=float(np.mean(times)), solves_per_s=len(times) / float(sum(times)),
                   step_times_s=times[:20])
    else:
        import multiprocessing
        # the reference's own test harness does the same (tests/conftest.py:7): numba's OpenMP layer does not
        # survive a plain fork()
        multiprocessing.set_start_method("forkserver", force=True)
        from bldfm.interface import run_bldfm_parallel, run_bldfm_single
        ntasks = a.tasks or 2 * a.cores
        cfg = _config2_cfg(ntasks)
        tower = cfg.towers[0]
        tower.x, tower.y = 2000.0, 2000.0
        run_bldfm_single(cfg, tower, met_index=0)            # warm the on-disk numba cache for the workers
        for _ in range(a.warmup):
            run_bldfm_parallel(cfg, max_workers=a.cores, parallel_over="both")
        times = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            res = run_bldfm_parallel(cfg, max_workers=a.cores, parallel_over="both")
            times.append(time.perf_counter() - t0)
            assert len(res["T"]) == ntasks
        out.update(reps=a.reps, tasks_per_rep=ntasks, step_times_s=times,
                   solves_per_s=ntasks * len(times) / float(sum(times)),
                   flx_sum=float(res["T"][0]["flx"].sum()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
