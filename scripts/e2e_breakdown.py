"""Where the end-to-end time of ONE steady_state_transport_solver call goes (BASELINE config 2, host numpy in/out):
host timers around the pieces of the Python call + the library's CUDA-event timings of the device work.

    python scripts/e2e_breakdown.py > profiles/r2_e2e_breakdown.json
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bldfm_b200
from bldfm_b200 import _lib, solver
from bench import config2

kw = config2()
L = _lib.lib()
geom = _lib.geometry(kw["srf_flx"].shape, kw["domain"], kw["modes"], None)
plan = bldfm_b200.get_fft_manager().plan(geom)
for _ in range(20):
    bldfm_b200.steady_state_transport_solver(**kw)

acc = {"c_call_enqueue": 0.0, "make_grid": 0.0, "wait_for_results": 0.0}
real_make_grid = solver.make_grid


class Timed:
    """Times two entry points of the library object without touching the others."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in ("bldfm_solve", "bldfm_plan_synchronize"):
            return fn
        key = "c_call_enqueue" if name == "bldfm_solve" else "wait_for_results"

        def wrapped(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            acc[key] += time.perf_counter() - t0
            return r
        return wrapped


def timed_grid(*a, **k):
    t0 = time.perf_counter()
    r = real_make_grid(*a, **k)
    acc["make_grid"] += time.perf_counter() - t0
    return r


N = 300
out = {}
for gmode in ("cow", "0"):
    bldfm_b200.config.GRID_COPY = gmode
    for _ in range(10):
        res = bldfm_b200.steady_state_transport_solver(**kw)
    t0 = time.perf_counter()
    for _ in range(N):
        res = bldfm_b200.steady_state_transport_solver(**kw)
    plain = (time.perf_counter() - t0) / N
    _lib._lib = Timed(L)
    solver.make_grid = timed_grid
    for k in acc:
        acc[k] = 0.0
    t0 = time.perf_counter()
    for _ in range(N):
        res = bldfm_b200.steady_state_transport_solver(**kw)
    total = (time.perf_counter() - t0) / N
    _lib._lib = L
    solver.make_grid = real_make_grid
    parts = {k: v / N * 1e6 for k, v in acc.items()}
    parts["python_before_and_after"] = total * 1e6 - sum(parts.values())
    out[f"grid_{gmode}"] = {"us_per_call_uninstrumented": plain * 1e6, "us_per_call_instrumented": total * 1e6,
                            "host_us": parts}
# device side of the same call (CUDA events in the library; the D2H is the difference to total_ms)
L.bldfm_plan_set_profiling(plan, 1)
tm = _lib.Timings()
rows = []
for _ in range(30):
    bldfm_b200.steady_state_transport_solver(**kw)
    _lib.check(L.bldfm_plan_last_timings(plan, C.byref(tm)))
    rows.append((tm.forward_ms, tm.march_ms, tm.inverse_ms, tm.total_ms))
L.bldfm_plan_set_profiling(plan, 0)
f, m, i, t = (float(np.median(c)) * 1e3 for c in zip(*rows))
out["device_us"] = {"h2d_params_and_forward": f, "march": m, "back_transform": i, "d2h_of_conc_and_flx": t - f - m - i,
                    "first_event_to_results_on_host": t}
out["note"] = ("make_grid runs while the GPU works (overlapped with `wait_for_results`); the call's critical path is "
               "python_before_and_after + c_call_enqueue + max(make_grid, device work) ~ device_us.first_event_to_results_on_host")
print(json.dumps(out, indent=1))
