"""Result delivery of ONE blocking steady_state_transport_solver call (BASELINE config 2): one D2H copy of the
adjacent conc/flx blocks on the copy stream (default) against the last kernel storing straight into the mapped
page-locked result (BLDFM_OUT_MAPPED + option BLDFM_B200_DIRECT_HOST).  Prints one JSON line per variant.

    python scripts/direct_host_probe.py > profiles/r2_direct_host_and_inline_copy.jsonl  (that record also holds a since-removed variant: copies on the compute stream)
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
from bldfm_b200 import _lib
from bench import config2

kw = config2()
L = _lib.lib()
geom = _lib.geometry(kw["srf_flx"].shape, kw["domain"], kw["modes"], None)
plan = bldfm_b200.get_fft_manager().plan(geom)
N = 300
ref = None
for f32 in (0, 1):
    bldfm_b200.config.DELIVER_FLOAT32 = bool(f32)
    for direct in (0, 64 << 20):
        _lib.set_option("BLDFM_B200_DIRECT_HOST", direct)
        for _ in range(20):
            res = bldfm_b200.steady_state_transport_solver(**kw)
        t0 = time.perf_counter()
        for _ in range(N):
            res = bldfm_b200.steady_state_transport_solver(**kw)
        us = (time.perf_counter() - t0) / N * 1e6
        L.bldfm_plan_set_profiling(plan, 1)
        tm = _lib.Timings()
        rows = []
        for _ in range(30):
            bldfm_b200.steady_state_transport_solver(**kw)
            _lib.check(L.bldfm_plan_last_timings(plan, C.byref(tm)))
            rows.append((tm.forward_ms, tm.march_ms, tm.inverse_ms, tm.total_ms))
        L.bldfm_plan_set_profiling(plan, 0)
        f, m, i, t = (float(np.median(c)) * 1e3 for c in zip(*rows))
        conc, flx = np.array(res[1]), np.array(res[2])
        if direct == 0:
            ref = (conc, flx)
        same = bool(np.array_equal(conc, ref[0]) and np.array_equal(flx, ref[1]))
        print(json.dumps({"deliver_float32": bool(f32), "direct_host_bytes": direct, "us_per_call": us,
                          "device_us": {"forward": f, "march": m, "back_transform": i, "after_back_transform": t - f - m - i,
                                        "total": t},
                          "bitwise_equal_to_copy_path": same}), flush=True)
_lib.set_option("BLDFM_B200_DIRECT_HOST", None)
