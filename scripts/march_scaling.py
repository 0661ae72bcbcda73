"""March kernel time vs number of levels (fixed overhead vs per-step cost), exact and fma modes."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bldfm_b200
from bldfm_b200 import _lib
from bldfm_b200.pbl_model import vertical_profiles

L = _lib.lib()
for mode in ("exact", "fma"):
    bldfm_b200.config.MARCH_MODE = mode
    rows = []
    for n in (8, 16, 32, 64, 128, 256):
        z, prof = vertical_profiles(n, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
        kw = dict(srf_flx=np.zeros((512, 512)), z=z, profiles=prof, domain=(4000.0, 4000.0), levels=n,
                  modes=(512, 512), meas_pt=(2000.0, 2000.0), footprint=True, precision="double")
        geom = _lib.geometry((512, 512), kw["domain"], kw["modes"], None)
        plan = bldfm_b200.get_fft_manager().plan(geom)
        for _ in range(3):
            bldfm_b200.steady_state_transport_solver(**kw)
        L.bldfm_plan_set_profiling(plan, 1)
        tm = _lib.Timings()
        t = []
        for _ in range(15):
            bldfm_b200.steady_state_transport_solver(**kw)
            L.bldfm_plan_last_timings(plan, C.byref(tm))
            t.append(tm.march_ms)
        L.bldfm_plan_set_profiling(plan, 0)
        rows.append((len(z) - 1, float(np.median(t))))
    S = np.array([r[0] for r in rows], float)
    T = np.array([r[1] for r in rows], float)
    b, a = np.polyfit(S, T, 1)
    print(json.dumps({"mode": mode, "fixed_us": round(a * 1e3, 2), "per_step_us": round(b * 1e3, 4),
                      "S104_ms": rows[3][1], "steps_ms": rows}))
