"""March kernel variants on BASELINE config 2: time per mode (exact / fma) with the per-solve table staged in
shared memory or carried in the kernel parameters, plus -- with BLDFM_B200_MARCH_TRACE=1 -- where the kernel's
fixed cost goes (per-CTA %globaltimer stamps).  One process per variant (the switches are read once).

    python scripts/march_variants.py            # drives the variants
"""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def one():
    import bldfm_b200
    from bldfm_b200 import _lib
    from bldfm_b200.pbl_model import vertical_profiles
    L = _lib.lib()
    n = int(os.environ.get("VAR_N", "64"))
    z, prof = vertical_profiles(n, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    kw = dict(srf_flx=np.zeros((512, 512)), z=z, profiles=prof, domain=(4000.0, 4000.0), levels=n,
              modes=(512, 512), meas_pt=(2000.0, 2000.0), footprint=True, precision="double")
    geom = _lib.geometry((512, 512), kw["domain"], kw["modes"], None)
    plan = bldfm_b200.get_fft_manager().plan(geom)
    for _ in range(5):
        bldfm_b200.steady_state_transport_solver(**kw)
    L.bldfm_plan_set_profiling(plan, 1)
    tm = _lib.Timings()
    t, tot = [], []
    for _ in range(30):
        bldfm_b200.steady_state_transport_solver(**kw)
        L.bldfm_plan_last_timings(plan, C.byref(tm))
        t.append(tm.march_ms)
        tot.append(tm.forward_ms + tm.march_ms + tm.inverse_ms)
    out = {"mode": bldfm_b200.config.MARCH_MODE, "big": os.environ.get("BLDFM_B200_MARCH_BIG", "1"),
           "S": len(z) - 1, "march_us": float(np.median(t)) * 1e3, "solve_us": float(np.median(tot)) * 1e3,
           "fma_used": int(L.bldfm_plan_last_march_mode(plan))}
    if os.environ.get("BLDFM_B200_MARCH_TRACE") == "1":
        buf = np.zeros((4096, 4), dtype=np.uint64)
        nc = C.c_int64(0)
        _lib.check(L.bldfm_plan_march_trace(plan, buf.ctypes.data, 4096, C.byref(nc)))
        b = buf[: nc.value].astype(np.int64)
        t0 = b[:, 0].min()
        out["trace_us"] = {"ctas": int(nc.value),
                           "start_skew_p50_p99": [float(np.percentile(b[:, 0] - t0, q)) * 1e-3 for q in (50, 99)],
                           "staging_p50": float(np.median(b[:, 1] - b[:, 0])) * 1e-3,
                           "loop_p50": float(np.median(b[:, 2] - b[:, 1])) * 1e-3,
                           "epilogue_p50": float(np.median(b[:, 3] - b[:, 2])) * 1e-3,
                           "cta_total_p50_p99": [float(np.percentile(b[:, 3] - b[:, 0], q)) * 1e-3 for q in (50, 99)],
                           "first_start_to_last_end": float(b[:, 3].max() - t0) * 1e-3}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if os.environ.get("VAR_CHILD") == "1":
        one()
    else:
        # 128-thread CTAs (7 per SM) against the lock-step 896-thread CTA (one per SM), with per-CTA traces
        for mode in ("exact", "fma"):
            for big in ("0", "1"):
                for trace in ("0", "1"):
                    env = dict(os.environ, VAR_CHILD="1", BLDFM_B200_MARCH=mode, BLDFM_B200_MARCH_BIG=big,
                               BLDFM_B200_MARCH_TRACE=trace)
                    subprocess.run([sys.executable, __file__], env=env, check=False)
        for n in (16, 128, 256):
            for big in ("0", "2"):
                env = dict(os.environ, VAR_CHILD="1", BLDFM_B200_MARCH="fma", BLDFM_B200_MARCH_BIG=big, VAR_N=str(n))
                subprocess.run([sys.executable, __file__], env=env, check=False)
