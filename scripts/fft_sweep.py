"""Sweep the pruned-FFT launch shape (transforms per CTA, threads) and print the back-transform time
of config 2 (1 level) and a config-3-like multi-level solve.  Each setting runs in a subprocess."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r'''
import sys, json, ctypes as C
import numpy as np
sys.path.insert(0, %r)
import bldfm_b200
from bldfm_b200 import _lib
from bench import config2
from scripts.bench_configs import config3
L = _lib.lib()
out = {}
for name, kw in (("cfg2", config2()), ("cfg3_512x65", config3(512, 64))):
    geom = _lib.geometry(kw["srf_flx"].shape, kw["domain"], kw["modes"], None)
    plan = bldfm_b200.get_fft_manager().plan(geom)
    for _ in range(3):
        bldfm_b200.steady_state_transport_solver(**kw)
    L.bldfm_plan_set_profiling(plan, 1)
    tm = _lib.Timings(); inv = []; mar = []
    for _ in range(10):
        bldfm_b200.steady_state_transport_solver(**kw)
        L.bldfm_plan_last_timings(plan, C.byref(tm)); inv.append(tm.inverse_ms); mar.append(tm.march_ms)
    out[name] = {"inverse_ms": float(np.median(inv)), "march_ms": float(np.median(mar))}
print(json.dumps(out))
''' % str(ROOT)

settings = [("default", {}), ("full", {"BLDFM_B200_FFT_FULL": "1"}), ("library", {"BLDFM_B200_FFT_LIBRARY": "1"})]
if "--sweep" in sys.argv:
    for cw in (1, 2, 4):
        for th in (128, 192, 256, 384):
            settings.append((f"cw{cw}_t{th}", {"BLDFM_FFT_CW": str(cw), "BLDFM_FFT_THREADS": str(th)}))
for name, env in settings:
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, cwd=str(ROOT))
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    print(name, line, flush=True)
