"""Calibration of BLDFM_MARCH_AUTO (needs a GPU): deviation of the fast marches (FMA-contracted shooting, downward
sweep) from the oracle as a function of the conditioning number kappa, next to the
bit-mirrored march.

Footprint 512x512, n=64, unstable MOST, domains 1000 ... 16000 m (kappa 15 ... 3.5), plus stable / neutral
profiles.  Prints one JSON line per case; the gate (bldfm_auto_kappa_limit, default 8.5) must keep every case
it admits below 1e-10 with a decade of margin.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bldfm_b200  # noqa: E402
from bldfm_b200 import _lib  # noqa: E402
from bldfm_b200.pbl_model import vertical_profiles  # noqa: E402
from conftest import rel_l2  # noqa: E402
from oracle import bldfm_oracle as O  # noqa: E402

O.build()
L = _lib.lib()
limit = L.bldfm_auto_kappa_limit()
cases = []
for dom in (1000.0, 1500.0, 2000.0, 2500.0, 3000.0, 4000.0, 6000.0, 8000.0, 16000.0):
    cases.append(("unstable L=-50", dom, dict(ustar=0.4, mol=-50.0), (-3.0, -4.0)))
for dom in (2000.0, 3000.0, 4000.0):
    cases.append(("stable L=+100", dom, dict(ustar=0.3, mol=100.0), (3.0, 1.0)))
    cases.append(("neutral", dom, dict(ustar=0.5), (5.0, 0.0)))
for name, dom, pk, wind in cases:
    z, prof = vertical_profiles(64, 10.0, wind, **pk)
    kw = dict(srf_flx=np.zeros((512, 512)), z=z, profiles=prof, domain=(dom, dom), levels=64, modes=(512, 512),
              meas_pt=(dom / 2, dom / 2), footprint=True)
    _, oc, of = O.solve(precision="double", nthreads=O.max_threads(), **kw)
    geom = _lib.geometry((512, 512), (dom, dom), (512, 512), None)
    prob, keep = _lib.make_problem(z, prof, kw["meas_pt"], 0.0)
    kap = C.c_double(0.0)
    _lib.check(L.bldfm_kappa(C.byref(geom), C.byref(prob), 64, C.byref(kap)))
    out = {"case": name, "domain_m": dom, "kappa": round(kap.value, 3), "kappa_oracle": round(O.kappa(z, prof, O.geometry((512, 512), (dom, dom), (512, 512), None), float(z[64])), 3),
           "auto_picks_a_fast_march": bool(kap.value <= limit)}
    for mode in ("exact", "fma", "sweep", "auto"):
        bldfm_b200.config.MARCH_MODE = mode
        _, c, f = bldfm_b200.steady_state_transport_solver(precision="double", **kw)
        out[f"{mode}_conc"] = rel_l2(c, oc)
        out[f"{mode}_flx"] = rel_l2(f, of)
    bldfm_b200.config.MARCH_MODE = "exact"
    out["predicted_fma_flx"] = 10 ** (0.468 * kap.value - 15.2)
    print(json.dumps(out), flush=True)
