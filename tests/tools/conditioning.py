"""Conditioning of linear shooting for the parity configs, CPU only.

For every golden case and for BASELINE config 2 prints kappa(z_m) -- growth exponent of the
auxiliary IVP solutions at the largest retained wavenumbers -- and the oracle's SELF-NOISE: the
rel-L2 change of its own output when alpha is multiplied by (1 + 2.2e-16).  A parity claim
<= 1e-10 is only meaningful where the self-noise is <= 1e-11.

    python scripts/conditioning.py [--full]     # --full adds config 2 at 512x512 (a few seconds)
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import SOLVE_CASES, load_case, rel_l2  # noqa: E402
from oracle import bldfm_oracle as O  # noqa: E402


def probe(name, kw, nthreads=1):
    kw = dict(kw, precision="double")
    if kw.get("analytic"):
        return
    g = O.geometry(np.asarray(kw["srf_flx"]).shape, kw["domain"], kw["modes"], kw.get("halo"))
    lv = np.atleast_1d(kw["levels"])
    zl = float(np.asarray(kw["z"])[int(lv.max())])
    kap = O.kappa(np.asarray(kw["z"]), kw["profiles"], g, zl)
    _, c0, f0 = O.solve(nthreads=nthreads, **kw)
    _, c1, f1 = O.solve(nthreads=nthreads, alpha_scale=1.0 + 2.2e-16, **kw)
    print(json.dumps({"case": name, "dx": g["dx"], "kappa_at_top_level": round(kap, 2),
                      "self_noise_conc": rel_l2(c1, c0), "self_noise_flx": rel_l2(f1, f0)}))


if __name__ == "__main__":
    O.build()
    for name in SOLVE_CASES:
        kw, _ = load_case(name)
        probe(name, kw)
    if "--full" in sys.argv:
        from bench import config2
        probe("BASELINE config 2 (512x512x64, domain 4000 m)", config2(), nthreads=O.max_threads())
        kw = config2()
        kw["domain"] = (1000.0, 1000.0)
        kw["meas_pt"] = (500.0, 500.0)
        probe("same at domain 1000 m (ill-conditioned, not a parity config)", kw, nthreads=O.max_threads())
