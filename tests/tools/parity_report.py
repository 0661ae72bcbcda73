"""Parity numbers of the CUDA path against the oracle / golden vectors for DESIGN.md (needs a GPU).

Prints, per golden case and for BASELINE config 2: rel-L2 of conc / flx in exact and fma march
mode, and for the spectral stage the number of elements that are bitwise equal to numpy's result.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bldfm_b200  # noqa: E402
from bldfm_b200.solver import spectral_fields  # noqa: E402
from conftest import SOLVE_CASES, load_case, rel_l2, rel_l2c  # noqa: E402
from oracle import bldfm_oracle as O  # noqa: E402

O.build()


def report(name, kw, nthreads=1):
    out = {"case": name}
    _, oc, of = O.solve(precision="double", nthreads=nthreads, **kw)
    otp, otq = O.solve(precision="double", nthreads=nthreads, return_spectral=True, **kw)
    for mode in ("exact", "fma"):
        bldfm_b200.config.MARCH_MODE = mode
        _, c, f = bldfm_b200.steady_state_transport_solver(precision="double", **kw)
        out[f"{mode}_conc"] = rel_l2(c, oc)
        out[f"{mode}_flx"] = rel_l2(f, of)
        tp, tq = spectral_fields(precision="double", **kw)
        out[f"{mode}_spec_p"] = rel_l2c(tp, otp)
        if mode == "exact":
            out["spec_bitwise_frac"] = float(np.mean((tp == otp) & (tq == otq)))
    bldfm_b200.config.MARCH_MODE = "exact"
    print(json.dumps(out), flush=True)


for name in SOLVE_CASES:
    kw, _ = load_case(name)
    report(name, kw)
from bench import config2  # noqa: E402
kw = config2()
kw.pop("precision")
report("BASELINE config 2", kw, nthreads=O.max_threads())
