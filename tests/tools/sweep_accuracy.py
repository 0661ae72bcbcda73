"""Accuracy of the downward sweep (bldfm_b200/csrc/march.cuh::sweep_body) and of the reference's linear shooting
(src/bldfm/solver.py:220-235) against the SAME discrete two-point problem solved in extended precision
(numpy longdouble, 64-bit mantissa; both formulations agree there to 1e-15 ... 1e-11) -- CPU only, numpy.

Footprint of BASELINE config 2 (unstable MOST, z_m = level 64 of 104 steps) on 4000 random modes, domains
4000 / 2000 / 1000 m (kappa = 7.1 / 10.3 / 15.3).  Expected: binary64 shooting is off by 7e-13 ... 2e-8 (it grows
with e^{2 kappa}: that is the reference's self-noise of SURVEY.md Appendix C), the binary64 sweep by ~1e-15 for
every kappa.  The GPU kernels are compared with the oracle in tests/test_gpu_parity.py; this script only
documents which of the two algorithms the round-off belongs to.

    python tests/tools/sweep_accuracy.py > profiles/r2_sweep_accuracy.jsonl
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from bldfm_b200.pbl_model import vertical_profiles  # noqa: E402
from oracle import bldfm_oracle as O  # noqa: E402


def run(dt, cdt, sweep, z, prof, g, lx, ly, ix, iy, L):
    u, v, Kx, Ky, Kz = [np.asarray(a, dt) for a in prof]
    zz = np.asarray(z, dt)
    Lx = np.asarray(lx[ix], dt)
    Ly = np.asarray(ly[iy], dt)
    S = len(z) - 1
    q0 = np.ones(len(ix), cdt) / dt(g["nxe"]) / dt(g["nye"])

    def coef(i):                                       # a, b, c of solver.py:357-364 (b with the code's sign)
        h = zz[i + 1] - zz[i]
        ki = dt(1) / Kz[i]
        T = (-(Kx[i] * Lx * Lx + Ky[i] * Ly * Ly)).astype(cdt) - cdt(1j) * (u[i] * Lx + v[i] * Ly)
        a = 1 - dt(0.5) * ki * T * h * h
        b = -ki * h - (dt(1) / dt(6)) * ki * ki * T * h ** 3
        c = T * h - (dt(1) / dt(6)) * ki * T * T * h ** 3
        return a, b, c

    kinv = dt(1) / Kz[S]
    eig = np.sqrt((Kx[S] * kinv * Lx * Lx + Ky[S] * kinv * Ly * Ly).astype(cdt)
                  + cdt(1j) * (u[S] * kinv * Lx + v[S] * kinv * Ly))
    lam = Kz[S] * eig
    if not sweep:                                      # two upward IVPs + alpha
        p1 = np.ones(len(ix), cdt)
        q1 = np.zeros(len(ix), cdt)
        p2 = np.zeros(len(ix), cdt)
        q2 = q0.copy()
        for i in range(S):
            if i == L:
                sn = (p1, q1, p2, q2)
            a, b, c = coef(i)
            p1, q1 = a * p1 + b * q1, c * p1 + a * q1
            p2, q2 = a * p2 + b * q2, c * p2 + a * q2
        al = -(q2 - lam * p2) / (q1 - lam * p1)
        return al * sn[0] + sn[2], al * sn[1] + sn[3]
    p = np.ones(len(ix), cdt)                          # one downward sweep with adj(M_i), det product below L
    q = lam.copy()
    D = np.ones(len(ix), cdt)
    for i in range(S - 1, -1, -1):
        a, b, c = coef(i)
        p, q = a * p - b * q, a * q - c * p
        if i == L:
            sn = (p, q)
        if i < L:
            D = D * (a * a - b * c)
    s = q0 * D / q
    return s * sn[0], s * sn[1]


def main():
    rng = np.random.default_rng(0)
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a - b, np.clongdouble)) / np.linalg.norm(b))  # noqa: E731
    for dom in (8000.0, 4000.0, 2000.0, 1000.0):
        z, prof = vertical_profiles(64, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
        g = O.geometry((512, 512), (dom, dom), (512, 512), None)
        lx, ly = O.wavenumbers(g)
        ix = rng.integers(0, 512, 4000)
        iy = rng.integers(0, 512, 4000)
        ok = (ix + iy) > 0
        ix, iy = ix[ok], iy[ok]
        a = (z, prof, g, lx, ly, ix, iy, 64)
        tp, tq = run(np.longdouble, np.clongdouble, True, *a)        # reference solution: extended-precision sweep
        xp, xq = run(np.longdouble, np.clongdouble, False, *a)
        fp, fq = run(np.float64, np.complex128, False, *a)
        sp, sq = run(np.float64, np.complex128, True, *a)
        print(json.dumps({"domain_m": dom, "kappa": round(O.kappa(z, prof, g, float(z[64])), 2),
                          "extended_precision_shooting_vs_sweep": [rel(xp, tp), rel(xq, tq)],
                          "binary64_shooting_error_p_q": [rel(fp, tp), rel(fq, tq)],
                          "binary64_sweep_error_p_q": [rel(sp, tp), rel(sq, tq)]}), flush=True)


if __name__ == "__main__":
    main()
