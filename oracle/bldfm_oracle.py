"""oracle/bldfm_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the C march in oracle/march.c) of the reference hot path
``bldfm.solver.steady_state_transport_solver`` (/root/reference/src/bldfm/solver.py:16-304)
and ``ivp_solver`` (solver.py:307-374).  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module, and only
as the checker.  ``bldfm_b200`` never imports it.

Parity pinning (see tests/test_oracle.py, tests/golden/make_golden.py):
  * the C march is bitwise-equal to the reference's numba ``ivp_solver`` on golden vectors
    generated from the unmodified reference in the build container;
  * ``solve`` reproduces the reference's own regression goldens
    (tests/references/source_area.npz, plume_3d.npz) and reference outputs on further
    seeded configs, to round-off.

The stages are written in this module's own structure (index maps instead of
fftshift/pad/ifftshift chains), but wherever round-off is amplified by the linear-shooting
combination the arithmetic is issued through the same numpy
operations, in the same order, as the cited reference lines.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import scipy.fft as _fft

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    """Compile oracle/march.c -> liboracle_march.so (gcc, -ffp-contract=off)."""
    so = _HERE / "liboracle_march.so"
    if force or not so.exists() or so.stat().st_mtime < (_HERE / "march.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle_march.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(str(build()))
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_ivp.restype = ctypes.c_int
        lib.oracle_ivp.argtypes = [
            ctypes.c_int64, dp, dp, ctypes.c_int, dp, dp, dp, dp, dp, dp,
            ctypes.c_int, ctypes.POINTER(ctypes.c_int64), dp, dp, dp, dp, dp, dp, ctypes.c_int,
        ]
        lib.oracle_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def ivp(fftpq, profiles, z, levels, Lx, Ly, nthreads: int = 1):
    """Restatement of ``ivp_solver`` (solver.py:307-374): returns (p_top, q_top, P, Q)."""
    p0 = np.ascontiguousarray(fftpq[0], dtype=np.complex128)
    q0 = np.ascontiguousarray(fftpq[1], dtype=np.complex128)
    u, v, Kx, Ky, Kz = (np.ascontiguousarray(a, dtype=np.float64) for a in profiles)
    z = np.ascontiguousarray(z, dtype=np.float64)
    lv = np.ascontiguousarray(np.atleast_1d(levels), dtype=np.int64)
    Lx = np.ascontiguousarray(Lx, dtype=np.float64)
    Ly = np.ascontiguousarray(Ly, dtype=np.float64)
    M = p0.shape[0]
    p_top = np.empty(M, np.complex128)
    q_top = np.empty(M, np.complex128)
    P = np.empty((len(lv), M), np.complex128)
    Q = np.empty((len(lv), M), np.complex128)
    rc = _lib().oracle_ivp(
        M, _dptr(p0), _dptr(q0), len(z), _dptr(z), _dptr(u), _dptr(v), _dptr(Kx), _dptr(Ky),
        _dptr(Kz), len(lv), lv.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dptr(Lx),
        _dptr(Ly), _dptr(p_top), _dptr(q_top), _dptr(P), _dptr(Q), int(nthreads))
    assert rc == 0
    return p_top, q_top, P, Q


def max_threads() -> int:
    return int(_lib().oracle_max_threads())


# ----------------------------------------------------------------------------------------
# geometry / index maps
# ----------------------------------------------------------------------------------------

def geometry(shape, domain, modes, halo):
    """Grid bookkeeping of solver.py:93-130 (pad widths, padded size, clamped modes)."""
    ny, nx = shape
    xmx, ymx = domain
    nlx, nly = modes
    if (nlx % 2 > 0) or (nly % 2 > 0):                      # :90-91
        raise ValueError("modes must consist of even numbers.")
    dx, dy = xmx / nx, ymx / ny                             # :98
    if halo is None:                                        # :108-109
        halo = max(xmx, ymx)
    px, py = int(halo / dx), int(halo / dy)                 # :112-113
    nxe, nye = nx + 2 * px, ny + 2 * py                     # :119-120
    if (nlx > nxe) or (nly > nye):                          # :122-127 clamp BOTH
        nlx, nly = nxe, nye
    dlx, dly = (nxe - nlx) // 2, (nye - nly) // 2           # :130
    return dict(nx=nx, ny=ny, dx=dx, dy=dy, halo=halo, px=px, py=py, nxe=nxe, nye=nye,
                nlx=nlx, nly=nly, dlx=dlx, dly=dly,
                nfx=nlx + 2 * dlx, nfy=nly + 2 * dly)       # transform size after :269-278


def wrap_index(nl, nf):
    """Truncated index i (fftfreq order, length nl) -> index in a length-nf spectrum.

    Equivalent to fftshift -> centre pad/slice -> ifftshift (solver.py:139-145, 265-278)
    whenever nf - nl is even: signed frequency f = i (i < nl/2) or i - nl, stored at f mod nf.
    """
    i = np.arange(nl)
    f = np.where(i < nl // 2, i, i - nl)
    return np.mod(f, nf)


def wavenumbers(g):
    """lx[nlx], ly[nly] of solver.py:148-153 (same numpy expressions)."""
    ilx = np.fft.fftfreq(g["nlx"], d=1.0 / g["nlx"])
    ily = np.fft.fftfreq(g["nly"], d=1.0 / g["nly"])
    lx = 2.0 * np.pi / g["dx"] / g["nxe"] * ilx
    ly = 2.0 * np.pi / g["dy"] / g["nye"] * ily
    return lx, ly


# ----------------------------------------------------------------------------------------
# full solve
# ----------------------------------------------------------------------------------------

def solve(srf_flx, z, profiles, domain, levels, modes=(512, 512), meas_pt=(0.0, 0.0),
          srf_bg_conc=0.0, footprint=False, analytic=False, halo=None, precision="single",
          nthreads: int = 1, return_spectral: bool = False, alpha_scale: float = 1.0):
    """Restatement of ``steady_state_transport_solver`` (solver.py:16-304) without the cache.

    Returns ((X, Y, Z), conc, flx) exactly like the reference (np.squeeze'd).
    """
    q0 = np.asarray(srf_flx)
    z = np.asarray(z, dtype=np.float64)
    u, v, Kx, Ky, Kz = profiles
    xmx, ymx = domain
    xm, ym = meas_pt
    nz = len(z)
    g = geometry(q0.shape, domain, modes, halo)
    nx, ny, px, py, nxe, nye = g["nx"], g["ny"], g["px"], g["py"], g["nxe"], g["nye"]
    nlx, nly, nfx, nfy = g["nlx"], g["nly"], g["nfx"], g["nfy"]
    halo = g["halo"]
    if (nxe - nlx) % 2 or (nye - nly) % 2:
        # solver.py:130,142 mis-slices (IndexError / shape drift) when the difference is odd.
        raise ValueError("padded grid size minus modes must be even.")
    if precision not in ("single", "double"):               # :187-188
        raise ValueError("precision must be single (default) or double.")

    lv = np.array([levels]) if np.ndim(levels) == 0 else np.asarray(levels)   # :102-103
    nlv = len(lv)
    wx, wy = wrap_index(nlx, nfx), wrap_index(nly, nfy)

    # --- K1-K3: source spectrum on the retained modes (solver.py:116,132-145)
    if footprint:
        tq0 = np.ones((nly, nlx), dtype=np.complex128) / nxe / nye             # :134
    else:
        padded = np.zeros((nye, nxe), dtype=q0.dtype)
        padded[py:py + ny, px:px + nx] = q0                                    # :116
        spec = _fft.fft2(padded, norm="forward", workers=nthreads)             # :136
        tq0 = spec[wy[:, None], wx[None, :]]                                   # :139-145

    # --- K4: wavenumbers, top eigenvalue (solver.py:148-174)
    lx, ly = wavenumbers(g)
    Lx, Ly = np.meshgrid(lx, ly)
    msk = np.ones((nly, nlx), dtype=bool)
    msk[0, 0] = False
    Lxm, Lym = Lx[msk], Ly[msk]
    kinv_top = 1.0 / Kz[nz - 1]
    kxk = Kx[nz - 1] * kinv_top
    kyk = Ky[nz - 1] * kinv_top
    eig = np.sqrt(kxk * Lxm ** 2 + kyk * Lym ** 2
                  + 1j * u[nz - 1] * kinv_top * Lxm + 1j * v[nz - 1] * kinv_top * Lym)

    cdt = np.complex64 if precision == "single" else np.complex128            # :177-185
    tp = np.zeros((nlv, nly, nlx), dtype=cdt)
    tq = np.zeros((nlv, nly, nlx), dtype=cdt)
    tp[0, 0, 0] = srf_bg_conc                                                  # :190
    tq[:, 0, 0] = tq0[0, 0]                                                    # :191

    if analytic:                                                               # :193-202
        h = z[lv] - z[0]
        tp[0, msk] = tq0[msk] * kinv_top / eig
        tp[:, 0, 0] = srf_bg_conc - tq0[0, 0] * kinv_top * h
        tq[:, msk] = tq0[msk] * np.exp(-eig * h)
        tp[:, msk] = tq[:, msk] * kinv_top / eig
    else:
        M = Lxm.shape[0]
        one = np.ones(M, np.complex128)
        zero = np.zeros(M, np.complex128)
        # --- K5: two IVPs (solver.py:220-226)
        p1, q1, P1, Q1 = ivp((one, zero), profiles, z, lv, Lxm, Lym, nthreads)
        p2, q2, P2, Q2 = ivp((zero, tq0[msk]), profiles, z, lv, Lxm, Lym, nthreads)
        # --- K6: shooting coefficient and combination (solver.py:228-235)
        alpha = -(q2 - Kz[nz - 1] * eig * p2) / (q1 - Kz[nz - 1] * eig * p1)
        if alpha_scale != 1.0:            # self-noise probe, never used in parity
            alpha = alpha * alpha_scale
        tp[:, msk] = alpha * P1 + P2
        tq[:, msk] = alpha * Q1 + Q2
        # --- K7: degenerate mode by trapezoid (solver.py:239-251)
        dz = np.diff(z)
        row = 0
        p00 = srf_bg_conc
        for i in range(nz - 1):
            if i in lv:
                tp[row, 0, 0] = p00
                row += 1
            p00 = p00 - tq0[0, 0] * dz[i] * (0.5 / Kz[i] + 0.5 / Kz[i + 1])
        if nz - 1 in lv:
            tp[row, 0, 0] = p00

    # --- K8: phase shift (solver.py:254-262)
    if footprint:
        sh = np.exp(1j * (Lx * (xm + halo) + Ly * (ym + halo)))
        tp = tp * sh
        tq = tq * sh
    elif xm ** 2 + ym ** 2 > 0.0:
        sh = np.exp(1j * (Lx * (xm - xmx / 2) + Ly * (ym - ymx / 2)))
        tp = tp * sh
        tq = tq * sh
    if return_spectral:
        return tp, tq

    # --- K9: scatter retained modes into the full spectrum (solver.py:265-278)
    fp = np.zeros((nlv, nfy, nfx), dtype=tp.dtype)
    fq = np.zeros((nlv, nfy, nfx), dtype=tq.dtype)
    fp[:, wy[:, None], wx[None, :]] = tp
    fq[:, wy[:, None], wx[None, :]] = tq

    # --- K10: transform back (solver.py:280-287)
    if footprint:
        p = _fft.fft2(fp, norm="backward", workers=nthreads).real
        q = _fft.fft2(fq, norm="backward", workers=nthreads).real
    else:
        p = _fft.ifft2(fp, norm="forward", workers=nthreads).real
        q = _fft.ifft2(fq, norm="forward", workers=nthreads).real

    # --- K11/K12: crop and grid (solver.py:289-298)
    conc = p[:, py:nye - py, px:nxe - px]
    flx = q[:, py:nye - py, px:nxe - px]
    x = np.linspace(0, xmx, nx, endpoint=False)
    y = np.linspace(0, ymx, ny, endpoint=False)
    Z, Y, X = np.meshgrid(z[lv], y, x, indexing="ij")
    return (np.squeeze(X), np.squeeze(Y), np.squeeze(Z)), np.squeeze(conc), np.squeeze(flx)


def kappa(z, profiles, g, z_level):
    """Conditioning number kappa(z) of linear shooting."""
    u, v, Kx, Ky, Kz = profiles
    lx = 2.0 * np.pi / (g["dx"] * g["nxe"]) * (g["nlx"] / 2)
    ly = 2.0 * np.pi / (g["dy"] * g["nye"]) * (g["nly"] / 2)
    dz = np.diff(z)
    i = np.nonzero(z[:-1] < z_level)[0]
    arg = (Kx[i] * lx ** 2 + Ky[i] * ly ** 2 + 1j * (np.abs(u[i]) * lx + np.abs(v[i]) * ly)) / Kz[i]
    return float(np.sum(np.sqrt(arg).real * dz[i]))
