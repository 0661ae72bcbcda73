"""Drop-in host mirror of ``bldfm.solver`` (src/bldfm/solver.py) on top of libbldfm_b200.

``steady_state_transport_solver`` keeps the reference signature, argument meaning, error messages,
output shapes/dtypes and cache behaviour (solver.py:16-304); the body between the argument checks
and the grid construction is ONE call into the CUDA library.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib
from . import config
from ._pinned import pool as _pinned_pool
from .fft_manager import get_fft_manager

logger = logging.getLogger("bldfm.solver")


def _flags(footprint, analytic, precision):
    if precision == "double":
        f = _lib.DOUBLE
    elif precision == "single":
        f = 0
    else:                                                                  # solver.py:187-188
        raise ValueError("precision must be single (default) or double.")
    if footprint:
        f |= _lib.FOOTPRINT
    if analytic:
        f |= _lib.ANALYTIC
    if config.MARCH_MODE == "fma":
        f |= _lib.MARCH_FMA
    elif config.MARCH_MODE != "exact":
        raise ValueError("bldfm_b200.config.MARCH_MODE must be 'exact' or 'fma'")
    if config.FFT_LIBRARY:
        f |= _lib.FFT_LIBRARY
    if config.FFT_FULL:
        f |= _lib.FFT_FULL
    if config.MARCH_FULL:
        f |= _lib.MARCH_FULL
    return f


_GEOM_CACHE = {}


def _geometry(shape, domain, modes, halo):
    key = (tuple(shape), float(domain[0]), float(domain[1]), int(modes[0]), int(modes[1]),
           None if halo is None else float(halo))
    g = _GEOM_CACHE.get(key)
    if g is None:
        g = _lib.geometry(shape, domain, modes, halo)
        if len(_GEOM_CACHE) < 256:
            _GEOM_CACHE[key] = g
    return g


_LEVEL_CACHE = {}


def _levels_array(levels):
    """(levels as the reference indexes with them, int64 copy for the C side)   solver.py:102-103"""
    if type(levels) is int:
        hit = _LEVEL_CACHE.get(levels)
        if hit is None:
            lv = np.array([levels])
            hit = (lv, np.ascontiguousarray(lv, dtype=np.int64))
            if len(_LEVEL_CACHE) < 1024:
                _LEVEL_CACHE[levels] = hit
        return hit
    lv = np.array([levels]) if np.ndim(levels) == 0 else np.asarray(levels)
    if lv.dtype.kind not in "iub":
        # numpy refuses float arrays as indices (z[levels], solver.py:296)
        raise IndexError("arrays used as indices must be of integer (or boolean) type")
    return lv, np.ascontiguousarray(lv, dtype=np.int64)


_LEVEL_PTR = {}


def _levels_ptr(lv64):
    """ctypes pointer to an int64 level array; memoised for the cached scalar-level arrays."""
    key = id(lv64)
    hit = _LEVEL_PTR.get(key)
    if hit is not None and hit[0] is lv64:
        return hit[1]
    p = lv64.ctypes.data_as(C.POINTER(C.c_int64))
    if len(_LEVEL_PTR) < 1024 and any(v[1] is lv64 for v in _LEVEL_CACHE.values()):
        _LEVEL_PTR[key] = (lv64, p)
    return p


_XY_CACHE = {}


def make_grid(z, lv, domain, nx, ny):
    """(X, Y, Z) of solver.py:293-298.  Values/shapes as the reference; see config.GRID_COPY."""
    xmx, ymx = domain
    zl = np.asarray(z)[lv]
    if config.GRID_COPY:
        x = np.linspace(0, xmx, nx, endpoint=False)
        y = np.linspace(0, ymx, ny, endpoint=False)
        Z, Y, X = np.meshgrid(zl, y, x, indexing="ij")
        return np.squeeze(X), np.squeeze(Y), np.squeeze(Z)
    nlv = len(zl)
    key = (float(xmx), float(ymx), nx, ny, nlv)
    xy = _XY_CACHE.get(key)
    if xy is None:
        x = np.linspace(0, xmx, nx, endpoint=False)
        y = np.linspace(0, ymx, ny, endpoint=False)
        x.setflags(write=False)
        y.setflags(write=False)
        shape = (nlv, ny, nx)
        xy = (np.squeeze(np.broadcast_to(x[None, None, :], shape)),
              np.squeeze(np.broadcast_to(y[None, :, None], shape)))
        if len(_XY_CACHE) < 64:
            _XY_CACHE[key] = xy
    Z = np.squeeze(np.broadcast_to(zl[:, None, None], (nlv, ny, nx)))
    return xy[0], xy[1], Z


def steady_state_transport_solver(
    srf_flx,
    z,
    profiles,
    domain,
    levels,
    modes=(512, 512),
    meas_pt=(0.0, 0.0),
    srf_bg_conc=0.0,
    footprint=False,
    analytic=False,
    halo=None,
    precision="single",
    cache=None,
):
    """Steady-state advection-diffusion solve on the GPU; see solver.py:31-74 for the arguments.

    Returns ``((X, Y, Z), conc, flx)`` exactly as the reference does (np.squeeze'd; float32 fields
    only for precision="single" without a phase shift, float64 otherwise).
    """
    if cache is not None and footprint:                                    # solver.py:77-80
        cached = cache.get(z, profiles, domain, modes, meas_pt, halo, precision)
        if cached is not None:
            return cached

    q0 = np.asarray(srf_flx)
    if q0.ndim != 2:
        raise ValueError("srf_flx must be a 2D array")
    ny, nx = q0.shape
    geom = _geometry(q0.shape, domain, modes, halo)                        # raises for odd modes (:90-91)
    if geom.clamped:                                                       # solver.py:122-127
        logger.info("Warning: Number of Fourier modes must not exeed number of grid cells.")
        logger.info("Setting both equal.")
    flags = _flags(footprint, analytic, precision)
    lv, lv64 = _levels_array(levels)
    nlv = len(lv64)

    prob, keep = _lib.make_problem(z, profiles, meas_pt, srf_bg_conc)
    # dtype rule of solver.py:177-185,254-262 (the C side applies the same one: bldfm_output_is_f32)
    f32 = precision == "single" and not footprint and not (prob.xm * prob.xm + prob.ym * prob.ym > 0.0)
    dt = np.float32 if f32 else np.float64
    both, pinned = _pinned_pool.empty2((2, nlv, ny, nx), dt)
    conc, flx = both[0], both[1]
    src = None
    if not footprint:
        src = _lib.as_f64(q0)

    plan = get_fft_manager().plan(geom)
    L = _lib.lib()
    # one address lookup for both outputs (taking an array's address from Python costs microseconds);
    # page-locked outputs: enqueue only, build the grid while the GPU works, then wait for the results
    base = both.ctypes.data
    rc = L.bldfm_solve(
        plan, C.byref(prob), _levels_ptr(lv64), nlv,
        None if src is None else _lib.ptr(src), flags | (_lib.ASYNC if pinned else 0), base, base + conc.nbytes)
    _lib.check(rc)
    try:
        grid = make_grid(z, lv, domain, nx, ny)
        result = (grid, np.squeeze(conc), np.squeeze(flx))
    finally:
        # never leave with the copy into `both` still in flight: the buffer returns to the pool when dropped
        if pinned:
            _lib.check(L.bldfm_plan_synchronize(plan))
    del keep

    if cache is not None and footprint:                                    # solver.py:301-302
        cache.put(z, profiles, domain, modes, meas_pt, halo, precision, *result)
    return result


def solve_batched(srf_flx, zs, profiles_list, domain, levels, modes=(512, 512), meas_pts=None,
                  srf_bg_conc=0.0, footprint=False, analytic=False, halo=None, precision="single",
                  wait=True):
    """Many (tower, met) conditions in ONE launch (C entry point ``bldfm_solve_batched``).

    ``zs[b]``, ``profiles_list[b]``, ``meas_pts[b]`` describe problem b; everything else is shared.
    Problems with byte-identical (z, profiles) share one vertical march.  Returns
    ``(conc, flx)`` of shape ``[B, nlv, ny, nx]`` (float32 only for precision="single" without any
    phase shift, like the single-problem solver).

    ``wait=False`` only enqueues the work: the returned arrays (pinned host memory) are valid after
    ``synchronize()``; the device->host copy of this batch then overlaps the kernels of the next one.
    Keep the returned arrays referenced until then -- a dropped array hands its buffer back to the pool.
    """
    q0 = np.asarray(srf_flx)
    ny, nx = q0.shape
    B = len(zs)
    if B < 1:
        raise ValueError("solve_batched needs at least one problem")
    if meas_pts is None:
        meas_pts = [(0.0, 0.0)] * B
    geom = _geometry(q0.shape, domain, modes, halo)
    if geom.clamped:
        logger.info("Warning: Number of Fourier modes must not exeed number of grid cells.")
        logger.info("Setting both equal.")
    flags = _flags(footprint, analytic, precision)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    bg = srf_bg_conc if np.ndim(srf_bg_conc) else [srf_bg_conc] * B
    probs, keep = [], []
    for b in range(B):
        p, k = _lib.make_problem(zs[b], profiles_list[b], meas_pts[b], bg[b])
        probs.append(p)
        keep.append(k)
    f32 = all(bool(_lib.lib().bldfm_output_is_f32(flags, p.xm, p.ym)) for p in probs)
    dt = np.float32 if f32 else np.float64
    conc = _pinned_pool.empty((B, nlv, ny, nx), dt)
    flx = _pinned_pool.empty((B, nlv, ny, nx), dt)
    src = None if footprint else _lib.as_f64(q0)
    parr = (_lib.Problem * B)(*probs)
    plan = get_fft_manager().plan(geom)
    if not wait:
        flags |= _lib.ASYNC
    _lib.check(_lib.lib().bldfm_solve_batched(
        plan, B, parr, lv64.ctypes.data_as(C.POINTER(C.c_int64)), nlv,
        None if src is None else _lib.ptr(src), flags, _lib.ptr(conc), _lib.ptr(flx)))
    del keep
    return conc, flx


def synchronize():
    """Wait for every enqueued solve / result copy on this process's plans (after ``wait=False``)."""
    mgr = get_fft_manager()
    for h in list(mgr._plans.values()):
        _lib.check(_lib.lib().bldfm_plan_synchronize(h))


def measure_batched(weight, srf_flx, zs, profiles_list, domain, levels, modes=(512, 512), meas_pts=None,
                    srf_bg_conc=0.0, footprint=False, analytic=False, halo=None, precision="single", wait=True):
    """``point_measurement`` (utils.py:80-92) fused on the device for a batch of solves.

    Returns ``(conc_w, flx_w)`` of shape ``[B, nlv]``: ``sum(conc[b, l] * weight)`` and
    ``sum(flx[b, l] * weight)``.  With ``footprint=True`` and ``weight`` = surface flux map, ``flx_w`` is
    the flux each tower measures -- 8 bytes per footprint cross PCIe instead of the 4 MB field.

    ``wait=False`` only enqueues the batch: the returned (pinned) arrays are valid after ``synchronize()``,
    and the host can prepare the next batch while this one runs.
    """
    q0 = np.asarray(srf_flx)
    B = len(zs)
    if meas_pts is None:
        meas_pts = [(0.0, 0.0)] * B
    w = _lib.as_f64(weight)
    if w.shape != q0.shape:
        raise ValueError("weight must have the shape of srf_flx")
    geom = _geometry(q0.shape, domain, modes, halo)
    flags = _flags(footprint, analytic, precision)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    bg = srf_bg_conc if np.ndim(srf_bg_conc) else [srf_bg_conc] * B
    probs, keep = [], []
    for b in range(B):
        p, k = _lib.make_problem(zs[b], profiles_list[b], meas_pts[b], bg[b])
        probs.append(p)
        keep.append(k)
    if wait:
        conc_w = np.empty((B, nlv))
        flx_w = np.empty((B, nlv))
    else:
        conc_w = _pinned_pool.empty((B, nlv), np.float64)
        flx_w = _pinned_pool.empty((B, nlv), np.float64)
        flags |= _lib.ASYNC
    src = None if footprint else _lib.as_f64(q0)
    parr = (_lib.Problem * B)(*probs)
    plan = get_fft_manager().plan(geom)
    _lib.check(_lib.lib().bldfm_solve_batched_measure(
        plan, B, parr, lv64.ctypes.data_as(C.POINTER(C.c_int64)), nlv,
        None if src is None else _lib.ptr(src), flags, _lib.ptr(w), _lib.ptr(conc_w), _lib.ptr(flx_w)))
    del keep
    return conc_w, flx_w


def spectral_fields(srf_flx, z, profiles, domain, levels, modes=(512, 512), meas_pt=(0.0, 0.0),
                    srf_bg_conc=0.0, footprint=False, analytic=False, halo=None,
                    precision="single"):
    """Parity hook: the combined, phase-shifted spectra (tfftp, tfftq) [nlv, nly, nlx] complex128
    as they stand before solver.py:265."""
    q0 = np.asarray(srf_flx)
    geom = _lib.geometry(q0.shape, domain, modes, halo)
    flags = _flags(footprint, analytic, precision)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    prob, keep = _lib.make_problem(z, profiles, meas_pt, srf_bg_conc)
    tp = np.empty((nlv, geom.nly, geom.nlx), dtype=np.complex128)
    tq = np.empty((nlv, geom.nly, geom.nlx), dtype=np.complex128)
    src = None if footprint else _lib.as_f64(q0)
    plan = get_fft_manager().plan(geom)
    _lib.check(_lib.lib().bldfm_solve_spectral(
        plan, C.byref(prob), lv64.ctypes.data_as(C.POINTER(C.c_int64)), nlv,
        None if src is None else _lib.ptr(src), flags, _lib.ptr(tp), _lib.ptr(tq)))
    del keep
    return tp, tq


def ivp_solver(fftpq, profiles, z, levels, Lx, Ly):
    """``ivp_solver`` of solver.py:307-374 on the GPU: returns (fftp_top, fftq_top, fftp, fftq)."""
    p0 = np.ascontiguousarray(fftpq[0], dtype=np.complex128).ravel()
    q0 = np.ascontiguousarray(fftpq[1], dtype=np.complex128).ravel()
    Lx = _lib.as_f64(Lx).ravel()
    Ly = _lib.as_f64(Ly).ravel()
    M = p0.shape[0]
    if not (q0.shape[0] == Lx.shape[0] == Ly.shape[0] == M):
        raise ValueError("fftpq, Lx and Ly must have the same number of modes")
    z = _lib.as_f64(z)
    u, v, Kx, Ky, Kz = (_lib.as_f64(a) for a in profiles)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    p_top = np.empty(M, np.complex128)
    q_top = np.empty(M, np.complex128)
    P = np.zeros((nlv, M), np.complex128)
    Q = np.zeros((nlv, M), np.complex128)
    flags = _lib.MARCH_FMA if config.MARCH_MODE == "fma" else 0
    _lib.check(_lib.lib().bldfm_march(
        config.DEVICE, M, _lib.ptr(p0), _lib.ptr(q0), len(z), _lib.ptr(z), _lib.ptr(u),
        _lib.ptr(v), _lib.ptr(Kx), _lib.ptr(Ky), _lib.ptr(Kz), nlv,
        lv64.ctypes.data_as(C.POINTER(C.c_int64)), _lib.ptr(Lx), _lib.ptr(Ly), flags,
        _lib.ptr(p_top), _lib.ptr(q_top), _lib.ptr(P), _lib.ptr(Q)))
    return p_top, q_top, P, Q
