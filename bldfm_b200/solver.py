"""Drop-in host mirror of ``bldfm.solver`` (src/bldfm/solver.py) on top of libbldfm_b200.

``steady_state_transport_solver`` keeps the reference signature, argument meaning, error messages,
output shapes/dtypes and cache behaviour (solver.py:16-304); the body between the argument checks
and the grid construction is ONE call into the CUDA library.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import logging
import mmap as _mmap

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
    elif config.MARCH_MODE == "sweep":
        f |= _lib.MARCH_SWEEP
    elif config.MARCH_MODE == "auto":
        f |= _lib.MARCH_AUTO
    elif config.MARCH_MODE != "exact":
        raise ValueError("bldfm_b200.config.MARCH_MODE must be 'exact', 'fma', 'sweep' or 'auto'")
    if config.DELIVER_FLOAT32:
        f |= _lib.DELIVER_F32
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
_Z_CACHE = {}
_VIEW_CACHE = {}
_COW_LIMIT = 64 << 20        # bytes per grid array up to which the copy-on-write mapping is used
_COW_CACHE_BYTES = 256 << 20  # memory files kept for recurring (geometry, level heights) combinations


class _CowBundle:
    """Constant arrays kept back to back in ONE anonymous memory file; ``views()`` maps the file PRIVATELY once and
    returns fresh, writable, independent numpy arrays over it for the price of a single mmap call (microseconds) --
    the pages are shared with the file until the caller actually writes to them."""

    def __init__(self, arrays):
        import os
        page = _mmap.PAGESIZE
        self.specs = []
        off = 0
        blobs = []
        for a in arrays:
            a = np.ascontiguousarray(a)
            self.specs.append((a.shape, a.dtype, off))
            blobs.append((off, a.tobytes()))
            off += -(-a.nbytes // page) * page
        self.nbytes = max(off, page)
        self.fd = os.memfd_create("bldfm_b200_grid", 0)
        os.ftruncate(self.fd, self.nbytes)
        with _mmap.mmap(self.fd, self.nbytes) as mm:
            for o, b in blobs:
                mm[o:o + len(b)] = b

    def views(self):
        mm = _mmap.mmap(self.fd, self.nbytes, flags=_mmap.MAP_PRIVATE, prot=_mmap.PROT_READ | _mmap.PROT_WRITE)
        return tuple(np.ndarray(shape, dtype, mm, off) for shape, dtype, off in self.specs)

    def __del__(self):
        try:
            import os
            os.close(self.fd)
        except Exception:
            pass


def make_grid(z, lv, domain, nx, ny, mode=None):
    """(X, Y, Z) of solver.py:293-298: same values and shapes as the reference's ``np.meshgrid`` + squeeze.

    ``config.GRID_COPY`` selects how they are produced:
      "cow" (default)  writable and independent like the reference's arrays: X and Y are private
                       copy-on-write mappings of a per-geometry constant, Z is filled afresh (that fill is
                       done by the caller while the GPU works);  arrays above 64 MB fall back to "0"
      "1"              ``np.meshgrid`` exactly as the reference (3 full arrays written per call)
      "0"              zero-copy READ-ONLY broadcast views (batched drivers; fastest)
    """
    mode = config.GRID_COPY if mode is None else mode
    xmx, ymx = domain
    zl = np.asarray(z)[lv]
    if mode == "0":
        # read-only views are immutable: one tuple per (geometry, level heights) serves every task that asks
        vkey = (float(xmx), float(ymx), nx, ny, zl.tobytes())
        hit = _VIEW_CACHE.get(vkey)
        if hit is not None:
            return hit
    if mode in (True, "1"):
        x = np.linspace(0, xmx, nx, endpoint=False)
        y = np.linspace(0, ymx, ny, endpoint=False)
        Z, Y, X = np.meshgrid(zl, y, x, indexing="ij")
        return np.squeeze(X), np.squeeze(Y), np.squeeze(Z)
    nlv = len(zl)
    cow = mode == "cow" and nlv * ny * nx * 8 <= _COW_LIMIT
    if cow:
        # footprints are taken at z[n] == meas_height, so the same level heights recur: X, Y and Z then come
        # out of one private mapping; new heights get X, Y from their bundle and a freshly filled Z
        zkey = (float(xmx), float(ymx), nx, ny, zl.tobytes())
        full = _Z_CACHE.get(zkey)
        if full is not None:
            return full.views()
    key = (float(xmx), float(ymx), nx, ny, nlv, cow)
    xy = _XY_CACHE.get(key)
    if xy is None:
        x = np.linspace(0, xmx, nx, endpoint=False)
        y = np.linspace(0, ymx, ny, endpoint=False)
        shape = (nlv, ny, nx)
        X = np.squeeze(np.broadcast_to(x[None, None, :], shape))
        Y = np.squeeze(np.broadcast_to(y[None, :, None], shape))
        if cow:
            xy = _CowBundle((X, Y))
        else:
            x.setflags(write=False)
            y.setflags(write=False)
            xy = (X, Y)
        if len(_XY_CACHE) < 64:
            _XY_CACHE[key] = xy
    if cow:
        Z = np.empty((nlv, ny, nx))
        Z[...] = zl[:, None, None]
        Z = np.squeeze(Z)
        X, Y = xy.views()
        # keep at most 32 bundles / 256 MB of them (oldest out first)
        bundle = _CowBundle((X, Y, Z))
        while _Z_CACHE and (len(_Z_CACHE) >= 32 or
                            sum(b.nbytes for b in _Z_CACHE.values()) + bundle.nbytes > _COW_CACHE_BYTES):
            _Z_CACHE.pop(next(iter(_Z_CACHE)))
        if bundle.nbytes <= _COW_CACHE_BYTES:
            _Z_CACHE[zkey] = bundle
        return X, Y, Z
    Z = np.squeeze(np.broadcast_to(zl[:, None, None], (nlv, ny, nx)))
    out = (xy[0], xy[1], Z)
    if mode == "0":
        if len(_VIEW_CACHE) >= 4096:
            _VIEW_CACHE.clear()
        _VIEW_CACHE[vkey] = out
    return out


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
    dt = np.float32 if (f32 or config.DELIVER_FLOAT32) else np.float64
    both, pinned, base = _pinned_pool.empty3((2, nlv, ny, nx), dt)
    conc, flx = both[0], both[1]
    src = None
    if not footprint:
        src = _lib.as_f64(q0)

    plan = get_fft_manager().plan(geom)
    L = _lib.lib()
    # one address lookup for both outputs (taking an array's address from Python costs microseconds);
    # page-locked outputs: enqueue only, build the grid while the GPU works, then wait for the results
    rc = L.bldfm_solve(
        plan, C.byref(prob), _levels_ptr(lv64), nlv,
        None if src is None else _lib.ptr(src), flags | ((_lib.ASYNC | _lib.OUT_MAPPED) if pinned else 0), base,
        base + conc.nbytes)
    _lib.check(rc)
    try:
        grid = make_grid(z, lv, domain, nx, ny)
        if nlv == 1 and ny > 1 and nx > 1:
            result = (grid, conc[0], flx[0])                    # what np.squeeze gives, without the calls
        else:
            result = (grid, np.squeeze(conc), np.squeeze(flx))
    finally:
        # never leave with the copy into `both` still in flight: the buffer returns to the pool when dropped
        if pinned:
            _lib.check(L.bldfm_plan_synchronize(plan))
    del keep

    if cache is not None and footprint:                                    # solver.py:301-302
        cache.put(z, profiles, domain, modes, meas_pt, halo, precision, *result)
    return result


def _problem_array(zs, profiles_list, meas_pts, bg):
    """ctypes array of ``bldfm_problem`` built problem by problem (general, per-problem Python cost)."""
    probs, keep = [], []
    for b in range(len(zs)):
        p, k = _lib.make_problem(zs[b], profiles_list[b], meas_pts[b], bg[b])
        probs.append(p)
        keep.append(k)
    parr = (_lib.Problem * len(probs))(*probs)
    return parr, keep


def _problem_view(problems):
    """(address, count, numpy view with the fields of bldfm_problem) of a problem array given either as the
    structured array of ``_lib.problems_from_batch`` or as a ctypes ``Problem`` array."""
    if isinstance(problems, np.ndarray):
        return problems.ctypes.data, len(problems), problems
    view = np.frombuffer(problems, dtype=_lib.PROBLEM_DTYPE)
    return C.addressof(problems), len(problems), view


def _is_f32(flags, view):
    """Per problem: would the reference return float32 fields (solver.py:177-185,254-262)?"""
    if flags & (_lib.DOUBLE | _lib.FOOTPRINT):
        return np.zeros(len(view), dtype=bool)
    return ~(view["xm"] * view["xm"] + view["ym"] * view["ym"] > 0.0)


def solve_batched(srf_flx, zs=None, profiles_list=None, domain=None, levels=None, modes=(512, 512), meas_pts=None,
                  srf_bg_conc=0.0, footprint=False, analytic=False, halo=None, precision="single",
                  wait=True, problems=None, out=None, out_pinned=False):
    """Many (tower, met) conditions in ONE launch (C entry point ``bldfm_solve_batched``).

    ``zs[b]``, ``profiles_list[b]``, ``meas_pts[b]`` describe problem b; everything else is shared.
    Problems with byte-identical (z, profiles) share one vertical march.  Returns
    ``(conc, flx)`` of shape ``[B, nlv, ny, nx]`` (float32 only for precision="single" without any
    phase shift, like the single-problem solver).

    ``problems=(array, keepalive)`` from ``_lib.problems_from_batch`` replaces ``zs/profiles_list/meas_pts``
    (no per-problem Python work).  ``out=(conc, flx)`` are caller-provided destination arrays
    ``[B, nlv, ny, nx]`` of the result dtype (e.g. slices of a shared-memory segment); ``out_pinned=True``
    states that they are page-locked, which ``wait=False`` needs.

    ``wait=False`` only enqueues the work: the returned arrays (pinned host memory) are valid after
    ``synchronize()``; the device->host copy of this batch then overlaps the kernels of the next one.
    Keep the returned arrays referenced until then -- a dropped array hands its buffer back to the pool.
    """
    q0 = np.asarray(srf_flx)
    ny, nx = q0.shape
    geom = _geometry(q0.shape, domain, modes, halo)
    if geom.clamped:
        logger.info("Warning: Number of Fourier modes must not exeed number of grid cells.")
        logger.info("Setting both equal.")
    flags = _flags(footprint, analytic, precision)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    if problems is None:
        B = len(zs)
        if B < 1:
            raise ValueError("solve_batched needs at least one problem")
        if meas_pts is None:
            meas_pts = [(0.0, 0.0)] * B
        bg = srf_bg_conc if np.ndim(srf_bg_conc) else [srf_bg_conc] * B
        parr, keep = _problem_array(zs, profiles_list, meas_pts, bg)
    else:
        parr, keep = problems
    addr, B, view = _problem_view(parr)
    if B < 1:
        raise ValueError("solve_batched needs at least one problem")
    is32 = _is_f32(flags, view)
    f32 = bool(is32.all())
    if is32.any() and not f32:
        # precision="single" with some towers at exactly (0,0) and some shifted: the reference returns
        # float32 fields for the former and float64 for the latter (solver.py:177-185,254-262).  One launch
        # holds one dtype, so the two kinds go out as two sub-batches; the merged array is float64 (the
        # float32 values are represented exactly) and the drivers hand each task its reference dtype back.
        if out is None:
            out = (np.empty((B, nlv, ny, nx), np.float64), np.empty((B, nlv, ny, nx), np.float64))
        for want in (True, False):
            idx = np.nonzero(is32 == want)[0]
            sub = np.ascontiguousarray(view[idx])
            c, f = solve_batched(srf_flx, domain=domain, levels=levels, modes=modes, footprint=footprint,
                                 analytic=analytic, halo=halo, precision=precision, wait=True,
                                 problems=(sub, keep))
            out[0][idx] = c
            out[1][idx] = f
        return out
    if out is not None and out[0].dtype == np.float64:
        flags &= ~_lib.DELIVER_F32        # caller-provided float64 destinations (shared result segment) win
    dt = np.float32 if (f32 or (flags & _lib.DELIVER_F32)) else np.float64
    if out is None:
        conc, pinned_c = _pinned_pool.empty2((B, nlv, ny, nx), dt)
        flx, pinned_f = _pinned_pool.empty2((B, nlv, ny, nx), dt)
        pinned = pinned_c and pinned_f
    else:
        conc, flx = out
        for a in (conc, flx):
            if a.shape != (B, nlv, ny, nx) or a.dtype != dt or not a.flags.c_contiguous:
                raise ValueError(f"out arrays must be C-contiguous {dt.__name__} of shape {(B, nlv, ny, nx)}")
        pinned = bool(out_pinned)
    src = None if footprint else _lib.as_f64(q0)
    plan = get_fft_manager().plan(geom)
    if not wait and pinned:
        flags |= _lib.ASYNC           # pageable destinations are copied synchronously (see BLDFM_ASYNC)
    _lib.check(_lib.lib().bldfm_solve_batched(
        plan, B, addr, _levels_ptr(lv64), nlv,
        None if src is None else _lib.ptr(src), flags, conc.ctypes.data, flx.ctypes.data))
    del keep
    return conc, flx


def synchronize_previous():
    """After several ``wait=False`` batches: wait for the batch BEFORE the most recent one (its results are then
    valid) while the most recent one may still be computing."""
    mgr = get_fft_manager()
    for h in list(mgr._plans.values()):
        _lib.check(_lib.lib().bldfm_plan_synchronize_previous(h))


def synchronize():
    """Wait for every enqueued solve / result copy on this process's plans (after ``wait=False``)."""
    mgr = get_fft_manager()
    for h in list(mgr._plans.values()):
        _lib.check(_lib.lib().bldfm_plan_synchronize(h))


def measure_batched(weight, srf_flx, zs=None, profiles_list=None, domain=None, levels=None, modes=(512, 512),
                    meas_pts=None, srf_bg_conc=0.0, footprint=False, analytic=False, halo=None, precision="single",
                    wait=True, problems=None):
    """``point_measurement`` (utils.py:80-92) fused on the device for a batch of solves.

    Returns ``(conc_w, flx_w)`` of shape ``[B, nlv]``: ``sum(conc[b, l] * weight)`` and
    ``sum(flx[b, l] * weight)``.  With ``footprint=True`` and ``weight`` = surface flux map, ``flx_w`` is
    the flux each tower measures -- 8 bytes per footprint cross PCIe instead of the 4 MB field.

    ``wait=False`` only enqueues the batch: the returned (pinned) arrays are valid after ``synchronize()``,
    and the host can prepare the next batch while this one runs.
    """
    q0 = np.asarray(srf_flx)
    w = _lib.as_f64(weight)
    if w.shape != q0.shape:
        raise ValueError("weight must have the shape of srf_flx")
    geom = _geometry(q0.shape, domain, modes, halo)
    flags = _flags(footprint, analytic, precision)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    if problems is None:
        B = len(zs)
        if meas_pts is None:
            meas_pts = [(0.0, 0.0)] * B
        bg = srf_bg_conc if np.ndim(srf_bg_conc) else [srf_bg_conc] * B
        parr, keep = _problem_array(zs, profiles_list, meas_pts, bg)
    else:
        parr, keep = problems
    addr, B, view = _problem_view(parr)
    is32 = _is_f32(flags, view)
    if is32.any() and not is32.all():
        raise ValueError("measure_batched: precision='single' batch mixes towers at (0,0) with shifted ones; "
                         "pass them as separate batches")
    if wait:
        conc_w = np.empty((B, nlv))
        flx_w = np.empty((B, nlv))
    else:
        conc_w = _pinned_pool.empty((B, nlv), np.float64)
        flx_w = _pinned_pool.empty((B, nlv), np.float64)
        flags |= _lib.ASYNC
    src = None if footprint else _lib.as_f64(q0)
    plan = get_fft_manager().plan(geom)
    _lib.check(_lib.lib().bldfm_solve_batched_measure(
        plan, B, addr, _levels_ptr(lv64), nlv,
        None if src is None else _lib.ptr(src), flags, _lib.ptr(w), _lib.ptr(conc_w), _lib.ptr(flx_w)))
    del keep
    return conc_w, flx_w


class FieldAccumulator:
    """Device-resident sums of solved fields per slot -- time aggregation of footprints without moving the
    individual fields to the host (examples/timeseries_example.py:46 does
    ``np.mean([r["flx"] for r in results], axis=0)`` on the host).

    ``add(...)`` solves a batch and adds problem b's ``conc``/``flx`` to slot ``slot_of[b]`` in problem order;
    ``fetch()`` returns the sums ``[nslots, nlv, ny, nx]`` (float64); ``device_pointers()`` exposes them for a
    cross-rank reduction.
    """

    def __init__(self, shape, domain, levels, nslots, modes=(512, 512), footprint=False, analytic=False,
                 halo=None, precision="single"):
        self.shape = tuple(shape)
        self.domain, self.levels, self.modes, self.halo = domain, levels, modes, halo
        self.footprint, self.analytic, self.precision = footprint, analytic, precision
        self.geom = _geometry(self.shape, domain, modes, halo)
        _, self.lv64 = _levels_array(levels)
        self.nlv = len(self.lv64)
        self.nslots = int(nslots)
        self.device = config.DEVICE
        self.count = np.zeros(self.nslots, dtype=np.int64)
        ny, nx = self.shape
        self.nbytes = self.nslots * self.nlv * ny * nx * 8
        L = _lib.lib()
        self._ptr = []
        for _ in range(2):
            p = C.c_void_p()
            _lib.check(L.bldfm_device_alloc(self.device, self.nbytes, C.byref(p)))
            self._ptr.append(p.value)
        self.reset()

    def reset(self):
        for p in self._ptr:
            _lib.check(_lib.lib().bldfm_device_memset(self.device, p, 0, self.nbytes))
        self.count[:] = 0

    def device_pointers(self):
        return tuple(self._ptr)

    def add(self, srf_flx, slot_of, problems=None, zs=None, profiles_list=None, meas_pts=None, srf_bg_conc=0.0):
        flags = _flags(self.footprint, self.analytic, self.precision) | _lib.ASYNC
        if problems is None:
            B = len(zs)
            if meas_pts is None:
                meas_pts = [(0.0, 0.0)] * B
            bg = srf_bg_conc if np.ndim(srf_bg_conc) else [srf_bg_conc] * B
            parr, keep = _problem_array(zs, profiles_list, meas_pts, bg)
        else:
            parr, keep = problems
        addr, B, view = _problem_view(parr)
        is32 = _is_f32(flags, view)
        if is32.any() and not is32.all():
            raise ValueError("FieldAccumulator.add: batch mixes float32 and float64 tasks; add them separately")
        slots = np.ascontiguousarray(slot_of, dtype=np.int32)
        if slots.shape != (B,):
            raise ValueError("slot_of must have one entry per problem")
        src = None if self.footprint else _lib.as_f64(np.asarray(srf_flx))
        plan = get_fft_manager().plan(self.geom)
        self._plan = plan
        _lib.check(_lib.lib().bldfm_solve_batched_accumulate(
            plan, B, addr, _levels_ptr(self.lv64), self.nlv, None if src is None else _lib.ptr(src), flags,
            _lib.ptr(slots), self.nslots, self._ptr[0], self._ptr[1]))
        np.add.at(self.count, slots[slots >= 0], 1)
        del keep

    def synchronize(self):
        plan = getattr(self, "_plan", None)
        if plan is not None:
            _lib.check(_lib.lib().bldfm_plan_synchronize(plan))

    def fetch(self):
        """(conc_sum, flx_sum) ``[nslots, nlv, ny, nx]`` float64 on the host (synchronises)."""
        L = _lib.lib()
        self.synchronize()
        ny, nx = self.shape
        out = []
        for p in self._ptr:
            a = np.empty((self.nslots, self.nlv, ny, nx), np.float64)
            _lib.check(L.bldfm_memcpy_d2h(self.device, _lib.ptr(a), p, self.nbytes))
            out.append(a)
        return tuple(out)

    def close(self):
        for p in self._ptr:
            _lib.lib().bldfm_device_free(self.device, C.c_void_p(p))
        self._ptr = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
