"""ctypes binding of libbldfm_b200.so (include/bldfm_b200.h).

There is no CPU compute fallback: if the library is missing or no CUDA device is usable, the
compute entry points raise (BldfmError / OSError) instead of silently running something else.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libbldfm_b200.so"

# status codes / flags (keep in sync with include/bldfm_b200.h)
OK = 0
ERR_ODD_MODES = -1
ERR_PRECISION = -2
ERR_INVALID = -3
ERR_CUDA = -4
ERR_CUFFT = -5
ERR_ALLOC = -6
ERR_ODD_PAD = -7
ERR_ANALYTIC_LEVELS = -8
ERR_LEVEL_RANGE = -9

FOOTPRINT = 0x001
ANALYTIC = 0x002
DOUBLE = 0x004
MARCH_FMA = 0x008
SRC_ON_DEVICE = 0x010
OUT_ON_DEVICE = 0x020
ASYNC = 0x040
FFT_LIBRARY = 0x080
FFT_FULL = 0x100
MARCH_FULL = 0x200
MARCH_AUTO = 0x400
DELIVER_F32 = 0x800
OUT_MAPPED = 0x1000
MARCH_SWEEP = 0x2000

# every symbol include/bldfm_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "bldfm_version", "bldfm_last_error_string", "bldfm_device_count", "bldfm_geometry_init",
    "bldfm_wavenumbers", "bldfm_output_is_f32", "bldfm_plan_create", "bldfm_plan_destroy",
    "bldfm_plan_stream", "bldfm_plan_synchronize", "bldfm_plan_launch_count",
    "bldfm_plan_set_profiling", "bldfm_plan_last_timings", "bldfm_plan_workspace_bytes",
    "bldfm_solve", "bldfm_solve_batched", "bldfm_solve_spectral", "bldfm_march",
    "bldfm_host_alloc", "bldfm_host_free", "bldfm_device_alloc", "bldfm_device_free",
    "bldfm_memcpy_d2h", "bldfm_memcpy_h2d", "bldfm_fp64_peak",
    "bldfm_solve_batched_measure", "bldfm_sharded_stage1", "bldfm_sharded_stage2", "bldfm_ipc_export", "bldfm_ipc_open", "bldfm_ipc_close",
    "bldfm_march_coverage",
    "bldfm_solve_batched_accumulate", "bldfm_kappa", "bldfm_sweep_admissible", "bldfm_auto_kappa_limit", "bldfm_plan_last_march_mode",
    "bldfm_device_memset", "bldfm_host_register", "bldfm_host_unregister",
    "bldfm_peer_signal", "bldfm_peer_wait", "bldfm_peer_status", "bldfm_plan_march_trace",
    "bldfm_plan_synchronize_previous", "bldfm_set_option", "bldfm_get_option",
]


class BldfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[bldfm_b200 {code}] {msg}")
        self.code = code
        self.msg = msg


class Geometry(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("px", C.c_int32), ("py", C.c_int32),
        ("nxe", C.c_int32), ("nye", C.c_int32), ("nlx", C.c_int32), ("nly", C.c_int32),
        ("nfx", C.c_int32), ("nfy", C.c_int32), ("clamped", C.c_int32), ("reserved", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("halo", C.c_double),
        ("xmax", C.c_double), ("ymax", C.c_double),
    ]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


_DP = C.POINTER(C.c_double)


class Problem(C.Structure):
    # the six profile pointers are declared void* (same ABI as const double*): plain integer
    # addresses can be assigned without a ctypes cast, which is what dominates the per-call cost
    _fields_ = [
        ("z", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("Kx", C.c_void_p), ("Ky", C.c_void_p),
        ("Kz", C.c_void_p),
        ("nz", C.c_int32), ("reserved", C.c_int32),
        ("xm", C.c_double), ("ym", C.c_double), ("srf_bg_conc", C.c_double),
    ]


class Timings(C.Structure):
    _fields_ = [("forward_ms", C.c_double), ("march_ms", C.c_double),
                ("inverse_ms", C.c_double), ("total_ms", C.c_double)]


_lib = None


def lib():
    """Load the shared library (raises OSError with build instructions if it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OSError(
            f"{LIB_PATH} not found: build it with `python -m bldfm_b200.build` "
            "(bldfm_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    GP = C.POINTER(Geometry)
    PP = C.POINTER(Problem)
    I64P = C.POINTER(C.c_int64)
    sig = {
        "bldfm_version": (C.c_char_p, []),
        "bldfm_last_error_string": (C.c_char_p, []),
        "bldfm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
        "bldfm_geometry_init": (C.c_int, [i32, i32, dbl, dbl, i32, i32, i32, dbl, GP]),
        "bldfm_wavenumbers": (C.c_int, [GP, _DP, _DP]),
        "bldfm_output_is_f32": (C.c_int, [C.c_int, dbl, dbl]),
        "bldfm_plan_create": (C.c_int, [GP, C.c_int, C.POINTER(vp)]),
        "bldfm_plan_destroy": (C.c_int, [vp]),
        "bldfm_plan_stream": (vp, [vp]),
        "bldfm_plan_synchronize": (C.c_int, [vp]),
        "bldfm_plan_launch_count": (i64, [vp]),
        "bldfm_plan_set_profiling": (C.c_int, [vp, C.c_int]),
        "bldfm_plan_last_timings": (C.c_int, [vp, C.POINTER(Timings)]),
        "bldfm_plan_workspace_bytes": (i64, [vp]),
        "bldfm_solve": (C.c_int, [vp, PP, I64P, i32, vp, C.c_int, vp, vp]),
        "bldfm_solve_batched": (C.c_int, [vp, i32, vp, I64P, i32, vp, C.c_int, vp, vp]),
        "bldfm_solve_spectral": (C.c_int, [vp, PP, I64P, i32, vp, C.c_int, vp, vp]),
        "bldfm_march": (C.c_int, [C.c_int, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, I64P,
                                  vp, vp, C.c_int, vp, vp, vp, vp]),
        "bldfm_march_coverage": (C.c_int, [C.POINTER(Geometry), i32, i32, i32, vp, C.POINTER(i64)]),
        "bldfm_host_alloc": (C.c_int, [i64, C.POINTER(vp)]),
        "bldfm_host_free": (C.c_int, [vp]),
        "bldfm_device_alloc": (C.c_int, [C.c_int, i64, C.POINTER(vp)]),
        "bldfm_device_free": (C.c_int, [C.c_int, vp]),
        "bldfm_memcpy_d2h": (C.c_int, [C.c_int, vp, vp, i64]),
        "bldfm_memcpy_h2d": (C.c_int, [C.c_int, vp, vp, i64]),
        "bldfm_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, _DP]),
        "bldfm_solve_batched_measure": (C.c_int, [vp, i32, vp, I64P, i32, vp, C.c_int, vp, vp, vp]),
        "bldfm_sharded_stage1": (C.c_int, [vp, PP, I64P, i32, vp, C.c_int, i32, i32, vp, vp, vp, vp]),
        "bldfm_sharded_stage2": (C.c_int, [vp, i32, C.c_int, i32, i32, vp, vp, vp, vp]),
        "bldfm_ipc_export": (C.c_int, [vp, C.c_char_p]),
        "bldfm_ipc_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(vp)]),
        "bldfm_ipc_close": (C.c_int, [C.c_int, vp]),
        "bldfm_solve_batched_accumulate": (C.c_int, [vp, i32, vp, I64P, i32, vp, C.c_int, vp, i32, vp, vp]),
        "bldfm_kappa": (C.c_int, [GP, PP, i32, _DP]),
        "bldfm_sweep_admissible": (C.c_int, [GP, PP, i32, C.POINTER(C.c_int32)]),
        "bldfm_auto_kappa_limit": (dbl, []),
        "bldfm_plan_last_march_mode": (C.c_int, [vp]),
        "bldfm_device_memset": (C.c_int, [C.c_int, vp, C.c_int, i64]),
        "bldfm_host_register": (C.c_int, [vp, i64]),
        "bldfm_host_unregister": (C.c_int, [vp]),
        "bldfm_peer_signal": (C.c_int, [vp, vp, i32, C.c_uint64]),
        "bldfm_peer_wait": (C.c_int, [vp, vp, i32, C.c_uint64, dbl]),
        "bldfm_peer_status": (C.c_int, [vp, C.POINTER(i32)]),
        "bldfm_plan_march_trace": (C.c_int, [vp, vp, i64, C.POINTER(i64)]),
        "bldfm_plan_synchronize_previous": (C.c_int, [vp]),
        "bldfm_set_option": (C.c_int, [C.c_char_p, i32]),
        "bldfm_get_option": (C.c_int, [C.c_char_p, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def set_option(name: str, value):
    """Override a tuning switch of the library at run time (``None``: back to the environment / default)."""
    check(lib().bldfm_set_option(name.encode(), -2**31 if value is None else int(value)))


def last_error() -> str:
    return lib().bldfm_last_error_string().decode("utf-8", "replace")


def check(rc: int):
    """Map a C status code to the exception the reference raises for the same condition."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_ODD_MODES, ERR_PRECISION, ERR_ODD_PAD, ERR_ANALYTIC_LEVELS):
        raise ValueError(msg)                      # solver.py:90-91, :187-188
    if rc == ERR_LEVEL_RANGE:
        raise IndexError(msg)                      # z[levels], solver.py:296
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_ALLOC:
        raise MemoryError(msg)
    raise BldfmError(rc, msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().bldfm_device_count(C.byref(n))
    if rc != OK:
        return 0
    return n.value


def geometry(shape, domain, modes, halo) -> Geometry:
    ny, nx = shape
    g = Geometry()
    check(lib().bldfm_geometry_init(int(nx), int(ny), float(domain[0]), float(domain[1]),
                                    int(modes[0]), int(modes[1]), 1 if halo is None else 0,
                                    0.0 if halo is None else float(halo), C.byref(g)))
    return g


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def dptr(a):
    return a.ctypes.data_as(_DP)


def _f64c(a):
    """float64 C-contiguous view/copy with a fast path for arrays that already are."""
    if type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def make_problem(z, profiles, meas_pt, srf_bg_conc):
    """Build a Problem struct; returns (struct, keepalive) -- keep `keepalive` referenced.

    z and the five profiles are packed into ONE [6][nz] float64 buffer: taking the address of a
    numpy array from Python costs ~2.5 us, six of them would dominate the call.
    """
    if len(profiles) != 5:
        raise ValueError("profiles must be (u, v, Kx, Ky, Kz)")
    n = len(z)
    try:
        # fast path: six float64 vectors of one length -> one C-level concatenate
        buf = np.concatenate((z, *profiles), dtype=np.float64, casting="unsafe")
        if buf.shape != (6 * n,):
            raise ValueError
    except (ValueError, TypeError):
        buf = np.empty((6, n), dtype=np.float64)
        try:
            buf[0] = z
            buf[1], buf[2], buf[3], buf[4], buf[5] = profiles
        except ValueError:
            raise ValueError("profiles must have the same length as z") from None
    base = buf.ctypes.data
    row = n * 8
    p = Problem(base, base + row, base + 2 * row, base + 3 * row, base + 4 * row, base + 5 * row, n, 0,
                float(meas_pt[0]), float(meas_pt[1]), float(srf_bg_conc))
    return p, buf


PROBLEM_DTYPE = np.dtype([("z", "<u8"), ("u", "<u8"), ("v", "<u8"), ("Kx", "<u8"), ("Ky", "<u8"), ("Kz", "<u8"),
                          ("nz", "<i4"), ("reserved", "<i4"), ("xm", "<f8"), ("ym", "<f8"), ("srf_bg_conc", "<f8")])
assert PROBLEM_DTYPE.itemsize == C.sizeof(Problem)


def problems_from_batch(batch, row_of_problem, xm, ym, srf_bg_conc=0.0):
    """``bldfm_problem[P]`` built without a Python loop: problem p uses row ``row_of_problem[p]`` of a
    ``pbl_model.ProfileBatch`` (``[B, 6, nzmax]`` buffer) and the measurement point ``(xm[p], ym[p])``.

    Returns ``(array, keepalive)``: a numpy structured array with the layout of ``Problem`` (pass
    ``array.ctypes.data`` as the ``bldfm_problem*``) and the objects that must stay referenced during the call.
    """
    buf = batch.buf
    if buf.dtype != np.float64 or not buf.flags.c_contiguous:
        raise ValueError("profile batch must be a C-contiguous float64 [B, 6, nzmax] buffer")
    rows = np.asarray(row_of_problem, dtype=np.int64)
    P = rows.shape[0]
    nzmax = buf.shape[2]
    base = buf.ctypes.data + rows.astype(np.uint64) * np.uint64(6 * nzmax * 8)
    arr = np.zeros(P, dtype=PROBLEM_DTYPE)
    for k, name in enumerate(("z", "u", "v", "Kx", "Ky", "Kz")):
        arr[name] = base + np.uint64(k * nzmax * 8)
    arr["nz"] = np.asarray(batch.nz, dtype=np.int64)[rows]
    arr["xm"] = xm
    arr["ym"] = ym
    arr["srf_bg_conc"] = srf_bg_conc
    return arr, (buf, arr)
