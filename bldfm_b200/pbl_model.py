"""Host-side input producer: MOST / MOSTM / CONSTANT / OAAHOC vertical profiles.

Same signature and return structure as the reference's ``bldfm.pbl_model.vertical_profiles``
(src/bldfm/pbl_model.py:8-204) so that callers (interface.py:77-95, examples, tests) are unchanged.
It is O(nz) work on 1-D arrays and stays on the host; the hot path
consumes its output.  ``vertical_profiles_batch`` is the vectorised form used by the batched
drivers: one call for B met conditions, row by row bitwise equal to the scalar function
(tests/test_profiles_batch.py).
"""

from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger("bldfm.pbl_model")

KAPPA = 0.4  # von Karman constant (pbl_model.py:61)

_CLOSURES = ("MOST", "CONSTANT", "MOSTM", "OAAHOC")


def _bad_closure(closure):
    return ValueError(
        f"Invalid closure type: {closure}. "
        "Supported closures are 'MOST', 'CONSTANT', and 'OAAHOC'."
    )


def psi(x):
    """Integrated Businger-Dyer stability correction for momentum (pbl_model.py:207-230)."""
    x = np.asarray(x, dtype=np.float64)
    stable = x > 0.0
    # (1-16x)^(1/4) through the complex power like the reference, real part taken -- evaluated only where
    # it is used (elementwise, so the values are the same; the complex power is the expensive part)
    xi = np.full(x.shape, np.nan)
    if x.ndim == 0:
        if not stable:
            xi = np.power(1.0 - 16.0 * x, 0.25, dtype=complex).real
    else:
        xi[~stable] = np.power(1.0 - 16.0 * x[~stable], 0.25, dtype=complex).real
    unstable_val = (
        -2.0 * np.log(0.5 * (1.0 + xi))
        - np.log(0.5 * (1.0 + xi**2))
        + 2.0 * np.arctan(xi)
        - 0.5 * np.pi
    )
    return np.where(stable, 5.0 * x, unstable_val)


def phi(x):
    """Businger-Dyer stability function for the eddy diffusivity (pbl_model.py:233-250)."""
    x = np.asarray(x, dtype=np.float64)
    stable = x > 0.0
    if x.ndim == 0:
        return np.where(stable, 1.0 + 5.0 * x, np.power(1.0 - 16.0 * x, -0.5, dtype=complex).real)
    out = 1.0 + 5.0 * x
    out[~stable] = np.power(1.0 - 16.0 * x[~stable], -0.5, dtype=complex).real
    return out


def _surface_layer(closure, zm, absum, ustar, z0, mol, tke):
    """Resolve (ustar, z0) and closure constants (pbl_model.py:66-101)."""
    extra = {}
    if closure in ("CONSTANT", "MOST", "MOSTM"):
        if z0 is None:
            z0 = zm * np.exp(-KAPPA * absum / ustar + psi(zm / mol))
        elif ustar is None:
            ustar = absum * KAPPA / (np.log(zm / z0) + psi(zm / mol))
        else:
            raise ValueError("Either z0 or ustar must be provided.")
    elif closure == "OAAHOC":
        cl, cm, ch = 0.845, 0.0856, 0.204
        if tke is None:
            logger.warning("No tke provided. Setting TKE to 1.0.")
            tke = 1.0
        tke = np.array(tke)[..., np.newaxis]
        z0 = zm * np.exp(-cm * cl * absum * np.sqrt(tke) / ustar**2)
        extra = dict(cl=cl, cm=cm, ch=ch, tke=tke)
    else:
        raise _bad_closure(closure)
    return ustar, z0, extra


def vertical_profiles(
    n,
    meas_height,
    wind,
    ustar=None,
    z0=None,
    mol=1e9,
    prsc=1.0,
    closure="MOST",
    domain_height=None,
    stretch=None,
    z0_min=0.001,
    z0_max=2.0,
    tke=None,
):
    """Vertical grid and (u, v, Kx, Ky, Kz) profiles; see pbl_model.py:22-56 for the arguments.

    Returns ``z, (u, v, Kx, Ky, Kz)`` with ``z[0] = z0``, ``z[n] = meas_height`` and the grid
    continuing to ``domain_height`` (default ``2*meas_height``), exponentially stretched.
    """
    zm = meas_height
    um, vm = wind
    absum = np.sqrt(um**2 + vm**2)
    ustar, z0, extra = _surface_layer(closure, zm, absum, ustar, z0, mol, tke)

    h = 2.0 * meas_height if stretch is None else stretch            # pbl_model.py:105-108
    zmx = 2.0 * meas_height if domain_height is None else domain_height

    bb = zm / (np.exp(-z0 / h) - np.exp(-zm / h))                      # :115-116
    aa = bb * np.exp(-z0 / h)
    zetamx = aa - bb * np.exp(-zmx / h)
    dzeta = zm / n
    zeta = np.arange(0.0, np.squeeze(zetamx).item() + dzeta, dzeta)    # :127
    z = -h * np.log(-(zeta - aa) / bb)                                 # :129

    if closure == "CONSTANT":                                          # :132-138
        Km = KAPPA * ustar * zm / prsc
        u = um * np.ones(len(z))
        v = vm * np.ones(len(z))
        K = Km * np.ones(len(z))
        Kx = Ky = Kz = K
    elif closure in ("MOST", "MOSTM"):                                 # :140-164
        absu = ustar / KAPPA * (np.log(z / z0) + psi(z / mol))
        u = um / absum * absu
        v = vm / absum * absu
        K = KAPPA * ustar * z / phi(z / mol) / prsc
        if closure == "MOST":
            Kx = Ky = Kz = K
        else:
            Kx = K * v**2 / (u**2 + v**2)
            Ky = K * u**2 / (u**2 + v**2)
            Kz = K
    else:  # OAAHOC                                                    # :166-176
        cl, cm, ch, tk = extra["cl"], extra["cm"], extra["ch"], extra["tke"]
        absu = ustar**2 / cm / cl / np.sqrt(tk) * np.log(z / z0)
        u = um / absum * absu
        v = vm / absum * absu
        K = ch * cl * z * np.sqrt(tk)
        Kx = Ky = Kz = K

    if logger.isEnabledFor(logging.INFO):
        logger.info("Stats from vertical_profiles")
        logger.info("z0    = %.3f m", z[0])
        logger.info("ustar = %.3f m s-1", ustar)
        logger.info("umax  = %.3f m s-1, vmax = %.3f m s-1, Kzmax = %.3f m2 s-1",
                    np.max(u), np.max(v), np.max(Kz))
    return z, (u, v, Kx, Ky, Kz)


class ProfileBatch:
    """B vertical grids + profiles packed as one ``[B, 6, nzmax]`` float64 buffer (rows z, u, v, Kx, Ky, Kz;
    entries beyond ``nz[b]`` are zero) -- the layout ``_lib.problems_from_batch`` hands to the C side
    without touching the rows from Python again."""

    __slots__ = ("buf", "nz")

    def __init__(self, buf, nz):
        self.buf = buf
        self.nz = nz

    def __len__(self):
        return self.buf.shape[0]

    def row(self, b):
        """(z, (u, v, Kx, Ky, Kz)) of row b, as ``vertical_profiles`` returns them."""
        n = int(self.nz[b])
        r = self.buf[b]
        return r[0, :n], (r[1, :n], r[2, :n], r[3, :n], r[4, :n], r[5, :n])


def _pow2(x):
    # the scalar function squares numpy/Python SCALARS with `**2`, i.e. libm pow(); numpy's ARRAY power
    # takes an x*x fast path for the exponent 2, which differs from pow() in the last bit for ~0.1 % of
    # the arguments -- so the B scalars are squared one by one, like the scalar function does
    return np.array([v ** 2 for v in np.asarray(x, dtype=np.float64)], dtype=np.float64)


def compute_wind_fields_batch(u_rot, wind_dir):
    """``compute_wind_fields`` (utils.py:7-27) for B rows; elementwise the same operations."""
    rad = np.deg2rad(np.asarray(wind_dir, dtype=np.float64))
    u_rot = np.asarray(u_rot, dtype=np.float64)
    return -u_rot * np.sin(rad), -u_rot * np.cos(rad)


def vertical_profiles_batch(n, meas_height, wind, ustar=None, z0=None, mol=1e9, prsc=1.0, closure="MOST",
                            domain_height=None, stretch=None, tke=None) -> ProfileBatch:
    """``vertical_profiles`` (pbl_model.py:58-204) for B met conditions in one pass of array operations.

    ``wind = (um[B], vm[B])``; ``ustar``, ``z0``, ``mol``, ``tke`` are scalars or ``[B]`` arrays;
    ``meas_height`` a scalar or ``[B]``.  Every per-row scalar is computed with the same operations in the
    same order as the scalar function and every profile with the same elementwise array operations, so
    row b equals ``vertical_profiles(...)`` of that row bit for bit.  The number of levels may differ
    between rows (it depends on z0 through ``np.arange``, pbl_model.py:127): rows are zero-padded.
    """
    um = np.atleast_1d(np.asarray(wind[0], dtype=np.float64))
    vm = np.atleast_1d(np.asarray(wind[1], dtype=np.float64))
    B = um.shape[0]

    def vec(x):
        return None if x is None else np.broadcast_to(np.asarray(x, dtype=np.float64), (B,))

    zm, mol_v, ustar_v, z0_v = vec(meas_height), vec(mol), vec(ustar), vec(z0)
    absum = np.sqrt(_pow2(um) + _pow2(vm))
    sq_tke = None
    if closure in ("CONSTANT", "MOST", "MOSTM"):
        if z0_v is None:
            if ustar_v is None:
                raise ValueError("Either z0 or ustar must be provided.")
            z0_v = zm * np.exp(-KAPPA * absum / ustar_v + psi(zm / mol_v))
        elif ustar_v is None:
            ustar_v = absum * KAPPA / (np.log(zm / z0_v) + psi(zm / mol_v))
        else:
            raise ValueError("Either z0 or ustar must be provided.")
    elif closure == "OAAHOC":
        cl, cm, ch = 0.845, 0.0856, 0.204
        if tke is None:
            logger.warning("No tke provided. Setting TKE to 1.0.")
            tke = 1.0
        sq_tke = np.sqrt(vec(tke))
        z0_v = zm * np.exp(-cm * cl * absum * sq_tke / _pow2(ustar_v))
    else:
        raise _bad_closure(closure)

    h = 2.0 * zm if stretch is None else vec(stretch)
    zmx = 2.0 * zm if domain_height is None else vec(domain_height)
    e0 = np.exp(-z0_v / h)
    bb = zm / (e0 - np.exp(-zm / h))
    aa = bb * e0
    zetamx = aa - bb * np.exp(-zmx / h)
    dzeta = zm / n
    # np.arange(0.0, stop, dzeta): length ceil((stop - 0.0)/dzeta), values 0.0 + i*dzeta
    with np.errstate(invalid="ignore"):
        nz = np.maximum(np.ceil((zetamx + dzeta) / dzeta), 0.0).astype(np.int64)   # arange is empty for stop <= 0
    nzmax = int(nz.max())
    col = lambda a: a[:, None]                                               # noqa: E731
    zeta = np.arange(nzmax, dtype=np.float64)[None, :] * col(dzeta)
    buf = np.zeros((B, 6, nzmax), dtype=np.float64)
    pad = np.arange(nzmax)[None, :] >= col(nz)
    with np.errstate(invalid="ignore", divide="ignore"):
        z = -col(h) * np.log(-(zeta - col(aa)) / col(bb))
        if pad.any():
            z = np.where(pad, col(zm), z)          # keep the padding finite for the operations below
        if closure == "CONSTANT":
            Km = KAPPA * ustar_v * zm / prsc
            ones = np.ones((B, nzmax))
            u, v = col(um) * ones, col(vm) * ones
            Kx = Ky = Kz = col(Km) * ones
        elif closure in ("MOST", "MOSTM"):
            zl = z / col(mol_v)
            absu = col(ustar_v / KAPPA) * (np.log(z / col(z0_v)) + psi(zl))
            u = col(um / absum) * absu
            v = col(vm / absum) * absu
            K = col(KAPPA * ustar_v) * z / phi(zl) / prsc
            if closure == "MOST":
                Kx = Ky = Kz = K
            else:
                Kx = K * v**2 / (u**2 + v**2)
                Ky = K * u**2 / (u**2 + v**2)
                Kz = K
        else:
            absu = col(_pow2(ustar_v) / cm / cl / sq_tke) * np.log(z / col(z0_v))
            u = col(um / absum) * absu
            v = col(vm / absum) * absu
            Kx = Ky = Kz = col(ch * cl) * z * col(sq_tke) if np.ndim(ch * cl) else (ch * cl) * z * col(sq_tke)
    for k, a in enumerate((z, u, v, Kx, Ky, Kz)):
        buf[:, k, :] = a
    if pad.any():
        buf[np.broadcast_to(pad[:, None, :], buf.shape)] = 0.0
    return ProfileBatch(buf, nz)
