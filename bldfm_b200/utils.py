"""Small host helpers with the reference's names and semantics (src/bldfm/utils.py:7-92).

They produce inputs for / consume outputs of the hot path and stay on the host.
"""

from __future__ import annotations

import numpy as np


def compute_wind_fields(u_rot, wind_dir):
    """Meteorological (speed, direction-from) -> (u, v) components (utils.py:7-27)."""
    rad = np.deg2rad(wind_dir)
    return -u_rot * np.sin(rad), -u_rot * np.cos(rad)


def ideal_source(nxy, domain, src_loc=None, shape="diamond"):
    """Synthetic surface-flux field: "diamond", "circle" or Gaussian "point" (utils.py:30-77)."""
    nx, ny = nxy
    xmx, ymx = domain
    dx = xmx / nx
    xs, ys = (xmx / 2, ymx / 2) if src_loc is None else src_loc
    X, Y = np.meshgrid(np.linspace(0.0, xmx, nx), np.linspace(0.0, ymx, ny))
    q0 = np.zeros([ny, nx])
    if shape == "diamond":
        q0 = np.where(np.abs(X - xs) + np.abs(Y - ys) < xmx / 12, 1.0, 0.0)
    if shape == "circle":
        q0 = np.where(np.sqrt((X - xs) ** 2 + (Y - ys) ** 2) < xmx / 12, 1.0, 0.0)
    if shape == "point":
        sig = 4.0 * dx
        rsq = (X - xs) ** 2 + (Y - ys) ** 2
        q0 = np.exp(-rsq / 2.0 / sig**2) / sig / np.sqrt(2.0 * np.pi)
    return q0


def point_measurement(f, g):
    """Footprint-weighted flux at the tower: sum(f*g) (utils.py:80-92)."""
    return np.sum(f * g)
