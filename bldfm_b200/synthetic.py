"""Synthetic benchmark inputs: a diurnal met timeseries and tower layouts.

Value-for-value restatement of the two generators the reference ships for demos and tests
(src/bldfm/synthetic.py:12-109 ``generate_synthetic_timeseries``, :112-185 ``generate_towers_grid``): same
random stream (``numpy.random.default_rng(seed)``, four normal draws in the order ustar, L, wind speed, wind
direction), same formulas, so that BASELINE config 4 (8 towers x 1440 half-hourly steps) is built from the
inputs SURVEY.md 8(d) names without the reference being importable.  Pinned bitwise by
the golden fixture tests/golden/synthetic.npz.
"""

from __future__ import annotations

from datetime import datetime, timedelta
from typing import List, Optional, Tuple

import numpy as np


def generate_synthetic_timeseries(n_timesteps: int = 48, dt_minutes: int = 30, start_time: str = "2024-01-01T00:00",
                                  ustar_range: Tuple[float, float] = (0.1, 0.8),
                                  mol_range: Tuple[float, float] = (-500.0, 500.0),
                                  wind_speed_range: Tuple[float, float] = (1.0, 8.0), wind_dir_mean: float = 270.0,
                                  wind_dir_std: float = 30.0, seed: Optional[int] = None) -> dict:
    """Dict with the keys of the met schema (ustar, mol, wind_speed, wind_dir as lists, timestamps)."""
    rng = np.random.default_rng(seed)
    hours = np.arange(n_timesteps) * dt_minutes / 60.0
    phase = 2.0 * np.pi * (hours % 24.0) / 24.0          # 0 at midnight, pi at noon
    day = np.sin(phase - np.pi / 2)                       # +1 at noon, -1 at midnight

    def swing(lo, hi, amp_frac, sigma):
        mean, amp = 0.5 * (lo + hi), amp_frac * (hi - lo)
        series = mean + amp * day
        series += rng.normal(0, sigma(amp), n_timesteps)
        return np.clip(series, lo, hi)

    ustar = swing(*ustar_range, 0.5, lambda amp: 0.05 * amp)
    # Obukhov length: unstable (negative) by day, stable by night, strongest around noon / midnight
    mol = np.where(day > 0, -1.0, 1.0) * (50.0 + 450.0 * np.abs(np.cos(phase - np.pi / 2)))
    mol += rng.normal(0, 20.0, n_timesteps)
    mol = np.clip(mol, mol_range[0], mol_range[1])
    wind_speed = swing(*wind_speed_range, 0.3, lambda amp: 0.5)
    wind_dir = (wind_dir_mean + rng.normal(0, wind_dir_std, n_timesteps)) % 360.0
    t0 = datetime.fromisoformat(start_time)
    stamps = [(t0 + timedelta(minutes=i * dt_minutes)).isoformat() for i in range(n_timesteps)]
    return {"ustar": ustar.tolist(), "mol": mol.tolist(), "wind_speed": wind_speed.tolist(),
            "wind_dir": wind_dir.tolist(), "timestamps": stamps}


def generate_towers_grid(n_towers: int = 4, center_lat: float = 50.9500, center_lon: float = 11.5860,
                         spacing_m: float = 500.0, z_m: float = 10.0, layout: str = "grid",
                         seed: Optional[int] = None) -> List[dict]:
    """Tower dicts (name, lat, lon, z_m) on a square grid, a transect or at random around a centre."""
    rng = np.random.default_rng(seed)
    per_m_lat = 1.0 / 111_320.0
    per_m_lon = 1.0 / (111_320.0 * np.cos(np.radians(center_lat)))
    if layout == "grid":
        side = int(np.ceil(np.sqrt(n_towers)))
        mid = (side - 1) / 2
        offsets = [((j - mid) * spacing_m, (i - mid) * spacing_m) for i in range(side) for j in range(side)][:n_towers]
    elif layout == "transect":
        offsets = [((i - (n_towers - 1) / 2) * spacing_m, 0.0) for i in range(n_towers)]
    elif layout == "random":
        half = spacing_m * np.sqrt(n_towers) / 2
        offsets = [(rng.uniform(-half, half), rng.uniform(-half, half)) for _ in range(n_towers)]
    else:
        raise ValueError(f"Unknown layout: {layout}. Use 'grid', 'transect', or 'random'.")
    return [{"name": f"tower_{chr(65 + i)}" if i < 26 else f"tower_{i}",
             "lat": round(center_lat + dy * per_m_lat, 6), "lon": round(center_lon + dx * per_m_lon, 6), "z_m": z_m}
            for i, (dx, dy) in enumerate(offsets)]
