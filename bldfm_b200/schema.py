"""Minimal configuration objects with the attribute names the reference's drivers read
(src/bldfm/config_parser.py:23-182).  The YAML parser, validation messages and lat/lon handling of
the reference are out of scope -- the reference's own ``BLDFMConfig``
objects can be passed to ``bldfm_b200.interface`` unchanged (duck typing); these classes exist so
that the drivers can be used and tested without the reference installed.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union


@dataclass
class Tower:
    name: str
    z_m: float
    x: float = 0.0
    y: float = 0.0
    lat: float = 0.0
    lon: float = 0.0


@dataclass
class Domain:
    nx: int
    ny: int
    xmax: float
    ymax: float
    nz: int
    modes: Tuple[int, int] = (512, 512)
    halo: Optional[float] = None
    output_levels: Optional[List[int]] = None
    full_output: bool = False


Num = Union[float, Sequence[float]]


@dataclass
class Met:
    ustar: Optional[Num] = None
    mol: Num = 1e9
    wind_speed: Num = 5.0
    wind_dir: Num = 270.0
    z0: Optional[float] = None
    timestamps: Optional[Sequence] = None

    @property
    def n_timesteps(self) -> int:
        for v in (self.ustar, self.wind_speed):
            if isinstance(v, (list, tuple)) or hasattr(v, "__len__"):
                return len(v)
        return 1

    def get_step(self, i: int) -> dict:
        def pick(v):
            if v is None:
                return None
            return v[i] if hasattr(v, "__len__") else v

        step = {k: pick(getattr(self, k)) for k in ("ustar", "mol", "wind_speed", "wind_dir")}
        if self.z0 is not None:
            step["z0"] = self.z0
        step["timestamp"] = self.timestamps[i] if self.timestamps is not None else i
        return step


@dataclass
class SolverOptions:
    closure: str = "MOST"
    precision: str = "single"
    footprint: bool = False
    surface_flux_shape: str = "diamond"
    analytic: bool = False
    src_loc: Optional[Tuple[float, float]] = None


@dataclass
class Parallel:
    num_threads: int = 1
    max_workers: int = 1
    use_cache: bool = False


@dataclass
class Config:
    domain: Domain
    towers: List[Tower]
    met: Met
    solver: SolverOptions = field(default_factory=SolverOptions)
    parallel: Parallel = field(default_factory=Parallel)
