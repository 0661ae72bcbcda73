"""bldfm_b200 -- B200-native (sm_100a) implementation of BLDFM's steady-state spectral solver.

Drop-in for the hot path of SchlutowSM2Group/BLDFM: ``steady_state_transport_solver`` with the
reference's signature (src/bldfm/solver.py:16-30), backed by hand-written CUDA kernels through the
C ABI of include/bldfm_b200.h.  No CPU fallback: without the built library and a CUDA device the
compute calls raise.
"""

from .solver import steady_state_transport_solver, ivp_solver, solve_batched, measure_batched  # noqa: F401
from .utils import compute_wind_fields, ideal_source, point_measurement  # noqa: F401
from .interface import (  # noqa: F401
    run_bldfm_single,
    run_bldfm_timeseries,
    run_bldfm_multitower,
    run_bldfm_parallel,
)
from .cache import GreensFunctionCache  # noqa: F401
from .fft_manager import get_fft_manager, reset_fft_manager  # noqa: F401
from . import config  # noqa: F401

__all__ = [
    "steady_state_transport_solver",
    "ivp_solver",
    "solve_batched",
    "measure_batched",
    "compute_wind_fields",
    "ideal_source",
    "point_measurement",
    "run_bldfm_single",
    "run_bldfm_timeseries",
    "run_bldfm_multitower",
    "run_bldfm_parallel",
    "GreensFunctionCache",
    "get_fft_manager",
    "reset_fft_manager",
    "config",
]
