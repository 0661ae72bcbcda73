"""Memory behaviour of the drop-in (restatement of the reference's tests/test_memory.py:64-154 for the CUDA
path): host RSS and DEVICE memory must not grow over repeated solves, and the three caches this build adds --
the pinned result pool (_pinned.py), the per-geometry plan cache with its LRU budget (fft_manager.py) and the
grow-only device workspaces of a plan -- must stay bounded under geometry churn.  Metrics are printed with the
reference's greppable ``MEMORY`` prefix."""
import ctypes as C
import gc

import numpy as np
import psutil
import pytest

pytestmark = pytest.mark.gpu

SINGLE_SOLVE_THRESHOLD_MB = 500      # tests/test_memory.py:26


def _rss_mb():
    return psutil.Process().memory_info().rss / 1024**2


def _dev_used_mb():
    import torch
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 1024**2


def _run_solve(B, footprint=True, nxy=(128, 64)):
    """One solver call at the reference's conftest scale (128x64, modes 128x64, nz=16; test_memory.py:34-60)."""
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    domain = (500.0, 250.0)
    srf = ideal_source(nxy, domain, (250.0, 125.0), shape="point")
    z, profs = vertical_profiles(16, 10.0, (5.0, 0.0), 0.4, closure="MOST")
    return B.steady_state_transport_solver(srf, z, profs, domain, 15, modes=nxy, footprint=footprint)


@pytest.fixture(scope="module")
def B(gpu_lib):
    import bldfm_b200
    return bldfm_b200


def test_single_solve_memory(B):
    _run_solve(B)                       # CUDA context + plan creation are one-off costs, like the numba JIT
    gc.collect()
    before = _rss_mb()
    _run_solve(B)
    delta = _rss_mb() - before
    print(f"\nMEMORY rss_before={before:.1f}MB rss_delta={delta:.1f}MB")
    assert delta < SINGLE_SOLVE_THRESHOLD_MB


def test_sequential_solves_no_leak(B):
    """10 sequential solves: host RSS growth < 20 % (test_memory.py:93-119), device memory and the plan's
    workspace constant, pinned buffers handed back to the pool."""
    from bldfm_b200 import _lib
    from bldfm_b200._pinned import pool
    for fp in (True, False):             # both kinds once: the source buffers of the non-footprint path too
        res = _run_solve(B, footprint=fp)
        del res
    gc.collect()
    geom = _lib.geometry((64, 128), (500.0, 250.0), (128, 64), None)
    plan = B.get_fft_manager().plan(geom)
    L = _lib.lib()
    ws0, dev0 = int(L.bldfm_plan_workspace_bytes(plan)), _dev_used_mb()
    rss = []
    for i in range(10):
        res = _run_solve(B, footprint=bool(i % 2))
        del res
        gc.collect()
        rss.append(_rss_mb())
    growth = rss[-1] / rss[0]
    print(f"\nMEMORY sequential_rss=[{', '.join(f'{v:.1f}' for v in rss)}]MB growth_ratio={growth:.3f} "
          f"device_used={_dev_used_mb():.0f}MB pinned_outstanding={pool.outstanding}")
    assert growth < 1.20
    assert int(L.bldfm_plan_workspace_bytes(plan)) == ws0
    assert abs(_dev_used_mb() - dev0) < 64.0
    assert pool.outstanding == 0          # every result buffer went back to the free list


def test_geometry_churn_is_bounded_by_the_plan_budget(B):
    """Changing geometries create plans; the LRU budget (config.MAX_WORKSPACE_BYTES) bounds what stays cached."""
    import torch
    mgr = B.get_fft_manager()
    old_budget = B.config.MAX_WORKSPACE_BYTES
    B.reset_fft_manager()
    mgr = B.get_fft_manager()
    torch.cuda.empty_cache()
    gc.collect()
    dev0, rss0 = _dev_used_mb(), _rss_mb()
    budget = 6 << 20
    B.config.MAX_WORKSPACE_BYTES = budget
    try:
        shapes = [(96 + 16 * k, 64 + 8 * k) for k in range(10)]
        peak_ws = 0
        for rep in range(3):
            for nxy in shapes:
                res = _run_solve(B, nxy=nxy)
                del res
                peak_ws = max(peak_ws, mgr.workspace_bytes())
        gc.collect()
        nplans = len(mgr._plans)
        print(f"\nMEMORY churn plans_cached={nplans} workspace={mgr.workspace_bytes() / 2**20:.1f}MB "
              f"peak_workspace={peak_ws / 2**20:.1f}MB device_delta={_dev_used_mb() - dev0:.0f}MB "
              f"rss_delta={_rss_mb() - rss0:.1f}MB")
        # the budget is checked before a plan is created: cached workspaces stay within budget + one plan
        # (the largest of these geometries holds ~10 MB: padded 720 x 680, both result sets, source buffers)
        assert peak_ws <= budget + (16 << 20)
        assert nplans < len(shapes)
        assert _dev_used_mb() - dev0 < 256.0
        assert _rss_mb() - rss0 < SINGLE_SOLVE_THRESHOLD_MB
    finally:
        B.config.MAX_WORKSPACE_BYTES = old_budget
        B.reset_fft_manager()


def test_pinned_pool_cap_falls_back_to_pageable_results(B, monkeypatch):
    """Live results beyond the pinned cap come back as ordinary arrays (same values), not as an error."""
    from bldfm_b200 import _pinned
    first = _run_solve(B)
    monkeypatch.setattr(_pinned, "MAX_OUTSTANDING", 1 << 16)
    held = [_run_solve(B) for _ in range(4)]
    for r in held:
        assert np.array_equal(r[1], first[1]) and np.array_equal(r[2], first[2])
    del held, first
    gc.collect()
    assert _pinned.pool.outstanding == 0
