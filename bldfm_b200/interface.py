"""Drivers with the reference's names, arguments and result dictionaries
(src/bldfm/interface.py:31-326), dispatching BATCHES of solves to the GPU.

The reference runs one Python-level solve per (tower, timestep) -- serially (``run_bldfm_timeseries``,
``run_bldfm_multitower``) or fanned out over a process pool (``run_bldfm_parallel``).  Here all
tasks of a call are grouped, the vertical march is computed once per (measurement height, met step)
and shared by the towers, chunks of problems go to ``bldfm_solve_batched`` in one
launch each, and under ``torchrun`` the march groups are spread over the GPUs of the node with a
single final gather (``distributed.py``).  ``config`` may be the reference's own ``BLDFMConfig`` or
``bldfm_b200.schema.Config``.
"""

from __future__ import annotations

import logging
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import distributed as _dist
from .pbl_model import vertical_profiles
from .solver import make_grid, solve_batched, steady_state_transport_solver, synchronize
from .utils import compute_wind_fields, ideal_source

logger = logging.getLogger("bldfm.interface")

MAX_CHUNK_BYTES = 128 << 20   # host bytes of (conc, flx) per batched launch


def _make_cache(config):
    """GreensFunctionCache when enabled (interface.py:22-28)."""
    if config.parallel.use_cache and config.solver.footprint:
        from .cache import GreensFunctionCache

        return GreensFunctionCache()
    return None


def _profiles_for(config, z_m, met_step):
    """Steps 1+2 of the workflow (interface.py:73-95): wind components, vertical profiles."""
    u_wind, v_wind = compute_wind_fields(met_step["wind_speed"], met_step["wind_dir"])
    z0_val = met_step.get("z0")
    kw = dict(n=config.domain.nz, meas_height=z_m, wind=(u_wind, v_wind), mol=met_step["mol"],
              closure=config.solver.closure)
    if z0_val is not None:                      # z0 takes precedence over ustar
        kw["z0"] = z0_val
    else:
        kw["ustar"] = met_step["ustar"]
    return vertical_profiles(**kw)


def _levels(config):
    dom = config.domain                          # interface.py:109-114
    if dom.output_levels:
        return dom.output_levels
    if dom.full_output:
        return list(range(dom.nz + 1))
    return dom.nz


def _surface_flux(config, surface_flux):
    if surface_flux is not None:
        return surface_flux
    dom, sol = config.domain, config.solver      # interface.py:98-106
    return ideal_source((dom.nx, dom.ny), (dom.xmax, dom.ymax), src_loc=sol.src_loc,
                        shape=sol.surface_flux_shape)


def _result(tower, met_step, grid, conc, flx):
    return {                                     # interface.py:130-138
        "grid": grid,
        "conc": conc,
        "flx": flx,
        "tower_name": tower.name,
        "tower_xy": (tower.x, tower.y),
        "timestamp": met_step["timestamp"],
        "params": met_step,
    }


def run_bldfm_single(config, tower, met_index: int = 0, surface_flux=None, cache=None) -> dict:
    """One solve for one tower at one timestep (interface.py:31-138)."""
    dom, sol = config.domain, config.solver
    met_step = config.met.get_step(met_index)
    z, profiles = _profiles_for(config, tower.z_m, met_step)
    grid, conc, flx = steady_state_transport_solver(
        srf_flx=_surface_flux(config, surface_flux), z=z, profiles=profiles,
        domain=(dom.xmax, dom.ymax), levels=_levels(config), modes=dom.modes,
        meas_pt=(tower.x, tower.y), footprint=sol.footprint, analytic=sol.analytic, halo=dom.halo,
        precision=sol.precision, cache=cache)
    return _result(tower, met_step, grid, conc, flx)


# ------------------------------------------------------------------------------------------------
# batched core
# ------------------------------------------------------------------------------------------------

def plan_tasks(config, tasks: Sequence[Tuple[int, int]]):
    """Group tasks (tower index, met index) by march key (z_m, met index).

    Returns (group_keys, task_group) with task_group[t] = index into group_keys.
    """
    keys: Dict[Tuple[float, int], int] = {}
    task_group = []
    for ti, mi in tasks:
        k = (float(config.towers[ti].z_m), int(mi))
        task_group.append(keys.setdefault(k, len(keys)))
    return list(keys.keys()), task_group


def solve_tasks(config, tasks: Sequence[Tuple[int, int]], surface_flux=None, cache=None) -> List[dict]:
    """Solve the given (tower index, met index) tasks on this process's GPU, batched."""
    dom, sol = config.domain, config.solver
    domain = (dom.xmax, dom.ymax)
    levels = _levels(config)
    lv = np.array([levels]) if np.ndim(levels) == 0 else np.asarray(levels)
    srf = _surface_flux(config, surface_flux)
    results: List[dict] = [None] * len(tasks)

    prof_cache: Dict[Tuple[float, int], tuple] = {}
    pending = []
    for t, (ti, mi) in enumerate(tasks):
        tower = config.towers[ti]
        met_step = config.met.get_step(mi)
        key = (float(tower.z_m), int(mi))
        if key not in prof_cache:
            prof_cache[key] = _profiles_for(config, tower.z_m, met_step)
        z, profiles = prof_cache[key]
        if cache is not None and sol.footprint:                       # solver.py:77-80
            hit = cache.get(z, profiles, domain, dom.modes, (tower.x, tower.y), dom.halo, sol.precision)
            if hit is not None:
                results[t] = _result(tower, met_step, *hit)
                continue
        pending.append((t, tower, met_step, z, profiles))

    nlv = len(lv)
    per_problem = 2 * nlv * dom.ny * dom.nx * 8
    chunk = max(1, min(256, MAX_CHUNK_BYTES // max(per_problem, 1)))
    # enqueue every chunk without waiting: the device->host copy of chunk k overlaps the kernels of
    # chunk k+1 and the host-side profile work of the chunks after it
    inflight = []
    for c0 in range(0, len(pending), chunk):
        part = pending[c0:c0 + chunk]
        conc, flx = solve_batched(
            srf, [p[3] for p in part], [p[4] for p in part], domain, levels, modes=dom.modes,
            meas_pts=[(p[1].x, p[1].y) for p in part], footprint=sol.footprint, analytic=sol.analytic,
            halo=dom.halo, precision=sol.precision, wait=False)
        inflight.append((part, conc, flx))
    synchronize()
    for part, conc, flx in inflight:
        for b, (t, tower, met_step, z, profiles) in enumerate(part):
            grid = make_grid(z, lv, domain, dom.nx, dom.ny)
            res = (grid, np.squeeze(conc[b]), np.squeeze(flx[b]))
            if cache is not None and sol.footprint:                   # solver.py:301-302
                cache.put(z, profiles, domain, dom.modes, (tower.x, tower.y), dom.halo, sol.precision, *res)
            results[t] = _result(tower, met_step, *res)
    return results


def run_bldfm_timeseries(config, tower, surface_flux=None) -> list:
    """All timesteps for one tower (interface.py:141-174), one batched launch per chunk."""
    n = config.met.n_timesteps
    logger.info("Running timeseries for tower '%s': %d timesteps", tower.name, n)
    ti = next(i for i, t in enumerate(config.towers) if t is tower or t == tower)
    return solve_tasks(config, [(ti, i) for i in range(n)], surface_flux, _make_cache(config))


def run_bldfm_multitower(config, surface_flux=None) -> dict:
    """All towers x all timesteps (interface.py:177-207); marches shared between towers."""
    n = config.met.n_timesteps
    logger.info("Running multitower: %d towers x %d timesteps", len(config.towers), n)
    # timestep-major task order keeps the towers of one met step in the same batch
    tasks = [(ti, mi) for mi in range(n) for ti in range(len(config.towers))]
    flat = solve_tasks(config, tasks, surface_flux, _make_cache(config))
    out = {t.name: [None] * n for t in config.towers}
    for (ti, mi), res in zip(tasks, flat):
        out[config.towers[ti].name][mi] = res
    return out


def run_bldfm_parallel(config, max_workers=None, parallel_over: str = "towers", surface_flux=None,
                       gather: bool = True) -> dict:
    """Multi-GPU counterpart of the reference's process pool (interface.py:241-326).

    Under ``torchrun`` (``torch.distributed`` initialised) the march groups are distributed over the
    ranks (``distributed.shard_groups``), every rank solves its share on its own GPU, and the cropped
    fields are gathered on rank 0 (other ranks return their local share only when ``gather=False``,
    else an empty dict).  Without a process group this is ``run_bldfm_multitower`` on one GPU.
    ``max_workers`` and ``parallel_over`` are accepted for compatibility; the strategies differ only
    in how the reference slices its task list, the results are identical (tests/test_parallel.py:78-90).
    """
    if parallel_over not in ("towers", "time", "both"):
        raise ValueError(f"Unknown parallel_over={parallel_over!r}. Choose 'towers', 'time', or 'both'.")
    if surface_flux is not None:
        logger.warning("surface_flux is ignored in parallel mode; an ideal source is generated "
                       "(interface.py:270-275).")
        surface_flux = None
    rank, ws = _dist.world()
    if ws == 1:
        return run_bldfm_multitower(config)

    n = config.met.n_timesteps
    towers = config.towers
    tasks = [(ti, mi) for mi in range(n) for ti in range(len(towers))]
    keys, task_group = plan_tasks(config, tasks)
    ntow = np.bincount(task_group, minlength=len(keys))
    assign = _dist.shard_groups(keys, 1.0 + 0.15 * ntow, ws)     # march dominates, towers add a little
    owner = _dist.owner_of_tasks(task_group, assign)
    mine = [t for t in range(len(tasks)) if owner[t] == rank]
    local = solve_tasks(config, [tasks[t] for t in mine], None, None)

    out = {t.name: [None] * n for t in towers}
    if not gather:
        for t, res in zip(mine, local):
            ti, mi = tasks[t]
            out[towers[ti].name][mi] = res
        return out

    def stack(key):
        if not local:
            return None
        return np.stack([np.asarray(r[key]) for r in local])

    probe = solve_shape(config)
    conc_l = stack("conc") if local else np.empty((0,) + probe[0], probe[1])
    flx_l = stack("flx") if local else np.empty((0,) + probe[0], probe[1])
    conc_all = _dist.gather_fields(conc_l, owner)
    flx_all = _dist.gather_fields(flx_l, owner)
    if rank != 0:
        return {}
    lv = _levels(config)
    lv = np.array([lv]) if np.ndim(lv) == 0 else np.asarray(lv)
    dom = config.domain
    for t, (ti, mi) in enumerate(tasks):
        tower = towers[ti]
        met_step = config.met.get_step(mi)
        z, _ = _profiles_for(config, tower.z_m, met_step)
        grid = make_grid(z, lv, (dom.xmax, dom.ymax), dom.nx, dom.ny)
        out[tower.name][mi] = _result(tower, met_step, grid, conc_all[t], flx_all[t])
    return out


def solve_shape(config):
    """(shape, dtype) of one task's squeezed conc/flx field."""
    from . import _lib

    dom, sol = config.domain, config.solver
    lv = _levels(config)
    nlv = 1 if np.ndim(lv) == 0 else len(lv)
    shape = tuple(s for s in (nlv, dom.ny, dom.nx) if s != 1)
    flags = (_lib.FOOTPRINT if sol.footprint else 0) | (_lib.DOUBLE if sol.precision == "double" else 0)
    t0 = config.towers[0]
    f32 = bool(_lib.lib().bldfm_output_is_f32(flags, float(t0.x), float(t0.y)))
    return shape, (np.float32 if f32 else np.float64)
