"""Drivers with the reference's names, arguments and result dictionaries
(src/bldfm/interface.py:31-326), dispatching BATCHES of solves to the GPU.

The reference runs one Python-level solve per (tower, timestep) -- serially (``run_bldfm_timeseries``,
``run_bldfm_multitower``) or fanned out over a process pool (``run_bldfm_parallel``).  Here all
tasks of a call are grouped, the vertical march is computed once per (measurement height, met step)
and shared by the towers, the profiles of all met steps come from ONE vectorised
``vertical_profiles_batch`` call, chunks of problems go to ``bldfm_solve_batched`` in one launch each,
and under ``torchrun`` the march groups are spread over the GPUs of the node with a single final
gather (``distributed.py``).  ``config`` may be the reference's own ``BLDFMConfig`` or
``bldfm_b200.schema.Config``.

Beyond the reference's drivers: ``run_bldfm_measure`` (tower fluxes ``sum(footprint * flux_map)``,
utils.py:80-92) and ``run_bldfm_aggregate`` (time-mean footprint per tower,
examples/timeseries_example.py:46) reduce on the device and move only the reduced result.
"""

from __future__ import annotations

import logging
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib
from . import config as _cfg
from . import distributed as _dist
from .pbl_model import ProfileBatch, compute_wind_fields_batch, vertical_profiles, vertical_profiles_batch
from .solver import (FieldAccumulator, make_grid, measure_batched, solve_batched, steady_state_transport_solver,
                     synchronize, synchronize_previous)
from .utils import compute_wind_fields, ideal_source

logger = logging.getLogger("bldfm.interface")

MAX_CHUNK_BYTES = 128 << 20   # host bytes of (conc, flx) per batched launch
# the batched drivers return thousands of result dicts: their grids are the shared zero-copy read-only views
# (values and shapes as the reference) unless the user asked for materialised grids (GRID_COPY = "1")


def _grid_mode():
    from . import config as _c
    return "1" if _c.GRID_COPY in (True, "1") else "0"


def _make_cache(config):
    """GreensFunctionCache when enabled (interface.py:22-28)."""
    if config.parallel.use_cache and config.solver.footprint:
        from .cache import GreensFunctionCache

        # the batched drivers produce entries far faster than np.savez writes them: write behind
        return GreensFunctionCache(background=True)
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


def profiles_batch(config, z_ms: Sequence[float], met_steps: Sequence[dict]) -> ProfileBatch:
    """Steps 1+2 (interface.py:73-95) for B (measurement height, met step) pairs in one vectorised pass;
    row b is bitwise what ``_profiles_for(config, z_ms[b], met_steps[b])`` returns."""
    ws = np.array([s["wind_speed"] for s in met_steps], dtype=np.float64)
    wd = np.array([s["wind_dir"] for s in met_steps], dtype=np.float64)
    mol = np.array([s["mol"] for s in met_steps], dtype=np.float64)
    um, vm = compute_wind_fields_batch(ws, wd)
    kw = dict(n=config.domain.nz, meas_height=np.asarray(z_ms, dtype=np.float64), wind=(um, vm), mol=mol,
              closure=config.solver.closure)
    z0s = [s.get("z0") for s in met_steps]
    if all(v is not None for v in z0s):         # z0 takes precedence over ustar (interface.py:77-95)
        kw["z0"] = np.array(z0s, dtype=np.float64)
    elif all(v is None for v in z0s):
        kw["ustar"] = np.array([s["ustar"] for s in met_steps], dtype=np.float64)
    else:
        raise ValueError("met steps must either all carry z0 or none")
    return vertical_profiles_batch(**kw)


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

def plan_tasks(config, tasks: Sequence[Tuple[object, int]]):
    """Group tasks (tower | tower index, met index) by march key (z_m, met index).

    Returns (group_keys, task_group) with task_group[t] = index into group_keys.
    """
    keys: Dict[Tuple[float, int], int] = {}
    task_group = []
    towers = config.towers
    for t, mi in tasks:
        if isinstance(t, (int, np.integer)):
            t = towers[t]
        task_group.append(keys.setdefault((float(t.z_m), int(mi)), len(keys)))
    return list(keys.keys()), task_group


class TaskBatch:
    """The pending tasks of one call, prepared for the device without per-task Python work: one
    ``ProfileBatch`` row per march group, one ``bldfm_problem`` per task.

    The profiles are computed lazily in blocks of ``BLOCK`` march groups (one vectorised
    ``vertical_profiles_batch`` call each), so that only the first block's host work is exposed: the
    launches are enqueue-only, and the following blocks are prepared while the device works.
    """

    BLOCK = 128

    def __init__(self, config, tasks):
        self.config = config
        towers = config.towers
        self.tasks = [(towers[t] if isinstance(t, (int, np.integer)) else t, int(mi)) for t, mi in tasks]
        self.keys, self.task_group = plan_tasks(config, self.tasks)
        self.task_group = np.asarray(self.task_group, dtype=np.int64)
        self.xm = np.array([t.x for t, _ in self.tasks], dtype=np.float64)
        self.ym = np.array([t.y for t, _ in self.tasks], dtype=np.float64)
        self.steps = {}
        self._blocks: Dict[int, ProfileBatch] = {}
        self.prep_seconds = 0.0
        dom, sol = config.domain, config.solver
        self.domain = (dom.xmax, dom.ymax)
        self.levels = _levels(config)
        self.lv = np.array([self.levels]) if np.ndim(self.levels) == 0 else np.asarray(self.levels)
        self.nlv = len(self.lv)
        self.solver_kw = dict(domain=self.domain, levels=self.levels, modes=dom.modes, footprint=sol.footprint,
                              analytic=sol.analytic, halo=dom.halo, precision=sol.precision)

    def step(self, mi):
        s = self.steps.get(mi)
        if s is None:
            s = self.steps[mi] = self.config.met.get_step(mi)
        return s

    def _block(self, b) -> ProfileBatch:
        pb = self._blocks.get(b)
        if pb is None:
            import time
            t0 = time.perf_counter()
            keys = self.keys[b * self.BLOCK:(b + 1) * self.BLOCK]
            pb = profiles_batch(self.config, [k[0] for k in keys], [self.step(k[1]) for k in keys])
            self._blocks[b] = pb
            self.prep_seconds += time.perf_counter() - t0
        return pb

    def problems(self, idx):
        """bldfm_problem array (+ keepalive) for the tasks ``idx`` (indices into self.tasks)."""
        idx = np.asarray(idx, dtype=np.int64)
        groups = self.task_group[idx]
        blocks = groups // self.BLOCK
        b0 = int(blocks[0]) if len(blocks) else 0
        if len(blocks) == 0 or bool((blocks == b0).all()):
            return _lib.problems_from_batch(self._block(b0), groups - b0 * self.BLOCK, self.xm[idx], self.ym[idx], 0.0)
        # the chunk spans blocks: the structs hold absolute addresses, so the pieces simply concatenate
        arr = np.zeros(len(idx), dtype=_lib.PROBLEM_DTYPE)
        keep = []
        for b in np.unique(blocks):
            sel = np.nonzero(blocks == b)[0]
            part, k = _lib.problems_from_batch(self._block(int(b)), groups[sel] - int(b) * self.BLOCK,
                                               self.xm[idx[sel]], self.ym[idx[sel]], 0.0)
            arr[sel] = part
            keep.append(k)
        return arr, (arr, keep)

    def is_f32(self):
        """Per task: float32 fields in the reference (precision="single", no phase shift)?"""
        sol = self.config.solver
        if sol.precision == "double" or sol.footprint:
            return np.zeros(len(self.tasks), dtype=bool)
        return ~(self.xm * self.xm + self.ym * self.ym > 0.0)

    def chunks(self, idx, per_task_bytes):
        """Split the task indices ``idx`` into launches of at most MAX_CHUNK_BYTES / 256 problems."""
        chunk = max(1, min(256, MAX_CHUNK_BYTES // max(per_task_bytes, 1)))
        return [idx[c0:c0 + chunk] for c0 in range(0, len(idx), chunk)]

    def row(self, t):
        g = int(self.task_group[t])
        return self._block(g // self.BLOCK).row(g % self.BLOCK)


def solve_tasks(config, tasks: Sequence[Tuple[object, int]], surface_flux=None, cache=None, out=None,
                out_pinned=False, build_results=True) -> List[dict]:
    """Solve the given (tower | tower index, met index) tasks on this process's GPU, batched.

    ``out=(conc, flx)``: float64 destination arrays ``[len(tasks), nlv, ny, nx]`` (e.g. this rank's block of
    a shared-memory segment) that receive the fields in task order; the result dicts then hold views.
    ``build_results=False`` (needs ``out``) only delivers the fields and returns an empty list.
    """
    dom, sol = config.domain, config.solver
    tb = TaskBatch(config, tasks)
    srf = _surface_flux(config, surface_flux)
    results: List[dict] = [None] * len(tb.tasks)
    domain = tb.domain

    pending = []
    for t, (tower, mi) in enumerate(tb.tasks):
        if cache is not None and sol.footprint:                       # solver.py:77-80
            z, profiles = tb.row(t)
            hit = cache.get(z, profiles, domain, dom.modes, (tower.x, tower.y), dom.halo, sol.precision)
            if hit is not None:
                results[t] = _result(tower, tb.step(mi), *hit)
                continue
        pending.append(t)

    nlv = tb.nlv
    per_task = 2 * nlv * dom.ny * dom.nx * 8
    is32 = tb.is_f32()
    # enqueue every chunk without waiting: the device->host copy of chunk k overlaps the kernels of
    # chunk k+1 and the host-side work of the chunks after it.  Tasks the reference answers in float32
    # (precision="single", tower at exactly (0,0)) and float64 ones go out as separate launches.
    def finish(part, conc, flx, direct):
        if not build_results and direct:
            return
        for b, t in enumerate(part):
            tower, mi = tb.tasks[t]
            z, profiles = tb.row(t)
            if out is not None and not direct:
                out[0][t] = conc[b]
                out[1][t] = flx[b]
            if not build_results:
                continue
            grid = make_grid(z, tb.lv, domain, dom.nx, dom.ny, mode=_grid_mode())
            res = (grid, np.squeeze(conc[b]), np.squeeze(flx[b]))
            if cache is not None and sol.footprint:                   # solver.py:301-302
                cache.put(z, profiles, domain, dom.modes, (tower.x, tower.y), dom.halo, sol.precision, *res)
            results[t] = _result(tower, tb.step(mi), *res)

    # the post-processing of chunk k (result dictionaries, cache entries handed to the writer thread) runs
    # while chunk k+1 computes; only pinned destinations are enqueue-only, so this needs no extra care for
    # pageable ones (their launch returns with the results in place)
    prev = None
    for want32 in (False, True):
        idx = [t for t in pending if bool(is32[t]) == want32]
        for part in tb.chunks(idx, per_task):
            dest = None
            if out is not None and not want32 and part == list(range(part[0], part[0] + len(part))):
                dest = (out[0][part[0]:part[0] + len(part)], out[1][part[0]:part[0] + len(part)])
            conc, flx = solve_batched(srf, problems=tb.problems(part), wait=False, out=dest,
                                      out_pinned=out_pinned, **tb.solver_kw)
            if prev is not None:
                synchronize_previous()
                finish(*prev)
            prev = (part, conc, flx, dest is not None)
    synchronize()
    if prev is not None:
        finish(*prev)
    if cache is not None and hasattr(cache, "flush"):
        cache.flush()        # like the reference, every entry is on disk when the driver returns
    return results if build_results else []


def run_bldfm_timeseries(config, tower, surface_flux=None) -> list:
    """All timesteps for one tower (interface.py:141-174), one batched launch per chunk.  ``tower`` may be any
    tower object (it need not be an entry of ``config.towers``, like in the reference)."""
    n = config.met.n_timesteps
    logger.info("Running timeseries for tower '%s': %d timesteps", tower.name, n)
    return solve_tasks(config, [(tower, i) for i in range(n)], surface_flux, _make_cache(config))


def _multitower_tasks(config):
    # timestep-major task order keeps the towers of one met step in the same batch
    n = config.met.n_timesteps
    return [(ti, mi) for mi in range(n) for ti in range(len(config.towers))]


def run_bldfm_multitower(config, surface_flux=None) -> dict:
    """All towers x all timesteps (interface.py:177-207); marches shared between towers."""
    n = config.met.n_timesteps
    logger.info("Running multitower: %d towers x %d timesteps", len(config.towers), n)
    tasks = _multitower_tasks(config)
    flat = solve_tasks(config, tasks, surface_flux, _make_cache(config))
    out = {t.name: [None] * n for t in config.towers}
    for (ti, mi), res in zip(tasks, flat):
        out[config.towers[ti].name][mi] = res
    return out


def _shard(config, tasks, delivered=False):
    """(rank, world, owner[t], my task indices): march groups spread over the ranks (distributed.shard_groups).

    ``delivered``: every field of the job goes to host memory, so with more than one rank the job is bound by
    the host links -- and those need not be equal for all GPUs of a box.  With ``config.LINK_AWARE_SHARDING``
    (``BLDFM_B200_LINK_AWARE=1``) the shares are then sized by the measured per-rank link rates
    (``distributed.link_rates``) and by the bytes a group delivers instead of by its compute cost."""
    rank, ws = _dist.world()
    keys, task_group = plan_tasks(config, tasks)
    ntow = np.bincount(task_group, minlength=len(keys))
    if delivered and ws > 1 and _cfg.LINK_AWARE_SHARDING:
        assign = _dist.shard_groups(keys, ntow.astype(float), ws, speeds=_dist.link_rates())
    else:
        assign = _dist.shard_groups(keys, 1.0 + 0.15 * ntow, ws)     # march dominates, towers add a little
    owner = _dist.owner_of_tasks(task_group, assign)
    mine = [t for t in range(len(tasks)) if owner[t] == rank]
    return rank, ws, owner, mine


# wall-clock phases of this rank's most recent gathering run_bldfm_parallel call (diagnostics; bench.py reports them)
LAST_PARALLEL_PHASES = {}


def run_bldfm_parallel(config, max_workers=None, parallel_over: str = "towers", surface_flux=None,
                       gather: bool = True) -> dict:
    """Multi-GPU counterpart of the reference's process pool (interface.py:241-326).

    Under ``torchrun`` (``torch.distributed`` initialised) the march groups are distributed over the
    ranks (``distributed.shard_groups``), every rank solves its share on its own GPU and copies its fields
    device->host over ITS OWN PCIe link straight into a page-locked shared-memory segment that rank 0 maps
    (``distributed.SharedResults``): the final gather is that segment plus one barrier -- no field crosses a
    PCIe link twice and none goes through another GPU.  Rank 0 returns the full result dict (arrays are views
    into the segment), the other ranks an empty dict; ``gather=False`` returns each rank's own share.
    Without a process group the same path runs with one rank (a page-locked segment of this process).
    ``max_workers`` and ``parallel_over`` are accepted for compatibility; the strategies differ only
    in how the reference slices its task list, the results are identical (tests/test_parallel.py:78-90).
    """
    if parallel_over not in ("towers", "time", "both"):
        raise ValueError(f"Unknown parallel_over={parallel_over!r}. Choose 'towers', 'time', or 'both'.")
    if surface_flux is not None:
        logger.warning("surface_flux is ignored in parallel mode; an ideal source is generated "
                       "(interface.py:270-275).")
        surface_flux = None
    n = config.met.n_timesteps
    towers = config.towers
    tasks = _multitower_tasks(config)
    rank, ws, owner, mine = _shard(config, tasks, delivered=gather)
    out = {t.name: [None] * n for t in towers}
    if not gather:
        local = solve_tasks(config, [tasks[t] for t in mine], None, None)
        for t, res in zip(mine, local):
            ti, mi = tasks[t]
            out[towers[ti].name][mi] = res
        return out

    shape, _ = solve_shape(config)
    lv = _levels(config)
    nlv = 1 if np.ndim(lv) == 0 else len(lv)
    dom = config.domain
    counts = np.bincount(owner, minlength=ws)
    import time as _time
    t_start = _time.perf_counter()
    seg = _dist.SharedResults.acquire((nlv, dom.ny, dom.nx), counts)
    conc_l, flx_l = seg.local_block()
    phases = LAST_PARALLEL_PHASES
    phases.clear()
    phases["acquire_segment_s"] = _time.perf_counter() - t_start

    # Rank 0 builds the result dictionaries (views into the segment, valid objects whatever the bytes are yet)
    # on a helper thread WHILE the GPUs solve and copy: its main thread spends that phase blocked in CUDA waits
    # with the GIL released, so the serial tail of the gather disappears behind the copies.
    deferred_cast = []
    failure = []

    def build():
        try:
            conc_all, flx_all = seg.all_blocks()           # rank-major: rank r's tasks in its own task order
            start = np.concatenate([[0], np.cumsum(counts)[:-1]])
            pos = np.empty(len(tasks), dtype=np.int64)     # task -> row of the rank-major segment
            fill = start.copy()
            for t in range(len(tasks)):
                pos[t] = fill[owner[t]]
                fill[owner[t]] += 1
            tb = TaskBatch(config, tasks)
            is32 = tb.is_f32()
            lvarr = tb.lv
            gmode = _grid_mode()
            domain = (dom.xmax, dom.ymax)
            squeeze = nlv == 1
            grids = {}                                     # march group -> grid tuple (towers of a group share z)
            for t, (tower, mi) in enumerate(tb.tasks):
                g = int(tb.task_group[t])
                grid = grids.get(g)
                if grid is None:
                    z, _ = tb.row(t)
                    grid = grids[g] = make_grid(z, lvarr, domain, dom.nx, dom.ny, mode=gmode)
                c, f = conc_all[pos[t]], flx_all[pos[t]]
                if squeeze:
                    c, f = c[0], f[0]
                if dom.ny == 1 or dom.nx == 1:
                    c, f = np.squeeze(c), np.squeeze(f)
                res = _result(tower, tb.step(mi), grid, c, f)
                if is32[t]:
                    deferred_cast.append(res)              # float32 tasks are cast once their bytes have landed
                out[tower.name][mi] = res
        except BaseException as e:                         # surfaced on the main thread
            failure.append(e)

    builder = None
    if rank == 0:
        import threading
        builder = threading.Thread(target=build, name="bldfm-result-builder")
        builder.start()
    try:
        t0 = _time.perf_counter()
        solve_tasks(config, [tasks[t] for t in mine], None, None, out=(conc_l, flx_l), out_pinned=seg.pinned,
                    build_results=False)
        phases["solve_and_copy_s"] = _time.perf_counter() - t0
        t0 = _time.perf_counter()
        seg.barrier()                                  # every rank's fields have landed in the segment
        phases["wait_for_other_ranks_s"] = _time.perf_counter() - t0
    finally:
        t0 = _time.perf_counter()
        if builder is not None:
            builder.join()
        phases["wait_for_result_builder_s"] = _time.perf_counter() - t0
        phases["total_s"] = _time.perf_counter() - t_start
    if rank != 0:
        return {}
    if failure:
        raise failure[0]
    for res in deferred_cast:
        res["conc"], res["flx"] = res["conc"].astype(np.float32), res["flx"].astype(np.float32)
    return out


def solve_shape(config):
    """(shape, dtype) of one task's squeezed conc/flx field (dtype of a shifted tower)."""
    dom, sol = config.domain, config.solver
    lv = _levels(config)
    nlv = 1 if np.ndim(lv) == 0 else len(lv)
    shape = tuple(s for s in (nlv, dom.ny, dom.nx) if s != 1)
    return shape, np.float64


# ------------------------------------------------------------------------------------------------
# reductions on the device
# ------------------------------------------------------------------------------------------------

def run_bldfm_measure(config, flux_map, chunk_groups: int = 8) -> dict:
    """Tower measurements without moving footprints: for every tower and timestep
    ``sum(conc * flux_map)`` and ``sum(flx * flux_map)`` -- ``point_measurement`` (utils.py:80-92) fused on
    the device.  Returns ``{tower_name: {"conc": [n_time(, nlv)], "flx": [...]}}`` on rank 0 (every rank
    when not distributed); under torchrun the march groups are sharded and only these scalars are gathered.
    """
    towers = config.towers
    n = config.met.n_timesteps
    tasks = _multitower_tasks(config)
    rank, ws, owner, mine = _shard(config, tasks)
    tb = TaskBatch(config, [tasks[t] for t in mine])
    srf = _surface_flux(config, None)
    is32 = tb.is_f32()
    ntow = max(1, len(towers))
    parts = []
    for want32 in (False, True):
        idx = [t for t in range(len(tb.tasks)) if bool(is32[t]) == want32]
        step = max(1, chunk_groups * ntow)
        for c0 in range(0, len(idx), step):
            part = idx[c0:c0 + step]
            cw, fw = measure_batched(flux_map, srf, problems=tb.problems(part), wait=False, **tb.solver_kw)
            parts.append((part, cw, fw))
    synchronize()
    nlv = tb.nlv
    loc = np.zeros((len(mine), 2, nlv))
    for part, cw, fw in parts:
        loc[part, 0] = cw
        loc[part, 1] = fw
    full = _dist.gather_small(loc, owner)
    if full is None:
        return {}
    out = {t.name: {"conc": np.empty((n, nlv)), "flx": np.empty((n, nlv))} for t in towers}
    for t, (ti, mi) in enumerate(tasks):
        out[towers[ti].name]["conc"][mi] = full[t, 0]
        out[towers[ti].name]["flx"][mi] = full[t, 1]
    if nlv == 1:
        for v in out.values():
            v["conc"], v["flx"] = v["conc"][:, 0], v["flx"][:, 0]
    return out


def run_bldfm_aggregate(config, chunk_groups: int = 8) -> dict:
    """Footprint climatology: the time-mean ``conc`` / ``flx`` field of every tower
    (examples/timeseries_example.py:46: ``np.mean([r["flx"] for r in results], axis=0)``), summed on the device
    -- one field per tower crosses PCIe instead of one per timestep.  Under torchrun every rank sums its own
    timesteps and the per-tower sums are reduced onto rank 0 over NCCL (the one collective of this driver).
    Returns ``{tower_name: {"grid": (X, Y), "conc": mean, "flx": mean, "n": n_time}}`` on rank 0.
    """
    towers = config.towers
    dom, sol = config.domain, config.solver
    n = config.met.n_timesteps
    tasks = _multitower_tasks(config)
    rank, ws, owner, mine = _shard(config, tasks)
    tb = TaskBatch(config, [tasks[t] for t in mine])
    srf = _surface_flux(config, None)
    acc = FieldAccumulator((dom.ny, dom.nx), tb.domain, tb.levels, len(towers), modes=dom.modes,
                           footprint=sol.footprint, analytic=sol.analytic, halo=dom.halo, precision=sol.precision)
    try:
        is32 = tb.is_f32()
        slot = np.array([tasks[t][0] for t in mine], dtype=np.int32)
        step = max(1, chunk_groups * max(1, len(towers)))
        for want32 in (False, True):
            idx = [t for t in range(len(tb.tasks)) if bool(is32[t]) == want32]
            for c0 in range(0, len(idx), step):
                part = idx[c0:c0 + step]
                acc.add(srf, slot[part], problems=tb.problems(part))
        if ws > 1:
            sums = _dist.reduce_device_sums(acc)
        else:
            sums = acc.fetch()
    finally:
        acc.close()
    if sums is None:
        return {}
    x = np.linspace(0, dom.xmax, dom.nx, endpoint=False)
    y = np.linspace(0, dom.ymax, dom.ny, endpoint=False)
    Y, X = np.meshgrid(y, x, indexing="ij")
    out = {}
    for ti, tower in enumerate(towers):
        out[tower.name] = {"grid": (X, Y), "conc": np.squeeze(sums[0][ti] / n), "flx": np.squeeze(sums[1][ti] / n),
                           "n": n}
    return out
