"""The non-headline BASELINE configs of bench.py (configs 3, 4, 5), each with parity flags.

Called by bench.py on every rank of a torchrun job (or the single process at N = 1); rank 0 gets the
dictionaries that go under "configs" in the JSON line.  Nothing here imports the oracle: checks against it
are passed in by bench.py as callables.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np


def _max_over_ranks(torch, dist, world, vals, device):
    t = torch.tensor(list(vals), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def _barrier(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def hbm_peak():
    import json
    from pathlib import Path
    try:
        return float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), \
            "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# config 3: 1024^2 x 129 levels, all levels written out (single GPU)
# ------------------------------------------------------------------------------------------------
def leg_config3(torch, local, reps=5):
    import bldfm_b200
    from bldfm_b200 import _lib
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source

    L = _lib.lib()
    n, nz = 1024, 128
    dom = (8000.0, 8000.0)
    z, prof = vertical_profiles(nz, 10.0, (6.0, 0.0), ustar=0.4)
    src = ideal_source((n, n), dom, src_loc=(2000.0, 4000.0), shape="point")
    levels = np.arange(0, nz + 1)
    geom = _lib.geometry((n, n), dom, (n, n), None)
    plan = bldfm_b200.get_fft_manager().plan(geom, local)
    prob, keep = _lib.make_problem(z, prof, (0.0, 0.0), 0.0)
    lv = np.ascontiguousarray(levels, dtype=np.int64)
    lvp = lv.ctypes.data_as(C.POINTER(C.c_int64))
    dev = f"cuda:{local}"
    nlv = len(lv)
    out_c = torch.empty((nlv, n, n), dtype=torch.float64, device=dev)
    out_f = torch.empty_like(out_c)
    d_src = torch.from_numpy(src).to(dev)
    flags = _lib.DOUBLE | _lib.OUT_ON_DEVICE | _lib.ASYNC | _lib.SRC_ON_DEVICE
    flags |= {"exact": 0, "fma": _lib.MARCH_FMA, "sweep": _lib.MARCH_SWEEP,
              "auto": _lib.MARCH_AUTO}[bldfm_b200.config.MARCH_MODE]

    def solve():
        _lib.check(L.bldfm_solve(plan, C.byref(prob), lvp, nlv, d_src.data_ptr(), flags, out_c.data_ptr(),
                                 out_f.data_ptr()))

    L.bldfm_plan_set_profiling(plan, 1)
    tm = _lib.Timings()
    rows = []
    for i in range(reps + 2):
        solve()
        _lib.check(L.bldfm_plan_last_timings(plan, C.byref(tm)))
        if i >= 2:
            rows.append((tm.forward_ms, tm.march_ms, tm.inverse_ms, tm.total_ms))
    L.bldfm_plan_set_profiling(plan, 0)
    fwd, march, inv, total = (float(np.median(c)) for c in zip(*rows))
    out_bytes = 2 * nlv * n * n * 8
    peak, src_peak = hbm_peak()
    # algorithmic bytes of the back-transform: half-plane spectra in, intermediate written + read, real fields out
    nrow = geom.nly // 2 + 1
    bt_bytes = 2 * nlv * (nrow * geom.nlx * 16 + 2 * nrow * n * 16 + n * n * 8)
    M = geom.nlx * geom.nly - 1
    S = len(z) - 1

    # parity: linearity in the source (size-independent property) and the level-0 flux == padded source mass
    c1 = out_c[nlv - 1].clone()
    d_src2 = d_src * 2.5
    _lib.check(L.bldfm_solve(plan, C.byref(prob), lvp, nlv, d_src2.data_ptr(), flags, out_c.data_ptr(), out_f.data_ptr()))
    torch.cuda.synchronize()
    lin = float((out_c[nlv - 1] - 2.5 * c1).abs().max() / c1.abs().max())
    mass = float(out_f[0].sum() / d_src2.sum())
    del keep
    res = {
        "workload": "BASELINE config 3: 1024x1024, n=128 (208 levels), modes 1024x1024, domain 8000 m (padded 3072^2), "
                    "neutral MOST, point source, all 129 levels to z_m written out (2.16 GB), FP64, non-footprint",
        "device_ms": total, "forward_ms": fwd, "march_ms": march, "backtransform_ms": inv,
        "solves_per_s": 1e3 / total, "mode_levels_per_s": M * S * 1e3 / total,
        "output_gbs": out_bytes / (total * 1e-3) * 1e-9,
        "roofline_backtransform": {"bound": "hbm", "unit": "GB/s", "achieved": bt_bytes / (inv * 1e-3) * 1e-9,
                                   "peak": peak, "peak_source": src_peak, "frac": bt_bytes / (inv * 1e-3) * 1e-9 / peak,
                                   "algorithmic_bytes": bt_bytes, "fields": 2 * nlv},
        "hbm_floor_ms_outputs_only": out_bytes / (peak * 1e9) * 1e3,
        # (the flux at level 0 IS the source, low-passed to the retained modes: its mass inside the cropped
        # domain differs from the source's by the truncation ripple, ~1e-7 here)
        "parity": {"linearity_rel_err": lin, "surface_flux_mass_ratio": mass,
                   "ok": bool(lin <= 1e-12 and abs(mass - 1.0) <= 1e-5),
                   "oracle": "256^2 x 33 replica vs oracle: tests/test_gpu_parity.py::test_baseline_config3_replica_against_oracle"},
    }
    return res


# ------------------------------------------------------------------------------------------------
# config 5: one oversized footprint solve, ky-slab sharded over the ranks
# ------------------------------------------------------------------------------------------------
def leg_config5(torch, dist, rank, world, local, oracle_check=None, reps=5, sizes=((2048, 256), (4096, 256))):
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.sharded import release_peer_buffers, steady_state_transport_solver_sharded

    dev = f"cuda:{local}"
    out = {}
    all_ok = True
    for n, nz in sizes:
        dom = 32000.0 * n / 4096
        z, prof = vertical_profiles(nz, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
        kw = dict(srf_flx=np.zeros((n, n)), z=z, profiles=prof, domain=(dom, dom), levels=nz, modes=(n, n),
                  meas_pt=(dom / 2, dom / 2), footprint=True, precision="double")
        M, S = n * n - 1, len(z) - 1
        entry = {"n": n, "levels": len(z), "padded": 3 * n, "mode_levels": M * S, "march_gflop_reference_count": 86.0 * M * S * 1e-9}
        # unsharded reference point on rank 0's GPU (also the N = 1 number)
        ref_c = ref_f = None
        if rank == 0:
            _, ref_c, ref_f = bldfm_b200.steady_state_transport_solver(**kw)
        if world == 1:
            from bldfm_b200 import _lib
            L = _lib.lib()
            geom = _lib.geometry((n, n), (dom, dom), (n, n), None)
            plan = bldfm_b200.get_fft_manager().plan(geom, local)
            L.bldfm_plan_set_profiling(plan, 1)
            tm = _lib.Timings()
            ts = []
            for i in range(reps + 1):
                bldfm_b200.steady_state_transport_solver(**kw)
                _lib.check(L.bldfm_plan_last_timings(plan, C.byref(tm)))
                if i:
                    ts.append(tm.forward_ms + tm.march_ms + tm.inverse_ms)
            L.bldfm_plan_set_profiling(plan, 0)
            entry["single_gpu_ms"] = float(np.median(ts))
            entry["solves_per_s"] = 1e3 / entry["single_gpu_ms"]
            entry["mode_levels_per_s"] = M * S * entry["solves_per_s"]
            entry["equal_to_unsharded"] = True
        else:
            best = None
            for name, fused in (("nccl", False), ("fused", True)):
                rows = []
                for i in range(reps + 1):
                    tm = {}
                    _barrier(torch, dist, world)
                    c, f = steady_state_transport_solver_sharded(fused=fused, gather=True, return_device=True,
                                                                 timings=tm, **kw)
                    if i:
                        rows.append((tm["stage1_ms"] + tm["exchange_ms"] + tm["stage2_ms"], tm["stage1_ms"],
                                     tm["exchange_ms"], tm["stage2_ms"], tm["gather_ms"]))
                med = [float(np.median(col)) for col in zip(*rows)]
                solve_ms, s1, ex, s2, ga = _max_over_ranks(torch, dist, world, med, dev)
                equal = True
                if rank == 0:
                    equal = bool(np.array_equal(ref_c, np.squeeze(c.cpu().numpy())) and
                                 np.array_equal(ref_f, np.squeeze(f.cpu().numpy())))
                sent = tm["exchange_bytes_sent"]
                entry[name] = {"ms": solve_ms, "stage1_ms": s1, "exchange_ms": ex, "stage2_ms": s2,
                               "allgather_of_results_ms": ga, "exchange_bytes_sent_per_rank": sent,
                               "equal_to_unsharded": equal}
                if not fused:
                    entry[name]["exchange_gbs_per_rank"] = sent / (ex * 1e-3) * 1e-9 if ex > 0 else None
                all_ok &= equal
                if best is None or solve_ms < best:
                    best = solve_ms
            entry["ms"] = best
            entry["solves_per_s"] = 1e3 / best
            entry["mode_levels_per_s"] = M * S * 1e3 / best
            entry["equal_to_unsharded"] = bool(entry["nccl"]["equal_to_unsharded"] and entry["fused"]["equal_to_unsharded"])
        if oracle_check is not None and n <= 2048 and rank == 0:
            entry["oracle_rel_l2"] = oracle_check(kw, ref_c, ref_f)
            entry["oracle_ok"] = bool(max(entry["oracle_rel_l2"]) <= 1e-10)
            all_ok &= entry["oracle_ok"]
        out["replica_2048" if n == 2048 else f"full_{n}"] = entry
        # free the big plans before the next size
        bldfm_b200.reset_fft_manager()
        torch.cuda.empty_cache()
    if world > 1:
        release_peer_buffers()
    out["workload"] = ("BASELINE config 5: one footprint solve 4096x4096, n=256 (414 levels), padded 12288^2, FP64, ky-slab "
                       "sharded over the ranks (half-plane rows), exchange = one grouped NCCL send/recv or fused "
                       "peer stores + device flags; 2048^2 replica (same dx and wavenumber range) checked against the oracle")
    out["ok"] = bool(all_ok)
    return out


def host_link_probe(torch, dist, rank, world, local, nbytes=256 << 20):
    """Device->host copy bandwidth into page-locked memory: rank 0 alone, then every rank at once.  The second
    figure is the platform's ceiling for anything that delivers fields to the host from all GPUs together
    (on this pool: ~57 GB/s for one GPU alone, ~143 GB/s for eight together; scripts/d2h_probe.py)."""
    dev = f"cuda:{local}"
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    def bw(aligned):
        best = 0.0
        for _ in range(3):
            if aligned:
                # every repetition starts on all ranks at once: otherwise a rank that is late copies alone and
                # reports the rate of an idle link (the GPUs of a box need not share the host links evenly:
                # profiles/r2_d2h_shared_probe_n8.json)
                _barrier(torch, dist, world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                h.copy_(d, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * nbytes / (e0.elapsed_time(e1) * 1e-3) * 1e-9)
        return best

    h.copy_(d)
    _barrier(torch, dist, world)
    alone = bw(False) if rank == 0 else 0.0
    _barrier(torch, dist, world)
    together = bw(True)
    t = torch.tensor([together], dtype=torch.float64, device=dev)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    a = torch.tensor([alone], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
    return {"d2h_gbs_one_gpu_alone": float(a.item()), "d2h_gbs_all_gpus_together_aggregate": float(t.item()),
            "d2h_gbs_all_gpus_together_min_per_gpu": float(tmin.item()), "bytes_per_copy": nbytes,
            "note": "pinned host memory, CUDA events, repetitions start on all ranks at once; the aggregate is the "
                    "ceiling of every delivered-to-host number"}


# ------------------------------------------------------------------------------------------------
# config 4: 8 towers x T met steps, independent solves sharded over the ranks, final gather timed
# ------------------------------------------------------------------------------------------------
def config4(T, ntow=8, n=512):
    """BASELINE config 4: towers = generate_towers_grid(8, layout="grid", spacing_m=500, z_m=10,
    seed=0), local x/y by the equirectangular map around the domain centre; met = generate_synthetic_timeseries(
    1440, seed=0) with |L| floored at 50 m (the synthetic noise can drive L towards 0, which is ill-conditioned);
    per-solve grid as config 2."""
    from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
    from bldfm_b200.synthetic import generate_synthetic_timeseries, generate_towers_grid
    met_d = generate_synthetic_timeseries(n_timesteps=T, seed=0)
    mol = np.array(met_d["mol"])
    mol = np.where(np.abs(mol) < 50.0, np.where(mol < 0, -50.0, 50.0), mol)
    tw = generate_towers_grid(n_towers=ntow, layout="grid", spacing_m=500, z_m=10.0, seed=0)
    lat0 = float(np.mean([t["lat"] for t in tw]))
    lon0 = float(np.mean([t["lon"] for t in tw]))
    R = 6_371_000.0
    towers = []
    for t in tw:
        x = R * np.radians(t["lon"] - lon0) * np.cos(np.radians(lat0)) + 2000.0
        y = R * np.radians(t["lat"] - lat0) + 2000.0
        towers.append(Tower(t["name"], t["z_m"], float(x), float(y), t["lat"], t["lon"]))
    met = Met(ustar=met_d["ustar"], mol=mol.tolist(), wind_speed=met_d["wind_speed"], wind_dir=met_d["wind_dir"],
              timestamps=met_d["timestamps"])
    dom = Domain(nx=n, ny=n, xmax=4000.0, ymax=4000.0, nz=64, modes=(n, n))
    return Config(dom, towers, met, SolverOptions(footprint=True, precision="double"), Parallel())


def leg_config4(torch, dist, rank, world, local, T=1440, reps=2, oracle_check=None, host_link=None):
    import bldfm_b200
    from bldfm_b200 import _lib, interface
    from bldfm_b200.utils import ideal_source

    dev = f"cuda:{local}"
    ntow, n = 8, 512
    # the delivered variant keeps every footprint on the host (rank 0): 2 x 2 MiB per footprint
    need = T * ntow * 2 * n * n * 8
    try:
        avail = int(next(ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")).split()[1]) * 1024
    except (OSError, StopIteration, ValueError):
        avail = 64 << 30
    note = None
    while need > 0.4 * avail and T > 90:
        T //= 2
        need = T * ntow * 2 * n * n * 8
        note = f"met steps reduced to {T}: the delivered fields must fit in 40 % of the host's available memory"
    cfg = config4(T, ntow, n)
    L = _lib.lib()
    nfoot = T * ntow
    flux_map = ideal_source((n, n), (4000.0, 4000.0), shape="circle") + 0.05
    res = {"workload": f"BASELINE config 4: {ntow} towers x {T} half-hourly met steps, footprints of config-2 size "
                       f"(512x512, 105 levels, FP64); {nfoot} footprints, {T} marches (one per met step, shared by the "
                       "towers); strong scaling: the work is fixed, the ranks share it",
           "footprints": nfoot, "marches": T, "n_gpus": world}
    if note:
        res["note"] = note

    def timed(fn, reps):
        best = None
        ret = None
        for _ in range(reps):
            ret = None                     # drop the previous result: its shared segment becomes reusable
            _barrier(torch, dist, world)
            t0 = time.perf_counter()
            ret = fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return _max_over_ranks(torch, dist, world, [best], dev)[0], ret

    # ---- device-resident rate: fields stay in HBM (CUDA events on the plan's stream around the whole job)
    tasks = interface._multitower_tasks(cfg)
    _, _, owner, mine = interface._shard(cfg, tasks)
    geom = _lib.geometry((n, n), (4000.0, 4000.0), (n, n), None)
    plan = bldfm_b200.get_fft_manager().plan(geom, local)
    stream = torch.cuda.ExternalStream(L.bldfm_plan_stream(plan), device=local)
    lv = np.array([64], dtype=np.int64)
    lvp = lv.ctypes.data_as(C.POINTER(C.c_int64))
    chunk = 8 * ntow
    dev_c = torch.empty((chunk, n, n), dtype=torch.float64, device=dev)
    dev_f = torch.empty_like(dev_c)
    flags = _lib.FOOTPRINT | _lib.DOUBLE | _lib.OUT_ON_DEVICE | _lib.ASYNC
    flags |= {"exact": 0, "fma": _lib.MARCH_FMA, "sweep": _lib.MARCH_SWEEP,
              "auto": _lib.MARCH_AUTO}[bldfm_b200.config.MARCH_MODE]

    def device_job():
        t0 = time.perf_counter()
        tb = interface.TaskBatch(cfg, [tasks[t] for t in mine])
        prep = time.perf_counter() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        idx = list(range(len(tb.tasks)))
        for c0 in range(0, len(idx), chunk):
            parr, keep = tb.problems(idx[c0:c0 + chunk])
            _lib.check(L.bldfm_solve_batched(plan, len(parr), parr.ctypes.data, lvp, 1, None, flags,
                                             dev_c.data_ptr(), dev_f.data_ptr()))
        e1.record(stream)
        _lib.check(L.bldfm_plan_synchronize(plan))
        return e0.elapsed_time(e1) * 1e-3, prep, prep + tb.prep_seconds

    device_job()
    _barrier(torch, dist, world)
    ds, prep, prep_all = device_job()
    ds, prep, prep_all = _max_over_ranks(torch, dist, world, [ds, prep, prep_all], dev)
    res["device_resident"] = {"s": ds, "footprints_per_s": nfoot / ds,
                              "host_prep_exposed_s": prep, "host_prep_exposed_over_device": prep / ds,
                              "host_prep_total_s": prep_all, "host_prep_total_over_device": prep_all / ds,
                              "host_prep": "task grouping (exposed) + vectorised wind/profile restatement in blocks of 128 "
                                           "met rows, computed while the device works on the previous block",
                              "timing": "CUDA events on the plan's stream around the rank's whole share, max over ranks"}

    # ---- delivered: run_bldfm_parallel, every footprint in host memory of rank 0, gather inside the timing
    interface.run_bldfm_parallel(cfg, parallel_over="both")            # warm-up: creates + page-locks the segment
    dt, full = timed(lambda: interface.run_bldfm_parallel(cfg, parallel_over="both"), reps)
    phases = dict(interface.LAST_PARALLEL_PHASES)                     # of the last repetition, per rank
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, phases)
    else:
        gathered = [phases]
    link = host_link if host_link is not None else host_link_probe(torch, dist, rank, world, local)
    res["host_link"] = link
    res["delivered"] = {"s": dt, "footprints_per_s": nfoot / dt, "bytes_to_host": need,
                        "host_gbs": need / dt * 1e-9,
                        "frac_of_host_link": need / dt * 1e-9 / max(link["d2h_gbs_all_gpus_together_aggregate"], 1e-9),
                        "phases_by_rank_last_rep": [{k: round(v, 4) for k, v in p.items()} for p in gathered],
                        "api": "bldfm_b200.run_bldfm_parallel(cfg, parallel_over='both'): result dict of every (tower, "
                               "timestep) on rank 0; each rank copies device->host over its own PCIe link into a "
                               "page-locked shared-memory segment; wall clock, barrier-bracketed, max over ranks"}

    # ---- reduced on the device: tower fluxes (measure) and time-mean footprints (aggregate)
    interface.run_bldfm_measure(cfg, flux_map)
    dtm, meas = timed(lambda: interface.run_bldfm_measure(cfg, flux_map), reps)
    res["measure"] = {"s": dtm, "footprints_per_s": nfoot / dtm,
                      "api": "bldfm_b200.interface.run_bldfm_measure: sum(footprint * flux_map) per tower and timestep"}
    interface.run_bldfm_aggregate(cfg)
    dta, agg = timed(lambda: interface.run_bldfm_aggregate(cfg), reps)
    res["aggregate"] = {"s": dta, "footprints_per_s": nfoot / dta,
                        "api": "bldfm_b200.interface.run_bldfm_aggregate: time-mean footprint per tower, summed on the "
                               "device, reduced over the ranks with one NCCL reduce"}

    # ---- parity (rank 0): sampled footprints bitwise vs single solves, measure vs host sums, aggregate vs np.mean
    ok = True
    if rank == 0:
        rng = np.random.default_rng(1)
        picks = [(int(rng.integers(ntow)), int(rng.integers(T))) for _ in range(6)] + [(0, 0), (ntow - 1, T - 1)]
        # (a launch of >= 4096 transforms takes the two-stage k_fft48 passes, a single solve k_fft24: same
        # mathematics, different rounding -- so the comparison with single solves is to 1e-12, not bitwise)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))          # noqa: E731
        worst_single = worst_meas = 0.0
        worst_oracle = None
        for k, (ti, mi) in enumerate(picks):
            tower = cfg.towers[ti]
            single = interface.run_bldfm_single(cfg, tower, met_index=mi)
            got = full[tower.name][mi]
            worst_single = max(worst_single, rel(got["flx"], single["flx"]), rel(got["conc"], single["conc"]))
            want = float(np.sum(got["flx"] * flux_map))
            worst_meas = max(worst_meas, abs(meas[tower.name]["flx"][mi] - want) / abs(want))
            if oracle_check is not None and k < 2:
                z, prof = interface._profiles_for(cfg, tower.z_m, cfg.met.get_step(mi))
                kwo = dict(srf_flx=np.zeros((n, n)), z=z, profiles=prof, domain=(4000.0, 4000.0), levels=64,
                           modes=(n, n), meas_pt=(tower.x, tower.y), footprint=True, precision="double")
                worst_oracle = max(worst_oracle or 0.0, *oracle_check(kwo, got["conc"], got["flx"]))
        bit = worst_single <= 1e-12
        t0 = cfg.towers[0].name
        mean = np.zeros((n, n))
        for r in full[t0]:
            mean += r["flx"]
        mean /= T
        agg_err = float(np.abs(agg[t0]["flx"] - mean).max() / np.abs(mean).max())
        mass = float(np.mean([full[t0][mi]["flx"].sum() for mi in range(0, T, max(1, T // 16))]))
        ok = bool(bit and worst_meas <= 1e-11 and agg_err <= 1e-12 and (worst_oracle is None or worst_oracle <= 1e-10))
        res["parity"] = {"sampled_footprints_max_rel_l2_vs_single_solves": worst_single, "samples": len(picks),
                         "sampled_footprints_max_rel_l2_vs_oracle": worst_oracle,
                         "measure_max_rel_err_vs_host_sum": worst_meas, "aggregate_max_rel_err_vs_host_mean": agg_err,
                         "mean_footprint_mass_in_domain": mass, "ok": ok}
    full = meas = agg = None
    res["ok"] = ok
    return res
