"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: sharding by march group, the final
gather and the run_bldfm_parallel protocol with a stand-in for the CUDA solve."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_groups_balanced_and_deterministic():
    from bldfm_b200.distributed import owner_of_tasks, shard_groups
    keys = [(10.0, i) for i in range(37)]
    costs = [1.0 + 0.15 * (1 + i % 4) for i in range(37)]
    a = shard_groups(keys, costs, 4)
    b = shard_groups(keys, costs, 4)
    assert a == b
    assert sorted(g for r in a for g in r) == list(range(37))
    loads = [sum(costs[g] for g in r) for r in a]
    assert max(loads) - min(loads) <= max(costs)
    owner = owner_of_tasks([0, 0, 1, 2, 36], a)
    assert owner[0] == owner[1]


def _fake_config():
    from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
    towers = [Tower("A", 10.0, 100.0, 200.0), Tower("B", 10.0, 300.0, 250.0), Tower("C", 5.0, 50.0, 60.0)]
    met = Met(ustar=[0.4, 0.5, 0.3, 0.45, 0.35], mol=[-50.0, -80.0, 100.0, 1e9, -200.0],
              wind_speed=[4.0, 5.0, 3.0, 6.0, 2.0], wind_dir=[270.0, 250.0, 200.0, 10.0, 90.0])
    return Config(Domain(nx=16, ny=12, xmax=400.0, ymax=300.0, nz=8, modes=(16, 12)), towers, met,
                  SolverOptions(footprint=True, precision="double"), Parallel())


def _fake_solve_tasks(config, tasks, surface_flux=None, cache=None, out=None, out_pinned=False, build_results=True):
    """Stand-in for the CUDA path: fields are a deterministic function of (tower, met index); like the real
    ``solve_tasks`` it delivers them into ``out`` (this rank's block of the shared-memory segment)."""
    from bldfm_b200 import interface
    res = []
    for k, (ti, mi) in enumerate(tasks):
        tower = config.towers[ti]
        step = config.met.get_step(mi)
        base = np.arange(12 * 16, dtype=np.float64).reshape(12, 16)
        conc, flx = base * (ti + 1) + mi, base - 7 * ti + mi * mi
        if out is not None:
            out[0][k, 0] = conc
            out[1][k, 0] = flx
        res.append(interface._result(tower, step, (None, None, None), conc, flx))
    return res


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from bldfm_b200 import interface
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        interface.solve_tasks = _fake_solve_tasks
        interface.make_grid = lambda *a, **k: (None, None, None)
        cfg = _fake_config()
        res = interface.run_bldfm_parallel(cfg, parallel_over="both")
        # a second call while the first result is alive must not overwrite it (fresh segment) ...
        res2 = interface.run_bldfm_parallel(cfg, parallel_over="towers")
        if rank == 0:
            a, b = res["A"][1]["conc"], res2["A"][1]["conc"]
            assert np.array_equal(a, b) and not np.shares_memory(a, b)
        # ... and once both are dropped the cached segment is reused
        del res2
        res3 = interface.run_bldfm_parallel(cfg, parallel_over="towers")
        if rank == 0:
            assert np.array_equal(res3["B"][4]["flx"], res["B"][4]["flx"])
        del res3
        part = interface.run_bldfm_parallel(cfg, parallel_over="time", gather=False)
        nloc = sum(r is not None for lst in part.values() for r in lst)
        if rank == 0:
            q.put(("full", {k: [(r["conc"], r["flx"], r["timestamp"]) for r in v] for k, v in res.items()}))
        else:
            assert res == {}
        q.put(("count", rank, nloc))
    finally:
        dist.destroy_process_group()


def test_run_bldfm_parallel_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(m[1] for m in msgs if m[0] == "full")
    counts = {m[1]: m[2] for m in msgs if m[0] == "count"}
    cfg = _fake_config()
    assert sum(counts.values()) == 15 and min(counts.values()) >= 5
    expect = _fake_solve_tasks(cfg, [(ti, mi) for ti in range(3) for mi in range(5)])
    k = 0
    for ti, tower in enumerate(cfg.towers):
        for mi in range(5):
            conc, flx, ts = full[tower.name][mi]
            assert np.array_equal(conc, expect[k]["conc"]) and np.array_equal(flx, expect[k]["flx"])
            assert ts == mi
            k += 1


def test_plan_tasks_groups_towers_by_height():
    from bldfm_b200.interface import plan_tasks
    cfg = _fake_config()
    tasks = [(ti, mi) for mi in range(5) for ti in range(3)]
    keys, tg = plan_tasks(cfg, tasks)
    assert len(keys) == 10                      # 2 distinct heights x 5 met steps
    assert tg[0] == tg[1] != tg[2]              # towers A and B (z_m = 10) share a march


def test_pin_to_local_cores_splits_the_allowed_cores():
    import os
    from bldfm_b200.distributed import pin_to_local_cores
    before = sorted(os.sched_getaffinity(0))
    try:
        assert pin_to_local_cores(0, 1) is None                    # nothing to split
        if len(before) >= 2:
            mine = pin_to_local_cores(1, 2)
            per = len(before) // 2
            assert mine == before[per:2 * per] and sorted(os.sched_getaffinity(0)) == mine
            os.sched_setaffinity(0, before)
            assert pin_to_local_cores(0, 2) == before[:per]
    finally:
        os.sched_setaffinity(0, before)


def test_run_bldfm_parallel_without_process_group_uses_a_local_segment(monkeypatch):
    """No torch.distributed: the same code path with one rank -- a process-local result segment, reused once the
    previous result is dropped, never overwritten while it is alive."""
    from bldfm_b200 import interface
    from bldfm_b200.distributed import SharedResults
    monkeypatch.setattr(interface, "solve_tasks", _fake_solve_tasks)
    monkeypatch.setattr(interface, "make_grid", lambda *a, **k: (None, None, None))
    cfg = _fake_config()
    r1 = interface.run_bldfm_parallel(cfg, parallel_over="both")
    r2 = interface.run_bldfm_parallel(cfg, parallel_over="time")
    expect = _fake_solve_tasks(cfg, [(ti, mi) for ti in range(3) for mi in range(5)])
    k = 0
    for tower in cfg.towers:
        for mi in range(5):
            for r in (r1, r2):
                assert np.array_equal(r[tower.name][mi]["conc"], expect[k]["conc"])
                assert np.array_equal(r[tower.name][mi]["flx"], expect[k]["flx"])
            k += 1
    assert not np.shares_memory(r1["A"][0]["conc"], r2["A"][0]["conc"])
    seg_ids = {id(s) for s in SharedResults._cache.values()}
    del r1, r2, r
    import gc
    gc.collect()
    r3 = interface.run_bldfm_parallel(cfg)
    assert {id(s) for s in SharedResults._cache.values()} == seg_ids      # the freed segment was reused
    assert np.array_equal(r3["C"][4]["flx"], expect[-1]["flx"])


def test_shard_groups_by_rank_speed():
    """Weighted LPT: a rank that works its share off 1.6x faster gets 1.6x the groups; equal speeds reproduce the
    unweighted assignment; the result is deterministic and covers every group once."""
    from bldfm_b200.distributed import shard_groups
    keys = list(range(1440))
    costs = [8.0] * 1440
    speeds = [11.6] * 4 + [18.6] * 4
    a = shard_groups(keys, costs, 8, speeds=speeds)
    assert sorted(g for r in a for g in r) == keys
    n = [len(r) for r in a]
    assert max(n[:4]) - min(n[:4]) <= 1 and max(n[4:]) - min(n[4:]) <= 1
    assert abs(n[4] / n[0] - 18.6 / 11.6) < 0.03
    finish = [8.0 * n[r] / speeds[r] for r in range(8)]
    assert max(finish) / min(finish) < 1.02
    assert a == shard_groups(keys, costs, 8, speeds=speeds)
    assert shard_groups(keys, costs, 8) == shard_groups(keys, costs, 8, speeds=[3.0] * 8)
    with pytest.raises(ValueError):
        shard_groups(keys, costs, 8, speeds=[1.0] * 7)
    with pytest.raises(ValueError):
        shard_groups(keys, costs, 2, speeds=[1.0, 0.0])


def test_link_aware_shares_of_a_delivered_job(monkeypatch):
    """interface._shard(delivered=True) with config.LINK_AWARE_SHARDING: shares follow the per-rank link rates and
    the bytes a group delivers; compute-bound jobs (delivered=False) and the default keep the equal split."""
    import bldfm_b200
    from bldfm_b200 import distributed as D, interface
    from scripts.bench_legs import config4
    cfg = config4(96, 8, 64)
    tasks = interface._multitower_tasks(cfg)
    monkeypatch.setattr(D, "world", lambda: (1, 4))
    monkeypatch.setattr(D, "link_rates", lambda: [10.0, 10.0, 20.0, 20.0])
    monkeypatch.setattr(bldfm_b200.config, "LINK_AWARE_SHARDING", False)
    _, ws, owner, mine = interface._shard(cfg, tasks, delivered=True)
    assert ws == 4 and np.bincount(owner, minlength=4).tolist() == [192] * 4
    monkeypatch.setattr(bldfm_b200.config, "LINK_AWARE_SHARDING", True)
    _, _, owner, mine = interface._shard(cfg, tasks, delivered=True)
    assert np.bincount(owner, minlength=4).tolist() == [128, 128, 256, 256]
    assert mine == [t for t in range(len(tasks)) if owner[t] == 1]
    _, _, owner, _ = interface._shard(cfg, tasks, delivered=False)
    assert np.bincount(owner, minlength=4).tolist() == [192] * 4
