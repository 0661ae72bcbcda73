"""Independent solves over two GPUs (first row): run_bldfm_parallel with the shared-memory
gather, run_bldfm_measure and run_bldfm_aggregate under NCCL, each compared with the single-GPU drivers.
Skipped on a one-GPU box (the host logic is covered on CPU with gloo in test_distributed_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _config():
    from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
    towers = [Tower("A", 10.0, 400.0, 400.0), Tower("B", 10.0, 560.0, 320.0), Tower("C", 6.0, 240.0, 480.0)]
    met = Met(ustar=[0.4, 0.5, 0.3, 0.45, 0.35], mol=[-50.0, -80.0, 100.0, 1e9, -200.0],
              wind_speed=[4.0, 5.0, 3.0, 6.0, 2.0], wind_dir=[270.0, 250.0, 200.0, 10.0, 90.0])
    dom = Domain(nx=64, ny=48, xmax=960.0, ymax=720.0, nz=16, modes=(64, 48))
    return Config(dom, towers, met, SolverOptions(footprint=True, precision="double"), Parallel())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bldfm_b200
        from bldfm_b200 import interface
        from bldfm_b200.distributed import SharedResults
        from bldfm_b200.utils import ideal_source
        bldfm_b200.config.DEVICE = rank
        cfg = _config()
        flux_map = ideal_source((64, 48), (960.0, 720.0), shape="circle") + 0.05
        res = {}
        par = interface.run_bldfm_parallel(cfg, parallel_over="both")
        meas = interface.run_bldfm_measure(cfg, flux_map, chunk_groups=1)
        agg = interface.run_bldfm_aggregate(cfg, chunk_groups=1)
        part = interface.run_bldfm_parallel(cfg, gather=False)
        nloc = sum(r is not None for lst in part.values() for r in lst)
        if rank == 0:
            seg = next(iter(SharedResults._cache.values()))
            res["segment_pinned"] = bool(seg.pinned)
            one = interface.run_bldfm_multitower(cfg)          # single-GPU drivers on this rank's device
            ok_fields = ok_meas = ok_agg = True
            for t in cfg.towers:
                for mi in range(5):
                    ok_fields &= np.array_equal(par[t.name][mi]["flx"], one[t.name][mi]["flx"])
                    ok_fields &= np.array_equal(par[t.name][mi]["conc"], one[t.name][mi]["conc"])
                    ok_fields &= par[t.name][mi]["timestamp"] == mi
                    want = float(np.sum(one[t.name][mi]["flx"] * flux_map))
                    ok_meas &= abs(meas[t.name]["flx"][mi] - want) <= 1e-12 * abs(want)
                mean = np.mean([r["flx"] for r in one[t.name]], axis=0)
                ok_agg &= bool(np.abs(agg[t.name]["flx"] - mean).max() <= 1e-14 * np.abs(mean).max())
            res.update(fields=bool(ok_fields), measure=bool(ok_meas), aggregate=bool(ok_agg))
        else:
            res["empty_on_other_ranks"] = par == {} and meas == {} and agg == {}
        q.put((rank, nloc, res))
    finally:
        dist.destroy_process_group()


def test_parallel_measure_aggregate_two_gpus(gpu_lib):
    if gpu_lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(n for _, n, _ in out) == 15 and min(n for _, n, _ in out) >= 5
    for rank, _, res in out:
        assert all(res.values()), (rank, res)
