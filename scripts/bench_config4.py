"""BASELINE config 4: multitower timeseries -- 8 towers x T half-hourly met conditions, footprints of
config-2 size (512x512, 105 levels, FP64), march groups distributed over the ranks of a torchrun job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        scripts/bench_config4.py [--steps 1440] [--towers 8] [--chunk 8]

Per rank: its share of the T march groups (one march feeds the 8 towers).  Three delivery modes:
  device   footprints stay in HBM (pure compute rate)
  host     every footprint is copied to pinned host memory (4 MB each) and dropped
  measure  only the tower fluxes sum(footprint * flux_map) leave the GPU (f-4)
Rank 0 prints one JSON line with footprints/s per mode (time = max over ranks, barrier-bracketed).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def synthetic_met(T, seed=0):
    """Diurnal synthetic met series in the spirit of the reference's generate_synthetic_timeseries
    (src/bldfm/synthetic.py:12-109): unstable days, stable nights, |L| floored at 50 m."""
    rng = np.random.default_rng(seed)
    hours = np.arange(T) * 0.5
    day = np.sin(2 * np.pi * (hours - 6.0) / 24.0)
    ustar = np.clip(0.45 + 0.3 * np.clip(day, 0, None) + 0.03 * rng.normal(size=T), 0.1, 0.8)
    mol = np.where(day > 0, -1.0, 1.0) * np.maximum(50.0, 500.0 * (1.0 - 0.9 * np.abs(day)) + 20.0 * rng.normal(size=T))
    ws = np.clip(4.5 + 2.5 * np.clip(day, 0, None) + 0.4 * rng.normal(size=T), 1.0, 8.0)
    wd = (270.0 + 30.0 * rng.normal(size=T)) % 360.0
    return ustar, mol, ws, wd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1440)
    ap.add_argument("--towers", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=8, help="march groups per batched launch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import bldfm_b200
    from bldfm_b200 import _lib
    from bldfm_b200.distributed import shard_groups
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import compute_wind_fields, ideal_source

    torch.cuda.set_device(local)
    bldfm_b200.config.DEVICE = local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T, NT = args.steps, args.towers
    ustar, mol, ws, wd = synthetic_met(T)
    towers = [(1000.0 + 500.0 * (i % 4), 1500.0 + 500.0 * (i // 4)) for i in range(NT)]
    dom = (4000.0, 4000.0)
    n = 512
    flux_map = ideal_source((n, n), dom, shape="circle") + 0.05
    mine = shard_groups(list(range(T)), [1.0] * T, world)[rank]

    L = _lib.lib()
    geom = _lib.geometry((n, n), dom, (n, n), None)
    plan = bldfm_b200.get_fft_manager().plan(geom, local)
    lv = np.array([64], dtype=np.int64)
    lvp = lv.ctypes.data_as(C.POINTER(C.c_int64))
    base = _lib.FOOTPRINT | _lib.DOUBLE
    B = args.chunk * NT
    dev_c = torch.empty((B, n, n), dtype=torch.float64, device="cuda")
    dev_f = torch.empty_like(dev_c)
    host_c = torch.empty((4, B, n, n), dtype=torch.float64).pin_memory()
    host_f = torch.empty((4, B, n, n), dtype=torch.float64).pin_memory()
    wq = torch.empty((4, B, 1), dtype=torch.float64).pin_memory()
    wf = torch.empty((4, B, 1), dtype=torch.float64).pin_memory()
    zeros = np.zeros((n, n))

    def problems(groups):
        probs, keep = [], []
        for gi in groups:
            u, v = compute_wind_fields(ws[gi], wd[gi])
            z, prof = vertical_profiles(64, 10.0, (u, v), ustar=ustar[gi], mol=mol[gi])
            for (tx, ty) in towers:
                p, k = _lib.make_problem(z, prof, (tx, ty), 0.0)
                probs.append(p)
                keep.append(k)
        return (_lib.Problem * len(probs))(*probs), keep, len(probs)

    def run(mode):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slot = 0
        for c0 in range(0, len(mine), args.chunk):
            parr, keep, nb = problems(mine[c0:c0 + args.chunk])
            if mode == "measure":
                # enqueue-only with a ring of 4 pinned result pairs: the host prepares the next chunk meanwhile
                if (c0 // args.chunk) % 4 == 3:
                    _lib.check(L.bldfm_plan_synchronize(plan))
                _lib.check(L.bldfm_solve_batched_measure(plan, nb, parr, lvp, 1, None, base | _lib.ASYNC,
                                                         _lib.ptr(flux_map), wq[slot].data_ptr(), wf[slot].data_ptr()))
                slot = (slot + 1) % 4
            elif mode == "host":
                # pinned ring of 4 result sets; enqueue-only: the D2H of this chunk overlaps the next compute
                if (c0 // args.chunk) % 4 == 3:
                    _lib.check(L.bldfm_plan_synchronize(plan))
                _lib.check(L.bldfm_solve_batched(plan, nb, parr, lvp, 1, None, base | _lib.ASYNC,
                                                 host_c[slot].data_ptr(), host_f[slot].data_ptr()))
                slot = (slot + 1) % 4
            else:
                _lib.check(L.bldfm_solve_batched(plan, nb, parr, lvp, 1, None, base | _lib.OUT_ON_DEVICE | _lib.ASYNC,
                                                 dev_c.data_ptr(), dev_f.data_ptr()))
        _lib.check(L.bldfm_plan_synchronize(plan))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # host-side profile preparation alone (for reference)
    t0 = time.perf_counter()
    for c0 in range(0, len(mine), args.chunk):
        problems(mine[c0:c0 + args.chunk])
    prep = time.perf_counter() - t0

    out = {"config": 4, "towers": NT, "met_steps": T, "footprints": T * NT, "marches": T, "n_gpus": world,
           "chunk_groups": args.chunk, "host_prep_s_rank0": prep}
    run("device")
    for mode in ("device", "host", "measure"):
        dt = run(mode)
        out[f"{mode}_s"] = dt
        out[f"{mode}_footprints_per_s"] = T * NT / dt
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
