"""BASELINE config 4, delivered variant only (every footprint in rank 0's host memory through run_bldfm_parallel),
under torchrun: wall time, per-rank phases, per-rank shares and link rates.  BLDFM_B200_LINK_AWARE=1 sizes the
shares by the measured host-link rate of every rank (distributed.link_rates).

    BLDFM_B200_LINK_AWARE=1 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        scripts/delivered_probe.py > profiles/r2_delivered_link_aware_n8.json
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

import bldfm_b200
from bldfm_b200 import distributed as D, interface
from scripts.bench_legs import config4

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bldfm_b200.config.DEVICE = local
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D.pin_to_local_cores()
T = int(os.environ.get("PROBE_STEPS", "1440"))
cfg = config4(T, 8, 512)
nfoot = T * 8


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


rates = D.link_rates() if bldfm_b200.config.LINK_AWARE_SHARDING else None
interface.run_bldfm_parallel(cfg, parallel_over="both")            # warm-up: creates + page-locks the segment
best, full = None, None
for _ in range(2):
    full = None
    sync()
    t0 = time.perf_counter()
    full = interface.run_bldfm_parallel(cfg, parallel_over="both")
    sync()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
t = torch.tensor([best], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
phases = [None] * world
if world > 1:
    dist.all_gather_object(phases, dict(interface.LAST_PARALLEL_PHASES))
else:
    phases = [dict(interface.LAST_PARALLEL_PHASES)]
tasks = interface._multitower_tasks(cfg)
_, _, owner, _ = interface._shard(cfg, tasks, delivered=True)
if rank == 0:
    # spot check: delivered fields equal single solves of the same (tower, met step)
    worst = 0.0
    names = [tw.name for tw in cfg.towers]
    for ti, mi in ((0, 0), (3, 17), (7, T - 1), (5, T // 2)):
        ref = bldfm_b200.run_bldfm_single(cfg, cfg.towers[ti], met_index=mi)
        got = full[names[ti]][mi]
        for k in ("conc", "flx"):
            worst = max(worst, float(np.linalg.norm(got[k] - ref[k]) / np.linalg.norm(ref[k])))
    s = float(t.item())
    print(json.dumps({"n_gpus": world, "footprints": nfoot, "link_aware": bool(bldfm_b200.config.LINK_AWARE_SHARDING),
                      "s": s, "footprints_per_s": nfoot / s, "host_gbs": nfoot * 2 * 512 * 512 * 8 / s * 1e-9,
                      "link_rates_gbs": rates, "footprints_per_rank": np.bincount(owner, minlength=world).tolist(),
                      "phases_by_rank_last_rep": [{k: round(v, 4) for k, v in p.items()} for p in phases],
                      "max_rel_l2_vs_single_solves": worst}))
