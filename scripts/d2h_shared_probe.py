"""Host-link probe for the result segment of run_bldfm_parallel (torchrun, one rank per GPU): every rank copies
1 GiB device->host, all ranks at once, into (a) its own cudaHostAlloc buffer and (b) its rows of the shared,
page-locked result segment (distributed.SharedResults: memfd + cudaHostRegister).  Per-rank GB/s for both.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_shared_probe.py
"""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import bldfm_b200
from bldfm_b200 import _lib, distributed as D

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bldfm_b200.config.DEVICE = local
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cores = D.pin_to_local_cores()
L = _lib.lib()
ITEMS = 256                                   # 256 x 2 MiB per half -> 1 GiB per rank over both halves
NB = ITEMS * 512 * 512 * 8
dev = torch.zeros(NB, dtype=torch.uint8, device="cuda")


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def rate(ptrs):
    best = 0.0
    for _ in range(3):
        sync()
        t0 = time.perf_counter()
        for p in ptrs:
            _lib.check(L.bldfm_memcpy_d2h(local, C.c_void_p(p), C.c_void_p(dev.data_ptr()), NB))
        dt = time.perf_counter() - t0
        best = max(best, len(ptrs) * NB / dt * 1e-9)
        sync()
    return best


own = C.c_void_p()
_lib.check(L.bldfm_host_alloc(NB, C.byref(own)))
r_own = rate([own.value, own.value])
seg = D.SharedResults.acquire((1, 512, 512), np.full(world, ITEMS))
c, f = seg.local_block()
r_seg = rate([c.ctypes.data, f.ctypes.data])
t = torch.tensor([r_own, r_seg, float(seg.pinned)], dtype=torch.float64, device="cuda")
allv = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, t)
else:
    allv = [t]
if rank == 0:
    print(json.dumps({"world": world, "bytes_per_rank_per_copy": NB, "cores_rank0": sorted(cores) if cores else None,
                      "own_cudaHostAlloc_gbs_per_rank": [round(float(v[0]), 2) for v in allv],
                      "shared_registered_segment_gbs_per_rank": [round(float(v[1]), 2) for v in allv],
                      "segment_page_locked": [bool(v[2]) for v in allv],
                      "note": "all ranks copy at once; synchronous cudaMemcpy, wall clock, best of 3"}))
