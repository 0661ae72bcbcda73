"""Sharding of independent (tower, timestep) solves over the GPUs of one node.

The reference fans independent solves out over a ``ProcessPoolExecutor`` and gets results back by
pickle (src/bldfm/interface.py:241-326).  Here the unit of distribution is a MARCH GROUP -- all
towers that share one vertical march (same measurement height and met step) -- so that a march is
never computed on two ranks.  There is no data-path collective: every rank solves its groups; the
only communication is the final gather of the cropped real fields (``gather_fields``), which uses
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import os
from typing import Hashable, List, Sequence, Tuple

import numpy as np


def world() -> Tuple[int, int]:
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_groups(keys: Sequence[Hashable], costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Assign group indices to ranks: longest-processing-time greedy on `costs`, ties by index.

    Deterministic (every rank computes the same assignment without communicating).
    Returns ``assign[rank] = sorted list of group indices``.
    """
    order = sorted(range(len(keys)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    assign: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        assign[r].append(i)
        load[r] += float(costs[i])
    return [sorted(a) for a in assign]


def owner_of_tasks(task_group: Sequence[int], assign: List[List[int]]) -> np.ndarray:
    """owner[t] = rank that solves task t, given each task's group index."""
    owner_of_group = {}
    for r, groups in enumerate(assign):
        for g in groups:
            owner_of_group[g] = r
    return np.array([owner_of_group[g] for g in task_group], dtype=np.int64)


def gather_fields(local: "np.ndarray", owner: np.ndarray, dst: int = 0, device=None):
    """Final gather: `local` holds this rank's tasks (in global task order restricted to this rank),
    shape [n_local, ...]; returns on rank `dst` the full [n_tasks, ...] array in task order, on the
    other ranks None.  One padded ``dist.gather`` -- the only collective of the distributed path.
    """
    import torch
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return local
    counts = np.bincount(owner, minlength=ws)
    nmax = int(counts.max())
    item_shape = tuple(local.shape[1:])
    backend = dist.get_backend()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((nmax,) + item_shape, dtype=torch.from_numpy(local[:0]).dtype, device=device)
    if len(local):
        buf[: len(local)].copy_(torch.from_numpy(np.ascontiguousarray(local)))
    recv = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, recv, dst=dst)
    if rank != dst:
        return None
    out = np.empty((len(owner),) + item_shape, dtype=local.dtype)
    for r in range(ws):
        idx = np.nonzero(owner == r)[0]
        if len(idx):
            out[idx] = recv[r][: len(idx)].cpu().numpy()
    return out
