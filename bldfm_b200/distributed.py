"""Sharding of independent (tower, timestep) solves over the GPUs of one node.

The reference fans independent solves out over a ``ProcessPoolExecutor`` and gets results back by
pickle (src/bldfm/interface.py:241-326).  Here the unit of distribution is a MARCH GROUP -- all
towers that share one vertical march (same measurement height and met step) -- so that a march is
never computed on two ranks.  There is no data-path collective: every rank solves its groups; the
only communication is the final gather of the cropped real fields.  On one node that gather is a
page-locked shared-memory segment (``SharedResults``): every rank copies its fields device->host over its
own PCIe link straight into the segment rank 0 maps, plus one barrier (``torch.distributed``: NCCL on GPUs,
gloo in the CPU tests).  ``gather_fields`` (a padded ``dist.gather``) remains for callers that hold their
fields in ordinary host arrays.
"""

from __future__ import annotations

import mmap
import os
import threading
import weakref
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


def pin_to_local_cores(local_rank=None, local_world=None):
    """One process per GPU on one node: give this rank its own equal slice of the host cores the job may use
    (``os.sched_setaffinity``).  Eight ranks that all float over the same cores disturb one another exactly
    where a sub-millisecond solve is sensitive -- the Python/ctypes path and the pinned-memory copies; call
    it before the first solve so that the pinned buffers are allocated (first-touched) from the pinned
    cores.  Returns the cores now in use, or None when there is nothing to split."""
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else int(local_rank)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))) \
        if local_world is None else int(local_world)
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    per = len(cores) // max(local_world, 1)
    if local_world <= 1 or per < 1:
        return None
    mine = cores[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    return mine


def shard_groups(keys: Sequence[Hashable], costs: Sequence[float], world_size: int,
                 speeds: Sequence[float] = None) -> List[List[int]]:
    """Assign group indices to ranks: longest-processing-time greedy on `costs`, ties by index.

    ``speeds`` (optional, one positive number per rank, the same on every rank): relative rate at which a rank
    works its share off -- e.g. its host-link bandwidth when every result is delivered to the host; a group goes
    to the rank that would finish it first.  Deterministic (every rank computes the same assignment without
    communicating).  Returns ``assign[rank] = sorted list of group indices``.
    """
    if speeds is None:
        speeds = [1.0] * world_size
    if len(speeds) != world_size or not all(float(v) > 0.0 for v in speeds):
        raise ValueError("speeds must hold one positive number per rank")
    order = sorted(range(len(keys)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    assign: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        c = float(costs[i])
        r = min(range(world_size), key=lambda k: ((load[k] + c) / float(speeds[k]), k))
        assign[r].append(i)
        load[r] += c
    return [sorted(a) for a in assign]


_LINK_RATES = {}


def link_rates(nbytes: int = 128 << 20):
    """Device->host copy rate of every rank into page-locked memory while ALL ranks copy at once (GB/s, the same
    list on every rank; measured once per process and world size, ~10 ms).  On a box whose GPUs do not share the
    host links evenly (measured on this pool: 11.6 GB/s for four of eight GPUs, 18.6 GB/s for the other four,
    profiles/r2_d2h_shared_probe_n8.json) these are the speeds a delivered-to-host job should be sharded by."""
    rank, ws = world()
    if ws in _LINK_RATES:
        return _LINK_RATES[ws]
    import ctypes as C
    import time
    from . import _lib, config
    import torch
    import torch.distributed as dist
    L = _lib.lib()
    dev, host = C.c_void_p(), C.c_void_p()
    _lib.check(L.bldfm_device_alloc(config.DEVICE, nbytes, C.byref(dev)))
    try:
        _lib.check(L.bldfm_host_alloc(nbytes, C.byref(host)))
        try:
            best = 0.0
            for rep in range(3):                      # first repetition warms the buffers up
                if ws > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                _lib.check(L.bldfm_memcpy_d2h(config.DEVICE, host, dev, nbytes))
                dt = time.perf_counter() - t0
                if rep:
                    best = max(best, nbytes / dt * 1e-9)
        finally:
            L.bldfm_host_free(host)
    finally:
        L.bldfm_device_free(config.DEVICE, dev)
    if ws > 1:
        out = [None] * ws
        dist.all_gather_object(out, float(best))
    else:
        out = [float(best)]
    _LINK_RATES[ws] = [max(float(v), 1e-3) for v in out]
    return _LINK_RATES[ws]


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


class SharedResults:
    """Final gather of one node without a collective on the data path (reference: results return to the parent
    by pickle through the pool's pipes, src/bldfm/interface.py:293-312).

    One anonymous shared-memory file (``memfd``; the other ranks open it through ``/proc/<pid>/fd``) holds
    ``conc`` and ``flx`` of ALL tasks, rank-major: ``[2][sum(counts)][*item_shape]`` float64, rank r owning the
    rows ``[start_r, start_r + counts[r])`` of both halves.  Every rank page-locks its own rows
    (``cudaHostRegister``) so that its device->host copies land there at full PCIe rate and asynchronously;
    rank 0 only reads.  Segments are cached per process and reused once no array handed out from them is
    alive any more (page-locking gigabytes costs more than the copies), otherwise a new one is created.
    """

    _cache: dict = {}
    _serial = 0

    def __init__(self, item_shape, counts, rank, ws, fd, creator_pid, creator_fd):
        self.item_shape = tuple(int(s) for s in item_shape)
        self.counts = np.asarray(counts, dtype=np.int64)
        self.rank, self.ws = rank, ws
        self.fd, self.creator_pid, self.creator_fd = fd, creator_pid, creator_fd
        self.item = int(np.prod(self.item_shape))
        self.total = int(self.counts.sum())
        self.start = np.concatenate([[0], np.cumsum(self.counts)[:-1]]).astype(np.int64)
        self.nbytes = max(mmap.PAGESIZE, 2 * self.total * self.item * 8)
        self.mm = mmap.mmap(fd, self.nbytes)
        self.busy = 0
        self._busy_lock = threading.RLock()
        self.pinned = False
        self._registered = []
        self._register()

    # ---- creation / reuse -------------------------------------------------------------------------
    @classmethod
    def acquire(cls, item_shape, counts):
        rank, ws = world()
        key = (ws, tuple(int(s) for s in item_shape), tuple(int(c) for c in counts))
        msg = [None]
        if rank == 0:
            seg = cls._cache.get(key)
            if seg is not None and seg.busy == 0:
                msg = [("reuse", seg.creator_pid, seg.creator_fd)]
            else:
                fd = _anon_file(2 * int(np.sum(counts)) * int(np.prod(item_shape)) * 8)
                msg = [("new", os.getpid(), fd)]
        if ws > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(msg, src=0)
        kind, pid, cfd = msg[0]
        seg = cls._cache.get(key)
        if kind == "reuse" and seg is not None and (seg.creator_pid, seg.creator_fd) == (pid, cfd):
            return seg
        if rank == 0:
            fd = cfd
        else:
            fd = os.open(f"/proc/{pid}/fd/{cfd}", os.O_RDWR)
        seg = cls(item_shape, counts, rank, ws, fd, pid, cfd)
        cls._cache[key] = seg        # a still-referenced older segment stays alive through its arrays
        seg.barrier()                # nobody writes before every rank has mapped the file
        return seg

    def _register(self):
        """Page-lock this rank's rows of both halves (best effort: without CUDA the copies are synchronous)."""
        try:
            from . import _lib
            L = _lib.lib()
            if _lib.device_count() < 1:
                return
        except OSError:
            return
        import ctypes as C
        base = C.addressof(C.c_char.from_buffer(self.mm))
        n = int(self.counts[self.rank]) * self.item * 8
        if n == 0:
            self.pinned = True
            return
        ok = True
        for half in range(2):
            off = (half * self.total + int(self.start[self.rank])) * self.item * 8
            lo = (off // mmap.PAGESIZE) * mmap.PAGESIZE
            hi = min(self.nbytes, -(-(off + n) // mmap.PAGESIZE) * mmap.PAGESIZE)
            if L.bldfm_host_register(C.c_void_p(base + lo), hi - lo) == _lib.OK:
                self._registered.append(base + lo)
            else:
                ok = False
        self.pinned = ok

    # ---- views ------------------------------------------------------------------------------------
    def _halves(self):
        # every view handed out keeps `root` alive (numpy collapses view chains onto the array made from the
        # buffer), so its finaliser fires exactly when the last of them is gone
        root = np.frombuffer(self.mm, dtype=np.float64, count=2 * self.total * self.item)
        with self._busy_lock:
            self.busy += 1
        weakref.finalize(root, self._released)
        return root.reshape((2, self.total) + self.item_shape)

    def _released(self):
        with self._busy_lock:
            self.busy -= 1

    def local_block(self):
        """(conc, flx) destination arrays ``[counts[rank], *item_shape]`` of this rank."""
        a = self._halves()
        s, n = int(self.start[self.rank]), int(self.counts[self.rank])
        return a[0, s:s + n], a[1, s:s + n]

    def all_blocks(self):
        """(conc, flx) of every task, rank-major ``[sum(counts), *item_shape]`` (meaningful after ``barrier``)."""
        a = self._halves()
        return a[0], a[1]

    def barrier(self):
        if self.ws > 1:
            import torch.distributed as dist

            dist.barrier()


def _anon_file(nbytes: int) -> int:
    """File descriptor of an anonymous shared-memory file of ``nbytes`` (memfd; /dev/shm as the fallback)."""
    nbytes = max(int(nbytes), mmap.PAGESIZE)
    try:
        fd = os.memfd_create("bldfm_b200_results", 0)
    except (AttributeError, OSError):
        import tempfile
        fd, path = tempfile.mkstemp(prefix="bldfm_b200_", dir="/dev/shm")
        os.unlink(path)
    os.ftruncate(fd, nbytes)
    return fd


def gather_small(local: "np.ndarray", owner: np.ndarray, dst: int = 0):
    """Gather small per-task results (``local[i]`` = this rank's i-th task in task order) onto ``dst`` in global
    task order; ``None`` elsewhere.  Meant for scalars per task (tower measurements), not fields."""
    rank, ws = world()
    if ws == 1:
        return local
    import torch.distributed as dist

    parts = [None] * ws if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local), parts, dst=dst)
    if rank != dst:
        return None
    out = np.empty((len(owner),) + tuple(local.shape[1:]), dtype=local.dtype)
    for r in range(ws):
        idx = np.nonzero(owner == r)[0]
        if len(idx):
            out[idx] = parts[r]
    return out


class _DevArray:
    """Minimal ``__cuda_array_interface__`` carrier so that torch can wrap a raw device pointer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def reduce_device_sums(acc, dst: int = 0):
    """Sum a ``solver.FieldAccumulator`` over the ranks onto ``dst`` (``dist.reduce`` over NCCL on the device
    buffers themselves -- the one collective of ``run_bldfm_aggregate``); returns the host arrays there, ``None``
    elsewhere."""
    import torch
    import torch.distributed as dist

    rank, ws = world()
    acc.synchronize()
    if ws > 1:
        dev = torch.device("cuda", acc.device)
        for p in acc.device_pointers():
            t = torch.as_tensor(_DevArray(p, acc.nbytes // 8), device=dev)
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(dev)
    return acc.fetch() if rank == dst else None
