"""Host-link probe (torchrun, one rank per GPU): device->host copy bandwidth into page-locked memory, every rank
alone and all ranks at once, with the pages placed (a) wherever the first touch puts them and (b) bound to the
NUMA node the rank's GPU hangs on (mbind).  Explains the ceiling of every "deliver the fields to the host" number.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_probe.py
"""
import ctypes
import json
import mmap
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NB = 1 << 30
dev = torch.empty(NB, dtype=torch.uint8, device="cuda")


def bdf(i):
    p = torch.cuda.get_device_properties(i)
    try:
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except AttributeError:
        return None


def gpu_node(i):
    try:
        return int(Path(f"/sys/bus/pci/devices/{bdf(i)}/numa_node").read_text())
    except (OSError, ValueError, TypeError):
        return -1


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def bw(host_ptr_tensor, reps=3):
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host_ptr_tensor.copy_(dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, NB / (e0.elapsed_time(e1) * 1e-3) * 1e-9)
    return best


def measure(host):
    host.copy_(dev)               # touch + warm
    alone = []
    for r in range(world):
        barrier()
        if r == rank:
            alone.append(bw(host))
        barrier()
    barrier()
    together = bw(host)
    t = torch.tensor([alone[0], together], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
    else:
        allv = [t]
    return [float(v[0]) for v in allv], [float(v[1]) for v in allv]


out = {"world": world}
info = {"gpu_numa_node": gpu_node(local), "cpus": sorted(os.sched_getaffinity(0))[:4]}
try:
    st = Path("/proc/self/status").read_text()
    info["mems_allowed"] = [ln.split(":")[1].strip() for ln in st.splitlines() if ln.startswith("Mems_allowed_list")][0]
    info["nodes"] = sorted(p.name for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
except OSError:
    pass
# (a) cudaHostAlloc, first touch
host = torch.empty(NB, dtype=torch.uint8).pin_memory()
a_alone, a_tog = measure(host)
del host
# (b) anonymous mapping bound to the GPU's NUMA node, then cudaHostRegister
res_b = None
node = info["gpu_numa_node"]
if node >= 0:
    libc = ctypes.CDLL(None, use_errno=True)
    mm = mmap.mmap(-1, NB)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    mask = ctypes.c_ulong(1 << node)
    MPOL_BIND, SYS_mbind = 2, 237
    rc = libc.syscall(SYS_mbind, ctypes.c_void_p(addr), ctypes.c_ulong(NB), MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(64), 0)
    info["mbind_rc"] = rc if rc == 0 else f"errno {ctypes.get_errno()}"
    if rc == 0:
        import numpy as np
        arr = np.frombuffer(mm, dtype=np.uint8)
        arr[::4096] = 1
        cudart = torch.cuda.cudart()
        r = cudart.cudaHostRegister(addr, NB, 0)
        info["register_rc"] = int(r)
        hostt = torch.from_numpy(arr)
        b_alone, b_tog = measure(hostt)
        res_b = (b_alone, b_tog)
        cudart.cudaHostUnregister(addr)
infos = [None] * world
if world > 1:
    dist.all_gather_object(infos, info)
else:
    infos = [info]
if rank == 0:
    out["ranks"] = infos
    out["first_touch"] = {"alone_gbs": a_alone, "together_gbs": a_tog, "aggregate_together_gbs": sum(a_tog)}
    if res_b:
        out["numa_bound"] = {"alone_gbs_rank0": res_b[0], "together_gbs": res_b[1], "aggregate_together_gbs": sum(res_b[1])}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
