"""BASELINE config 5: one large footprint solve, ky-slab sharded over the ranks of a torchrun job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        scripts/bench_sharded.py [--n 4096] [--nz 256] [--reps 5]

Prints one JSON line (rank 0): per-solve device time (max over ranks, CUDA events on the plan's
stream) for the NCCL all-to-all variant and the fused peer-store variant, with and without the final
all-gather.  With G=1 it is the unsharded reference point.
"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--nz", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="compare against the unsharded solve on rank 0's GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import bldfm_b200
    from bldfm_b200 import _lib
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.sharded import release_peer_buffers, steady_state_transport_solver_sharded

    torch.cuda.set_device(local)
    bldfm_b200.config.DEVICE = local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    dom = 32000.0 * n / 4096
    z, prof = vertical_profiles(args.nz, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    kw = dict(srf_flx=np.zeros((n, n)), z=z, profiles=prof, domain=(dom, dom), levels=args.nz, modes=(n, n),
              meas_pt=(dom / 2, dom / 2), footprint=True, precision="double")
    M = n * n - 1
    S = len(z) - 1

    geom = _lib.geometry((n, n), (dom, dom), (n, n), None)
    plan = bldfm_b200.get_fft_manager().plan(geom, local)
    stream = torch.cuda.ExternalStream(_lib.lib().bldfm_plan_stream(plan), device=local)

    def run(fused, gather):
        # device time of the whole sharded solve: CUDA events on the plan's stream (the exchange is
        # enqueued on it as well), median over the repetitions, max over ranks
        times = []
        for i in range(args.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c, f = steady_state_transport_solver_sharded(fused=fused, gather=gather, return_device=True, **kw)
            e1.record(stream)
            torch.cuda.synchronize()
            if i:
                times.append(e0.elapsed_time(e1) * 1e-3)
        t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), c, f

    out = {"config": 5, "n": n, "nz_levels": len(z), "padded": 3 * n, "n_gpus": world,
           "mode_levels": M * S, "march_gflop": 86.0 * M * S * 1e-9}
    variants = [("nccl_a2a", False)] + ([("fused_p2p", True)] if world > 1 else [])
    for name, fused in variants:
        for gather in (False, True):
            t, c, f = run(fused, gather)
            out[f"{name}{'_gather' if gather else ''}_ms"] = t * 1e3
    out["solves_per_s_best"] = 1e3 / min(v for k, v in out.items() if k.endswith("_ms"))
    out["mode_levels_per_s_best"] = out["solves_per_s_best"] * M * S
    if args.check:
        _, c, f = run(False, True)
        if rank == 0:
            _, c0, f0 = bldfm_b200.steady_state_transport_solver(**kw)
            out["equal_to_unsharded"] = bool(np.array_equal(c0, np.squeeze(c.cpu().numpy())) and
                                             np.array_equal(f0, np.squeeze(f.cpu().numpy())))
            out["flx_sum"] = float(f0.sum())
    release_peer_buffers()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
