"""ky-slab sharded solve: world_size 1 on any GPU box, world_size 2 (NCCL all-to-all and
the fused peer-store transpose) when two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _problem(n=96, nz=16, footprint=True):
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.utils import ideal_source
    z, prof = vertical_profiles(nz, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    dom = (n * 7.8125, n * 7.8125)
    src = np.zeros((n, n)) if footprint else ideal_source((n, n), dom, src_loc=(dom[0] * 0.6, dom[1] * 0.7), shape="circle")
    return dict(srf_flx=src, z=z, profiles=prof, domain=dom, levels=[3, nz], modes=(n, n),
                meas_pt=(n * 3.9, n * 3.1), footprint=footprint, precision="double")


def _plain_and_sharded(kw, full, **skw):
    """(plain result, sharded result) in the default half-plane / real-output mode, or -- full=True --
    in the cross-check mode (every row marched, full complex passes)."""
    import bldfm_b200
    from bldfm_b200.sharded import steady_state_transport_solver_sharded
    bldfm_b200.config.FFT_FULL = full
    bldfm_b200.config.MARCH_FULL = full
    try:
        plain = bldfm_b200.steady_state_transport_solver(**kw)
        shard = steady_state_transport_solver_sharded(**skw, **kw)
    finally:
        bldfm_b200.config.FFT_FULL = False
        bldfm_b200.config.MARCH_FULL = False
    return plain, shard


def test_sharded_world1_equals_plain(gpu_lib):
    import bldfm_b200
    for footprint in (True, False):
        kw = _problem(footprint=footprint)
        # same kernels, same arithmetic per transform: bitwise equal to the plain solver in either mode
        for full in (False, True):
            (g0, c0, f0), (g1, c1, f1) = _plain_and_sharded(kw, full)
            assert np.array_equal(c0, c1) and np.array_equal(f0, f1), (footprint, full)
            for a, b in zip(g0, g1):
                assert np.array_equal(a, b)
        # and the two modes agree to round-off
        _, c2, f2 = bldfm_b200.steady_state_transport_solver(**kw)
        assert np.abs(c2 - c1).max() <= 1e-12 * np.abs(c1).max()
        assert np.abs(f2 - f1).max() <= 1e-12 * np.abs(f1).max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


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
        from bldfm_b200.sharded import release_peer_buffers, steady_state_transport_solver_sharded
        bldfm_b200.config.DEVICE = rank
        res = {}
        for full in (False, True):
            (_, c0, f0), (_, c1, f1) = _plain_and_sharded(_problem(footprint=False), full)
            res[f"non-footprint full={full}"] = bool(np.array_equal(c0, c1) and np.array_equal(f0, f1))
            kw = _problem()
            for fused in (False, True):
                for rep in range(2):
                    (_, c0, f0), (_, c1, f1) = _plain_and_sharded(kw, full, fused=fused)
                res[f"full={full} fused={fused}"] = bool(np.array_equal(c0, c1) and np.array_equal(f0, f1))
                _, (_, cs, fs) = _plain_and_sharded(kw, full, fused=fused, gather=False)
                nxl = c0.shape[-1] // world
                res[f"slab full={full} fused={fused}"] = bool(np.array_equal(cs, c0[..., rank * nxl:(rank + 1) * nxl]))
        # uneven blocks: 66 modes -> 34 half-plane rows -> blocks of 17; 50 -> 26 rows
        kw = _problem(n=100)
        kw["modes"] = (50, 66)
        (_, c0, f0), (_, c1, f1) = _plain_and_sharded(kw, False)
        res["truncated modes"] = bool(np.array_equal(c0, c1) and np.array_equal(f0, f1))
        release_peer_buffers()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_sharded_world2_nccl_and_fused(gpu_lib):
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
    for rank, res in out:
        assert all(res.values()), (rank, res)
