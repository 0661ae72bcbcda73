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


@pytest.mark.parametrize("G", [2, 4, 8])
def test_emulated_ranks_equal_plain_bitwise(gpu_lib, G):
    """The G rank programs (stage 1 -> exchange -> stage 2) run back to back on this one GPU: same kernels
    and launch geometry per rank as a real G-GPU run, so equality here pins the ky-slab decomposition itself
    on a single-GPU box; the NCCL / peer-store transport is what test_sharded_world2_* adds."""
    import bldfm_b200
    from bldfm_b200.sharded import solve_sharded_emulated
    for footprint in (True, False):
        kw = _problem(n=96, footprint=footprint)
        _, c0, f0 = bldfm_b200.steady_state_transport_solver(**kw)
        kw.pop("precision")
        c1, f1 = solve_sharded_emulated(G, **kw)
        assert np.array_equal(c0, c1) and np.array_equal(f0, f1), (G, footprint)
    # uneven row blocks and truncated modes: 66 modes -> 34 half-plane rows; 8 ranks -> blocks of 5 (last: 4)
    kw = _problem(n=96)
    kw["modes"] = (48, 66)
    _, c0, f0 = bldfm_b200.steady_state_transport_solver(**kw)
    kw.pop("precision")
    c1, f1 = solve_sharded_emulated(G, **kw)
    assert np.array_equal(c0, c1) and np.array_equal(f0, f1), G


def test_baseline_config5_replica_2048_sharded_against_oracle(gpu_lib, oracle):
    """BASELINE config 5 at the 2048^2 replica SURVEY.md 8(d) allows (same dx = 7.8 m and largest wavenumber
    as the 4096^2 x 256 case, n = 256 -> 414 levels): the 8-rank ky-slab decomposition against the oracle
    (<= 1e-10 rel-L2) and bitwise against the unsharded solve."""
    import bldfm_b200
    from bldfm_b200.pbl_model import vertical_profiles
    from bldfm_b200.sharded import solve_sharded_emulated
    from conftest import rel_l2
    n, nz = 2048, 256
    dom = 32000.0 * n / 4096
    z, prof = vertical_profiles(nz, 10.0, (-3.0, -4.0), ustar=0.4, mol=-50.0)
    kw = dict(srf_flx=np.zeros((n, n)), z=z, profiles=prof, domain=(dom, dom), levels=nz, modes=(n, n),
              meas_pt=(dom / 2, dom / 2), footprint=True)
    c1, f1 = solve_sharded_emulated(8, **kw)
    _, c0, f0 = bldfm_b200.steady_state_transport_solver(precision="double", **kw)
    assert np.array_equal(c0, c1) and np.array_equal(f0, f1)
    _, oc, of = oracle.solve(precision="double", nthreads=oracle.max_threads(), **kw)
    assert rel_l2(c1, oc) <= 1e-10 and rel_l2(f1, of) <= 1e-10
    assert abs(f1.sum() - of.sum()) <= 1e-12 * abs(of.sum())


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
