"""One oversized solve sharded by ky-slab over the GPUs of a node.

Every rank (one process per GPU, ``torch.distributed`` initialised with NCCL) calls
``steady_state_transport_solver_sharded`` with the SAME arguments.  Rank r

  1. marches its block of the half-plane rows ``ky in [0, nly/2]`` (modes are independent, and the spectra
     of a real source are conjugate-symmetric, so the other half-plane is never marched) and x-transforms
     them (``bldfm_sharded_stage1``).  Blocks hold ``Rp = ceil((nly/2+1)/G)`` rows;
  2. exchanges column blocks with all peers -- the ONLY collective of the solve: either an
     ``all_to_all_single`` per field over NCCL/NVLink, or (``fused=True``) no collective at all: the
     x-transform kernel stores its output straight into the peers' receive buffers through CUDA-IPC
     mapped pointers, so the transpose rides on the kernel's own stores;
  3. y-transforms its ``nx/G`` columns into real slabs (``bldfm_sharded_stage2``; real-output pass, two
     columns per complex transform).

``config.MARCH_FULL`` selects the cross-check variant instead: every retained row is marched
(``nly/G`` rows per rank) and both passes are full complex transforms.

The slabs are optionally all-gathered into the full fields.  float64.  In non-footprint mode every rank
holds the whole source and computes only its own rows of the source spectrum (the x-pass over the ny
source rows is replicated on every rank -- about 1 % of the march work -- so the forward transform
needs no exchange at all).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import config
from .distributed import world
from .fft_manager import get_fft_manager
from .solver import _flags, _geometry, _levels_array, make_grid


class _PeerBuffers:
    """Receive buffers allocated with cudaMalloc and mapped into every peer through CUDA IPC."""

    def __init__(self, nbytes_each, device, rank, nranks):
        import torch.distributed as dist

        L = _lib.lib()
        self.device, self.rank, self.nranks, self.nbytes = device, rank, nranks, nbytes_each
        self.local = []
        handles = []
        for _ in range(2):
            p = C.c_void_p()
            _lib.check(L.bldfm_device_alloc(device, nbytes_each, C.byref(p)))
            self.local.append(p.value)
            h = C.create_string_buffer(64)
            _lib.check(L.bldfm_ipc_export(p, h))
            handles.append(h.raw)
        gathered = [None] * nranks
        dist.all_gather_object(gathered, handles)
        self.opened = []
        self.peer = [[None] * nranks, [None] * nranks]   # [p|q][rank] -> device pointer valid here
        for r in range(nranks):
            for k in range(2):
                if r == rank:
                    self.peer[k][r] = self.local[k]
                else:
                    q = C.c_void_p()
                    _lib.check(L.bldfm_ipc_open(device, gathered[r][k], C.byref(q)))
                    self.opened.append(q.value)
                    self.peer[k][r] = q.value

    def close(self):
        L = _lib.lib()
        for q in self.opened:
            L.bldfm_ipc_close(self.device, C.c_void_p(q))
        for p in self.local:
            L.bldfm_device_free(self.device, C.c_void_p(p))
        self.opened, self.local = [], []


_peer_cache = {}


def _peer_buffers(key, nbytes, device, rank, nranks):
    pb = _peer_cache.get(key)
    if pb is None or pb.nbytes < nbytes:
        if pb is not None:
            pb.close()
        pb = _PeerBuffers(nbytes, device, rank, nranks)
        _peer_cache[key] = pb
    return pb


def release_peer_buffers():
    for pb in _peer_cache.values():
        pb.close()
    _peer_cache.clear()


def steady_state_transport_solver_sharded(srf_flx, z, profiles, domain, levels, modes=(512, 512),
                                          meas_pt=(0.0, 0.0), srf_bg_conc=0.0, footprint=True,
                                          halo=None, precision="double", gather=True, fused=False,
                                          return_device=False):
    """ky-slab sharded version of ``steady_state_transport_solver`` (same arguments).

    Returns ``(grid, conc, flx)`` like the single-GPU solver when ``gather`` is true (on every rank);
    otherwise the rank's column slab ``[..., ny, nx/G]`` with ``grid`` restricted to it.
    """
    import torch
    import torch.distributed as dist

    if precision != "double":
        raise ValueError("sharded solve: precision must be 'double'")
    rank, G = world()
    dev_index = config.DEVICE
    device = torch.device("cuda", dev_index)
    q0 = np.asarray(srf_flx)
    ny, nx = q0.shape
    geom = _geometry(q0.shape, domain, modes, halo)
    flags = _flags(footprint, False, precision) | _lib.ASYNC
    src = None if footprint else _lib.as_f64(q0)
    srcp = None if src is None else _lib.ptr(src)
    lv, lv64 = _levels_array(levels)
    nlv = len(lv64)
    herm = not config.MARCH_FULL
    if nx % G or (not herm and geom.nly % G):
        raise ValueError("sharded solve needs nx (and nly with MARCH_FULL) divisible by the number of ranks")
    nxl = nx // G
    # rows of one rank's block in the exchange; the receiver sees [G*rows][nx/G] per field
    rows = -(-(geom.nly // 2 + 1) // G) if herm else geom.nly // G
    if herm and rank * rows >= geom.nly // 2 + 1:
        raise ValueError("sharded solve: more ranks than row blocks of the half-plane")
    L = _lib.lib()
    plan = get_fft_manager().plan(geom, dev_index)
    stream = torch.cuda.ExternalStream(L.bldfm_plan_stream(plan), device=device)
    prob, keep = _lib.make_problem(z, profiles, meas_pt, srf_bg_conc)
    lvp = lv64.ctypes.data_as(C.POINTER(C.c_int64))

    out = torch.empty((2, nlv, ny, nxl), dtype=torch.float64, device=device)
    field_elems = nlv * G * rows * nxl            # complex elements of one of p / q on the receiver

    with torch.cuda.stream(stream):
        if fused and G > 1:
            pb = _peer_buffers((dev_index, G), field_elems * 16, dev_index, rank, G)
            # pointer tables: where THIS rank's row block starts inside every peer's receive buffer
            off = rank * rows * nxl * 16
            tab = torch.tensor([[pb.peer[k][r] + off for r in range(G)] for k in range(2)],
                               dtype=torch.int64, device=device)
            send = torch.empty(1, dtype=torch.complex128, device=device)    # unused placeholder
            dist.barrier()                                                 # peers finished reading (WAR)
            _lib.check(L.bldfm_sharded_stage1(plan, C.byref(prob), lvp, nlv, srcp, flags, rank, G,
                                              send.data_ptr(), send.data_ptr(),
                                              tab[0].data_ptr(), tab[1].data_ptr()))
            stream.synchronize()                                           # my stores have landed
            dist.barrier()                                                 # ... and everybody's
            recv_p, recv_q = pb.local
        else:
            send = torch.empty((2, nlv, G, rows, nxl), dtype=torch.complex128, device=device)
            _lib.check(L.bldfm_sharded_stage1(plan, C.byref(prob), lvp, nlv, srcp, flags, rank, G,
                                              send[0].data_ptr(), send[1].data_ptr(), None, None))
            if G > 1:
                recv = torch.empty_like(send)
                for k in range(2):
                    for l in range(nlv):
                        dist.all_to_all_single(recv[k, l], send[k, l])
            else:
                recv = send
            recv_p, recv_q = recv[0].data_ptr(), recv[1].data_ptr()
        _lib.check(L.bldfm_sharded_stage2(plan, nlv, flags, rank, G, recv_p, recv_q,
                                          out[0].data_ptr(), out[1].data_ptr()))
        if gather and G > 1:
            full = torch.empty((G, 2, nlv, ny, nxl), dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(full, out)
            out = full.permute(1, 2, 3, 0, 4).reshape(2, nlv, ny, nx)
    stream.synchronize()
    del keep

    if return_device:
        return out[0], out[1]
    conc = np.squeeze(out[0].cpu().numpy())
    flx = np.squeeze(out[1].cpu().numpy())
    grid = make_grid(z, lv, domain, nx, ny)
    if not (gather or G == 1):
        sl = slice(rank * nxl, (rank + 1) * nxl)
        grid = tuple(a[..., sl] for a in grid)
    return grid, conc, flx
