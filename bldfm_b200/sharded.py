"""One oversized solve sharded by ky-slab over the GPUs of a node.

Every rank (one process per GPU, ``torch.distributed`` initialised with NCCL) calls
``steady_state_transport_solver_sharded`` with the SAME arguments.  Rank r

  1. marches its block of the half-plane rows ``ky in [0, nly/2]`` (modes are independent, and the spectra
     of a real source are conjugate-symmetric, so the other half-plane is never marched) and x-transforms
     them (``bldfm_sharded_stage1``).  Blocks hold ``Rp = ceil((nly/2+1)/G)`` rows;
  2. exchanges column blocks with all peers -- the ONLY collective of the solve: either one grouped
     send/receive exchange for all fields over NCCL/NVLink, or (``fused=True``) no collective at all: the
     x-transform kernel stores its output straight into the peers' receive buffers through CUDA-IPC
     mapped pointers, so the transpose rides on the kernel's own stores, and per-peer flags in device
     memory (no host barrier) order the y-transform behind them;
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
    """Receive buffers of the fused transpose: ONE cudaMalloc per rank, mapped into every peer through CUDA IPC.

    Layout: ``[set 0: p | q][set 1: p | q][flags: 2 sets x G uint64]``.  The two buffer sets alternate from
    solve to solve, which makes the write-after-read hazard vanish without any extra handshake: rank r stores
    into set ``s % 2`` of peer d during solve s; d last read that set in solve s-2, and r has already seen d's
    flag of solve s-1, which d raised after its stage 2 of solve s-2 (stream order on d).
    """

    def __init__(self, nbytes_each, device, rank, nranks):
        import torch.distributed as dist

        L = _lib.lib()
        self.device, self.rank, self.nranks = device, rank, nranks
        self.nbytes = (int(nbytes_each) + 255) // 256 * 256
        self.flag_off = 4 * self.nbytes
        total = self.flag_off + 2 * nranks * 8
        p = C.c_void_p()
        _lib.check(L.bldfm_device_alloc(device, total, C.byref(p)))
        self.base = p.value
        _lib.check(L.bldfm_device_memset(device, p, 0, total))
        h = C.create_string_buffer(64)
        _lib.check(L.bldfm_ipc_export(p, h))
        gathered = [None] * nranks
        dist.all_gather_object(gathered, h.raw)
        self.opened = []
        self.peer_base = [None] * nranks        # base of every rank's allocation as mapped HERE
        for r in range(nranks):
            if r == rank:
                self.peer_base[r] = self.base
            else:
                q = C.c_void_p()
                _lib.check(L.bldfm_ipc_open(device, gathered[r], C.byref(q)))
                self.opened.append(q.value)
                self.peer_base[r] = q.value
        self.seq = 0
        dist.barrier()                          # every mapping exists and every flag array is zeroed

    def recv(self, which, k):
        """Local receive buffer of buffer set `which` (0/1) and field kind k (0: p, 1: q)."""
        return self.base + (2 * which + k) * self.nbytes

    def peer_recv(self, r, which, k):
        return self.peer_base[r] + (2 * which + k) * self.nbytes

    def flags(self, which):
        return self.base + self.flag_off + which * self.nranks * 8

    def peer_flag_slot(self, r, which):
        """This rank's slot in rank r's flag array of buffer set `which`."""
        return self.peer_base[r] + self.flag_off + (which * self.nranks + self.rank) * 8

    def close(self):
        L = _lib.lib()
        for q in self.opened:
            L.bldfm_ipc_close(self.device, C.c_void_p(q))
        if self.base:
            L.bldfm_device_free(self.device, C.c_void_p(self.base))
        self.opened, self.base = [], None


_peer_cache = {}


def _peer_buffers(key, nbytes, device, rank, nranks):
    pb = _peer_cache.get(key)
    if pb is None or pb.nbytes < nbytes:
        if pb is not None:
            import torch
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier()                      # nobody still stores into the buffers about to be freed
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
                                          return_device=False, timings=None):
    """ky-slab sharded version of ``steady_state_transport_solver`` (same arguments).

    Returns ``(grid, conc, flx)`` like the single-GPU solver when ``gather`` is true (on every rank);
    otherwise the rank's column slab ``[..., ny, nx/G]`` with ``grid`` restricted to it.

    Nothing on the data path synchronises with the host: the NCCL variant is ONE grouped send/recv exchange
    for all ``2*nlv`` fields, stream-ordered behind stage 1; the fused variant has no collective -- stage 1
    stores into the peers' buffers and a flag per peer (``bldfm_peer_signal`` / ``bldfm_peer_wait``) orders
    stage 2 behind everybody's stores on the device.  ``timings`` (a dict) receives CUDA-event times
    ``stage1_ms / exchange_ms / stage2_ms / gather_ms`` of this rank.
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
    L = _lib.lib()
    plan = get_fft_manager().plan(geom, dev_index)
    stream = torch.cuda.ExternalStream(L.bldfm_plan_stream(plan), device=device)
    prob, keep = _lib.make_problem(z, profiles, meas_pt, srf_bg_conc)
    lvp = lv64.ctypes.data_as(C.POINTER(C.c_int64))

    out = torch.empty((2, nlv, ny, nxl), dtype=torch.float64, device=device)
    field_elems = nlv * G * rows * nxl            # complex elements of one of p / q on the receiver
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timings is not None else None

    def mark(i):
        if ev is not None:
            ev[i].record(stream)

    with torch.cuda.stream(stream):
        mark(0)
        if fused and G > 1:
            pb = _peer_buffers((dev_index, G), field_elems * 16, dev_index, rank, G)
            pb.seq += 1
            which = pb.seq & 1
            # pointer tables (device): where THIS rank's row block starts inside every peer's receive buffer,
            # and this rank's slot in every peer's flag array
            off = rank * rows * nxl * 16
            tab = torch.tensor([[pb.peer_recv(r, which, k) + off for r in range(G)] for k in range(2)] +
                               [[pb.peer_flag_slot(r, which) for r in range(G)]], dtype=torch.int64, device=device)
            _lib.check(L.bldfm_sharded_stage1(plan, C.byref(prob), lvp, nlv, srcp, flags, rank, G,
                                              pb.recv(which, 0), pb.recv(which, 1),
                                              tab[0].data_ptr(), tab[1].data_ptr()))
            mark(1)
            _lib.check(L.bldfm_peer_signal(plan, tab[2].data_ptr(), G, pb.seq))
            _lib.check(L.bldfm_peer_wait(plan, pb.flags(which), G, pb.seq, 0.0))
            recv_p, recv_q = pb.recv(which, 0), pb.recv(which, 1)
            keep = (keep, tab)
        else:
            send = torch.empty((2, nlv, G, rows, nxl), dtype=torch.complex128, device=device)
            _lib.check(L.bldfm_sharded_stage1(plan, C.byref(prob), lvp, nlv, srcp, flags, rank, G,
                                              send[0].data_ptr(), send[1].data_ptr(), None, None))
            mark(1)
            if G > 1:
                # ONE grouped exchange (a single NCCL group of sends/receives) for all 2*nlv fields
                recv = torch.empty_like(send)
                ops = []
                for k in range(2):
                    for l in range(nlv):
                        for r in range(G):
                            ops.append(dist.P2POp(dist.isend, send[k, l, r], r))
                            ops.append(dist.P2POp(dist.irecv, recv[k, l, r], r))
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                       # stream-ordered on the plan's stream, not a host wait
            else:
                recv = send
            recv_p, recv_q = recv[0].data_ptr(), recv[1].data_ptr()
        mark(2)
        _lib.check(L.bldfm_sharded_stage2(plan, nlv, flags, rank, G, recv_p, recv_q,
                                          out[0].data_ptr(), out[1].data_ptr()))
        mark(3)
        if gather and G > 1:
            full = torch.empty((G, 2, nlv, ny, nxl), dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(full, out)
            out = full.permute(1, 2, 3, 0, 4).reshape(2, nlv, ny, nx)
        mark(4)
    stream.synchronize()
    del keep
    if fused and G > 1:
        st = C.c_int32(0)
        _lib.check(L.bldfm_peer_status(plan, C.byref(st)))
        if st.value:
            raise RuntimeError(f"sharded solve: rank {st.value - 1} never delivered its rows (peer wait timed out)")
    if timings is not None:
        for i, name in enumerate(("stage1_ms", "exchange_ms", "stage2_ms", "gather_ms")):
            timings[name] = ev[i].elapsed_time(ev[i + 1])
        timings["exchange_bytes_sent"] = 0 if G == 1 else 2 * nlv * (G - 1) * rows * nxl * 16

    if return_device:
        return out[0], out[1]
    conc = np.squeeze(out[0].cpu().numpy())
    flx = np.squeeze(out[1].cpu().numpy())
    grid = make_grid(z, lv, domain, nx, ny)
    if not (gather or G == 1):
        sl = slice(rank * nxl, (rank + 1) * nxl)
        grid = tuple(a[..., sl] for a in grid)
    return grid, conc, flx


def solve_sharded_emulated(G, srf_flx, z, profiles, domain, levels, modes=(512, 512), meas_pt=(0.0, 0.0),
                           srf_bg_conc=0.0, footprint=True, halo=None):
    """The G rank programs of ``steady_state_transport_solver_sharded`` run one after the other on ONE GPU,
    the exchange replaced by the equivalent device-side permutation: same kernels, same launch geometry
    per rank (``bldfm_sharded_stage1`` / ``stage2`` with ``rank = 0..G-1``), so the result is bit-identical
    to a real G-GPU run.  Lets a single-GPU box verify the ky-slab decomposition for any G; returns
    ``(conc, flx)`` float64 ``[nlv, ny, nx]`` (squeezed) on the host.
    """
    import torch

    dev_index = config.DEVICE
    device = torch.device("cuda", dev_index)
    q0 = np.asarray(srf_flx)
    ny, nx = q0.shape
    geom = _geometry(q0.shape, domain, modes, halo)
    flags = _flags(footprint, False, "double") | _lib.ASYNC
    src = None if footprint else _lib.as_f64(q0)
    srcp = None if src is None else _lib.ptr(src)
    _, lv64 = _levels_array(levels)
    nlv = len(lv64)
    herm = not config.MARCH_FULL
    if nx % G or (not herm and geom.nly % G):
        raise ValueError("sharded solve needs nx (and nly with MARCH_FULL) divisible by the number of ranks")
    nxl = nx // G
    rows = -(-(geom.nly // 2 + 1) // G) if herm else geom.nly // G
    L = _lib.lib()
    plan = get_fft_manager().plan(geom, dev_index)
    stream = torch.cuda.ExternalStream(L.bldfm_plan_stream(plan), device=device)
    prob, keep = _lib.make_problem(z, profiles, meas_pt, srf_bg_conc)
    lvp = lv64.ctypes.data_as(C.POINTER(C.c_int64))
    with torch.cuda.stream(stream):
        # send[src rank][p|q][level][dst rank][row][col]; ranks beyond the last row block send nothing
        send = torch.zeros((G, 2, nlv, G, rows, nxl), dtype=torch.complex128, device=device)
        for r in range(G):
            _lib.check(L.bldfm_sharded_stage1(plan, C.byref(prob), lvp, nlv, srcp, flags, r, G,
                                              send[r, 0].data_ptr(), send[r, 1].data_ptr(), None, None))
        # all_to_all_single: rank d receives block d of every source rank, ordered by source rank
        recv = send.permute(3, 1, 2, 0, 4, 5).contiguous()            # [dst][p|q][level][src][row][col]
        out = torch.empty((2, nlv, ny, nx), dtype=torch.float64, device=device)
        slab = torch.empty((G, 2, nlv, ny, nxl), dtype=torch.float64, device=device)
        for d in range(G):
            _lib.check(L.bldfm_sharded_stage2(plan, nlv, flags, d, G, recv[d, 0].data_ptr(), recv[d, 1].data_ptr(),
                                              slab[d, 0].data_ptr(), slab[d, 1].data_ptr()))
        out = slab.permute(1, 2, 3, 0, 4).reshape(2, nlv, ny, nx)
    stream.synchronize()
    del keep
    return np.squeeze(out[0].cpu().numpy()), np.squeeze(out[1].cpu().numpy())
