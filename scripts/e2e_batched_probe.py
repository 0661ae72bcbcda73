"""Where does the pipelined batched end-to-end path spend its time?  (host call time per chunk, total
time, raw D2H rate of the same buffers)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bldfm_b200
from bldfm_b200.pbl_model import vertical_profiles
from bench import config2
kw = config2()
for CH in (4, 8, 16):
    zs, pls, mps = [], [], []
    for b in range(CH):
        zb, pb = vertical_profiles(64, 10.0, (-3.0 - 0.02 * b, -4.0 + 0.01 * b), ustar=0.4 + 0.001 * b, mol=-50.0 - 0.5 * b)
        zs.append(zb); pls.append(pb); mps.append(kw["meas_pt"])
    bkw = dict(domain=kw["domain"], levels=kw["levels"], modes=kw["modes"], meas_pts=mps, footprint=True, precision="double")
    held = [None] * 3
    for i in range(4):
        held[i % 3] = bldfm_b200.solve_batched(kw["srf_flx"], zs, pls, wait=False, **bkw)
    bldfm_b200.solver.synchronize()
    n = 40
    calls = []
    t0 = time.perf_counter()
    for i in range(n):
        tc = time.perf_counter()
        held[i % 3] = bldfm_b200.solve_batched(kw["srf_flx"], zs, pls, wait=False, **bkw)
        calls.append(time.perf_counter() - tc)
    bldfm_b200.solver.synchronize()
    tot = time.perf_counter() - t0
    # blocking variant
    t1 = time.perf_counter()
    for i in range(10):
        held[i % 3] = bldfm_b200.solve_batched(kw["srf_flx"], zs, pls, wait=True, **bkw)
    blk = (time.perf_counter() - t1) / 10
    print(f"chunk {CH}: async {tot / n * 1e3:.3f} ms/chunk ({n * CH / tot:.0f} solves/s), host call median {np.median(calls) * 1e3:.3f} ms "
          f"max {np.max(calls) * 1e3:.3f}; blocking {blk * 1e3:.3f} ms/chunk; bytes/chunk {CH * 4.19:.1f} MB "
          f"-> {n * CH * 4.194304e6 / tot * 1e-9:.1f} GB/s")
x = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
h = torch.empty(16 << 20, dtype=torch.uint8, pin_memory=True)
for _ in range(3): h.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): h.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print("raw D2H 16 MiB pinned:", 20 * 16.777 / (time.perf_counter() - t0) * 1e-3, "GB/s")
