"""Back-transform kernels in the throughput regime: one batched launch of G march groups x T towers of
config-2 size (2*G*T fields of 1536^2 padded -> 512^2), device-resident.  Prints the per-stage times.
Used under ncu to capture k_fft_h with full grids (profiles/r1i_fft24_batched_ncu.txt, profiles/r2_fft24_batched_ncu.txt)."""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bldfm_b200
from bldfm_b200 import _lib
from bldfm_b200.pbl_model import vertical_profiles

ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=8)
ap.add_argument("--towers", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--nz", type=int, default=64)
args = ap.parse_args()
L = _lib.lib()
n = args.n
dom = (4000.0 * n / 512, 4000.0 * n / 512)
geom = _lib.geometry((n, n), dom, (n, n), None)
plan = bldfm_b200.get_fft_manager().plan(geom, 0)
probs, keeps = [], []
for g in range(args.groups):
    z, prof = vertical_profiles(args.nz, 10.0, (-3.0 - 0.02 * g, -4.0 + 0.01 * g), ustar=0.4 + 0.001 * g, mol=-50.0 - 0.5 * g)
    for t in range(args.towers):
        p_, k_ = _lib.make_problem(z, prof, (dom[0] * (0.3 + 0.05 * t), dom[1] * (0.6 - 0.03 * t)), 0.0)
        probs.append(p_)
        keeps.append(k_)
B = len(probs)
parr = (_lib.Problem * B)(*probs)
lv = np.array([args.nz], dtype=np.int64)
lvp = lv.ctypes.data_as(C.POINTER(C.c_int64))
out_c = torch.empty((B, n, n), dtype=torch.float64, device="cuda:0")
out_f = torch.empty_like(out_c)
flags = _lib.FOOTPRINT | _lib.DOUBLE | _lib.OUT_ON_DEVICE | _lib.ASYNC
L.bldfm_plan_set_profiling(plan, 1)
tm = _lib.Timings()
rows = []
for i in range(args.reps + 2):
    _lib.check(L.bldfm_solve_batched(plan, B, parr, lvp, 1, None, flags, out_c.data_ptr(), out_f.data_ptr()))
    _lib.check(L.bldfm_plan_last_timings(plan, C.byref(tm)))
    if i >= 2:
        rows.append((tm.march_ms, tm.inverse_ms, tm.total_ms))
r = np.median(np.array(rows), axis=0)
print(json.dumps({"n": n, "groups": args.groups, "towers": args.towers, "fields": 2 * B, "march_ms": r[0], "inverse_ms": r[1],
                  "total_ms": r[2], "inverse_us_per_field": r[1] * 1e3 / (2 * B),
                  "footprints_per_s": B / (r[2] * 1e-3)}))
