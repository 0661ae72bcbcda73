import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bldfm_b200
from bench import config2
kw = config2()
def t(label, n=100):
    for _ in range(5): bldfm_b200.steady_state_transport_solver(**kw)
    t0=time.perf_counter()
    for _ in range(n): r=bldfm_b200.steady_state_transport_solver(**kw)
    print(label, (time.perf_counter()-t0)/n*1e3, "ms")
t("plain")
import torch
t("after import torch")
torch.cuda.set_device(0); x=torch.zeros(10,device="cuda"); torch.cuda.synchronize()
t("after torch cuda init")
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda"); flush.zero_(); torch.cuda.synchronize()
t("after flush alloc")
import ctypes as C
from bldfm_b200 import _lib
L=_lib.lib()
pk=C.c_double(0); L.bldfm_fp64_peak(0,0,20000,C.byref(pk)); print(pk.value)
t("after fp64 peak")
geom=_lib.geometry((512,512),(4000.,4000.),(512,512),None)
plan=bldfm_b200.get_fft_manager().plan(geom,0)
L.bldfm_plan_set_profiling(plan,1)
t("profiling on")
L.bldfm_plan_set_profiling(plan,0)
t("profiling off")
