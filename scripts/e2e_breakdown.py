"""Where does the end-to-end (host in / host out) time of one config-2 solve go?"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bldfm_b200
from bench import config2

kw = config2()
for _ in range(5):
    bldfm_b200.steady_state_transport_solver(**kw)
n = 200
t0 = time.perf_counter()
for _ in range(n):
    r = bldfm_b200.steady_state_transport_solver(**kw)
print("ms per call", (time.perf_counter() - t0) / n * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    r = bldfm_b200.steady_state_transport_solver(**kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
