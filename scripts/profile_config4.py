import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bldfm_b200
from bldfm_b200.schema import Config, Domain, Met, Parallel, SolverOptions, Tower
T = 48
rng = np.random.default_rng(0)
hours = np.arange(T) * 0.5
ustar = (0.45 + 0.25 * np.sin(2 * np.pi * (hours - 6) / 24) + 0.02 * rng.normal(size=T)).clip(0.1, 0.8)
mol = np.where(np.sin(2 * np.pi * (hours - 6) / 24) > 0, -1.0, 1.0) * (50.0 + 400.0 * rng.random(T))
ws = (4.5 + 3.0 * np.sin(2 * np.pi * (hours - 8) / 24) + 0.3 * rng.normal(size=T)).clip(1.0, 8.0)
wd = (270.0 + 30.0 * rng.normal(size=T)) % 360.0
towers = [Tower(f"T{i}", 10.0, 1000.0 + 500.0 * (i % 4), 1500.0 + 500.0 * (i // 4)) for i in range(8)]
cfg = Config(Domain(nx=512, ny=512, xmax=4000.0, ymax=4000.0, nz=64, modes=(512, 512)), towers,
             Met(ustar=list(ustar), mol=list(mol), wind_speed=list(ws), wind_dir=list(wd)),
             SolverOptions(footprint=True, precision="double"), Parallel())
for _ in range(2):
    res = bldfm_b200.run_bldfm_multitower(cfg)
t0 = time.perf_counter(); res = bldfm_b200.run_bldfm_multitower(cfg); print("wall", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); res = bldfm_b200.run_bldfm_multitower(cfg); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
