import os, sys, time
sys.path.insert(0, os.getcwd())
import dual_threshold_optimization_b200 as dto
from dual_threshold_optimization_b200 import synthetic as S
pairs = []
for q in range(240):
    a1, b1, a2, b2 = S.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    pairs.append((dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2), 6000))
for n in (2, 40, 40, 40, 240, 240, 240, 40):
    t0 = time.perf_counter(); dto.run_pairs(pairs[:n], 1000); dt = time.perf_counter() - t0
    print(f"run_pairs({n:3d} pairs): {dt*1e3:8.2f} ms  ({n/dt:8.1f} pairs/s)", flush=True)
