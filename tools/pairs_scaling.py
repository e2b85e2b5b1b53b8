"""How much do extra host workers (contexts) per GPU help the batched-pair driver? (config 4 shape)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dual_threshold_optimization_b200 as dto
from tests import helpers as H
pairs = []
for q in range(64):
    a1, b1, a2, b2 = H.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    pairs.append((dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2), 6000))
dto.run_pairs(pairs[:4], 1000)
for workers in (1, 2, 3, 4, 6, 8):
    t0 = time.perf_counter()
    r = dto.run_pairs(pairs, 1000, devices=[0] * workers, seed=1)
    dt = time.perf_counter() - t0
    print(f"workers {workers}: {len(pairs) / dt:.1f} pairs/s ({1e3 * dt / len(pairs):.2f} ms/pair)")
