"""run_pairs with one host context (fewer than 8 pairs per call) vs four contexts per device, same process and box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dual_threshold_optimization_b200 as dto
from tests import helpers as H
pairs = []
for q in range(56):
    a1, b1, a2, b2 = H.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    pairs.append((dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2), 6000))
dto.run_pairs(pairs[:7], 1000)
for rep in range(3):
    t0 = time.perf_counter()
    for k in range(0, 56, 7):
        dto.run_pairs(pairs[k:k + 7], 1000)  # 7 pairs per call -> one context
    t1 = time.perf_counter()
    dto.run_pairs(pairs, 1000)               # 56 pairs -> four contexts
    t2 = time.perf_counter()
    print(f"rep {rep}: one context {56 / (t1 - t0):.0f} pairs/s (incl. 8 context start-ups), four contexts {56 / (t2 - t1):.0f} pairs/s; cores {os.cpu_count()}, load {os.getloadavg()}")
