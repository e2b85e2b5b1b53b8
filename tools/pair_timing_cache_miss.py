import sys, os, time
sys.path.insert(0, "/root/repo")
import dual_threshold_optimization_b200 as dto
from tests import helpers as H
eng = dto.Engine(0)
eng.set_option("table_cache", 0)
pairs = []
for q in range(6):
    a1, b1, a2, b2 = H.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    pairs.append((dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2)))
def t(fn):
    t0 = time.perf_counter(); r = fn(); return r, (time.perf_counter() - t0) * 1e3
for rep in range(2):
  for q, (x, y) in enumerate(pairs):
    _, tl = t(lambda: eng.load_lists(x, y, 6000))
    _, tu = t(lambda: eng.run_unpermuted())
    _, tp = t(lambda: eng.run_permuted_philox(q, 0, 1000))
    print(f"rep {rep} pair {q}: load(miss) {tl:.2f} ms, unpermuted {tu:.2f} ms, 1000 perms {tp:.2f} ms")
