"""Where does an end-to-end step (bench.py e2e leg) spend its time?  N = 20 000, 100 000 permutations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dual_threshold_optimization_b200 as dto
from dual_threshold_optimization_b200.stat_operations import empirical_pvalue_struct
from tests import helpers as H

ids1, r1, ids2, r2 = H.synthetic_pair(20000, 20000, 0.25)
l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
eng = dto.Engine(0)
P = 100000
def step(k):
    t = [time.perf_counter()]
    eng.load_lists(l1, l2, 20000); t.append(time.perf_counter())
    rec0 = eng.run_unpermuted(); t.append(time.perf_counter())
    recs = eng.run_permuted_philox(5, k * P, P); t.append(time.perf_counter())
    allrec = np.concatenate([np.asarray([rec0], dtype=recs.dtype), recs]); t.append(time.perf_counter())
    emp = empirical_pvalue_struct(allrec).empirical_pvalue; t.append(time.perf_counter())
    return np.diff(t) * 1e3, eng.stats()["last_run_ms"]
step(0)
for k in range(1, 4):
    d, dev = step(k)
    print("load %.2f  unpermuted %.2f  permuted(host call) %.2f [device %.2f]  concat %.2f  epilogue %.2f  total %.2f ms" % (*d[:3], dev, d[3], d[4], d.sum()))
