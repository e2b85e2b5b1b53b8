"""Distributional check of the device permutation null against the oracle's null (north star: KS test on the null of
the minimum p).  Device: Philox-sorted pairings; oracle: Fisher-Yates (xoshiro256**) permutations through the integer-mode
restatement on all host threads.  Writes a JSON report."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy import stats
import dual_threshold_optimization_b200 as dto
from tests import helpers as H
from tests.helpers import O

out = {}
for N, P_dev, P_cpu in ((2000, 200000, 20000), (6000, 200000, 6000), (20000, 200000, 1500)):
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, 0.25)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    eng = dto.Engine(0)
    eng.load_lists(l1, l2, N)
    t0 = time.perf_counter()
    dev = eng.run_permuted_philox(123, 0, P_dev, want_records=False, want_minp=True)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu = O.run_single_node(o1, o2, N, [1] * P_cpu, os.cpu_count(), seed=99, mode=1)["pvalue"]
    t_cpu = time.perf_counter() - t0
    ks = stats.ks_2samp(np.log(dev), np.log(cpu))
    qs = [0.001, 0.01, 0.05, 0.25, 0.5, 0.75, 0.95]
    out[f"N={N}"] = {
        "device_permutations": P_dev, "oracle_permutations": P_cpu, "device_seconds": t_dev, "oracle_seconds": t_cpu,
        "ks_statistic": float(ks.statistic), "ks_pvalue": float(ks.pvalue),
        "quantiles": {str(q): {"device": float(np.quantile(dev, q)), "oracle": float(np.quantile(cpu, q))} for q in qs},
    }
    eng.close()
print(json.dumps(out, indent=1))
