"""Diagnostic: per-permutation cost distribution of the scan kernel (run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dual_threshold_optimization_b200 as dto
from tests import helpers as H

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
ids1, r1, ids2, r2 = H.synthetic_pair(N, N, 0.25)
l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
eng = dto.Engine(0)
eng.set_option("task_stats", 1)
eng.set_option("batch", P)
eng.load_lists(l1, l2, N)
rec = eng.run_permuted_philox(1, 0, P)
ts = eng.last_batch_task_stats(P).astype(np.float64)
cyc = ts[:, 3] * 16
q = [0, 10, 50, 90, 99, 99.9, 100]
out = {"N": N, "P": P, "scan_ms": eng.stats()["last_scan_kernel_ms"]}
for name, col in (("screened", ts[:, 0]), ("refined", ts[:, 1]), ("exact", ts[:, 2]), ("cycles", cyc)):
    out[name] = {"mean": float(col.mean()), **{f"p{x}": float(np.percentile(col, x)) for x in q}}
out["sum_cycles_over_warps"] = float(cyc.sum())
worst = np.argsort(-cyc)[:5]
out["worst"] = [{"task": int(t), "cycles": float(cyc[t]), "screened": int(ts[t, 0]), "refined": int(ts[t, 1]), "minp": float(rec[t]["pvalue"]),
                 "rank1": int(rec[t]["rank1"]), "rank2": int(rec[t]["rank2"]),
                 "final_level": int(ts[t, 4])} for t in worst]
out["lptab_entries"] = eng.stats()["lptab_entries"]
print(json.dumps(out, indent=1))
