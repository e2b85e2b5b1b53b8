"""One-off deep parity report (more permutations than the test-suite affords): the performance path (device Philox
pairings), the same permutations fed back as host indices, and the integer oracle on those indices must agree on EVERY
record.  Writes a JSON report (profiles/r02_deep_parity.json).  Usage: python tools/deep_parity.py [scale]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import dual_threshold_optimization_b200 as dto
from tests import helpers as H
from tests.helpers import O

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
eng = dto.Engine(0)
out = {}


def run(name, ids1, r1, ids2, r2, bg, P, seed):
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    pop = dto.compute_population_size(l1, l2, dto.FeatureList(bg) if bg is not None else None)
    eng.load_lists(l1, l2, pop)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    lf = O.ln_factorial_table(pop)
    n = len(ids1)
    eng.reset_stats()
    t0 = time.perf_counter()
    fast = eng.run_permuted_philox(seed, 1, P)
    t_dev = time.perf_counter() - t0
    st = eng.stats()
    mism = {"fast_vs_host_indices": 0, "integer_fields_vs_oracle": 0, "p_outside_1e-12": 0, "host_settled_p_not_bit_exact": 0}
    worst = 0.0
    t_cpu = 0.0
    B = 2000
    for b0 in range(0, P, B):
        m = min(B, P - b0)
        p1 = np.tile(np.arange(n, dtype=np.uint32), (m, 1))
        p2 = np.empty((m, n), dtype=np.uint32)
        for t in range(m):
            p2[t, eng.philox_pairing(seed, 1 + b0 + t)] = slot
        host = eng.run_permuted_indices(p1, p2)
        f = fast[b0:b0 + m]
        mism["fast_vs_host_indices"] += int((f.tobytes() != host.tobytes()) and sum(f[i].tobytes() != host[i].tobytes() for i in range(m)))
        t0 = time.perf_counter()
        ref = O.best_batch(o1, o2, pop, p1, p2, slot, lf)
        t_cpu += time.perf_counter() - t0
        bad = np.zeros(m, dtype=bool)
        for fld in ("rank1", "rank2", "set1_len", "set2_len", "intersection_size", "population_size"):
            bad |= f[fld].astype(np.int64) != ref[fld].astype(np.int64)
        mism["integer_fields_vs_oracle"] += int(bad.sum())
        rel = np.abs(f["pvalue"] - ref["pvalue"]) / np.maximum(ref["pvalue"], 1e-300)
        rel[(f["pvalue"] == 0) & (ref["pvalue"] == 0)] = 0.0
        worst = max(worst, float(rel.max()))
        mism["p_outside_1e-12"] += int((rel > 1e-12).sum())
        on_host = (f["flags"] & dto._capi.FLAG_HOST_PVALUE) != 0
        mism["host_settled_p_not_bit_exact"] += int((f["pvalue"][on_host] != ref["pvalue"][on_host]).sum())
    out[name] = {"features": n, "population": int(pop), "threshold_pairs": int(eng.shape[0] * eng.shape[1]), "permutations": P,
                 "mismatches": mism, "worst_relative_p_difference": worst, "tie_sets_settled_on_host": int(st["tasks_tie_resolved"]),
                 "dense_path_tasks": int(st["tasks_full"]), "device_seconds": t_dev, "oracle_seconds_all_host_threads": t_cpu,
                 "oracle_threads": os.cpu_count()}
    print(name, json.dumps(out[name]), flush=True)


ids1, r1, ids2, r2 = H.synthetic_pair(6000, 6000, 0.25)
run("c2_N6000", ids1, r1, ids2, r2, None, int(20000 * scale), 61)
ids1, r1, ids2, r2 = H.synthetic_pair(20000, 20000, 0.25)
run("c3_N20000", ids1, r1, ids2, r2, None, int(8000 * scale), 62)
f1, fr1, f2, fr2, bg = H.background_subset_pair(60000, 40000, 60000, 0.3)
run("c5_universe60000_background40000", f1, fr1, f2, fr2, bg, int(1000 * scale), 63)
ids1, r1, ids2, r2 = H.synthetic_pair(6000, 7, None, tied_frac=0.05)
run("c2_N6000_null_5pct_ties", ids1, r1, ids2, r2, None, int(6000 * scale), 64)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_deep_parity.json"), "w"), indent=1)
