"""Runs every BASELINE.json config shape once on cuda:0 and prints throughput (reduced permutation counts where noted)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import dual_threshold_optimization_b200 as dto
from tests import helpers as H

out = {}
eng = dto.Engine(0)


def timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t0


# a fresh box starts with idle clocks and cold module loads: spin the GPU for ~2 s first so the figures are steady-state
_w1, _w2, _w3, _w4 = H.synthetic_pair(6000, 6000, 0.25)
eng.load_lists(dto.RankedFeatureList.from_(_w1, _w2), dto.RankedFeatureList.from_(_w3, _w4), 6000)
_t0 = time.perf_counter()
while time.perf_counter() - _t0 < 2.0:
    eng.run_permuted_philox(1, 0, 20000, want_records=False, want_minp=True)

# C1: test_data, 1 000 permutations through the host API (run_single_node + epilogue)
ids1, r1, ids2, r2, bg = H.load_test_data()
l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
N = dto.compute_population_size(l1, l2, dto.FeatureList(bg))
tasks = [dto.Task(0, False)] + [dto.Task(i, True) for i in range(1, 1001)]
dto.run_single_node(tasks, l1, l2, N, 1)
res, dt = timed(lambda: dto.run_single_node(tasks, l1, l2, N, 1))
out["C1_test_data_1000perms"] = {"seconds": dt, "tasks_per_s": 1001 / dt, "json": dto.empirical_pvalue(res)}

# C2: N = 6 000, 10 000 permutations
ids1, r1, ids2, r2 = H.synthetic_pair(6000, 6000, 0.25)
l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
_, dt_load = timed(lambda: eng.load_lists(l1, l2, 6000))
eng.run_permuted_philox(6000, 0, 10000)
_, dt = timed(lambda: eng.run_permuted_philox(6000, 10000, 10000))
st = eng.stats()
out["C2_N6000_10000perms"] = {"seconds": dt, "perms_per_s": 10000 / dt, "load_lists_s": dt_load, "grid": list(eng.shape[:2]), "lptab_entries": st["lptab_entries"],
                              "scan_ms": st["last_scan_kernel_ms"], "sort_ms": st["last_sigma_kernel_ms"]}

# C4: pairs of N = 6 000 with 1 000 permutations each (40 pairs here; the config has 2 000)
pairs = []
for q in range(40):
    a1, b1, a2, b2 = H.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    x, y = dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2)
    pairs.append((x, y, 6000))
dto.run_pairs(pairs[:2], 1000)
res, dt = timed(lambda: dto.run_pairs(pairs, 1000))
out["C4_40pairs_N6000_1000perms"] = {"seconds": dt, "pairs_per_s": len(pairs) / dt, "perms_per_s": len(pairs) * 1001 / dt,
                                     "extrapolated_2000_pairs_s": 2000 * dt / len(pairs), "example": res[1]}
# 240 DISTINCT list pairs (every pair pays its own id join and set-up; the per-context start-up is amortised)
many = list(pairs)
for q in range(40, 240):
    a1, b1, a2, b2 = H.synthetic_pair(6000, 1 + q, 0.3 if q % 2 else None)
    many.append((dto.RankedFeatureList.from_(a1, b1), dto.RankedFeatureList.from_(a2, b2), 6000))
res, dt = timed(lambda: dto.run_pairs(many, 1000))
out["C4_240pairs_N6000_1000perms"] = {"seconds": dt, "pairs_per_s": len(many) / dt, "perms_per_s": len(many) * 1001 / dt,
                                      "extrapolated_2000_pairs_s": 2000 * dt / len(many)}

# C5: 60 000-id universe filtered to a 40 000-feature background (ranks keep gaps), 20 000 permutations here
rng = np.random.default_rng(60000)
U, B = 60000, 40000
ids1, r1, ids2, r2 = H.synthetic_pair(U, 60000, 0.25)
keep = np.zeros(U, dtype=bool)
keep[rng.choice(U, size=B, replace=False)] = True
idx = {g: i for i, g in enumerate(ids1)}
m1 = [i for i, g in enumerate(ids1) if keep[idx[g]]]
m2 = [i for i, g in enumerate(ids2) if keep[idx[g]]]
l1 = dto.RankedFeatureList.from_([ids1[i] for i in m1], r1[m1])
l2 = dto.RankedFeatureList.from_([ids2[i] for i in m2], r2[m2])
bgl = dto.FeatureList([ids1[i] for i in m1])
N5 = dto.compute_population_size(l1, l2, bgl)
_, dt_load = timed(lambda: eng.load_lists(l1, l2, N5))
rec0 = eng.run_unpermuted()
eng.run_permuted_philox(60000, 0, 4000)
_, dt = timed(lambda: eng.run_permuted_philox(60000, 4000, 20000))
st = eng.stats()
out["C5_universe60000_bg40000_20000perms"] = {"seconds": dt, "perms_per_s": 20000 / dt, "load_lists_s": dt_load, "population": N5, "grid": list(eng.shape[:2]),
                                             "lptab_entries": st["lptab_entries"], "scan_ms": st["last_scan_kernel_ms"], "sort_ms": st["last_sigma_kernel_ms"],
                                             "unpermuted": [int(rec0["rank1"]), int(rec0["rank2"]), float(rec0["pvalue"])]}
print(json.dumps(out, indent=1))
