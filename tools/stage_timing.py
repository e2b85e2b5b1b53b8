"""Where the time of setting a problem up goes (load_lists / set_problem), per config shape, on cuda:0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench
import dual_threshold_optimization_b200 as dto

out = {}
eng = dto.Engine(0)
for cfg in ("c2", "c3", "c5"):
    ids1, r1, ids2, r2, bg = bench.workload_lists(cfg)
    t0 = time.perf_counter()
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    t_lists = time.perf_counter() - t0
    pop = dto.compute_population_size(l1, l2, dto.FeatureList(bg) if bg is not None else None)
    res = {"ranked_list_from_s": t_lists}
    for label, cache in (("first_load_s", 1), ("second_load_cache_hit_s", 1), ("reload_no_cache_s", 0)):
        eng.set_option("table_cache", cache)
        t0 = time.perf_counter()
        eng.load_lists(l1, l2, pop)
        res[label] = time.perf_counter() - t0
    eng.set_option("table_cache", 1)
    t0 = time.perf_counter()
    u = eng.run_unpermuted()
    res["run_unpermuted_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    u = eng.run_unpermuted()
    res["run_unpermuted_again_s"] = time.perf_counter() - t0
    res["lptab_entries"] = eng.stats()["lptab_entries"]
    out[cfg] = res
print(json.dumps(out, indent=1))
