import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import dual_threshold_optimization_b200 as dto
from dual_threshold_optimization_b200 import synthetic as S
import torch
ids1, r1, ids2, r2 = S.synthetic_pair(20000, 20000, 0.25)
l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
eng = dto.Engine(0)
eng.load_lists(l1, l2, 20000)
for P in (12500, 25000, 50000):
    d = torch.zeros(P, dtype=torch.float64, device="cuda:0")
    for w in (12, 11, 10, 9, 8):
        eng.set_option("warps_per_cta", w)
        for _ in range(3):
            eng.run_permuted_philox_device(1, 1, P, d.data_ptr())
        ms = []
        for s in range(10):
            eng.run_permuted_philox_device(1, 1 + s * P, P, d.data_ptr())
            ms.append(eng.stats()["last_scan_kernel_ms"])
        print(f"P={P} warps/CTA={w}: scan {np.mean(ms):.3f} ms (min {np.min(ms):.3f})  waves {P/(296*w):.2f}", flush=True)
