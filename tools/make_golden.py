"""Regenerates tests/golden/oracle_vectors.json: outputs of the CPU oracle (which reproduces every reference golden
bit-exactly, tests/test_oracle_goldens.py) on small seeded inputs, used as committed fixtures by the GPU tests.
The reference itself is Rust and cannot run in this image, so these are oracle-generated (DESIGN.md "Oracle")."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers as H  # noqa: E402
from tests.helpers import O  # noqa: E402


def case(name, ids1, r1, ids2, r2, bg, n_perm, seed):
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    N = O.compute_population_size(o1, o2, bg)
    slot = O.slot_map(o1, o2)
    g = O.grid_int(o1, o2, N, slot, want_logp=True)
    out = {"name": name, "population": N, "n1": len(ids1), "n2": len(ids2), "T1": int(o1.thresholds.size), "T2": int(o2.thresholds.size),
           "thresholds1_tail": [int(x) for x in o1.thresholds[-3:]],
           "unpermuted_best": {k: (float(v) if k == "pvalue" else int(v)) for k, v in g.best.items()},
           "overlap_checksum": int(g.overlap.astype(np.uint64).sum()), "overlap_diag": [int(g.overlap[i, min(i, g.overlap.shape[1] - 1)]) for i in range(0, g.overlap.shape[0], max(1, g.overlap.shape[0] // 8))],
           "p_hex_corner": float(g.p[-1, -1]).hex(), "perm_seed": seed, "permuted_best": []}
    p1, p2 = H.perms(len(ids1), n_perm, seed), H.perms(len(ids2), n_perm, seed + 1)
    for t in range(n_perm):
        b = O.grid_int(o1, o2, N, slot, p1[t], p2[t], want_overlap=False, want_p=False).best
        out["permuted_best"].append({k: (float(v) if k == "pvalue" else int(v)) for k, v in b.items()})
    return out


def main():
    cases = []
    ids1, r1, ids2, r2, bg = H.load_test_data()
    cases.append(case("test_data", ids1, r1, ids2, r2, bg, 16, 100))
    ids1, r1, ids2, r2 = H.ten_gene_case()
    cases.append(case("ten_gene", ids1, r1, ids2, r2, None, 8, 100))
    ids1, r1, ids2, r2 = H.synthetic_pair(300, 3, 0.25, tied_frac=0.2)
    cases.append(case("synthetic_300_ties", ids1, r1, ids2, r2, None, 8, 100))
    ids1, r1, ids2, r2, uni = H.background_case(400, 330, 290, 5)
    cases.append(case("background_330_290", ids1, r1, ids2, r2, uni, 8, 100))
    ids1, r1, ids2, r2 = H.synthetic_pair(2000, 13, None)
    cases.append(case("null_2000", ids1, r1, ids2, r2, None, 6, 100))
    with open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json"), "w") as f:
        json.dump({"generator": "tools/make_golden.py", "cases": cases}, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
