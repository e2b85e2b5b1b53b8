"""Pins the CPU oracle against every golden the reference's own tests hold for the hot path,
then checks its internal consistency (faithful string/HashSet form == integer histogram form, cached == uncached,
independent accuracy vs scipy).  No GPU needed."""
import json
import os

import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import O

G = json.load(open(os.path.join(H.GOLDEN, "reference_goldens.json")))


def test_hypergeometric_bit_exact_goldens():
    for v in G["hypergeometric_pvalue_exact"]:
        assert O.hypergeometric_pvalue(v["N"], v["K"], v["n"], v["k"]) == v["p"], v["src"]  # assert_eq! in the reference
        lf = O.ln_factorial_table(v["N"])
        assert O.hypergeometric_pvalue_cached(lf, v["N"], v["K"], v["n"], v["k"]) == v["p"]
    for v in G["hypergeometric_pvalue_1e-10"]:
        assert abs(O.hypergeometric_pvalue(v["N"], v["K"], v["n"], v["k"]) - v["p"]) < 1e-10, v["src"]


def test_direct_tail_not_one_minus_cdf():
    """the goldens discriminate the direct upper-tail sum from 1 - cdf."""
    from scipy import stats

    one_minus_cdf = 1.0 - stats.hypergeom.cdf(9, 1000, 50, 60)
    assert one_minus_cdf != 0.00044068070222441115
    assert abs(one_minus_cdf / 0.00044068070222441115 - 1) < 1e-6


def test_optimize_ten_gene_golden():
    ids1, r1, ids2, r2 = H.ten_gene_case()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    b = O.optimize_faithful(o1, o2, 10)
    g = G["optimize_ten_gene"]
    assert (int(b["rank1"]), int(b["rank2"])) == (g["rank1"], g["rank2"])
    assert float(b["pvalue"]) == g["pvalue"]  # assert_eq! in the reference
    assert O.grid_int(o1, o2, 10).best["pvalue"] == g["pvalue"]


def test_readme_transcript_golden():
    ids1, r1, ids2, r2, bg = H.load_test_data()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    N = O.compute_population_size(o1, o2, bg)
    grid = O.process_threshold_pairs_faithful(o1, o2, N)
    assert grid.size == 900
    best = O.argmin_tiebreak(grid)
    out = O.final_json(np.array([best]))
    g = G["readme_test_data"]
    for k in ("rank1", "rank2", "set1_len", "set2_len", "unpermuted_intersection_size", "population_size"):
        assert out[k] == g[k]
    assert out["unpermuted_pvalue"] == g["unpermuted_pvalue"] and out["fdr"] == g["fdr"]
    assert out["empirical_pvalue"] == 1.0  # no permuted results -> 1.0 (empirical_pvalue.rs:140-152)


def test_epilogue_goldens():
    f = G["fdr"]
    assert abs(O.fdr(*f["args"]) - f["value"]) < f["tol"]
    e = G["empirical_pvalue"]
    u = e["unpermuted"]
    assert O.empirical_pvalue(e["permuted_pvalues"], u["pvalue"]) == e["json"]["empirical_pvalue"]
    assert O.fdr(u["set1_len"], u["set2_len"], u["intersection_size"], u["population_size"], 0.8) == e["json"]["fdr"]
    assert O.empirical_pvalue([], 0.5) == 1.0
    assert np.isnan(O.fdr(1, 1, 1, 1, 0.0))  # panics in the reference


def test_threshold_series_and_last_threshold_bug():
    for max_rank, (count, last) in ((int(k), v) for k, v in G["threshold_counts"].items() if not k.startswith("_")):
        l = O.OracleRankedList.make([f"g{i}" for i in range(max_rank)], np.arange(1, max_rank + 1))
        assert l.thresholds.size == count
        assert l.thresholds[0] == 1
        if last is not None:
            assert l.thresholds[-1] == last
        if max_rank > 200:
            assert l.thresholds[-1] < max_rank  # the no-op "set final threshold to max rank" (ranked.rs:370-372)
    # f64 arithmetic of the recurrence
    t = O.OracleRankedList.make([str(i) for i in range(400)], np.arange(1, 401)).thresholds
    assert list(t[:100]) == list(range(1, 101)) and t[100] == 102 and t[101] == 104  # floor(100*1.01+1) = 102
    assert O.OracleRankedList.make([], []).thresholds.size == 0


def test_stable_sort_and_rank_zero():
    l = O.OracleRankedList.make(["a", "b", "c", "d"], [2, 0, 2, 1])
    assert l.ids == ["b", "d", "a", "c"] and list(l.ranks) == [0, 1, 2, 2]
    assert list(l.thresholds) == [1, 2]
    with pytest.raises(ValueError):
        O.OracleRankedList.make(["a"], [1, 2])


@pytest.mark.parametrize("case", ["test_data", "ten_gene", "ties", "background", "maxrank"])
def test_faithful_equals_integer_form(case):
    if case == "test_data":
        ids1, r1, ids2, r2, bg = H.load_test_data()
    elif case == "ten_gene":
        ids1, r1, ids2, r2 = H.ten_gene_case()
        bg = None
    elif case == "ties":
        ids1, r1, ids2, r2 = H.synthetic_pair(120, 3, 0.3, tied_frac=0.3)
        bg = None
    elif case == "background":
        ids1, r1, ids2, r2, bg = H.background_case(160, 130, 110, 8)
    else:
        ids1, r1, ids2, r2 = H.synthetic_pair(260, 5, 0.2)
        bg = None
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    N = O.compute_population_size(o1, o2, bg)
    slot = O.slot_map(o1, o2)
    for perm in (None, 1, 2):
        p1 = None if perm is None else H.perms(len(ids1), 1, perm)[0]
        p2 = None if perm is None else H.perms(len(ids2), 1, 50 + perm)[0]
        f = O.process_threshold_pairs_faithful(o1, o2, N, p1, p2)
        g = O.grid_int(o1, o2, N, slot, p1, p2)
        assert np.array_equal(f["intersection_size"].reshape(g.overlap.shape), g.overlap)
        assert np.array_equal(f["pvalue"].reshape(g.p.shape), g.p)  # bit-exact incl. the early exit
        assert np.array_equal(f["set1_len"].reshape(g.overlap.shape)[:, 0], np.searchsorted(o1.ranks, o1.thresholds, side="right"))
        b = O.argmin_tiebreak(f)
        assert {k: b[k] for k in b.dtype.names} == g.best


def test_tiebreak_rule():
    """optimize_main.rs:73-116: min p -> max intersection -> smallest (rank1, rank2)."""
    rec = np.zeros(5, dtype=O.RECORD_DTYPE)
    rec["rank1"] = [1, 1, 2, 2, 3]
    rec["rank2"] = [5, 6, 1, 2, 1]
    rec["intersection_size"] = [3, 4, 4, 2, 4]
    rec["pvalue"] = [0.1, 0.01, 0.01, 0.01, 0.01]
    b = O.argmin_tiebreak(rec)
    assert (int(b["rank1"]), int(b["rank2"])) == (1, 6)
    rec["intersection_size"] = [3, 1, 1, 1, 1]
    b = O.argmin_tiebreak(rec)
    assert (int(b["rank1"]), int(b["rank2"])) == (1, 6)


def test_oracle_accuracy_against_scipy():
    """Independent accuracy reference (not the parity target): statrs-order sum vs scipy's hypergeom.sf."""
    from scipy import stats

    rng = np.random.default_rng(0)
    lf = O.ln_factorial_table(20000)
    for _ in range(200):
        N = int(rng.integers(50, 20000))
        K, n = int(rng.integers(1, N)), int(rng.integers(1, N))
        lo, hi = max(0, K + n - N), min(K, n)
        k = int(rng.integers(lo, hi + 1))
        p = O.hypergeometric_pvalue_cached(lf, N, K, n, k)
        q = stats.hypergeom.sf(k - 1, N, K, n)
        if q > 1e-290:
            assert abs(np.log(p) - np.log(q)) <= 1e-9 * max(1.0, abs(np.log(q))), (N, K, n, k, p, q)
        lp = O.hypergeometric_log_pvalue(lf, N, K, n, k)
        if p > 0:
            assert abs(lp - np.log(p)) <= 1e-10 * max(1.0, abs(lp))


def test_underflow_is_exact_zero_and_not_clamped():
    assert O.hypergeometric_pvalue(6000, 3000, 3000, 3000) == 0.0
    p = O.hypergeometric_pvalue(6000, 100, 100, 100)
    assert 3.2e-220 < p < 3.3e-220
    assert O.hypergeometric_pvalue(10, 5, 5, 0) == 1.0
    assert np.isnan(O.hypergeometric_pvalue(10, 11, 5, 1))  # Hypergeometric::new fails -> panic in the reference


def test_committed_oracle_vectors_are_current():
    """tests/golden/oracle_vectors.json (tools/make_golden.py) still matches the oracle."""
    V = json.load(open(os.path.join(H.GOLDEN, "oracle_vectors.json")))
    by = {c["name"]: c for c in V["cases"]}
    ids1, r1, ids2, r2, bg = H.load_test_data()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    c = by["test_data"]
    g = O.grid_int(o1, o2, c["population"])
    assert int(g.overlap.astype(np.uint64).sum()) == c["overlap_checksum"]
    assert float(g.p[-1, -1]).hex() == c["p_hex_corner"]
    p1, p2 = H.perms(30, 16, c["perm_seed"]), H.perms(30, 16, c["perm_seed"] + 1)
    for t in range(16):
        b = O.grid_int(o1, o2, c["population"], None, p1[t], p2[t]).best
        assert {k: (float(v) if k == "pvalue" else int(v)) for k, v in b.items()} == c["permuted_best"][t]


def test_run_single_node_threads_and_modes():
    ids1, r1, ids2, r2, bg = H.load_test_data()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    tasks = [0] + [1] * 20
    a = O.run_single_node(o1, o2, 30, tasks, 3, seed=5, mode=0)
    b = O.run_single_node(o1, o2, 30, tasks, 1, seed=5, mode=1)
    assert np.array_equal(a, b)  # faithful == integer mode, independent of the thread count
    assert a[0]["permuted"] == 0 and a[0]["pvalue"] == 0.15632183908046102 and all(a[1:]["permuted"] == 1)
    out = O.final_json(a)
    assert 0.0 <= out["empirical_pvalue"] <= 1.0


def test_numpy_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (the checker's restatement used to replay the device generator)."""
    def words(*a):
        return [int(x) for x in H.philox4x32_10(*a)]

    assert words(0, 0, 0, 0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert words(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert words(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
