"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): overlap counts and optimal threshold pairs bit-exact; p-values within 1e-9
relative in log p (here they agree to ~1e-13: the device sums the same ln_factorial table in statrs' order and
differs only in exp()'s last ulp).  EVERY record is compared on (rank1, rank2, set sizes, overlap): where the reference's
`==` / `<` between two p-values could hinge on that last ulp, the library re-evaluates the tie set on the host with the
host libm (flag TIE_RESOLVED), so there is no record the tests exempt."""
import numpy as np
import pytest

import dual_threshold_optimization_b200 as dto
from tests import helpers as H
from tests.helpers import O

pytestmark = pytest.mark.gpu

LOGP_RTOL = 1e-9  # north_star tolerance, relative in log p


def load(engine, ids1, r1, ids2, r2, background=None):
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    bg = dto.FeatureList(background) if background is not None else None
    N = dto.compute_population_size(l1, l2, bg)
    engine.load_lists(l1, l2, N)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    assert O.compute_population_size(o1, o2, background) == N
    assert np.array_equal(l1.thresholds(), o1.thresholds) and np.array_equal(l2.thresholds(), o2.thresholds)
    return o1, o2, N, O.slot_map(o1, o2)


def check_grid(engine, o1, o2, N, slot, perm1=None, perm2=None):
    ov, pv, lp = engine.grid_debug(perm1, perm2, want_p=True, want_logp=True)
    ref = O.grid_int(o1, o2, N, slot, perm1, perm2, want_logp=True)
    assert np.array_equal(ov, ref.overlap), "overlap counts must be bit-exact"
    zero = ref.p == 0.0
    assert np.array_equal(pv == 0.0, zero), "underflow-to-zero cells must coincide"
    nz = ~zero
    assert np.allclose(pv[nz], ref.p[nz], rtol=1e-12, atol=0.0)
    fin = np.isfinite(ref.logp)
    assert np.array_equal(np.isfinite(lp), fin)
    tol = LOGP_RTOL * np.maximum(np.abs(ref.logp[fin]), 1e-300) + 1e-13
    assert np.all(np.abs(lp[fin] - ref.logp[fin]) <= tol), "log p outside the 1e-9 relative tolerance"
    return ref


CASES = {
    "test_data": lambda: H.load_test_data(),
    "ten_gene": lambda: H.ten_gene_case() + (None,),
    "synthetic_300_ties": lambda: H.synthetic_pair(300, 3, 0.25, tied_frac=0.2) + (None,),
    "background_330_290": lambda: H.background_case(400, 330, 290, 5),
    "maxrank_1000": lambda: H.synthetic_pair(1000, 9, 0.3) + (None,),
    "null_2000": lambda: H.synthetic_pair(2000, 13, None) + (None,),
    "identical_1500": lambda: H.synthetic_pair(1500, 1, 0.0) + (None,),
}


@pytest.mark.parametrize("name", list(CASES))
def test_grid_and_unpermuted_optimum(engine, name):
    ids1, r1, ids2, r2, bg = CASES[name]()
    o1, o2, N, slot = load(engine, ids1, r1, ids2, r2, bg)
    ref = check_grid(engine, o1, o2, N, slot)
    rec = engine.run_unpermuted()
    H.assert_record_matches(rec, ref.best)
    assert not (int(rec["flags"]) & dto._capi.FLAG_PERMUTED)
    # and through the scan kernel (identity indices): same optimum
    i1, i2 = np.arange(len(ids1), dtype=np.uint32)[None, :], np.arange(len(ids2), dtype=np.uint32)[None, :]
    rk = engine.run_permuted_indices(i1, i2)[0]
    H.assert_record_matches(rk, ref.best)
    if len(ids1) <= 400:  # the reference-faithful (string/HashSet/uncached) oracle agrees too
        fb = O.optimize_faithful(o1, o2, N)
        H.assert_record_matches(rec, {k: fb[k] for k in fb.dtype.names})


def test_reference_goldens_through_the_gpu(engine):
    # dto/optimize_main.rs:168-170
    ids1, r1, ids2, r2 = H.ten_gene_case()
    load(engine, ids1, r1, ids2, r2)
    rec = engine.run_unpermuted()
    assert (int(rec["rank1"]), int(rec["rank2"])) == (3, 4)
    assert float(rec["pvalue"]) == pytest.approx(0.33333333333333337, rel=1e-15)
    # README.md:174-184
    ids1, r1, ids2, r2, bg = H.load_test_data()
    load(engine, ids1, r1, ids2, r2, bg)
    rec = engine.run_unpermuted()
    assert [int(rec[f]) for f in ("rank1", "rank2", "set1_len", "set2_len", "intersection_size", "population_size")] == [24, 13, 24, 13, 12, 30]
    assert float(rec["pvalue"]) == pytest.approx(0.15632183908046102, rel=1e-14)
    # stat_operations/hypergeometric_pvalue.rs:27-31 and :56-87
    p = engine.hypergeometric_pvalues([1000, 6060, 3, 3, 3, 3, 3, 10], [50, 5808, 1, 1, 1, 2, 2, 1], [60, 154, 1, 2, 3, 1, 2, 1], [10, 153, 1, 1, 1, 1, 2, 0])
    want = [0.00044068070222441115, 0.010413637619010246, 1 / 3, 2 / 3, 1.0, 2 / 3, 1 / 3, 1.0]
    assert np.allclose(p, want, rtol=1e-13, atol=0)


@pytest.mark.parametrize("name,P", [("test_data", 64), ("ten_gene", 32), ("background_330_290", 48), ("synthetic_300_ties", 48), ("null_2000", 24), ("identical_1500", 8)])
def test_permuted_host_indices_match_oracle(engine, name, P):
    ids1, r1, ids2, r2, bg = CASES[name]()
    o1, o2, N, slot = load(engine, ids1, r1, ids2, r2, bg)
    p1, p2 = H.perms(len(ids1), P, 100), H.perms(len(ids2), P, 200)
    recs = engine.run_permuted_indices(p1, p2)
    for t in range(P):
        ob = O.grid_int(o1, o2, N, slot, p1[t], p2[t]).best
        assert int(recs[t]["flags"]) & dto._capi.FLAG_PERMUTED
        H.assert_record_matches(recs[t], ob)
        if int(recs[t]["flags"]) & dto._capi.FLAG_HOST_PVALUE:  # settled on the host: the reference's very bits
            assert float(recs[t]["pvalue"]) == float(ob["pvalue"])
    # a few full grids under permutation
    for t in range(min(P, 3)):
        check_grid(engine, o1, o2, N, slot, p1[t], p2[t])


@pytest.mark.parametrize("N,P,sigma,ties", [(6000, 6, 0.25, 0.0), (6000, 6, None, 0.0), (20000, 3, None, 0.0), (6000, 4, 0.25, 0.05)])
def test_baseline_sizes_match_oracle(engine, N, P, sigma, ties):
    """configs C2 / C3 of BASELINE.json at full feature counts (few tasks: the oracle needs ~seconds per task), plus the
    tied variant of SURVEY 8(d): 5 % of the features collapsed into tie groups."""
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, sigma, tied_frac=ties)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    if not ties:
        assert (o1.thresholds.size, o2.thresholds.size) == ({6000: 469, 20000: 589}[N],) * 2
    lf = O.ln_factorial_table(pop)
    rec = engine.run_unpermuted()
    H.assert_record_matches(rec, O.grid_int(o1, o2, pop, slot, lf=lf, want_overlap=False, want_p=False).best)
    p1, p2 = H.perms(N, P, 1), H.perms(N, P, 2)
    recs = engine.run_permuted_indices(p1, p2)
    for t in range(P):
        ob = O.grid_int(o1, o2, pop, slot, p1[t], p2[t], lf=lf, want_overlap=False, want_p=False).best
        H.assert_record_matches(recs[t], ob)
    ov, _, _ = engine.grid_debug(p1[0], p2[0], want_p=False)
    assert np.array_equal(ov, O.grid_int(o1, o2, pop, slot, p1[0], p2[0], lf=lf, want_p=False).overlap)


@pytest.mark.parametrize("name", ["null_2000", "background_330_290", "test_data"])
def test_device_philox_permutations_replay_through_oracle(engine, name):
    ids1, r1, ids2, r2, bg = CASES[name]()
    o1, o2, N, slot = load(engine, ids1, r1, ids2, r2, bg)
    P = 40
    recs = engine.run_permuted_philox(seed=42, first_perm_id=1000, P=P)
    n_common = int((slot >= 0).sum())
    for t in range(0, P, 3):
        pairing = engine.philox_pairing(42, 1000 + t)
        paired = pairing[pairing != 0xFFFFFFFF]
        assert paired.size == n_common and np.unique(paired).size == paired.size and paired.max() < len(ids2)
        p1, p2 = H.perms_from_pairing(pairing, slot, len(ids2))
        ob = O.grid_int(o1, o2, N, slot, p1, p2).best
        H.assert_record_matches(recs[t], ob)
    # results are a pure function of (seed, permutation id): independent of batching / sharding
    a = engine.run_permuted_philox(42, 1000, 17)
    b = engine.run_permuted_philox(42, 1017, P - 17)
    assert np.array_equal(np.concatenate([a, b]), recs)
    assert not np.array_equal(engine.run_permuted_philox(43, 1000, P)["pvalue"], recs["pvalue"])


def test_device_pairing_is_uniform_small_n(engine):
    """All 4! pairings of a 4-feature problem are equally likely (chi-square over 4800 permutation ids)."""
    from scipy import stats

    ids = ["a", "b", "c", "d"]
    r = np.array([1, 2, 3, 4], dtype=np.uint32)
    load(engine, ids, r, ids, r)
    counts = {}
    for t in range(4800):
        key = tuple(int(x) for x in engine.philox_pairing(5, t))
        counts[key] = counts.get(key, 0) + 1
    assert len(counts) == 24
    chi = sum((c - 200.0) ** 2 / 200.0 for c in counts.values())
    assert stats.chi2.sf(chi, 23) > 1e-4


def test_device_pairing_position_marginals(engine):
    """n=512: every (position, partner) cell is hit uniformly: chi-square on the 512x512 -> 16x16 coarse table."""
    from scipy import stats

    n = 512
    ids1, r1, ids2, r2 = H.synthetic_pair(n, 2, None)
    load(engine, ids1, r1, ids2, r2)
    T = 600
    table = np.zeros((16, 16))
    fixed = 0
    for t in range(T):
        pr = engine.philox_pairing(99, t).astype(np.int64)
        assert np.array_equal(np.sort(pr), np.arange(n))
        np.add.at(table, (np.arange(n) // 32, pr // 32), 1)
        fixed += int((pr == np.arange(n)).sum())
    exp = T * n / 256.0
    chi = ((table - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi, 15 * 15) > 1e-4
    assert abs(fixed / T - 1.0) < 0.25  # fixed points of a uniform permutation ~ Poisson(1)


def test_null_distribution_ks_against_oracle_null(engine):
    """north_star: on-device Philox permutations are checked distributionally (KS on the null of min p)
    against the reference-semantics null (oracle with numpy permutations)."""
    from scipy import stats

    N = 1500
    ids1, r1, ids2, r2 = H.synthetic_pair(N, 21, 0.25)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    dev = engine.run_permuted_philox(seed=7, first_perm_id=0, P=4000, want_records=False, want_minp=True)
    lf = O.ln_factorial_table(pop)
    p1, p2 = H.perms(N, 300, 11), H.perms(N, 300, 12)
    ref = np.array([O.grid_int(o1, o2, pop, slot, p1[t], p2[t], lf=lf, want_overlap=False, want_p=False).best["pvalue"] for t in range(300)])
    ks = stats.ks_2samp(np.log(dev), np.log(ref))
    assert ks.pvalue > 1e-3, ks
    # and the same test between two device seeds is also unremarkable
    dev2 = engine.run_permuted_philox(seed=8, first_perm_id=0, P=4000, want_records=False, want_minp=True)
    assert stats.ks_2samp(dev, dev2).pvalue > 1e-3


def test_underflow_plateau_tiebreak(engine):
    """Strongly concordant lists: the reference's p underflows to exactly 0.0 in many cells and the winner is
    decided by the integer tie-break (largest intersection, then smallest (rank1, rank2))."""
    N = 3000
    ids1, r1, ids2, r2 = H.synthetic_pair(N, 4, 0.0)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    ref = O.grid_int(o1, o2, pop, slot)
    assert (ref.p == 0.0).sum() > 100
    rec = engine.run_unpermuted()
    assert float(rec["pvalue"]) == 0.0
    H.assert_record_matches(rec, ref.best)
    # the same task through the warp-per-permutation scan kernel (identity "permutation"): plateau logic of K1
    ident = np.arange(N, dtype=np.uint32)[None, :]
    rk = engine.run_permuted_indices(ident, ident)[0]
    assert float(rk["pvalue"]) == 0.0
    H.assert_record_matches(rk, ref.best)
    # near-identical: 2% of list 2 perturbed
    ids1, r1, ids2, r2 = H.synthetic_pair(N, 4, 0.002)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    want = O.grid_int(o1, o2, pop, slot).best
    H.assert_record_matches(engine.run_unpermuted(), want)
    H.assert_record_matches(engine.run_permuted_indices(ident, ident)[0], want)


def test_degenerate_and_error_paths(engine):
    # no common genes: every cell short-circuits to p = 1.0 -> first threshold pair wins (dense path)
    l1 = dto.RankedFeatureList.from_(["a", "b", "c"], [1, 2, 3])
    l2 = dto.RankedFeatureList.from_(["x", "y", "z"], [1, 2, 3])
    engine.load_lists(l1, l2, 6)
    rec = engine.run_unpermuted()
    assert (int(rec["rank1"]), int(rec["rank2"]), int(rec["intersection_size"]), float(rec["pvalue"])) == (1, 1, 0, 1.0)
    o1, o2 = H.oracle_lists(["a", "b", "c"], [1, 2, 3], ["x", "y", "z"], [1, 2, 3])
    H.assert_record_matches(rec, O.optimize_faithful(o1, o2, 6))
    # optimize_main.rs doctest :20-52: different features, population 4, permuted -> a Best record
    l1 = dto.RankedFeatureList.from_(["gene1", "gene2", "gene3"], [1, 2, 3])
    l2 = dto.RankedFeatureList.from_(["gene2", "gene3", "gene4"], [1, 2, 3])
    r = dto.optimize(l1, l2, True, 4, engine=engine)
    assert r.permuted and r.population_size == 4
    # Hypergeometric::new panics when a set is larger than the population
    with pytest.raises(dto.DtoPanic):
        engine.load_lists(l1, l2, 2)
    # empty list -> no thresholds -> unwrap panic in the reference
    with pytest.raises(dto.DtoPanic):
        engine.load_lists(dto.RankedFeatureList.from_([], []), l2, 4)
    # the standalone evaluator maps the same condition to the same panic (hypergeometric_pvalue.rs:40-41)
    with pytest.raises(dto.DtoPanic):
        engine.hypergeometric_pvalues([10, 10], [3, 11], [4, 4], [1, 1])
    with pytest.raises(dto.DtoError):
        engine.set_option("levels", 1)  # fewer than two screen levels certify nothing
    # not a permutation
    engine.load_lists(l1, l2, 4)
    with pytest.raises(dto.DtoError):
        engine.run_permuted_indices(np.array([[0, 0, 1]], np.uint32), np.array([[0, 1, 2]], np.uint32))
    with pytest.raises(dto.DtoError):  # out-of-range entries are reported, never dereferenced
        engine.run_permuted_indices(np.array([[0, 1, 2]], np.uint32), np.array([[0, 7, 4000000000]], np.uint32))
    assert engine.run_permuted_indices(np.array([[2, 0, 1]], np.uint32), np.array([[1, 2, 0]], np.uint32)).size == 1  # context still healthy
    # duplicate ids are rejected (documented deviation)
    with pytest.raises(dto.DtoError):
        engine.load_lists(dto.RankedFeatureList.from_(["a", "a"], [1, 2]), l2, 4)


def test_run_single_node_and_epilogue(engine):
    ids1, r1, ids2, r2, bg = H.load_test_data()
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    N = dto.compute_population_size(l1, l2, dto.FeatureList(bg))
    tasks = [dto.Task(0, False)] + [dto.Task(i, True) for i in range(1, 1001)]  # config C1: 1 000 permutations
    res = dto.run_single_node(tasks, l1, l2, N, 4)
    assert len(res) == 1001 and not res[0].permuted and all(r.permuted for r in res[1:])
    out = dto.empirical_pvalue(res)
    assert {k: out[k] for k in ("rank1", "rank2", "set1_len", "set2_len", "unpermuted_intersection_size", "population_size", "fdr")} == {
        "rank1": 24, "rank2": 13, "set1_len": 24, "set2_len": 13, "unpermuted_intersection_size": 12, "population_size": 30, "fdr": 0.0}
    assert out["unpermuted_pvalue"] == pytest.approx(0.15632183908046102, rel=1e-14)
    # the empirical p of the oracle's own null (numpy permutations) agrees statistically
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    p1, p2 = H.perms(30, 1000, 3), H.perms(30, 1000, 4)
    ref_null = np.array([O.grid_int(o1, o2, N, slot, p1[t], p2[t], want_overlap=False, want_p=False).best["pvalue"] for t in range(1000)])
    ref_emp = float((ref_null <= 0.15632183908046102).mean())
    assert abs(out["empirical_pvalue"] - ref_emp) < 0.07


def test_cli_end_to_end(built):
    import json
    import os
    import subprocess

    td = os.path.join(H.GOLDEN, "test_data")
    cli = dto._capi.CLI_PATH
    r = subprocess.run([cli, "-1", f"{td}/ranklist1.csv", "-2", f"{td}/ranklist2.csv", "-b", f"{td}/background.txt", "-p", "5", "-t", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert list(out) == sorted(out)  # serde_json without preserve_order prints keys alphabetically
    assert out["rank1"] == 24 and out["rank2"] == 13 and out["unpermuted_intersection_size"] == 12 and out["fdr"] == 0.0
    assert '"unpermuted_pvalue": 0.15632183908046102' in r.stdout and '"fdr": 0.0' in r.stdout
    assert "Permutations: 5" in r.stderr and "threshold lists" in r.stderr and ": 900" in r.stderr
    # -m shards over every visible GPU (the single-box replacement of the MPI mode); --seed makes the null reproducible
    args = [cli, "-1", f"{td}/ranklist1.csv", "-2", f"{td}/ranklist2.csv", "-p", "2000", "-m", "--seed", "7"]
    a = subprocess.run(args, capture_output=True, text=True, timeout=300)
    b = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and a.stdout == b.stdout and "Multi-node mode: enabled" in a.stderr
    oa = json.loads(a.stdout)
    assert oa["population_size"] == 30 and oa["rank1"] == 24 and 0.5 < oa["empirical_pvalue"] <= 1.0
    # the reference's panics surface as a non-zero exit with the reference's message
    bad = subprocess.run([cli, "-1", f"{td}/ranklist1.csv", "-2", f"{td}/background.txt"], capture_output=True, text=True, timeout=60)
    assert bad.returncode != 0 and "Invalid format in file" in bad.stderr


def test_committed_golden_vectors(engine):
    """tests/golden/oracle_vectors.json (tools/make_golden.py): committed fixtures, independent of the oracle build."""
    import json
    import os

    V = json.load(open(os.path.join(H.GOLDEN, "oracle_vectors.json")))
    for c in V["cases"]:
        ids1, r1, ids2, r2, bg = CASES[c["name"]]()
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        assert (l1.thresholds().size, l2.thresholds().size) == (c["T1"], c["T2"])
        assert [int(x) for x in l1.thresholds()[-3:]] == c["thresholds1_tail"]
        engine.load_lists(l1, l2, c["population"])
        ov, pv, _ = engine.grid_debug()
        assert int(ov.astype(np.uint64).sum()) == c["overlap_checksum"]
        assert float(pv[-1, -1]) == pytest.approx(float.fromhex(c["p_hex_corner"]), rel=1e-13, abs=0)
        H.assert_record_matches(engine.run_unpermuted(), c["unpermuted_best"])
        P = len(c["permuted_best"])
        p1, p2 = H.perms(len(ids1), P, c["perm_seed"]), H.perms(len(ids2), P, c["perm_seed"] + 1)
        recs = engine.run_permuted_indices(p1, p2)
        for t in range(P):
            H.assert_record_matches(recs[t], c["permuted_best"][t])


def test_full_size_properties_c3(engine):
    """BASELINE configs[2] (N = 20 000, 589 x 589) at scale, through size-independent properties:
    determinism across batch sizes (checksum of records), every record self-consistent under re-evaluation on
    the device, minimum really minimal against the dense grid, set sizes = #{ranks <= threshold}."""
    N, P = 20000, 30000
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, 0.25)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    engine.load_lists(l1, l2, N)
    assert engine.shape[:2] == (589, 589)
    engine.reset_stats()
    engine.set_option("batch", 0)
    a = engine.run_permuted_philox(20000, 0, P)
    engine.set_option("batch", 4099)
    b = engine.run_permuted_philox(20000, 0, P)
    engine.set_option("batch", 0)
    assert np.array_equal(a, b)
    st = engine.stats()
    assert st["tasks_full"] == 0
    # re-evaluate every record's p from its own (N, K, n, k) with the standalone device kernel
    p = engine.hypergeometric_pvalues(np.full(P, N), a["set1_len"], a["set2_len"], a["intersection_size"])
    on_host = (a["flags"] & dto._capi.FLAG_HOST_PVALUE) != 0  # tie sets settled with the host libm: equal up to exp()'s last ulp
    assert np.array_equal(p[~on_host], a["pvalue"][~on_host])
    assert np.allclose(p[on_host], a["pvalue"][on_host], rtol=1e-13, atol=0)
    assert st["tasks_tie_resolved"] == 2 * int(on_host.sum()) and on_host.mean() < 0.02  # both runs counted
    t1, t2 = l1.thresholds(), l2.thresholds()
    assert np.array_equal(a["set1_len"], np.searchsorted(l1.ranks(), a["rank1"], side="right"))
    assert np.array_equal(a["set2_len"], np.searchsorted(l2.ranks(), a["rank2"], side="right"))
    assert np.all(np.isin(a["rank1"], t1)) and np.all(np.isin(a["rank2"], t2))
    assert np.all((a["pvalue"] > 0) & (a["pvalue"] < 0.5))
    # the dense grid of a few of these permutations: the record is the argmin with the reference tie-break
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    for t in (0, 777, P - 1):
        pairing = engine.philox_pairing(20000, t)
        assert np.array_equal(np.sort(pairing), np.arange(N))
        p1, p2 = H.perms_from_pairing(pairing, slot, N)
        ov, pv, _ = engine.grid_debug(p1, p2)
        i, j = np.unravel_index(np.argmin(pv), pv.shape)
        assert pv[i, j] == pytest.approx(float(a[t]["pvalue"]), rel=1e-13)
        assert int(ov[np.searchsorted(t1, a[t]["rank1"]), np.searchsorted(t2, a[t]["rank2"])]) == int(a[t]["intersection_size"])
    # null calibration: P(min p <= x) is monotone and the empirical p of the null median is ~0.5
    med = np.median(a["pvalue"])
    assert abs(float((a["pvalue"] <= med).mean()) - 0.5) < 0.01


def test_background_subset_config_c5_small(engine):
    """configs[4] semantics at reduced size: lists filtered to a background subset keep their original ranks
    (gaps), population = |background| (hard part 5)."""
    rng = np.random.default_rng(60000)
    U, B = 3000, 2000
    uni = H.ids_for(U, "t")
    bg_idx = np.sort(rng.choice(U, size=B, replace=False))
    ids1, r1, ids2, r2 = H.synthetic_pair(U, 60000, 0.3)
    keep = set(uni[i] for i in bg_idx)
    ids1 = [f"t{x[1:]}" for x in ids1]
    ids2 = [f"t{x[1:]}" for x in ids2]
    m1 = [i for i, g in enumerate(ids1) if g in keep]
    m2 = [i for i, g in enumerate(ids2) if g in keep]
    f1, fr1 = [ids1[i] for i in m1], r1[m1]
    f2, fr2 = [ids2[i] for i in m2], r2[m2]
    o1, o2, N, slot = load(engine, f1, fr1, f2, fr2, [uni[i] for i in bg_idx])
    assert N == B and len(f1) == B
    ref = check_grid(engine, o1, o2, N, slot)
    H.assert_record_matches(engine.run_unpermuted(), ref.best)
    p1, p2 = H.perms(B, 12, 5), H.perms(B, 12, 6)
    recs = engine.run_permuted_indices(p1, p2)
    for t in range(12):
        H.assert_record_matches(recs[t], O.grid_int(o1, o2, N, slot, p1[t], p2[t], want_overlap=False, want_p=False).best)


def test_batched_pairs_driver(engine):
    """BASELINE configs[3] at reduced size: many independent list pairs, each = unpermuted optimum + permutation null
    + epilogue; sharding over contexts does not change any result."""
    pairs, want = [], []
    for q in range(6):
        ids1, r1, ids2, r2 = H.synthetic_pair(300 + 40 * q, 1000 + q, 0.35 if q % 2 else None)
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        pairs.append((l1, l2, dto.compute_population_size(l1, l2, None)))
        o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
        want.append(O.grid_int(o1, o2, len(ids1)).best)
    a = dto.run_pairs(pairs, 400, devices=[0], seed=9)
    b = dto.run_pairs(pairs, 400, devices=[0, 0, 0], seed=9)  # three contexts on the same GPU: same sharding code path
    assert a == b
    for q, (res, ob) in enumerate(zip(a, want)):
        assert (res["rank1"], res["rank2"], res["set1_len"], res["set2_len"], res["unpermuted_intersection_size"]) == (
            int(ob["rank1"]), int(ob["rank2"]), int(ob["set1_len"]), int(ob["set2_len"]), int(ob["intersection_size"]))
        assert res["unpermuted_pvalue"] == pytest.approx(float(ob["pvalue"]), rel=1e-12)
        assert res["fdr"] == O.fdr(int(ob["set1_len"]), int(ob["set2_len"]), int(ob["intersection_size"]), pairs[q][2], 0.8)
        assert 0.0 <= res["empirical_pvalue"] <= 1.0
        if q % 2:  # concordant pairs are significant against their own null
            assert res["empirical_pvalue"] <= 0.01
    # null pairs: the empirical p of a null pair is roughly uniform -- at least not all tiny
    assert max(a[q]["empirical_pvalue"] for q in (0, 2, 4)) > 0.05
    assert dto.run_pairs(pairs, 400, seed=10) != a
    # concurrency stress: four contexts on one GPU working on problems of different shapes at the same time
    for rep in range(3):
        assert dto.run_pairs(pairs * 3, 150, devices=[0, 0, 0, 0], seed=rep) == dto.run_pairs(pairs * 3, 150, devices=[0], seed=rep)


def test_table_cache_reuses_tables_only_when_set_sizes_match(engine):
    """The screen / log-p tables depend on (population, set sizes per threshold) only: a second list pair with the same
    tie-free rank structure reuses them (stats.table_cache_hits), a pair with ties rebuilds them, and records are
    identical with the cache switched off."""
    def records(cache):
        eng = dto.Engine(0)
        eng.set_option("table_cache", cache)
        out = []
        for seed, tied in ((1, 0.0), (2, 0.0), (3, 0.2), (4, 0.0)):
            ids1, r1, ids2, r2 = H.synthetic_pair(1200, seed, 0.3 if seed % 2 else None, tied_frac=tied)
            l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
            eng.load_lists(l1, l2, 1200)
            recs = eng.run_permuted_philox(5, 0, 300)
            out.append((eng.run_unpermuted().tobytes(), recs.tobytes()))
            if seed == 2:
                o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
                H.assert_record_matches(eng.run_unpermuted(), O.grid_int(o1, o2, 1200).best)
        return out, eng.stats()["table_cache_hits"]

    on, hits_on = records(1)
    off, hits_off = records(0)
    assert on == off
    assert hits_off == 0
    assert hits_on == 1  # pair 2 reuses pair 1's tables; the tied pair 3 and pair 4 (after 3) rebuild


def test_generic_and_packed_screen_agree(engine, built):
    """The scan kernel has two screen code paths: packed 15-bit (set sizes <= 32766) and generic 16-bit (longer lists).
    Both must give identical records; the generic one is also what lists longer than 32 766 features use."""
    ids1, r1, ids2, r2 = H.synthetic_pair(3000, 77, 0.3)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    engine.load_lists(l1, l2, 3000)
    a = engine.run_permuted_philox(5, 0, 3000)
    with dto.Engine(0) as e2:
        e2.set_option("packed_screen", 0)
        e2.load_lists(l1, l2, 3000)
        b = e2.run_permuted_philox(5, 0, 3000)
        ident = np.arange(3000, dtype=np.uint32)[None, :]
        u = e2.run_permuted_indices(ident, ident)[0]
    assert np.array_equal(a, b)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    H.assert_record_matches(u, O.grid_int(o1, o2, 3000).best)


@pytest.mark.parametrize("N,T", [(40000, 659), (65534, 708)])
def test_long_lists_generic_path(engine, N, T):
    """Lists longer than 32 766 features (generic 16-bit screen; at the 65 534-feature maximum the pairing kernel also
    moves its arrival slots to the global scratch): a few permutations replayed through the oracle."""
    ids1, r1, ids2, r2 = H.synthetic_pair(N, 4, None)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    assert engine.shape[:2] == (T, T)
    lf = O.ln_factorial_table(pop)
    recs = engine.run_permuted_philox(3, 0, 64)
    for t in (0, 63):
        pairing = engine.philox_pairing(3, t)
        assert np.array_equal(np.sort(pairing), np.arange(N))
        p1, p2 = H.perms_from_pairing(pairing, slot, N)
        ob = O.grid_int(o1, o2, pop, slot, p1, p2, lf=lf, want_overlap=False, want_p=False).best
        H.assert_record_matches(recs[t], ob)


def test_huge_rank_values_many_thresholds(engine):
    """Ranks are arbitrary u32 values (gaps, rank 0): a 300-feature list whose ranks reach 5e8 has ~1 650 thresholds
    per list (CH = 64 kernel variant, mostly empty rows)."""
    rng = np.random.default_rng(5)
    n = 300
    ids = H.ids_for(n, "h")
    r1 = np.sort(rng.integers(0, 500_000_000, size=n)).astype(np.uint32)
    r2 = rng.permutation(np.sort(rng.integers(1, 400_000_000, size=n))).astype(np.uint32)
    o1, o2, N, slot = load(engine, ids, r1, list(ids), r2)
    assert o1.thresholds.size > 1500 and o2.thresholds.size > 1500
    ref = O.grid_int(o1, o2, N, slot)
    ov, pv, _ = engine.grid_debug()
    assert np.array_equal(ov, ref.overlap)
    assert np.allclose(pv, ref.p, rtol=1e-12, atol=0)
    H.assert_record_matches(engine.run_unpermuted(), ref.best)
    p1, p2 = H.perms(n, 10, 1), H.perms(n, 10, 2)
    recs = engine.run_permuted_indices(p1, p2)
    for t in range(10):
        H.assert_record_matches(recs[t], O.grid_int(o1, o2, N, slot, p1[t], p2[t], want_overlap=False, want_p=False).best)
    ph = engine.run_permuted_philox(1, 0, 50)
    assert np.all((ph["pvalue"] > 0) & (ph["pvalue"] <= 1.0))


def test_lookup_table_accuracy(engine):
    """The scan kernel trusts the per-problem log-p table to ~1e-8 when it shortlists candidates for the exact stage:
    bound its error against the oracle's log-sum-exp on random cells of the N = 20 000 problem."""
    N = 20000
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, 0.25)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    lf = O.ln_factorial_table(pop)
    c1 = np.searchsorted(o1.ranks, o1.thresholds, side="right")
    c2 = np.searchsorted(o2.ranks, o2.thresholds, side="right")
    rng = np.random.default_rng(3)
    rows = rng.integers(0, 589, size=4000)
    cols = rng.integers(0, 589, size=4000)
    K, n = c1[rows], c2[cols]
    mean = K * n / N
    sd = np.sqrt(np.maximum(K * n * (N - K) * (N - n) / (N ** 2 * (N - 1.0)), 1e-9))
    ks = np.clip(np.round(mean + rng.uniform(-1.0, 5.0, size=4000) * sd + rng.integers(0, 2, size=4000)), 1, np.minimum(K, n)).astype(np.int64)
    got = engine.table_logp(rows, cols, ks)
    inside = ~np.isnan(got)
    assert inside.sum() > 1500
    worst = 0.0
    for x in np.nonzero(inside)[0]:
        want = O.hypergeometric_log_pvalue(lf, pop, int(K[x]), int(n[x]), int(ks[x]))
        worst = max(worst, abs(got[x] - want))
        assert np.log(1.5e-5) - 1e-3 <= want <= np.log(0.95) + 1e-3  # the tabulated range is tau_32 .. tau_1
    assert worst < 1e-9, worst


@pytest.mark.parametrize("n", [30, 1000, 6000, 20000])
def test_device_generator_is_philox_sort(engine, n):
    """The exported device permutation equals an independent numpy replay: Philox4x32-10 keys sorted on (key bits,
    secondary key, index).  This pins both the Philox implementation and the shared-memory sort."""
    ids1, r1, ids2, r2 = H.synthetic_pair(n, 8, None)
    load(engine, ids1, r1, ids2, r2)
    for seed, pid in ((0, 0), (12345678901234567, 7), (2 ** 63 + 5, 2 ** 40 + 3)):
        got = engine.philox_pairing(seed, pid)
        assert np.array_equal(got, H.expected_pairing_identical(n, seed, pid)), (n, seed, pid)


def test_rowwise_fast_pairing_equals_exact_pairing_at_scale(engine):
    """The performance path of the pairing kernel only resolves the ROW of every position (positions inside one
    threshold row are interchangeable); its records must equal, bit for bit, those of the exactly ranked permutation
    with the same id (exported, then fed back through the host-indices path).  N = 20 000, 48 ids."""
    N = 20000
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, 0.25)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    fast = engine.run_permuted_philox(77, 500, 48)
    ident = np.arange(N, dtype=np.uint32)
    p1 = np.tile(ident, (48, 1))
    p2 = np.empty((48, N), dtype=np.uint32)
    for t in range(48):
        pairing = engine.philox_pairing(77, 500 + t)  # exact path: pos2_of_pos1
        # unpermuted list 1 (perm1 = identity): gene at slot j must sit at list-2 position pairing[j]
        p2[t, pairing] = slot
    exact = engine.run_permuted_indices(p1, p2)
    assert np.array_equal(fast, exact)


def test_fuzz_tiny_problems_against_faithful_oracle(engine):
    """60 random tiny problems (1-40 features per list, ties, rank 0, partial gene overlap, arbitrary population):
    dense grid, unpermuted optimum and a few host-index permutations against the reference-faithful oracle form."""
    rng = np.random.default_rng(2024)
    checked = 0
    for case in range(60):
        n1, n2 = int(rng.integers(1, 41)), int(rng.integers(1, 41))
        uni = H.ids_for(60, "z")
        ids1 = [uni[i] for i in rng.choice(60, size=n1, replace=False)]
        ids2 = [uni[i] for i in rng.choice(60, size=n2, replace=False)]
        r1 = rng.integers(0, max(2, n1 + 5), size=n1).astype(np.uint32)
        r2 = rng.integers(0, max(2, 2 * n2), size=n2).astype(np.uint32)
        if r1.max() == 0 or r2.max() == 0:
            continue  # max rank 0 -> no thresholds -> the reference panics (covered elsewhere)
        pop = int(max(n1, n2) + rng.integers(0, 30))
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        engine.load_lists(l1, l2, pop)
        o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
        f = O.process_threshold_pairs_faithful(o1, o2, pop)
        ov, pv, _ = engine.grid_debug()
        assert np.array_equal(ov.ravel(), f["intersection_size"]), case
        assert np.allclose(pv.ravel(), f["pvalue"], rtol=1e-12, atol=0), case
        fb = O.argmin_tiebreak(f)
        H.assert_record_matches(engine.run_unpermuted(), {k: fb[k] for k in fb.dtype.names})
        P = 6
        p1, p2 = H.perms(n1, P, case), H.perms(n2, P, 1000 + case)
        recs = engine.run_permuted_indices(p1, p2)
        for t in range(P):
            fbp = O.argmin_tiebreak(O.process_threshold_pairs_faithful(o1, o2, pop, p1[t], p2[t]))
            H.assert_record_matches(recs[t], {k: fbp[k] for k in fbp.dtype.names})
        ph = engine.run_permuted_philox(case, 0, 8)
        assert np.all(ph["pvalue"] <= 1.0 + 1e-12) and np.all(ph["pvalue"] >= 0.0)
        checked += 1
    assert checked >= 50


def _host_indices_from_device_pairings(engine, seed, first, P, slot, n):
    """(perm1, perm2) rows (permuted.rs index semantics) that realise the device's Philox pairings first .. first+P-1 for
    identical gene sets: list 1 stays put, the gene at list-1 slot j is shown by list 2 at the position paired with j."""
    p1 = np.tile(np.arange(n, dtype=np.uint32), (P, 1))
    p2 = np.empty((P, n), dtype=np.uint32)
    for t in range(P):
        pairing = engine.philox_pairing(seed, first + t)
        p2[t, pairing] = slot
    return p1, p2


def _assert_all_records_match(got, ref):
    """every record: integer fields bit-exact, p within 1e-12 relative; host-settled records carry the oracle's bits"""
    for f in ("rank1", "rank2", "set1_len", "set2_len", "intersection_size", "population_size"):
        bad = np.nonzero(got[f].astype(np.int64) != ref[f].astype(np.int64))[0]
        assert bad.size == 0, (f, bad[:5], got[bad[:5]], ref[bad[:5]])
    gp, rp = got["pvalue"], ref["pvalue"]
    assert np.array_equal(gp == 0.0, rp == 0.0)
    assert np.all(np.abs(gp - rp) <= 1e-12 * rp)
    on_host = (got["flags"] & dto._capi.FLAG_HOST_PVALUE) != 0
    assert np.array_equal(gp[on_host], rp[on_host])
    assert np.all((got["flags"] & dto._capi.FLAG_PERMUTED) != 0)


@pytest.mark.parametrize("N,P,sigma", [(6000, 2000, 0.25), (20000, 2000, 0.25)])
def test_deep_parity_at_baseline_sizes(engine, N, P, sigma):
    """configs C2 / C3 at full size, 2 000 permutations each, three ways that must agree on EVERY record: (1) the
    performance path (on-device Philox pairing, row-wise sort, scan kernel), (2) the same permutations exported and fed
    back as host indices (parity path: compose kernel + scan kernel), (3) the integer oracle on those host indices."""
    ids1, r1, ids2, r2 = H.synthetic_pair(N, N, sigma)
    o1, o2, pop, slot = load(engine, ids1, r1, ids2, r2)
    lf = O.ln_factorial_table(pop)
    engine.reset_stats()
    fast = engine.run_permuted_philox(N + 1, 10**6, P)
    st = engine.stats()
    p1, p2 = _host_indices_from_device_pairings(engine, N + 1, 10**6, P, slot, N)
    host = engine.run_permuted_indices(p1, p2)
    assert np.array_equal(fast, host)
    ref = O.best_batch(o1, o2, pop, p1, p2, slot, lf)
    _assert_all_records_match(fast, ref)
    # plain numpy permutations of BOTH lists too (perm1 != identity), fewer of them
    q1, q2 = H.perms(N, 200, 31), H.perms(N, 200, 32)
    _assert_all_records_match(engine.run_permuted_indices(q1, q2), O.best_batch(o1, o2, pop, q1, q2, slot, lf))
    print(f"N={N}: {P} permutations, tie sets settled on the host: {st['tasks_tie_resolved']}, dense-path tasks: {st['tasks_full']}")


def test_deep_parity_background_subset_c5_shape(engine):
    """configs[4] at its real shape: 60 000-id universe ranked by both lists, filtered to a 40 000-id background (ranks
    keep their gaps, so consecutive thresholds repeat set sizes), population 40 000; 200 device permutations replayed as
    host indices and through the oracle, plus 56 numpy permutations of both lists."""
    rng = np.random.default_rng(60000)
    U, B = 60000, 40000
    ids1, r1, ids2, r2 = H.synthetic_pair(U, 60000, 0.3)
    keep = np.zeros(U, dtype=bool)
    keep[rng.choice(U, size=B, replace=False)] = True
    uni = H.ids_for(U)
    idx_of = {g: i for i, g in enumerate(uni)}
    m1 = [i for i, g in enumerate(ids1) if keep[idx_of[g]]]
    m2 = [i for i, g in enumerate(ids2) if keep[idx_of[g]]]
    f1, fr1 = [ids1[i] for i in m1], r1[m1]
    f2, fr2 = [ids2[i] for i in m2], r2[m2]
    o1, o2, N, slot = load(engine, f1, fr1, f2, fr2, [uni[i] for i in np.nonzero(keep)[0]])
    assert N == B and len(f1) == B and engine.shape[0] >= 690
    lf = O.ln_factorial_table(N)
    H.assert_record_matches(engine.run_unpermuted(), O.grid_int(o1, o2, N, slot, lf=lf, want_overlap=False, want_p=False).best)
    P = 200
    fast = engine.run_permuted_philox(5, 0, P)
    p1, p2 = _host_indices_from_device_pairings(engine, 5, 0, P, slot, B)
    assert np.array_equal(fast, engine.run_permuted_indices(p1, p2))
    _assert_all_records_match(fast, O.best_batch(o1, o2, N, p1, p2, slot, lf))
    q1, q2 = H.perms(B, 56, 41), H.perms(B, 56, 42)
    _assert_all_records_match(engine.run_permuted_indices(q1, q2), O.best_batch(o1, o2, N, q1, q2, slot, lf))


def test_huge_population_margins(engine):
    """Population 5e6 (lists of 2 500 features inside a huge background): ln-factorials reach 7e7, where one ulp is
    1.5e-8 -- the certification margins of the screen scale with ulp(lf[N]) (Problem::refine_eps), so records still
    match the oracle on every field."""
    n, pop = 2500, 5_000_000
    ids1, r1, ids2, r2 = H.synthetic_pair(n, 77, 0.3)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    engine.load_lists(l1, l2, pop)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    lf = O.ln_factorial_table(pop)
    H.assert_record_matches(engine.run_unpermuted(), O.grid_int(o1, o2, pop, slot, lf=lf, want_overlap=False, want_p=False).best)
    p1, p2 = H.perms(n, 96, 1), H.perms(n, 96, 2)
    _assert_all_records_match(engine.run_permuted_indices(p1, p2), O.best_batch(o1, o2, pop, p1, p2, slot, lf))


def test_tie_sets_are_settled_like_the_reference(engine):
    """Symmetric problem (both lists rank the same genes 1..n, so cell (i, j) and its mirror (j, i) have mathematically
    equal p whenever their overlaps agree): many permutations have a tie set of several cells whose order hangs on the
    last ulp of exp().  Every record must still be the oracle's pick, and the settled ones carry its exact p."""
    n = 40
    ids1, r1, ids2, r2 = H.synthetic_pair(n, 5, None)
    o1, o2, N, slot = load(engine, ids1, r1, ids2, r2)
    P = 3000
    p1, p2 = H.perms(n, P, 7), H.perms(n, P, 8)
    engine.reset_stats()
    recs = engine.run_permuted_indices(p1, p2)
    st = engine.stats()
    _assert_all_records_match(recs, O.best_batch(o1, o2, N, p1, p2, slot))
    settled = int(((recs["flags"] & dto._capi.FLAG_TIE_RESOLVED) != 0).sum())
    assert settled == st["tasks_tie_resolved"] and settled >= 5, settled  # measured: 20 of 3 000
    # the epilogue's `<=` against the unpermuted p is the reference's as well (empirical_pvalue.rs:160-165)
    un = engine.run_unpermuted()
    all_recs = np.concatenate([np.array([un], dtype=recs.dtype), recs])
    ob = O.grid_int(o1, o2, N, slot).best
    ref = O.best_batch(o1, o2, N, p1, p2, slot)
    want = float((ref["pvalue"] <= float(ob["pvalue"])).mean())
    assert dto.empirical_pvalue(all_recs)["empirical_pvalue"] == want
    assert float(un["pvalue"]) == float(ob["pvalue"]) and int(un["flags"]) & dto._capi.FLAG_HOST_PVALUE


def _pair_reference(l1, l2, pop, P, seed):
    """what dto_b200_run_pairs promises for one pair: the CLI run on [unpermuted, P permuted tasks with ids 1..P]"""
    tasks = [dto.Task(0, False)] + [dto.Task(i, True) for i in range(1, P + 1)]
    return dto.empirical_pvalue(dto.run.run_single_node_records(tasks, l1, l2, pop, 1, [0], seed))


def test_pair_groups_equal_pair_by_pair_runs(engine):
    """config 4 driver: consecutive pairs with one rank structure share a launch (their unpermuted tasks ride through
    the scan kernel); a pair of another size, a tied pair and a pair with differing gene sets break the groups.  Every
    result must equal the pair's own CLI-style run with seed + q * 0x9E37..., whatever the grouping."""
    specs = [(400, 1, 0.3, 0.0), (400, 2, None, 0.0), (400, 3, 0.0, 0.0), (400, 4, 0.25, 0.0), (350, 5, 0.3, 0.0), (400, 6, None, 0.0),
             (400, 7, 0.3, 0.2), (400, 8, None, 0.0), (400, 9, 0.35, 0.0)]
    pairs = []
    for n, s, sg, tied in specs:
        ids1, r1, ids2, r2 = H.synthetic_pair(n, s, sg, tied_frac=tied)
        if s % 2 == 0:  # different gene order in the two lists: the gene map really differs from pair to pair
            order = np.random.default_rng(s).permutation(n)
            ids2 = [ids2[i] for i in order]
            r2 = r2[order]
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        pairs.append((l1, l2, dto.compute_population_size(l1, l2, None)))
    ids1, r1, ids2, r2, bg = H.background_case(500, 380, 360, 11)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    pairs.insert(4, (l1, l2, dto.compute_population_size(l1, l2, dto.FeatureList(bg))))
    P, seed = 300, 99
    got = dto.run_pairs(pairs, P, devices=[0], seed=seed)
    for q, (l1, l2, pop) in enumerate(pairs):
        want = _pair_reference(l1, l2, pop, P, (seed + q * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        assert got[q] == want, (q, got[q], want)
    # the unpermuted optimum of every pair against the oracle (p bit-exact: evaluated with the host libm)
    for q, (n, s, sg, tied) in enumerate(specs[:4]):
        ids1, r1, ids2, r2 = H.synthetic_pair(n, s, sg, tied_frac=tied)
        if s % 2 == 0:
            order = np.random.default_rng(s).permutation(n)
            ids2 = [ids2[i] for i in order]
            r2 = r2[order]
        o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
        ob = O.grid_int(o1, o2, n).best
        assert (got[q]["rank1"], got[q]["rank2"], got[q]["unpermuted_intersection_size"]) == (int(ob["rank1"]), int(ob["rank2"]), int(ob["intersection_size"]))
        assert got[q]["unpermuted_pvalue"] == float(ob["pvalue"])
    # P = 0: only the unpermuted task, empirical p = 1.0 by definition (empirical_pvalue.rs:150-158)
    z = dto.run_pairs(pairs[:3], 0, devices=[0], seed=seed)
    assert all(r["empirical_pvalue"] == 1.0 for r in z) and [r["rank1"] for r in z] == [r["rank1"] for r in got[:3]]


def test_run_tasks_with_ids_shards_like_the_unsharded_run(engine):
    """Task.id is the Philox id (src/run/task.rs:3-9): two processes' worth of task slices reproduce the single run."""
    ids1, r1, ids2, r2 = H.synthetic_pair(800, 12, 0.3)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    tasks = [dto.Task(0, False)] + [dto.Task(i, True) for i in range(1, 501)]
    whole = dto.run.run_single_node_records(tasks, l1, l2, 800, 1, [0], 5)
    a = dto.run.run_single_node_records(tasks[:201], l1, l2, 800, 1, [0], 5)
    b = dto.run.run_single_node_records(tasks[201:], l1, l2, 800, 1, [0], 5)
    assert np.array_equal(np.concatenate([a, b]), whole)
    # non-contiguous ids, unpermuted task in the middle
    mixed = [tasks[7], tasks[3], tasks[0], tasks[500], tasks[499]]
    m = dto.run.run_single_node_records(mixed, l1, l2, 800, 1, [0], 5)
    assert np.array_equal(m, whole[[7, 3, 0, 500, 499]])


def test_product_allgather_single_rank(engine):
    """dto_b200_allgather_minima with a one-rank communicator made by the library's own helpers (the plumbing a host
    application without torch uses); the multi-rank form is exercised by test_multi_gpu_* and bench.py under torchrun."""
    import ctypes as C

    import torch

    capi = dto._capi
    ids1, r1, ids2, r2 = H.synthetic_pair(600, 3, None)
    load(engine, ids1, r1, ids2, r2)
    uid = (C.c_char * 128)()
    capi.check(capi.lib().dto_b200_nccl_unique_id(uid))
    comm = C.c_void_p()
    capi.check(capi.lib().dto_b200_nccl_comm_create(C.byref(comm), 1, uid, 0, 0))
    P = 257
    d_minp = torch.zeros(P, dtype=torch.float64, device="cuda:0")
    d_all = torch.zeros(P, dtype=torch.float64, device="cuda:0")
    engine.run_permuted_philox_device(3, 1, P, d_minp.data_ptr())
    capi.check(capi.lib().dto_b200_allgather_minima(engine.ctx, comm, C.c_void_p(d_minp.data_ptr()), C.c_void_p(d_all.data_ptr()), P))
    want = engine.run_permuted_philox(3, 1, P)["pvalue"]
    assert np.array_equal(d_all.cpu().numpy(), want)
    capi.check(capi.lib().dto_b200_nccl_comm_destroy(comm))


def _n_gpus():
    return dto.device_count()


@pytest.mark.skipif("_n_gpus() < 2")
def test_multi_gpu_in_process_equals_one_gpu(engine):
    """run_single_node(devices=[0, 1, ..]) -- what `-m` of the CLI maps to, replacing multi_node.rs:114-161 on one box --
    returns the records of the one-GPU run; so does dto_b200_run_pairs sharded by pair."""
    import json
    import os
    import subprocess

    G = min(_n_gpus(), 8)
    ids1, r1, ids2, r2 = H.synthetic_pair(3000, 21, 0.3)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    tasks = [dto.Task(0, False)] + [dto.Task(i, True) for i in range(1, 4001)]
    one = dto.run.run_single_node_records(tasks, l1, l2, 3000, 1, [0], 17)
    many = dto.run.run_single_node_records(tasks, l1, l2, 3000, 1, list(range(G)), 17)
    assert np.array_equal(one, many)
    pairs = []
    for q in range(2 * G + 1):
        a, b, c, d = H.synthetic_pair(500, 100 + q, 0.3 if q % 2 else None)
        x, y = dto.RankedFeatureList.from_(a, b), dto.RankedFeatureList.from_(c, d)
        pairs.append((x, y, 500))
    assert dto.run_pairs(pairs, 200, devices=[0], seed=3) == dto.run_pairs(pairs, 200, devices=list(range(G)), seed=3)
    td = os.path.join(H.GOLDEN, "test_data")
    base = [dto._capi.CLI_PATH, "-1", f"{td}/ranklist1.csv", "-2", f"{td}/ranklist2.csv", "-p", "3000", "--seed", "7"]
    a = subprocess.run(base, capture_output=True, text=True, timeout=300)
    b = subprocess.run(base + ["-m"], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0 and json.loads(a.stdout) == json.loads(b.stdout)


def _allgather_worker(rank, world, uid_bytes, q):
    import ctypes as C

    import torch

    capi = dto._capi
    try:
        ids1, r1, ids2, r2 = H.synthetic_pair(700, 4, None)
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        with dto.Engine(rank) as eng:
            eng.load_lists(l1, l2, 700)
            comm = C.c_void_p()
            uid = (C.c_char * 128).from_buffer_copy(uid_bytes)
            capi.check(capi.lib().dto_b200_nccl_comm_create(C.byref(comm), world, uid, rank, rank))
            P = 300
            torch.cuda.set_device(rank)
            d_minp = torch.zeros(P, dtype=torch.float64, device=f"cuda:{rank}")
            d_all = torch.zeros(P * world, dtype=torch.float64, device=f"cuda:{rank}")
            eng.run_permuted_philox_device(9, 1 + rank * P, P, d_minp.data_ptr())
            capi.check(capi.lib().dto_b200_allgather_minima(eng.ctx, comm, C.c_void_p(d_minp.data_ptr()), C.c_void_p(d_all.data_ptr()), P))
            q.put((rank, d_all.cpu().numpy()))
            capi.check(capi.lib().dto_b200_nccl_comm_destroy(comm))
    except Exception as e:  # surface the failure to the parent instead of hanging it
        q.put((rank, repr(e)))


@pytest.mark.skipif("_n_gpus() < 2")
def test_multi_gpu_product_allgather_over_nccl(engine):
    """One process per GPU, raw ncclComm made with the library's helpers (no torch.distributed): every rank ends up with
    the minima of all ranks, equal to the single-GPU run over the whole id range."""
    import ctypes as C
    import multiprocessing as mp

    capi = dto._capi
    world = 2
    uid = (C.c_char * 128)()
    capi.check(capi.lib().dto_b200_nccl_unique_id(uid))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allgather_worker, args=(r, world, bytes(uid.raw), q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    ids1, r1, ids2, r2 = H.synthetic_pair(700, 4, None)
    load(engine, ids1, r1, ids2, r2)
    want = engine.run_permuted_philox(9, 1, 300 * world)["pvalue"]
    for r in range(world):
        assert isinstance(got[r], np.ndarray), got[r]
        assert np.array_equal(got[r], want)


def test_debug_feature_sets_and_tie_notices(engine, built, tmp_path):
    """debug=true rebuilds FeatureSets::Both on the host (process_threshold_pairs.rs:111-115): the sets of every cell
    intersect in exactly the overlap the device counted -- also under permutation.  And the unpermuted record flags the
    situations in which the reference prints its tie notices (optimize_main.rs:85-107), which the CLI then prints."""
    import subprocess

    ids1, r1, ids2, r2, bg = H.load_test_data()
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    p1, p2 = dto.PermutedRankedFeatureList(l1, seed=1), dto.PermutedRankedFeatureList(l2, seed=2)
    for perm in (False, (p1, p2)):
        recs = dto.optimize(l1, l2, perm, 30, debug=True, engine=engine)
        assert len(recs) == 900
        for r in recs[::37]:
            s1, s2 = r.feature_sets.both()
            assert len(s1) == r.set1_len and len(s2) == r.set2_len
            assert len(set(f.id() for f in s1) & set(f.id() for f in s2)) == r.intersection_size
        if perm:
            assert [f.id() for f in recs[5 * 30].feature_sets.set1()] == [f.id() for f in p1.get_feature_set_by_threshold(int(l1.thresholds()[5]))]
    assert dto.optimize(l1, l2, False, 30, engine=engine).feature_sets is None
    # identical lists: the p-value underflows to 0.0 on a plateau of cells, several with the same (largest) overlap? no:
    # overlap grows along the diagonal, so exactly one cell has the largest one -> first notice only
    n = 3000
    a, ra, b, rb = H.synthetic_pair(n, 4, 0.0)
    la, lb = dto.RankedFeatureList.from_(a, ra), dto.RankedFeatureList.from_(b, rb)
    engine.load_lists(la, lb, n)
    rec = engine.run_unpermuted()
    o1, o2 = H.oracle_lists(a, ra, b, rb)
    ref = O.grid_int(o1, o2, n)
    zero = ref.p == 0.0
    kmax = int(ref.overlap[zero].max())
    f = int(rec["flags"])
    assert bool(f & dto._capi.FLAG_TIE_MINP) == (int(zero.sum()) > 1)
    assert bool(f & dto._capi.FLAG_TIE_OVERLAP) == (int((ref.overlap[zero] == kmax).sum()) > 1)
    # ranks with gaps repeat set sizes at consecutive thresholds: the minimum is shared by cells with equal (K, n, k)
    ids = H.ids_for(60, "q")
    rg = (np.arange(60, dtype=np.uint32) * 7 + 3)
    f1, f2 = tmp_path / "a.csv", tmp_path / "b.csv"
    f1.write_text("".join(f"{g},{r}\n" for g, r in zip(ids, rg)))
    f2.write_text("".join(f"{g},{r}\n" for g, r in zip(ids, rg)))
    out = subprocess.run([dto._capi.CLI_PATH, "-1", str(f1), "-2", str(f2), "-p", "10", "--seed", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    o = H.oracle_lists(ids, rg, ids, rg)
    fr = O.process_threshold_pairs_faithful(o[0], o[1], 60)
    pmin = fr["pvalue"].min()
    ties = fr[fr["pvalue"] == pmin]
    assert ties.size > 1 and "Multiple results with the same minimum p-value (%.15f)" % pmin in out.stderr
    kmx = ties["intersection_size"].max()
    assert ("Multiple results with the same maximum intersection size (%d)" % kmx in out.stderr) == (int((ties["intersection_size"] == kmx).sum()) > 1)
