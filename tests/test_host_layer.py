"""Host layer of the product (C++ behind the C ABI): collections, readers, population size, epilogue, JSON.
None of this needs a GPU; the oracle is only the checker."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import dual_threshold_optimization_b200 as dto
from dual_threshold_optimization_b200 import _capi as capi
from dual_threshold_optimization_b200.stat_operations import empirical_pvalue_struct, final_json
from tests import helpers as H
from tests.helpers import O

G = json.load(open(os.path.join(H.GOLDEN, "reference_goldens.json")))


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


def test_ranked_list_from_matches_oracle():
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 30, 257, 6000):
        ids = [f"g{i}" for i in range(n)]
        ranks = rng.integers(0, max(2, 2 * n), size=n).astype(np.uint32)
        l = dto.RankedFeatureList.from_(ids, ranks)
        o = O.OracleRankedList.make(ids, ranks)
        assert l.ids() == o.ids and np.array_equal(l.ranks(), o.ranks) and np.array_equal(l.thresholds(), o.thresholds)
        assert len(l) == n and l.is_empty() == (n == 0)
    l = dto.RankedFeatureList.from_([f"g{i}" for i in range(20000)], np.arange(1, 20001))
    assert l.thresholds().size == 589 and l.thresholds()[-1] == 19850  # last threshold NOT forced to the max rank
    with pytest.raises(ValueError):
        dto.RankedFeatureList.from_(["a", "b"], [1])
    assert [f.id() for f in l.get_feature_set_by_threshold(3)] == ["g0", "g1", "g2"]


def test_permuted_view_semantics():
    l = dto.RankedFeatureList.from_(["gene1", "gene2", "gene3"], [1, 2, 3])
    p = dto.PermutedRankedFeatureList(l, indices=[2, 0, 1])
    assert [f.id() for f in p.get_feature_set_by_threshold(2)] == ["gene3", "gene1"]  # ranks stay, genes move (permuted.rs:95-99)
    assert len(dto.PermutedRankedFeatureList(l, seed=3).get_feature_set_by_threshold(2)) == 2
    with pytest.raises(ValueError):
        dto.PermutedRankedFeatureList(l, indices=[0, 0, 1])


def test_readers(tmp_path):
    td = os.path.join(H.GOLDEN, "test_data")
    l1 = dto.read_ranked_feature_list_from_csv(os.path.join(td, "ranklist1.csv"))
    assert len(l1) == 30 and l1.ids()[0] == "gene11" and l1.thresholds().size == 30
    bg = dto.read_feature_list_from_file(os.path.join(td, "background.txt"))
    assert len(bg) == 30
    p = tmp_path / "x.csv"
    p.write_text(" geneA , 2\r\ngeneB,1\n")
    l = dto.read_ranked_feature_list_from_csv(str(p))
    assert l.ids() == ["geneB", "geneA"] and list(l.ranks()) == [1, 2]
    p.write_text("geneA,1,extra\n")
    with pytest.raises(dto.DtoPanic, match="Invalid format in file .* at line 1"):
        dto.read_ranked_feature_list_from_csv(str(p))
    p.write_text("geneA\n")
    with pytest.raises(dto.DtoPanic, match="Invalid format"):
        dto.read_ranked_feature_list_from_csv(str(p))
    p.write_text("geneA,1.5\n")
    with pytest.raises(dto.DtoPanic, match="Invalid rank value"):
        dto.read_ranked_feature_list_from_csv(str(p))
    p.write_text("geneA,4294967297\n")  # usize parse ok, `as u32` truncates to 1
    assert list(dto.read_ranked_feature_list_from_csv(str(p)).ranks()) == [1]
    with pytest.raises(dto.DtoError):
        dto.read_ranked_feature_list_from_csv(str(tmp_path / "missing.csv"))
    b = tmp_path / "bg.txt"
    b.write_text("g1\n\n g2 \ng1\n")
    assert dto.read_feature_list_from_file(str(b)).ids() == ["g1", "", "g2", "g1"]  # blanks/duplicates count


def test_compute_population_size_rules():
    l1 = dto.RankedFeatureList.from_(["gene1", "gene2", "gene3"], [1, 2, 3])
    l2 = dto.RankedFeatureList.from_(["gene2", "gene3", "gene4"], [1, 2, 3])
    same = dto.RankedFeatureList.from_(["gene3", "gene1", "gene2"], [1, 2, 3])
    assert dto.compute_population_size(l1, same, None) == 3
    assert dto.compute_population_size(l1, l2, dto.FeatureList(["gene1", "gene2", "gene3", "gene4"])) == 4
    with pytest.raises(dto.DtoPanic, match="If no background is provided, the feature lists must have identical genes."):
        dto.compute_population_size(l1, l2, None)
    with pytest.raises(dto.DtoPanic, match="in the first ranked feature list are not in the background"):
        dto.compute_population_size(l1, l2, dto.FeatureList(["gene2", "gene3", "gene4"]))
    with pytest.raises(dto.DtoPanic, match="in the second ranked feature list are not in the background"):
        dto.compute_population_size(l1, l2, dto.FeatureList(["gene1", "gene2", "gene3"]))
    ids1, r1, ids2, r2, bg = H.load_test_data()
    a, b = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    assert dto.compute_population_size(a, b, dto.FeatureList(bg)) == 30 == dto.compute_population_size(a, b, None)
    # the id join runs on the per-list hash index: repeated ids of list 1 each count (FeatureList::intersect keeps every
    # list-1 item whose id occurs in list 2, feature_list.rs:278-286), ids are compared as whole strings, and the index
    # of a 20 000-feature list agrees with a plain Python set
    dup1 = dto.RankedFeatureList.from_(["a", "a", "b"], [1, 2, 3])
    assert dto.compute_population_size(dup1, dto.RankedFeatureList.from_(["a", "b", "c"], [1, 2, 3]), None) == 3
    # ... which is why the reference accepts this pair too: 3 list-1 items found, both lists 3 long ("ab" is never looked up)
    assert dto.compute_population_size(dup1, dto.RankedFeatureList.from_(["a", "b", "ab"], [1, 2, 3]), None) == 3
    with pytest.raises(dto.DtoPanic, match="identical genes"):
        dto.compute_population_size(dto.RankedFeatureList.from_(["a", "ab", "b"], [1, 2, 3]), dto.RankedFeatureList.from_(["a", "b", "c"], [1, 2, 3]), None)
    big1, rb1, big2, rb2 = H.synthetic_pair(20000, 5, None)
    x, y = dto.RankedFeatureList.from_(big1, rb1), dto.RankedFeatureList.from_(big2[:-1] + ["not-there"], rb2)
    with pytest.raises(dto.DtoPanic, match="identical genes"):
        dto.compute_population_size(x, y, None)
    assert dto.compute_population_size(x, dto.RankedFeatureList.from_(big2, rb2), None) == 20000 == len(set(big1) & set(big2))


def test_fdr_and_empirical_goldens():
    f = G["fdr"]
    assert abs(dto.fdr(*f["args"]) - f["value"]) < f["tol"]
    assert dto.fdr(24, 13, 12, 30, 0.8) == 0.0 and dto.fdr(5, 5, 0, 10, 0.8) == 0.0
    with pytest.raises(dto.DtoPanic, match="Sensitivity must be greater than 0"):
        dto.fdr(1, 1, 1, 1, 0.0)
    e = G["empirical_pvalue"]
    u = e["unpermuted"]
    recs = [dto.OptimizationResultRecord(u["rank1"], u["rank2"], u["set1_len"], u["set2_len"], u["population_size"], u["intersection_size"], u["pvalue"], False)]
    for p, k in zip(e["permuted_pvalues"], (8, 9, 7)):
        recs.append(dto.OptimizationResultRecord(1, 2, 50, 40, 100, k, p, True))
    assert dto.empirical_pvalue(recs) == e["json"]
    assert dto.empirical_pvalue(recs[:1])["empirical_pvalue"] == 1.0
    with pytest.raises(dto.DtoPanic, match="No unpermuted result found"):
        dto.empirical_pvalue(recs[1:])
    # against the oracle on random inputs
    rng = np.random.default_rng(2)
    for _ in range(50):
        b, r, k, N = (int(x) for x in rng.integers(0, 500, size=4))
        assert dto.fdr(b, r, k, N + 1, 0.8) == O.fdr(b, r, k, N + 1, 0.8)


def test_json_matches_serde_pretty():
    fin = capi.FinalResult(24, 13, 24, 13, 30, 12, 0.15632183908046102, 0.8, 0.0)
    s = final_json(fin)
    want = ('{\n  "empirical_pvalue": 0.8,\n  "fdr": 0.0,\n  "population_size": 30,\n  "rank1": 24,\n  "rank2": 13,\n  "set1_len": 24,\n'
            '  "set2_len": 13,\n  "unpermuted_intersection_size": 12,\n  "unpermuted_pvalue": 0.15632183908046102\n}')
    assert s == want  # README.md:174-184 byte for byte
    for val, txt in ((1.0, "1.0"), (1e-7, "1e-7"), (1.5e-10, "1.5e-10"), (0.00001, "0.00001"), (123456.0, "123456.0"), (1e16, "1e16"),
                     (1.2345e22, "1.2345e22"), (0.1, "0.1"), (5e-324, "5e-324"), (0.3333333333333333, "0.3333333333333333")):
        fin.unpermuted_pvalue = val
        assert f'"unpermuted_pvalue": {txt}\n' in final_json(fin), (val, final_json(fin))
        assert json.loads(final_json(fin))["unpermuted_pvalue"] == val


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(H.ROOT, "include", "dto_b200.h")).read()
    declared = set(re.findall(r"\b(dto_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert C.sizeof(capi.Record) == 40
    assert b"sm_100a" in L.dto_b200_version()


def test_no_cpu_fallback_without_gpu():
    """Without a usable CUDA device every compute entry point fails loudly (no oracle / CPU path behind the API)."""
    n = C.c_int(-1)
    rc = capi.lib().dto_b200_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dto.DtoError, match="no CUDA device|CUDA"):
        dto.Engine(0)
    l = dto.RankedFeatureList.from_(["a", "b"], [1, 2])
    with pytest.raises(dto.DtoError):
        dto.run_single_node([dto.Task(0, False)], l, l, 2, 1)
    src = open(os.path.join(H.ROOT, "dual_threshold_optimization_b200", "_capi.py")).read()
    for mod in os.listdir(os.path.join(H.ROOT, "dual_threshold_optimization_b200")):
        if mod.endswith(".py"):
            assert "oracle" not in open(os.path.join(H.ROOT, "dual_threshold_optimization_b200", mod)).read().replace("no oracle", ""), mod
    assert "oracle" not in src


def test_c_abi_argument_validation():
    """Entry points validate their arguments before touching CUDA and report through dto_b200_last_error()."""
    L = capi.lib()
    assert L.dto_b200_fdr(1, 1, 1, 1, 0.8, None) == capi.ERR_INVALID
    assert b"null" in L.dto_b200_last_error()
    assert L.dto_b200_ranked_list_from(None, None, 0, None) == capi.ERR_INVALID
    assert L.dto_b200_create(None, 0) == capi.ERR_INVALID
    assert L.dto_b200_set_problem(None, None, 0, None, 0, None, 0, None, 0, None, 0) == capi.ERR_INVALID
    assert L.dto_b200_run_unpermuted(None, None) == capi.ERR_INVALID
    assert L.dto_b200_device_count(None) == capi.ERR_INVALID
    assert L.dto_b200_empirical_pvalue(None, 3, None) == capi.ERR_INVALID
    assert L.dto_b200_run_single_node(None, None, 1, None, 1, None, 0, 0, None) == capi.ERR_INVALID
    assert L.dto_b200_run_pairs(None, None, None, 1, 1, None, 0, 0, None) == capi.ERR_INVALID
    assert L.dto_b200_run_pairs(None, None, None, 0, 1, None, 0, 0, None) == capi.OK  # zero pairs: nothing to do
    assert L.dto_b200_compute_population_size(None, None, None, None) == capi.ERR_INVALID
    h = C.c_void_p()
    assert L.dto_b200_read_ranked_list_csv(b"/nonexistent/file.csv", C.byref(h)) == capi.ERR_IO
    L.dto_b200_destroy(None)  # no-op
    L.dto_b200_ranked_list_free(None)
    L.dto_b200_feature_list_free(None)


def test_empirical_pvalue_compares_ulp_close_values_on_the_host():
    """empirical_pvalue.rs:160-165 counts permuted p <= unpermuted p with values of ONE evaluator.  Permuted records
    normally carry device-evaluated p (CUDA exp, last-ulp noise): simulate that noise on records whose true p equals the
    unpermuted one (same cell), mirrors it ((K, n, k) -> (n, K, k): mathematically equal, different bits), or sits well
    away.  The count must be the oracle's, whatever the noise did."""
    N = 30
    lf = O.ln_factorial_table(N)
    cells = [(24, 13, 12), (13, 24, 12), (20, 15, 11), (15, 20, 11), (24, 13, 11), (10, 10, 6), (12, 9, 7), (9, 12, 7)]
    rng = np.random.default_rng(0)
    for (Ku, nu, ku) in cells[:4]:
        pu = O.hypergeometric_pvalue_cached(lf, N, Ku, nu, ku)
        recs = np.zeros(1 + 400, dtype=capi.RECORD_DTYPE)
        recs[0] = (Ku, nu, Ku, nu, ku, capi.FLAG_HOST_PVALUE, N, pu)
        want = 0
        for t in range(1, recs.size):
            K, n, k = cells[int(rng.integers(0, len(cells)))]
            p = O.hypergeometric_pvalue_cached(lf, N, K, n, k)
            want += p <= pu
            noisy = p
            for _ in range(int(rng.integers(0, 4))):
                noisy = np.nextafter(noisy, [0.0, 1.0][int(rng.integers(0, 2))])
            recs[t] = (K, n, K, n, k, capi.FLAG_PERMUTED, N, noisy)
        fin = empirical_pvalue_struct(recs)
        assert fin.empirical_pvalue == want / 400.0
        assert fin.unpermuted_pvalue == pu


def test_oracle_batch_equals_per_task_oracle():
    ids1, r1, ids2, r2 = H.synthetic_pair(400, 13, None)
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    p1, p2 = H.perms(400, 24, 1), H.perms(400, 24, 2)
    b = O.best_batch(o1, o2, 400, p1, p2, num_threads=3)
    for t in range(24):
        ob = O.grid_int(o1, o2, 400, None, p1[t], p2[t], want_overlap=False, want_p=False).best
        assert all(int(b[t][f]) == int(ob[f]) for f in ("rank1", "rank2", "set1_len", "set2_len", "intersection_size"))
        assert float(b[t]["pvalue"]) == float(ob["pvalue"])


def test_limits_query_and_feature_list_reader(tmp_path):
    lim = (C.c_uint64 * 4)()
    capi.check(capi.lib().dto_b200_get_limits(C.byref(lim)))
    assert list(lim)[:3] == [65534, 2048, 1 << 27]
    p = tmp_path / "bg.txt"
    p.write_text(" a \r\nb\n\n c\n")
    bg = dto.read_feature_list_from_file(str(p))
    assert bg.ids() == ["a", "b", "", "c"]  # every line counts, blank ones too (read_feature_list_from_file.rs:48-52)


def test_host_evaluator_is_pinned_to_the_reference_goldens_and_the_oracle():
    """The product's own host evaluation of hypergeometric_pvalue (what settles tie sets and the epilogue's `<=`): bit-exact
    on the reference-held goldens (hypergeometric_pvalue.rs:27-31, :56-87) and on random quadruples against the oracle."""
    L = capi.lib()

    def host_p(N, K, n, k):
        out = C.c_double()
        capi.check(L.dto_b200_hypergeometric_pvalue_host(N, K, n, k, C.byref(out)))
        return out.value

    assert host_p(1000, 50, 60, 10) == 0.00044068070222441115
    assert host_p(6060, 5808, 154, 153) == 0.010413637619010246
    assert host_p(10, 1, 1, 0) == 1.0 and host_p(3, 1, 3, 1) == 1.0
    rng = np.random.default_rng(5)
    for N in (30, 3000, 20000):
        lf = O.ln_factorial_table(N)
        for _ in range(300):
            K, n = int(rng.integers(0, N + 1)), int(rng.integers(0, N + 1))
            k = int(rng.integers(0, min(K, n) + 1))
            assert host_p(N, K, n, k) == O.hypergeometric_pvalue_cached(lf, N, K, n, k), (N, K, n, k)
    with pytest.raises(dto.DtoPanic):
        host_p(10, 11, 3, 1)


def test_cli_without_gpu_fails_loudly_and_prints_version():
    """The CLI drop-in (src/main.rs flags): --version / --help work anywhere; a run without a usable GPU exits non-zero with
    the library's message instead of computing anything on the CPU."""
    import subprocess

    import torch

    cli = capi.CLI_PATH
    v = subprocess.run([cli, "--version"], capture_output=True, text=True, timeout=60)
    assert v.returncode == 0 and "dual_threshold_optimization 2.0.1" in v.stdout and "sm_100a" in v.stdout
    h = subprocess.run([cli, "--help"], capture_output=True, text=True, timeout=60)
    assert h.returncode == 0 and "--ranked-list1" in h.stdout and "--permutations" in h.stdout and "--multi-node" in h.stdout
    miss = subprocess.run([cli, "-1", "a.csv"], capture_output=True, text=True, timeout=60)
    assert miss.returncode == 2 and "--ranked-list2" in miss.stderr
    if torch.cuda.is_available():
        return
    td = os.path.join(H.GOLDEN, "test_data")
    r = subprocess.run([cli, "-1", f"{td}/ranklist1.csv", "-2", f"{td}/ranklist2.csv", "-p", "10"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and not r.stdout.strip()
    assert "no CUDA device" in r.stderr or "CUDA" in r.stderr
    assert "Ranked list 1:" in r.stderr and "Permutations: 10" in r.stderr  # the run-information lines come first, as in main.rs:78-88
