"""Pins the CPU oracle against the reference's own fixtures and known answers."""
import math

import numpy as np
import pytest
import scipy.special as sp

from oracle.oracle import OracleError, dense_to_csc


def rel(a, b):
    return abs(a - b) / abs(b)


# ------------------------------------------------------------------ deterministic libm vs libm/scipy
def test_log_exp_within_one_ulp(oracle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-3, 1e3, 200000), 10.0 ** rng.uniform(-300, 300, 100000),
                        np.arange(1, 5000, dtype=np.float64), [1.0, 2.0, 0.5, 1e-310]])
    got, want = oracle.log(x), np.log(x)
    ulp = np.spacing(np.abs(want))
    assert np.all(np.abs(got - want) <= ulp)
    assert oracle.log(np.array([1.0]))[0] == 0.0
    e = np.concatenate([rng.uniform(-745, 709, 300000), -10.0 ** rng.uniform(-20, 2.8, 100000), [0.0, -0.0]])
    got, want = oracle.exp(e), np.exp(e)
    assert np.all(np.abs(got - want) <= np.spacing(want))
    assert oracle.exp(np.array([0.0]))[0] == 1.0
    assert oracle.exp(np.array([-800.0]))[0] == 0.0


def test_lgamma_matches_scipy(oracle):
    rng = np.random.default_rng(2)
    x = np.concatenate([np.arange(1, 20001, dtype=np.float64), rng.uniform(0.01, 10, 50000),
                        rng.uniform(10, 1e5, 100000), 10.0 ** rng.uniform(1, 12, 100000),
                        np.arange(0, 3000) + 0.5, [4934720.0, 4934721.0, 94906265.0, 94906266.0, 1e17, 2e17]])
    got, want = oracle.lgamma(x), sp.gammaln(x)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 5e-15, (err.max(), x[err.argmax()])
    # exact integers R returns: lgamma(1) = lgamma(2) = 0
    assert oracle.lgamma(np.array([1.0, 2.0])).tolist() == [0.0, 0.0]
    assert math.isinf(oracle.lgamma(np.array([0.0]))[0])


# ------------------------------------------------------------------ tests/testthat/test-matrixSums.R:5-76
def _rowsum(mat, lab, n):
    out = np.zeros((n, mat.shape[1]), dtype=mat.dtype)
    for i, g in enumerate(lab):
        out[g - 1] += mat[i]
    return out


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_matrixSums_known_answers(oracle, dtype):
    mat = np.asfortranarray(np.tile(np.arange(1, 6), 20).reshape(10, 10, order="F").astype(dtype))  # matrix(seq(5),10,10)
    lab = np.repeat([1, 2], 5).astype(np.int32)
    want_r = _rowsum(mat, lab, 2)
    want_c = _rowsum(mat.T, lab, 2).T
    assert np.array_equal(oracle.rowSumByGroup(mat, lab, 2), want_r)
    assert np.array_equal(oracle.colSumByGroup(mat, lab, 2), want_c)
    # Change variants: start from the sums under `lab`, move to lab2
    lab2 = np.array([1, 2, 1, 2, 1, 2, 1, 2, 1, 2], dtype=np.int32)
    px = np.asfortranarray(want_r.copy())
    r = oracle.rowSumByGroupChange(mat, px, lab2, lab, 2)
    assert r is px and np.array_equal(px, _rowsum(mat, lab2, 2))
    px = np.asfortranarray(want_c.copy())
    oracle.colSumByGroupChange(mat, px, lab2, lab, 2)
    assert np.array_equal(px, _rowsum(mat.T, lab2, 2).T)
    # error contract (:14-15,24-28,37-38,47-51,61-62,70-71)
    with pytest.raises(OracleError):
        oracle.rowSumByGroup(mat, lab, -1)  # not a factor
    with pytest.raises(OracleError):
        oracle.rowSumByGroup(mat, lab[:9], 2)  # wrong length
    with pytest.raises(OracleError):
        oracle.colSumByGroup(mat, lab[:9], 2)
    with pytest.raises(OracleError):
        oracle.rowSumByGroupChange(mat, np.asfortranarray(want_r.copy()), lab2, lab, 2, 3)  # level mismatch
    with pytest.raises(OracleError):
        oracle.colSumByGroupChange(mat, np.asfortranarray(np.zeros((3, 2), dtype=dtype)), lab2, lab, 2)  # px dims


def test_sparse_aggregation_equals_dense(oracle, gold_cg):
    counts = gold_cg["counts"]
    csc = dense_to_csc(counts)
    z, y = gold_cg["z"], gold_cg["y"]
    d = counts.astype(np.float64)
    assert np.array_equal(oracle.rowSumByGroupSparse(csc, y, 10), oracle.rowSumByGroup(d, y, 10))
    assert np.array_equal(oracle.colSumByGroupSparse(csc, z, 5), oracle.colSumByGroup(d, z, 5))
    rng = np.random.default_rng(0)
    z2 = rng.integers(1, 6, z.size).astype(np.int32)
    y2 = rng.integers(1, 11, y.size).astype(np.int32)
    px = oracle.colSumByGroupSparse(csc, z, 5)
    oracle.colSumByGroupChangeSparse(csc, px, z2, z, 5)
    assert np.array_equal(px, oracle.colSumByGroupSparse(csc, z2, 5))
    px = oracle.rowSumByGroupSparse(csc, y, 10)
    oracle.rowSumByGroupChangeSparse(csc, px, y2, y, 10)
    assert np.array_equal(px, oracle.rowSumByGroupSparse(csc, y2, 10))
    for bad in (lambda: oracle.colSumByGroupSparse(csc, z[:-1], 5), lambda: oracle.colSumByGroupSparse(csc, z, 4),
                lambda: oracle.rowSumByGroupSparse(csc, y, 101), lambda: oracle.rowSumByGroupSparse(csc, y * 0, 10)):
        with pytest.raises(OracleError):
            bad()


# ------------------------------------------------------------------ log-likelihood goldens
def _cg_ll(oracle, counts, s, z, y, K, L, a=1.0, b=1.0, d=1.0, g=1.0):
    p = oracle.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, K, L)
    return oracle.cCGCalcLL(K, L, p["mCPByS"], p["nTSByCP"], p["nByG"], p["nByTS"], p["nGByTS"], p["nS"], p["nG"],
                            a, b, d, g), p


def test_cCGCalcLL_golden(oracle, oracle_ld, gold_cg):
    # long-double accumulation (x86 R) reproduces the stored value BIT FOR BIT; plain-double mode to a few ulp
    ll, _ = _cg_ll(oracle_ld, gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"], 5, 10)
    assert ll == float(gold_cg["finalLogLik"])  # data/celdaCGMod.rda: -1215541.0958389007
    ll, _ = _cg_ll(oracle, gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"], 5, 10)
    assert rel(ll, float(gold_cg["finalLogLik"])) < 5e-15
    assert float(gold_cg["finalLogLik"]) == -1215541.0958389007
    assert float(gold_cg["completeLogLik"].max()) == float(gold_cg["finalLogLik"])


def test_cCCalcLL_golden(oracle, oracle_ld, gold_c):
    for o in (oracle, oracle_ld):
        p = o.cCDecomposeCounts(dense_to_csc(gold_c["counts"]), gold_c["s_mod"], gold_c["z"], 10)
        ll = o.cCCalcLL(p["mCPByS"], p["nGByCP"], 10, p["nS"], p["nG"], 1.0, 1.0)
        assert float(gold_c["finalLogLik"]) == -1273077.9330206949  # data/celdaCMod.rda
        if o is oracle_ld:
            assert ll == -1273077.9330206949  # bit-identical in x86-R accumulation mode
        assert rel(ll, -1273077.9330206949) < 5e-15


def test_cGCalcLL_golden(oracle, oracle_ld, gold_g):
    for o in (oracle, oracle_ld):
        p = o.cGDecomposeCounts(dense_to_csc(gold_g["counts"]), gold_g["y"], 10)
        ll = o.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], p["nM"], p["nG"], 10, 1.0, 1.0, 1.0)
        if o is oracle_ld:
            assert ll == -290669.0461321392  # data/celdaGMod.rda, bit-identical
        assert rel(ll, -290669.0461321392) < 5e-15


def test_grid_search_models_golden(oracle, oracle_ld, gold_cg, gold_grid):
    """9 celda_CG models (K 4-6 x L 9-11) of data/celdaCGGridSearchRes.rda on the celdaCGSim counts."""
    nexact = 0
    for m in range(9):
        K, L = int(gold_grid["K"][m]), int(gold_grid["L"][m])
        ll_ld, _ = _cg_ll(oracle_ld, gold_cg["counts"], gold_grid["s"], gold_grid["z"][m], gold_grid["y"][m], K, L)
        nexact += ll_ld == gold_grid["finalLogLik"][m]
        assert rel(ll_ld, gold_grid["finalLogLik"][m]) < 4e-16
        ll, _ = _cg_ll(oracle, gold_cg["counts"], gold_grid["s"], gold_grid["z"][m], gold_grid["y"][m], K, L)
        assert rel(ll, gold_grid["finalLogLik"][m]) < 5e-15
        assert rel(ll, gold_grid["runParams_logLikelihood"][m]) < 1e-12
    assert nexact >= 7, nexact  # an older R/libm produced these; most still reproduce bit for bit


def test_perplexity_statistical_pin(oracle, gold_cg, gold_grid):
    """Stored perplexities are of RESAMPLED counts (45.20-45.29 for K=5,L=10): statistical pin only."""
    m = [i for i in range(9) if gold_grid["K"][i] == 5 and gold_grid["L"][i] == 10][0]
    csc = dense_to_csc(gold_cg["counts"])
    p = oracle.cCGDecomposeCounts(csc, gold_grid["s"], gold_grid["z"][m], gold_grid["y"][m], 5, 10)
    px = oracle.perplexity_CG(csc, p["mCPByS"], p["nTSByCP"], p["nByG"], gold_grid["y"][m], 5, 10, gold_grid["s"],
                              1.0, 1.0, 1.0)
    stored = gold_grid["perplexity"][m]
    assert stored.min() - 0.2 < px < stored.max() + 0.2
    assert abs(px - 45.2538) < 1e-3


# ------------------------------------------------------------------ Delta-probs == Delta-LL (SURVEY.md section 8c iv)
def test_z_probs_equal_LL_differences(oracle, gold_cg):
    counts, s, z, y = gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"]
    K, L = 5, 10
    base, p = _cg_ll(oracle, counts, s, z, y, K, L)
    nM = counts.shape[1]
    ix = np.arange(1, nM + 1, dtype=np.int32)
    out = oracle.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0,
                                  ix, np.zeros(nM), doSample=False)
    assert np.array_equal(out["z"], z) and np.array_equal(out["nGByCP"], p["nTSByCP"])
    for cell in (0, 17, 200, 424):
        lls = []
        for j in range(1, K + 1):
            z2 = z.copy()
            z2[cell] = j
            lls.append(_cg_ll(oracle, counts, s, z2, y, K, L)[0])
        d_ll = np.array(lls) - lls[0]
        d_pr = out["probs"][:, cell] - out["probs"][0, cell]
        assert np.abs(d_ll - d_pr).max() < 5e-9  # LL round-off on |LL| ~ 1.2e6


def test_y_probs_equal_LL_differences_and_simple_variant(oracle, gold_cg):
    counts, s, z, y = gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"]
    K, L = 5, 10
    base, p = _cg_ll(oracle, counts, s, z, y, K, L)
    nG = counts.shape[0]
    ix = np.arange(1, nG + 1, dtype=np.int32)
    lgbeta = oracle.lgamma(np.arange(0, p["nCP"].max() + 1) + 1.0)
    lggamma = oracle.lgamma(np.arange(0, nG + L + 1) + 1.0)
    lgdelta = np.concatenate([[np.nan], oracle.lgamma(np.arange(1, nG + L + 1) * 1.0)])
    out_t = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                    1.0, ix, np.zeros(nG), lgbeta, lggamma, lgdelta, doSample=False)
    out = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                  1.0, ix, np.zeros(nG), doSample=False)
    assert np.array_equal(out["probs"], out_t["probs"])  # table look-up == direct evaluation, bit for bit
    for gene in (0, 13, 57, 99):
        lls = []
        for l in range(1, L + 1):
            y2 = y.copy()
            y2[gene] = l
            lls.append(_cg_ll(oracle, counts, s, z, y2, K, L)[0])
        d_ll = np.array(lls) - lls[0]
        d_pr = out["probs"][:, gene] - out["probs"][0, gene]
        assert np.abs(d_ll - d_pr).max() < 5e-9
        simple = oracle.cG_calcGibbsProbY_Simple(p["nGByCP"], p["nGByTS"], p["nTSByCP"], p["nByTS"], p["nByG"], y, L,
                                                 gene + 1, 1.0, 1.0, 1.0)
        d_s = simple - simple[0]
        assert np.abs(d_s - d_pr).max() < 5e-9  # src/cG_calcGibbsProbY.cpp:5-6 "sanity check" variants agree


# ------------------------------------------------------------------ sampler semantics
def test_sampler_semantics(oracle):
    ll = np.log(np.array([0.1, 0.6, 0.3]))
    # sorted descending: 0.6 (label 2), 0.3 (label 3), 0.1 (label 1); cum = .6, .9
    assert oracle.sampleLl(ll, 0.59)[0] == 2
    assert oracle.sampleLl(ll, 0.61)[0] == 3
    assert oracle.sampleLl(ll, 0.95)[0] == 1
    assert oracle.sampleLl(np.array([-1e4, 0.0, -1e4]), 0.999999)[0] == 2  # underflowed entries never drawn
    assert oracle.sampleLl(np.array([3.0]), 0.5)[0] == 1


def _one_iteration(o, csc, s, z, y, K, L, ixy, uy, ixz, uz):
    nr, nc = csc[0], csc[1]
    p = o.cCGDecomposeCounts(csc, s, z, y, K, L)
    oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nr, 1.0, 1.0, 1.0,
                            ixy, uy)
    nTSByC = o.rowSumByGroupChangeSparse(csc, p["nTSByC"].copy(order="F"), oy["y"], y, L)
    oz = o.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, s, K, L, nc, 1.0, 1.0, ixz, uz)
    return p, oy, nTSByC, oz


def test_sweeps_preserve_invariants_and_ld_mode_agrees(oracle, oracle_ld, gold_cg):
    """Tables updated incrementally by a sampled sweep equal tables recomputed from scratch; the long-double
    accumulation mode (x86 R) draws the same labels as the default double mode on the fixture."""
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    rng = np.random.default_rng(7)
    nG, nM = counts.shape
    z = rng.integers(1, K + 1, nM).astype(np.int32)
    y = rng.integers(1, L + 1, nG).astype(np.int32)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    csc = dense_to_csc(counts)
    res = []
    for o in (oracle, oracle_ld):
        p, oy, nTSByC, oz = _one_iteration(o, csc, s, z, y, K, L, ixy, uy, ixz, uz)
        p2 = o.cCGDecomposeCounts(csc, s, oz["z"], oy["y"], K, L)
        assert np.array_equal(p2["nTSByC"], nTSByC)
        assert np.array_equal(p2["nTSByCP"], oz["nGByCP"])
        assert np.array_equal(p2["mCPByS"], oz["mCPByS"])
        assert np.array_equal(p2["nCP"], oz["nCP"])
        assert np.array_equal(p2["nByTS"], oy["nByTS"]) and np.array_equal(p2["nGByTS"], oy["nGByTS"])
        assert (oz["z"] != z).sum() > 50 and (oy["y"] != y).sum() > 10
        res.append((oy["y"], oz["z"], oy["probs"], oz["probs"]))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.allclose(res[0][3], res[1][3], rtol=1e-10, atol=0)  # sums of ~1e5-magnitude terms cancel to ~1e3


def test_em_step(oracle, gold_c):
    counts, s, z = gold_c["counts"], gold_c["s_mod"], gold_c["sim_z"]
    K = 10
    csc = dense_to_csc(counts)
    p = oracle.cCDecomposeCounts(csc, s, z, K)
    a = oracle.cCCalcEMProbZ(csc, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"], p["nM"], 1.0, 1.0)
    b = oracle.cCCalcEMProbZ(counts.astype(np.float64), p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K,
                             p["nG"], p["nM"], 1.0, 1.0)
    assert np.array_equal(a["z"], b["z"])
    assert np.allclose(a["probs"], b["probs"], rtol=1e-12)
    p2 = oracle.cCDecomposeCounts(csc, s, a["z"], K)
    assert np.array_equal(p2["nGByCP"], a["nGByCP"]) and np.array_equal(p2["mCPByS"], a["mCPByS"])
    # numpy cross-check of the formula (A.2)
    phi = np.log((p["nGByCP"] + 1.0) / (p["nGByCP"].sum(0) + counts.shape[0]))
    theta = np.log((p["mCPByS"] + 1.0) / (p["mCPByS"].sum(0) + K))
    want = phi.T @ counts + theta[:, s - 1]
    assert np.allclose(a["probs"], want, rtol=1e-11)
    assert np.array_equal(a["z"], want.argmax(0) + 1)


def test_walker_alias_branch_of_sample_int(oracle):
    """More than 200 likely categories: sample.int takes Walker's alias method (R random.c walker_ProbSampleReplace).
    The restatement must (1) return a category whose alias cell contains u * n, checked against an independent numpy
    construction of the tables, and (2) sample every category with its probability (exact: cell masses sum to p)."""
    rng = np.random.default_rng(3)
    n = 260
    ll = np.log(rng.dirichlet(np.full(n, 40.0)))           # diffuse: n * p > 0.1 for every category
    p = np.exp(ll - ll.max())
    p = p / p.sum()
    p = p / p[p > 0].sum()
    assert (n * p > 0.1).sum() > 200

    def tables(p):
        q = p * n
        HL, a = np.empty(n, dtype=np.int64), np.arange(n)
        H, Lp = -1, n
        for i in range(n):
            if q[i] < 1.0:
                H += 1
                HL[H] = i
            else:
                Lp -= 1
                HL[Lp] = i
        if H >= 0 and Lp < n:
            for k in range(n - 1):
                i, j = HL[k], HL[Lp]
                a[i] = j
                q[j] += q[i] - 1
                if q[j] < 1.0:
                    Lp += 1
                if Lp >= n:
                    break
        return q, a

    q, a = tables(p.copy())
    mass = np.zeros(n)                                      # probability of each category under the alias tables
    for k in range(n):
        mass[k] += min(q[k], 1.0) / n
        mass[a[k]] += max(1.0 - q[k], 0.0) / n
    assert np.allclose(mass, p, atol=1e-12)
    for u in np.concatenate([rng.random(400), [1e-300, 0.5, 1 - 1e-16]]):
        got, _ = oracle.sampleLl(ll, u)
        rU = u * n
        k = min(int(rU), n - 1)
        want = (k if rU < q[k] + k else a[k]) + 1
        assert got == want


def test_committed_fixtures_equal_a_fresh_extraction_from_the_reference(tmp_path):
    """tests/golden/*.npz are what tests/golden/make_golden.py extracts from /root/reference/data/*.rda today (skipped on
    a box without the reference): the fixtures the parity tests stand on were not edited by hand."""
    import importlib.util
    import os
    if not os.path.isdir("/root/reference/data"):
        pytest.skip("the reference is not present on this box")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.main(str(tmp_path))
    for name in ("celdaCG", "celdaC", "celdaG", "celdaCGGrid"):
        fresh, kept = np.load(tmp_path / f"{name}.npz"), np.load(os.path.join(here, f"{name}.npz"))
        assert sorted(fresh.files) == sorted(kept.files), name
        for k in fresh.files:
            assert np.array_equal(fresh[k], kept[k]), (name, k)
