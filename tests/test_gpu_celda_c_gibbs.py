"""GPU parity for the celda_C Gibbs z sweep on raw counts (a8, celda_C form): every gene is scanned per
(cell, population), zeros included, exactly like .cCCalcGibbsProbZ (R/celda_C.R:664-712)."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _oracle_c_iteration(o, O, counts, s, z, K, ix, u, alpha, beta):
    p = o.cCDecomposeCounts(O, s, z, K)
    oz = o.cCCalcGibbsProbZ(counts.astype(np.float64), p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K,
                            counts.shape[0], counts.shape[1], alpha, beta, ix, u)
    ll = o.cCCalcLL(oz["mCPByS"], oz["nGByCP"], K, p["nS"], counts.shape[0], alpha, beta)
    return oz, ll


@pytest.mark.parametrize("start", ["random", "fitted"])
def test_celda_C_gibbs_sweeps_bit_exact_fixture(oracle, gold_c, start):
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    nG, nM = counts.shape
    rng = np.random.default_rng(17)
    z = rng.integers(1, K + 1, nM).astype(np.int32) if start == "random" else gold_c["z"].copy()
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, model=cb.MODEL_C)
    ch.set_state(z=z)
    for it in range(3):
        ix, u = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
        ref, ll_ref = _oracle_c_iteration(oracle, O, counts, s, z, K, ix, u, 1.0, 1.0)
        ll = ch.iterate(None, None, ix, u)
        zg, _ = ch.get_state()
        assert np.array_equal(zg, ref["z"]), (it, (zg != ref["z"]).sum())
        t = ch.tables()
        assert np.array_equal(t["nGByCP"], ref["nGByCP"]) and np.array_equal(t["mCPByS"], ref["mCPByS"])
        assert np.array_equal(t["nCP"], ref["nCP"])
        pz, _ = ch.probs()
        assert np.array_equal(pz, ref["probs"])
        assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        z = ref["z"]
    ch.close()


def test_celda_C_gibbs_sparse_midsize_two_chunks(oracle):
    """Sparse counts, K = 40 (> 32: two population chunks), 3 samples, non-default priors, doSample = FALSE too."""
    rng = np.random.default_rng(23)
    G, C, K = 700, 1500, 40
    z_true = rng.integers(1, K + 1, C)
    z_true[:K] = np.arange(1, K + 1)
    phi = rng.dirichlet(np.ones(G) * 0.05, K)
    counts = np.zeros((G, C), dtype=np.int32)
    for c in range(C):
        counts[:, c] = rng.multinomial(rng.integers(40, 300), phi[z_true[c] - 1])
    counts[counts.sum(1) == 0, 0] = 1
    assert (counts == 0).mean() > 0.8
    s = rng.integers(1, 4, C).astype(np.int32)
    s[:3] = [1, 2, 3]
    z = z_true.astype(np.int32).copy()
    flip = rng.random(C) < 0.3
    z[flip] = rng.integers(1, K + 1, int(flip.sum()))
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    alpha, beta = 0.7, 0.3
    ch = cb.Chain(A, s, K=K, alpha=alpha, beta=beta, model=cb.MODEL_C)
    ch.set_state(z=z)
    for it in range(2):
        ix, u = rng.permutation(C).astype(np.int32) + 1, rng.random(C)
        ref, ll_ref = _oracle_c_iteration(oracle, O, counts, s, z, K, ix, u, alpha, beta)
        ll = ch.iterate(None, None, ix, u)
        zg, _ = ch.get_state()
        assert np.array_equal(zg, ref["z"]), (it, (zg != ref["z"]).sum())
        t = ch.tables()
        assert np.array_equal(t["nGByCP"], ref["nGByCP"]) and np.array_equal(t["mCPByS"], ref["mCPByS"])
        pz, _ = ch.probs()
        assert np.array_equal(pz, ref["probs"])
        assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        z = ref["z"]
    # doSample = FALSE: probabilities only, state untouched (clusterProbability path)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    ix, u = np.arange(1, C + 1, dtype=np.int32), np.zeros(C)
    ref = oracle.cCCalcGibbsProbZ(counts.astype(np.float64), p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, G, C,
                                  alpha, beta, ix, u, doSample=False)
    ch.sweep_z_gibbs(ix, u, doSample=False)
    pz, _ = ch.probs()
    assert np.array_equal(pz, ref["probs"])
    zg, _ = ch.get_state()
    assert np.array_equal(zg, z)
    ch.close()


def test_celda_C_gibbs_philox_replay(oracle, gold_c):
    """Philox mode: the device draws order and uniforms; replaying them through the oracle gives the same chain."""
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    nG, nM = counts.shape
    z = np.random.default_rng(2).integers(1, K + 1, nM).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, model=cb.MODEL_C)
    ch.set_state(z=z)
    ch.seed(12345, 0)
    for it in range(2):
        ll = ch.iterate(None, None, None, None)
        ix, u = ch.get_draws(1)
        ref, ll_ref = _oracle_c_iteration(oracle, O, counts, s, z, K, ix, u, 1.0, 1.0)
        zg, _ = ch.get_state()
        assert np.array_equal(zg, ref["z"])
        assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        z = ref["z"]
    ch.close()


def test_model_drivers_celda_C_gibbs_and_celda_G(oracle, gold_c, gold_g):
    """celda_C(algorithm = "Gibbs") and celda_G chain loops: the reported best log-likelihood is the oracle's
    log-likelihood of the returned labels, and it improves on the random start."""
    from celda_b200 import model
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    res = model.celda_C(A, s, K, algorithm="Gibbs", maxIter=12, nchains=2, seed=7, zInitialize="random")
    m = res["model"]
    p = oracle.cCDecomposeCounts(O, s, m["z"], K)
    want = oracle.cCCalcLL(p["mCPByS"], p["nGByCP"], K, p["nS"], p["nG"], 1.0, 1.0)
    assert abs(m["finalLogLik"] - want) <= 1e-9 * abs(want)
    assert m["finalLogLik"] > m["completeLogLik"][0]
    assert m["finalLogLik"] == max(c["finalLogLik"] for c in res["chains"])

    counts, L = gold_g["counts"], 10
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    res = model.celda_G(A, L, maxIter=8, nchains=2, seed=7, yInitialize="random")
    m = res["model"]
    p = oracle.cGDecomposeCounts(O, m["y"], L)
    want = oracle.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], counts.shape[1], counts.shape[0], L,
                           1.0, 1.0, 1.0)
    assert abs(m["finalLogLik"] - want) <= 1e-9 * abs(want)
    assert m["finalLogLik"] > m["completeLogLik"][0]
