"""GPU parity for perplexity (a14) and the celda_G chain's decompose / log-likelihood."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def test_perplexity_CG(oracle, gold_cg, gold_grid):
    counts = gold_cg["counts"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    m = [i for i in range(9) if gold_grid["K"][i] == 5 and gold_grid["L"][i] == 10][0]
    z, y, s = gold_grid["z"][m], gold_grid["y"][m], gold_grid["s"]
    ch = cb.Chain(A, s, K=5, L=10)
    ch.set_state(z, y)
    got = ch.perplexity()
    p = oracle.cCGDecomposeCounts(O, s, z, y, 5, 10)
    want = oracle.perplexity_CG(O, p["mCPByS"], p["nTSByCP"], p["nByG"], y, 5, 10, s, 1.0, 1.0, 1.0)
    assert abs(got - want) <= 1e-9 * want
    assert abs(got - 45.2538) < 1e-3  # SURVEY.md 8c statistical pin (stored resampled values 45.20-45.29)
    assert isinstance(got, float)  # tests/testthat/test-celda_CG.R:189-191 is.numeric(perplexity(model))
    ch.close()
    ch = cb.Chain(A, s, K=5, L=10, alpha=0.5, beta=0.3, delta=2.0, gamma=1.0)
    ch.set_state(z, y)
    want = oracle.perplexity_CG(O, p["mCPByS"], p["nTSByCP"], p["nByG"], y, 5, 10, s, 0.5, 0.3, 2.0)
    assert abs(ch.perplexity() - want) <= 1e-9 * want
    ch.close()


def test_perplexity_C(oracle, gold_c):
    counts, s, z = gold_c["counts"], gold_c["s_mod"], gold_c["z"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=10, model=cb.MODEL_C)
    ch.set_state(z=z)
    p = oracle.cCDecomposeCounts(O, s, z, 10)
    want = oracle.perplexity_C(O, p["mCPByS"], p["nGByCP"], 10, s, 1.0, 1.0)
    assert abs(ch.perplexity() - want) <= 1e-9 * want
    ch.close()


def test_celda_G_chain_ll_and_perplexity(oracle, gold_g):
    counts, y = gold_g["counts"], gold_g["y"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, None, L=10, model=cb.MODEL_G)
    ch.set_state(y=y)
    t = ch.tables()
    p = oracle.cGDecomposeCounts(O, y, 10)
    for k in ("nTSByC", "nByG", "nByTS", "nGByTS"):
        assert np.array_equal(t[k], p[k]), k
    ll = ch.loglik()
    assert abs(ll - (-290669.0461321392)) <= 1e-12 * abs(ll)  # data/celdaGMod.rda finalLogLik
    want = oracle.perplexity_G(O, p["nTSByC"], p["nByG"], y, 10, 1.0, 1.0)
    assert abs(ch.perplexity() - want) <= 1e-9 * want
    ch.close()


def test_perplexity_on_new_counts_and_resampling(oracle, gold_cg, gold_grid):
    """perplexity(newCounts = ...) (R/perplexity.R:233-325) and .resamplePerplexity(doResampling = TRUE)
    (:449-476, .resampleCountMatrix :764-775) for a celda_CG model."""
    from celda_b200 import model
    counts = gold_cg["counts"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    m = [i for i in range(9) if gold_grid["K"][i] == 5 and gold_grid["L"][i] == 10][0]
    z, y, s = gold_grid["z"][m], gold_grid["y"][m], gold_grid["s"]
    ch = cb.Chain(A, s, K=5, L=10)
    ch.set_state(z, y)
    p = oracle.cCGDecomposeCounts(O, s, z, y, 5, 10)
    # (1) another matrix with its own sparsity pattern
    rng = np.random.default_rng(8)
    other = rng.poisson(counts * 0.7 + 0.05).astype(np.int64)
    other[0, other.sum(0) == 0] = 1
    want = oracle.perplexity_CG(dense_to_csc(other), p["mCPByS"], p["nTSByCP"], p["nByG"], y, 5, 10, s, 1.0, 1.0, 1.0)
    got = ch.perplexity_new(cb.CSC.from_dense(other))
    assert abs(got - want) <= 1e-9 * want
    assert abs(ch.perplexity_new(A) - ch.perplexity()) <= 1e-12 * got
    # (2) resamples: column totals and support preserved, values follow the column proportions; the oracle reproduces
    # the perplexity of the very matrix the device drew
    perp, x = ch.perplexity_resampled(seed=2024, numResample=3, return_counts=True)
    assert perp.shape == (3,) and len(set(perp.tolist())) == 3
    res = np.zeros_like(counts)
    cols = np.repeat(np.arange(A.ncol), np.diff(A.p))
    res[A.i, cols] = x
    assert np.array_equal(res.sum(0), counts.sum(0)) and (res[counts == 0] == 0).all() and (x >= 0).all()
    want = oracle.perplexity_CG(dense_to_csc(res), p["mCPByS"], p["nTSByCP"], p["nByG"], y, 5, 10, s, 1.0, 1.0, 1.0)
    assert abs(perp[2] - want) <= 1e-9 * want
    big = counts >= 30  # z-scores of resampled entries against Binomial(colSum, prop): unit variance, no bias
    N = np.broadcast_to(counts.sum(0), counts.shape)[big].astype(float)
    pr = counts[big] / N
    zs = (res[big] - counts[big]) / np.sqrt(N * pr * (1 - pr))
    assert abs(zs.mean()) < 0.05 and 0.9 < zs.std() < 1.1
    # (3) the stored perplexities of the reference's grid search came from resampled counts: 45.20-45.29 for K=5, L=10
    perp5 = ch.perplexity_resampled(seed=7, numResample=5)
    assert 45.0 < perp5.mean() < 45.5
    same = ch.perplexity_resampled(seed=7, numResample=5)
    assert np.array_equal(perp5, same)  # counter-based generator: reproducible
    pr2, mean2 = model.resamplePerplexity(ch, doResampling=True, numResample=5, seed=7)
    assert np.array_equal(pr2, perp5) and abs(mean2 - perp5.mean()) < 1e-12
    pr1, mean1 = model.resamplePerplexity(ch)
    assert pr1.shape == (1,) and pr1[0] == ch.perplexity() == mean1
    ch.close()
