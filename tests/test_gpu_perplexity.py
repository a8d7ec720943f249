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
