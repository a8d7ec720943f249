"""Modules with one gene and empty modules in the y sweeps, and lg_delta[0] = NA (R/celda_CG.R:429,
src/cG_calcGibbsProbY.cpp:233-243): with the +1 pseudo-gene of .cCGDecomposeCounts (R/celda_CG.R:916) entry 0 is never
indexed and the sweeps run as usual; a caller that passes nGByTS WITHOUT the pseudo-gene reaches entry 0 for a
singleton module, the probability is NA and the draw fails like sample.int ("NA in probability vector")."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _labels_with_edges(rng, nG, L):
    y = rng.integers(3, L + 1, nG).astype(np.int32)  # modules 1 and 2 start empty
    y[5] = 1                                          # module 1: exactly one gene; module 2 stays empty
    for l in range(3, L + 1):                         # every other module populated
        y[10 + l] = l
    return y


def test_cg_y_sweep_singleton_and_empty_module(oracle, gold_cg):
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(2)
    z = rng.integers(1, K + 1, nM).astype(np.int32)
    y = _labels_with_edges(rng, nG, L)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
    assert p["nGByTS"][0] == 2 and p["nGByTS"][1] == 1  # one gene + pseudo-gene; pseudo-gene only
    for it in range(2):
        ix, u = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
        ref = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                      1.0, ix, u)
        ch.sweep_y(ix, u)
        _, yg = ch.get_state()
        _, py = ch.probs()
        assert np.array_equal(yg, ref["y"]) and np.array_equal(py, ref["probs"])
        assert np.isfinite(py).all()
        y = ref["y"]
        p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
    ch.close()


def test_g_y_sweep_singleton_and_empty_module(oracle, gold_g):
    counts, L = gold_g["counts"], 10
    nG, nM = counts.shape
    rng = np.random.default_rng(3)
    y = _labels_with_edges(rng, nG, L)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, None, L=L, model=cb.MODEL_G)
    ch.set_state(y=y)
    p = oracle.cGDecomposeCounts(O, y, L)
    ix, u = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ref = oracle.cGCalcGibbsProbY(counts.astype(np.float64), p["nTSByC"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0,
                                  1.0, 1.0, ix, u)
    ch.sweep_y(ix, u)
    _, yg = ch.get_state()
    _, py = ch.probs()
    assert np.array_equal(yg, ref["y"]) and np.array_equal(py, ref["probs"]) and np.isfinite(py).all()
    ch.close()


def test_lgdelta_entry_zero_is_NA(oracle, gold_cg):
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(4)
    z = rng.integers(1, K + 1, nM).astype(np.int32)
    y = _labels_with_edges(rng, nG, L)
    p = oracle.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, K, L)
    bare = (p["nGByTS"] - 1).astype(np.int32)  # tabulate(y, L) without the pseudo-gene: module 1 -> 1, module 2 -> 0
    gene = 6  # 1-based index of the gene that is alone in module 1: lg_delta[g - 1] = lg_delta[0] = NA
    row = p["nGByCP"][gene - 1]
    want = oracle.cG_CalcGibbsProbY(gene, row, p["nTSByCP"], p["nByTS"], bare, p["nByG"], y, L, 1.0, 1.0, 1.0)
    got = cb.cG_CalcGibbsProbY(gene, row, p["nTSByCP"], p["nByTS"], bare, p["nByG"], y, L, nG, 1.0, 1.0, 1.0)
    assert np.isnan(want[0]) and np.isnan(got[0])          # the gene's own (singleton) module
    assert np.isnan(want[1]) and np.isnan(got[1])          # the empty module: lg_delta[g] with g = 0
    assert np.array_equal(got[2:], want[2:]) and np.isfinite(got[2:]).all()
    ix, u = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    with pytest.raises(cb.CeldaCudaError, match="NA in probability vector"):
        cb.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], bare, p["nByG"], y, L, nG, 1.0, 1.0, 1.0, ix, u)


def test_walker_alias_draws_more_than_200_likely_modules(oracle):
    """L = 240 modules and very low counts: nearly every gene's module vector has > 200 entries with L * p > 0.1, so
    sample.int draws through Walker's alias tables (R/celda_functions.R:1-11 -> sample.int).  Labels, tables and
    probabilities stay bit-identical to the oracle, in the celda_CG form and on the raw counts (celda_G)."""
    rng = np.random.default_rng(17)
    G, C, K, L = 720, 48, 3, 240
    counts = rng.poisson(0.4, (G, C)).astype(np.int32)
    counts[counts.sum(1) == 0, 0] = 1
    counts[0, counts.sum(0) == 0] = 1
    s = np.ones(C, dtype=np.int32)
    z = rng.integers(1, K + 1, C).astype(np.int32)
    y = (rng.permutation(G) % L + 1).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    # celda_CG form
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
    ix, u = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
    ref = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, G, 1.0, 1.0, 1.0, ix, u)
    ch.sweep_y(ix, u)
    _, yg = ch.get_state()
    _, py = ch.probs()
    q = np.exp(py - py.max(axis=0))
    q /= q.sum(axis=0)
    assert ((L * q > 0.1).sum(axis=0) > 200).mean() > 0.5     # the Walker branch is the common case here
    assert np.array_equal(yg, ref["y"]) and np.array_equal(py, ref["probs"])
    ch.close()
    # celda_G form
    ch = cb.Chain(A, None, L=L, model=cb.MODEL_G)
    ch.set_state(y=y)
    pg = oracle.cGDecomposeCounts(O, y, L)
    refg = oracle.cGCalcGibbsProbY(counts.astype(np.float64), pg["nTSByC"], pg["nByTS"], pg["nGByTS"], pg["nByG"], y, L, G,
                                   1.0, 1.0, 1.0, ix, u)
    ch.sweep_y(ix, u)
    _, yg = ch.get_state()
    assert np.array_equal(yg, refg["y"]) and np.array_equal(ch.probs()[1], refg["probs"])
    ch.close()
