"""GPU parity for the celda_G y sweep on raw counts (configs[3] form): every column is scanned, zeros included,
exactly like cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:212-225)."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _oracle_g_iteration(o, O, counts, y, L, ix, u, hyper):
    b, d, g = hyper
    p = o.cGDecomposeCounts(O, y, L)
    oy = o.cGCalcGibbsProbY(counts.astype(np.float64), p["nTSByC"], p["nByTS"], p["nGByTS"], p["nByG"], y, L,
                            counts.shape[0], b, d, g, ix, u)
    ll = o.cGCalcLL(oy["nTSByC"], oy["nByTS"], p["nByG"], oy["nGByTS"], counts.shape[1], counts.shape[0], L, b, d, g)
    return oy, ll


@pytest.mark.parametrize("start", ["random", "fitted"])
def test_celda_G_sweeps_bit_exact_fixture(oracle, gold_g, start):
    counts, L = gold_g["counts"], 10
    nG, nM = counts.shape
    rng = np.random.default_rng(5)
    y = rng.integers(1, L + 1, nG).astype(np.int32) if start == "random" else gold_g["y"].copy()
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, None, L=L, model=cb.MODEL_G)
    ch.set_state(y=y)
    for it in range(3):
        ix, u = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
        ref, ll_ref = _oracle_g_iteration(oracle, O, counts, y, L, ix, u, (1.0, 1.0, 1.0))
        ll = ch.iterate(ix, u, None, None)
        _, yg = ch.get_state()
        assert np.array_equal(yg, ref["y"]), (it, (yg != ref["y"]).sum())
        t = ch.tables()
        assert np.array_equal(t["nTSByC"], ref["nTSByC"]) and np.array_equal(t["nByTS"], ref["nByTS"])
        assert np.array_equal(t["nGByTS"], ref["nGByTS"])
        _, py = ch.probs()
        assert np.array_equal(py, ref["probs"])
        assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        y = ref["y"]
    ch.close()


def test_celda_G_sweep_sparse_midsize_two_chunks(oracle):
    """Sparse counts (many zero columns per gene), L = 40 (> 32: two module chunks), non-default priors."""
    rng = np.random.default_rng(13)
    G, C, L = 160, 900, 40
    y_true = rng.integers(1, L + 1, G)
    y_true[:L] = np.arange(1, L + 1)
    phi = rng.dirichlet(np.ones(L) * 0.3, C)            # per-cell module usage
    counts = np.zeros((G, C), dtype=np.int32)
    for c in range(C):
        mod = rng.multinomial(rng.integers(50, 200), phi[c])
        for l in np.flatnonzero(mod):
            idx = np.flatnonzero(y_true == l + 1)
            counts[idx, c] += rng.multinomial(mod[l], np.ones(idx.size) / idx.size).astype(np.int32)
    counts[counts.sum(1) == 0, 0] = 1
    counts[0, counts.sum(0) == 0] = 1
    assert (counts == 0).mean() > 0.6
    y = y_true.astype(np.int32).copy()
    flip = rng.random(G) < 0.3
    y[flip] = rng.integers(1, L + 1, int(flip.sum()))
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    hyper = (0.5, 2.0, 3.0)
    ch = cb.Chain(A, None, L=L, beta=hyper[0], delta=hyper[1], gamma=hyper[2], model=cb.MODEL_G)
    ch.set_state(y=y)
    for it in range(2):
        ix, u = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
        ref, ll_ref = _oracle_g_iteration(oracle, O, counts, y, L, ix, u, hyper)
        ll = ch.iterate(ix, u, None, None)
        _, yg = ch.get_state()
        assert np.array_equal(yg, ref["y"]), (it, (yg != ref["y"]).sum())
        t = ch.tables()
        assert np.array_equal(t["nTSByC"], ref["nTSByC"])
        _, py = ch.probs()
        assert np.array_equal(py, ref["probs"])
        assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        assert (ref["y"] != y).sum() > 0
        y = ref["y"]
    # doSample = FALSE leaves everything in place and still fills probs
    ch.sweep_y(np.arange(1, G + 1), np.zeros(G), doSample=False)
    _, y2 = ch.get_state()
    assert np.array_equal(y2, y)
    ch.close()
