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


def test_celda_G_scan_mode_count_columns_only():
    """Opt-in scan mode 1 (celda_chain_set_scan_mode): the columns where a gene has no count are skipped.  The
    probabilities differ from the exact scan by rounding only, and whenever the reported margin exceeds the reported
    bound the labels are those of the exact scan."""
    rng = np.random.default_rng(21)
    G, C, L = 200, 1500, 40
    y_true = rng.integers(1, L + 1, G)
    y_true[:L] = np.arange(1, L + 1)
    counts = np.zeros((G, C), dtype=np.int32)
    for c in range(C):
        mod = rng.multinomial(rng.integers(40, 160), rng.dirichlet(np.ones(L) * 0.3))
        for l in np.flatnonzero(mod):
            idx = np.flatnonzero(y_true == l + 1)
            counts[idx, c] += rng.multinomial(mod[l], np.ones(idx.size) / idx.size).astype(np.int32)
    counts[counts.sum(1) == 0, 0] = 1
    counts[0, counts.sum(0) == 0] = 1
    y = y_true.astype(np.int32).copy()
    flip = rng.random(G) < 0.3
    y[flip] = rng.integers(1, L + 1, int(flip.sum()))
    A = cb.CSC.from_dense(counts)
    exact, fast = cb.Chain(A, None, L=L, model=cb.MODEL_G), cb.Chain(A, None, L=L, model=cb.MODEL_G)
    exact.set_state(y=y)
    fast.set_state(y=y)
    fast.set_scan_mode(1)
    with pytest.raises(cb.CeldaCudaError):
        fast.scan_margin()  # no sweep yet
    for it in range(3):
        ix, u = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
        le, lf = exact.iterate(ix, u, None, None), fast.iterate(ix, u, None, None)
        ratio, unproven = fast.scan_margin()
        assert ratio >= 0.0 and unproven >= 0
        pe, pf = exact.probs()[1], fast.probs()[1]
        assert np.max(np.abs(pe - pf)) <= 1e-9  # the skipped pairs move a log-probability by rounding only
        if unproven == 0:
            assert ratio > 1.0
            assert np.array_equal(exact.get_state()[1], fast.get_state()[1])
            assert abs(le - lf) <= 1e-12 * abs(le)
        else:  # a draw too close to call: continue both chains from the exact labels
            fast.set_state(y=exact.get_state()[1])
    assert unproven == 0  # on this problem the margins are orders of magnitude above the bounds
    fast.set_scan_mode(0)   # back to the exact scan: bit-identical again
    ix, u = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
    exact.iterate(ix, u, None, None)
    fast.iterate(ix, u, None, None)
    assert np.array_equal(exact.probs()[1], fast.probs()[1])
    cg = cb.Chain(A, np.ones(C, np.int32), K=2, L=L)
    with pytest.raises(cb.CeldaCudaError):
        cg.set_scan_mode(1)  # only the celda_G sweep scans raw counts
    cg.close()
    exact.close()
    fast.close()
