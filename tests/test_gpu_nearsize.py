"""Parity near the sizes of configs[2] and configs[3] (the fixtures top out at a few thousand cells): one EM z step on
200k cells x 5k genes, K = 50, and one celda_G y sweep on 20k cells with L = 200, both against the oracle."""
import numpy as np
import pytest

import celda_b200 as cb

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_em_step_200k_cells_bit_exact(oracle):
    """configs[2] shape at a fifth of its cells: .cCCalcEMProbZ (R/celda_C.R:717-756) -- every label (first-max ties
    included), all K x nM probabilities, the re-decomposed tables; log-likelihood to 1e-9."""
    from bench_support import simulate_cells_c
    K, S, G, C = 50, 10, 5000, 200000
    d = simulate_cells_c(12345, S, C // S, G, K, (500, 3000))
    rng = np.random.default_rng(11)
    z = d["z"].copy()
    flip = rng.random(C) < 0.1
    z[flip] = rng.integers(1, K + 1, int(flip.sum()))
    p32 = d["p"].astype(np.int32)
    O = (G, C, p32, d["i"], d["x"])
    ch = cb.Chain(cb.CSC(G, C, p32, d["i"], d["x"]), d["s"], K=K, model=cb.MODEL_C)
    ch.set_state(z=z)
    p = oracle.cCDecomposeCounts(O, d["s"], z, K)
    t = ch.tables()
    assert np.array_equal(t["nGByCP"], p["nGByCP"]) and np.array_equal(t["mCPByS"], p["mCPByS"])
    ref = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, d["s"], K, G, C, 1.0, 1.0)
    ll = ch.iterate_em()
    zg, _ = ch.get_state()
    assert np.array_equal(zg, ref["z"]), int((zg != ref["z"]).sum())
    pz, _ = ch.probs()
    assert np.array_equal(pz, ref["probs"])
    t = ch.tables()
    assert np.array_equal(t["nGByCP"], ref["nGByCP"]) and np.array_equal(t["mCPByS"], ref["mCPByS"])
    assert np.array_equal(t["nCP"], ref["nCP"])
    want = oracle.cCCalcLL(ref["mCPByS"], ref["nGByCP"], K, S, G, 1.0, 1.0)
    assert abs(ll - want) <= 1e-9 * abs(want)
    assert (zg != z).sum() > 1000
    ch.close()


@pytest.mark.timeout(900)
def test_celda_G_sweep_20k_cells_L200_bit_exact(oracle):
    """configs[3] shape (L = 200, seven module chunks, 20 000 cells = 157 column tiles) on 640 genes: one Gibbs y sweep
    on the raw counts, zeros scanned like cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:212-225)."""
    rng = np.random.default_rng(29)
    G, C, L = 640, 20000, 200
    y_true = rng.integers(1, L + 1, G)
    y_true[:L] = np.arange(1, L + 1)
    size = np.bincount(y_true, minlength=L + 1)[1:]
    usage = rng.dirichlet(np.full(L, 0.05), C)           # per-cell module usage: a few modules per cell => sparse counts
    N = rng.integers(200, 800, C)
    pg = usage[:, y_true - 1] / size[y_true - 1]          # C x G: gene probabilities per cell
    counts = np.stack([rng.multinomial(N[c], pg[c] / pg[c].sum()) for c in range(C)], axis=1).astype(np.int32)
    counts[counts.sum(1) == 0, 0] = 1
    counts[0, counts.sum(0) == 0] = 1
    y = y_true.astype(np.int32).copy()
    flip = rng.random(G) < 0.2
    y[flip] = rng.integers(1, L + 1, int(flip.sum()))
    from oracle.oracle import dense_to_csc
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, None, L=L, model=cb.MODEL_G)
    ch.set_state(y=y)
    ix, u = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
    p = oracle.cGDecomposeCounts(O, y, L)
    ref = oracle.cGCalcGibbsProbY(counts.astype(np.float64), p["nTSByC"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, G, 1.0,
                                  1.0, 1.0, ix, u)
    ll_ref = oracle.cGCalcLL(ref["nTSByC"], ref["nByTS"], p["nByG"], ref["nGByTS"], C, G, L, 1.0, 1.0, 1.0)
    ll = ch.iterate(ix, u, None, None)
    _, yg = ch.get_state()
    assert np.array_equal(yg, ref["y"]), int((yg != ref["y"]).sum())
    t = ch.tables()
    assert np.array_equal(t["nTSByC"], ref["nTSByC"]) and np.array_equal(t["nByTS"], ref["nByTS"])
    assert np.array_equal(t["nGByTS"], ref["nGByTS"])
    _, py = ch.probs()
    assert np.array_equal(py, ref["probs"])
    assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
    assert (yg != y).sum() > 20
    ch.close()
