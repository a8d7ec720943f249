"""GPU parity for the EM z step (a9) and the celda_C chain (decompose, LL, EM iteration)."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def test_celda_C_chain_decompose_ll_golden(oracle, gold_c):
    counts, s, z = gold_c["counts"], gold_c["s_mod"], gold_c["z"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=10, model=cb.MODEL_C)
    ch.set_state(z=z)
    t = ch.tables()
    p = oracle.cCDecomposeCounts(O, s, z, 10)
    for k in ("mCPByS", "nGByCP", "nCP", "nByC"):
        assert np.array_equal(t[k], p[k]), k
    ll = ch.loglik()
    assert abs(ll - (-1273077.9330206949)) / abs(ll) < 1e-12  # data/celdaCMod.rda finalLogLik
    ch.close()


@pytest.mark.parametrize("start", ["truth", "random"])
def test_celda_C_em_iterations_bit_exact(oracle, gold_c, start):
    """configs[2] at fixture size: celda_C, zAlgorithm = EM."""
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    rng = np.random.default_rng(3)
    z = gold_c["sim_z"].copy() if start == "truth" else rng.integers(1, K + 1, counts.shape[1]).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, model=cb.MODEL_C, alpha=1.0, beta=1.0)
    ch.set_state(z=z)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    m, n, ncp = p["mCPByS"], p["nGByCP"], p["nCP"]
    for it in range(4):
        ref = oracle.cCCalcEMProbZ(O, m, n, p["nByC"], ncp, z, s, K, p["nG"], p["nM"], 1.0, 1.0)
        ll = ch.iterate_em()
        zg, _ = ch.get_state()
        assert np.array_equal(zg, ref["z"]), (it, (zg != ref["z"]).sum())
        t = ch.tables()
        assert np.array_equal(t["nGByCP"], ref["nGByCP"]) and np.array_equal(t["mCPByS"], ref["mCPByS"])
        assert np.array_equal(t["nCP"], ref["nCP"])
        pz, _ = ch.probs()
        assert np.array_equal(pz, ref["probs"])
        want = oracle.cCCalcLL(ref["mCPByS"], ref["nGByCP"], K, p["nS"], p["nG"], 1.0, 1.0)
        assert abs(ll - want) <= 1e-9 * abs(want)
        z, m, n, ncp = ref["z"], ref["mCPByS"], ref["nGByCP"], ref["nCP"]
    ch.close()


def test_em_no_sample_and_first_max_ties(oracle, gold_c):
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 12  # populations 11, 12 empty: identical columns => exact ties
    z = gold_c["sim_z"].copy()
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, model=cb.MODEL_C)
    ch.set_state(z=z)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    ch.sweep_z_em(doSample=False)
    ref = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"], p["nM"], 1.0, 1.0,
                               doSample=False)
    pz, _ = ch.probs()
    assert np.array_equal(pz, ref["probs"])
    assert np.array_equal(pz[10], pz[11])  # the tie is real
    zg, _ = ch.get_state()
    assert np.array_equal(zg, z)
    ch.sweep_z_em()
    ref = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"], p["nM"], 1.0, 1.0)
    zg, _ = ch.get_state()
    assert np.array_equal(zg, ref["z"])
    ch.close()


def test_celda_CG_em_iteration(oracle, gold_cg):
    """celda_CG with algorithm = "EM" (the reference default, R/celda_CG.R:96): Gibbs y sweep + EM z step."""
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(17)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    for it in range(3):
        ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
        p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
        oy = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                     1.0, ixy, uy)
        nTSByC = oracle.rowSumByGroupChangeSparse(O, p["nTSByC"].copy(order="F"), oy["y"], y, L)
        oz = oracle.cCCalcEMProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0)
        want = oracle.cCGCalcLL(K, L, oz["mCPByS"], oz["nGByCP"], p["nByG"], oy["nByTS"], oy["nGByTS"], p["nS"], nG, 1.0,
                                1.0, 1.0, 1.0)
        ll = ch.iterate_em(ixy, uy)
        zg, yg = ch.get_state()
        assert np.array_equal(yg, oy["y"]) and np.array_equal(zg, oz["z"])
        t = ch.tables()
        assert np.array_equal(t["nTSByCP"], oz["nGByCP"]) and np.array_equal(t["nCP"], oz["nCP"])
        assert np.array_equal(t["nGByCP"], oracle.colSumByGroupSparse(O, oz["z"], K))
        pz, _ = ch.probs()
        assert np.array_equal(pz, oz["probs"])
        assert abs(ll - want) <= 1e-9 * abs(want)
        z, y = oz["z"], oy["y"]
    ch.close()


@pytest.mark.parametrize("K", [20, 45])
def test_celda_C_em_tiled_kernel_large(oracle, K):
    """>= 4096 cells takes the shared-memory-tiled contraction (several gene tiles when G*K*8 > 200 KB)."""
    rng = np.random.default_rng(100 + K)
    G, C, S = 1500, 6000, 4
    phi = rng.dirichlet(np.ones(G) * 0.05, K)
    z_true = rng.integers(1, K + 1, C)
    counts = np.zeros((G, C), dtype=np.int32)
    for c in range(C):
        counts[:, c] = rng.multinomial(rng.integers(100, 400), phi[z_true[c] - 1])
    counts[counts.sum(1) == 0, 0] = 1
    s = rng.integers(1, S + 1, C).astype(np.int32)
    s[:S] = np.arange(1, S + 1)
    z = z_true.astype(np.int32).copy()
    flip = rng.random(C) < 0.3
    z[flip] = rng.integers(1, K + 1, int(flip.sum()))
    z[:K] = np.arange(1, K + 1)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ch = cb.Chain(A, s, K=K, model=cb.MODEL_C, alpha=0.7, beta=0.4)
    ch.set_state(z=z)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    m, n, ncp = p["mCPByS"], p["nGByCP"], p["nCP"]
    for it in range(2):
        ref = oracle.cCCalcEMProbZ(O, m, n, p["nByC"], ncp, z, s, K, G, C, 0.7, 0.4)
        ch.iterate_em()
        zg, _ = ch.get_state()
        assert np.array_equal(zg, ref["z"]), (it, (zg != ref["z"]).sum())
        pz, _ = ch.probs()
        assert np.array_equal(pz, ref["probs"])
        t = ch.tables()
        assert np.array_equal(t["nGByCP"], ref["nGByCP"]) and np.array_equal(t["mCPByS"], ref["mCPByS"])
        z, m, n, ncp = ref["z"], ref["mCPByS"], ref["nGByCP"], ref["nCP"]
    ch.close()


def test_cluster_probability_and_factorize(oracle, gold_cg):
    """Cold callers of the hot path on the fitted fixture model: tests/testthat/test-celda_CG.R
    :22-41 -- factorizeMatrix self-consistency and dimensions, clusterProbability rows summing to 1."""
    from celda_b200 import model
    counts, s, z, y = gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"]
    K, L = 5, 10
    ch = cb.Chain(cb.CSC.from_dense(counts), s, K=K, L=L)
    ch.set_state(z, y)
    cp = model.clusterProbability(ch)
    assert cp["zProbability"].shape == (counts.shape[1], K) and cp["yProbability"].shape == (counts.shape[0], L)
    assert np.allclose(cp["zProbability"].sum(1), 1.0, atol=1e-10) and np.allclose(cp["yProbability"].sum(1), 1.0, atol=1e-10)
    z2, y2 = ch.get_state()
    assert np.array_equal(z2, z) and np.array_equal(y2, y)
    f = model.factorizeMatrix(ch)
    assert f["counts"]["cellPopulation"].shape == (L, K) and f["counts"]["module"].shape == (counts.shape[0], L)
    assert f["counts"]["cell"].sum() == counts.sum() and f["counts"]["module"].sum() == counts.sum()
    assert np.allclose(f["proportions"]["cell"].sum(0), 1.0) and np.allclose(f["posterior"]["module"].sum(0), 1.0)
    assert np.allclose(f["counts"]["cell"], f["proportions"]["cell"] * counts.sum(0)[None, :])
    # the posterior factors are exactly what the perplexity kernel uses
    M = np.log(f["posterior"]["module"] @ f["posterior"]["cellPopulation"])
    inner = M.T @ counts + np.log(f["posterior"]["sample"])[:, s - 1]
    mx = inner.max(0)
    px = np.exp(-((mx + np.log(np.exp(inner - mx).sum(0))).sum() / counts.sum()))
    assert abs(px - ch.perplexity()) <= 1e-9 * px
    ch.close()
