"""GPU parity for the sweep-granularity entry points that keep the R closures' signatures and the
stateless log-likelihoods."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def test_stateless_loglik_goldens(oracle, gold_cg, gold_c, gold_g):
    O = dense_to_csc(gold_cg["counts"])
    p = oracle.cCGDecomposeCounts(O, gold_cg["s_mod"], gold_cg["z"], gold_cg["y"], 5, 10)
    ll = cb.cCGCalcLL(5, 10, p["mCPByS"], p["nTSByCP"], p["nByG"], p["nByTS"], p["nGByTS"], p["nS"], p["nG"], 1.0, 1.0,
                      1.0, 1.0)
    assert abs(ll - (-1215541.0958389007)) <= 1e-12 * abs(ll)
    assert ll < 0  # tests/testthat/test-celda_CG.R:60-94
    O = dense_to_csc(gold_c["counts"])
    p = oracle.cCDecomposeCounts(O, gold_c["s_mod"], gold_c["z"], 10)
    ll = cb.cCCalcLL(p["mCPByS"], p["nGByCP"], gold_c["s_mod"], gold_c["z"], 10, p["nS"], p["nG"], 1.0, 1.0)
    assert abs(ll - (-1273077.9330206949)) <= 1e-12 * abs(ll)
    O = dense_to_csc(gold_g["counts"])
    p = oracle.cGDecomposeCounts(O, gold_g["y"], 10)
    ll = cb.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], p["nM"], p["nG"], 10, 1.0, 1.0, 1.0)
    assert abs(ll - (-290669.0461321392)) <= 1e-12 * abs(ll)
    # non-default hyper-parameters against the oracle
    ll = cb.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], p["nM"], p["nG"], 10, 0.3, 2.5, 4.0)
    ref = oracle.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], p["nM"], p["nG"], 10, 0.3, 2.5, 4.0)
    assert abs(ll - ref) <= 1e-12 * abs(ref)


def test_closure_signature_sweeps_bit_exact(oracle, gold_cg):
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(31)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    p = oracle.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, K, L)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    got = cb.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0, 1.0,
                              ixy, uy)
    ref = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                  1.0, ixy, uy)
    for k in ("nTSByC", "nGByTS", "nByTS", "y", "probs"):
        assert np.array_equal(got[k], ref[k]), k
    got = cb.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0,
                              ixz, uz)
    ref = oracle.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0,
                                  ixz, uz)
    for k in ("mCPByS", "nGByCP", "nCP", "z", "probs"):
        assert np.array_equal(got[k], ref[k]), k
    # inputs are not modified (R copy semantics)
    assert np.array_equal(p["nTSByCP"], oracle.colSumByGroup(p["nTSByC"], z, K))
    # celda_C form of the z sweep: the raw (densified) count matrix against nGByCP, R = nG
    dense = counts.astype(np.float64)
    got = cb.cCCalcGibbsProbZ(dense, p["mCPByS"], p["nGByCP"], p["nByC"], p["nGByCP"].sum(0), z, s, K, nG, nM, 0.5, 0.2,
                              ixz, uz)
    ref = oracle.cCCalcGibbsProbZ(dense, p["mCPByS"], p["nGByCP"], p["nByC"], p["nGByCP"].sum(0), z, s, K, nG, nM, 0.5,
                                  0.2, ixz, uz)
    for k in ("mCPByS", "nGByCP", "nCP", "z", "probs"):
        assert np.array_equal(got[k], ref[k]), k


def test_closure_signature_em(oracle, gold_c):
    counts, s, z, K = gold_c["counts"], gold_c["s_mod"], gold_c["sim_z"], 10
    O = dense_to_csc(counts)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    ref = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"], p["nM"], 1.0, 1.0)
    for cnt in (cb.CSC.from_dense(counts), counts.astype(np.float64)):
        got = cb.cCCalcEMProbZ(cnt, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"], p["nM"], 1.0, 1.0)
        for k in ("mCPByS", "nGByCP", "nCP", "z"):
            assert np.array_equal(got[k], ref[k]), k
        assert np.allclose(got["probs"], ref["probs"], rtol=1e-12, atol=0)
    with pytest.raises(cb.CeldaCudaError, match="Division by 0"):
        cb.cCCalcEMProbZ(counts.astype(np.float64), p["mCPByS"] * 0, p["nGByCP"], p["nByC"], p["nCP"], z, s, K, p["nG"],
                         p["nM"], 0.0, 1.0)


def test_closure_z_sweep_celda_C_shape_beyond_shared_memory(oracle):
    """.cCCalcGibbsProbZ with counts = a raw G x C matrix whose G x K table does not fit shared memory (G = 2600 at
    K = 30, R/celda_C.R:427-440): the stateless call runs on the raw-count kernel; labels, tables and probabilities
    equal the oracle's dense evaluation bit for bit."""
    rng = np.random.default_rng(44)
    G, C, K, S = 2600, 120, 30, 2
    counts = rng.poisson(0.15, (G, C)).astype(np.int32)
    counts[counts.sum(1) == 0, 0] = 1
    counts[0, counts.sum(0) == 0] = 1
    s = (np.arange(C) % S + 1).astype(np.int32)
    z = (rng.permutation(C) % K + 1).astype(np.int32)
    p = oracle.cCDecomposeCounts(dense_to_csc(counts), s, z, K)
    ix, u = rng.permutation(C).astype(np.int32) + 1, rng.random(C)
    dense = counts.astype(np.float64)
    for do_sample in (True, False):
        got = cb.cCCalcGibbsProbZ(dense, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, G, C, 1.0, 1.0, ix, u,
                                  doSample=do_sample)
        ref = oracle.cCCalcGibbsProbZ(dense, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, G, C, 1.0, 1.0, ix, u,
                                      doSample=do_sample)
        for k in ("mCPByS", "nGByCP", "nCP", "z", "probs"):
            assert np.array_equal(got[k], ref[k]), (do_sample, k)
    assert (ref["z"] == z).all() and (got["z"] == z).all()  # the last pass did not sample
