"""recursiveSplitCell / recursiveSplitModule on the device (celda_b200/recursive.py; R/recursiveSplit.R:452-870,
1250-1650) on the reference's celdaCGSim fixture: the K / L sequence, and every model's log-likelihood equal to the
oracle's collapsed log-likelihood of the labels it returns."""
import numpy as np
import pytest

import celda_b200 as cb
from celda_b200 import recursive
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _ll_C(o, O, s, z, K, nS):
    p = o.cCDecomposeCounts(O, s, z, K)
    return o.cCCalcLL(p["mCPByS"], p["nGByCP"], K, nS, p["nG"], 1.0, 1.0)


def test_recursiveSplitCell_counts_form(oracle, gold_cg):
    counts, s = gold_cg["counts"], gold_cg["s"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    res = recursive.recursiveSplitCell(A, s, initialK=3, maxK=6, seed=7)
    Ks = [p["K"] for p in res["runParams"]]
    assert Ks == [3, 4, 5, 6]
    for mod, p in zip(res["resList"], res["runParams"]):
        z = np.asarray(mod["z"], np.int32)
        assert z.shape == (counts.shape[1],) and np.unique(z).size == p["K"]
        ref = _ll_C(oracle, O, s, z, p["K"], int(s.max()))
        assert abs(p["log_likelihood"] - ref) <= 1e-9 * abs(ref)
    assert res["perplexity"].shape == (4, 1) and np.all(np.isfinite(res["perplexity"]))
    # splitting one population at a time towards the simulated K = 5 improves the fit
    assert res["runParams"][2]["log_likelihood"] > res["runParams"][0]["log_likelihood"]


def test_recursiveSplitCell_with_module_labels(oracle, gold_cg):
    counts, s = gold_cg["counts"], gold_cg["s"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    res = recursive.recursiveSplitCell(A, s, initialK=3, maxK=5, yInit=gold_cg["sim_y"], seed=5, perplexity=False)
    assert [p["K"] for p in res["runParams"]] == [3, 4, 5] and all(p["L"] == 10 for p in res["runParams"])
    for mod, p in zip(res["resList"], res["runParams"]):
        z, y = np.asarray(mod["z"], np.int32), np.asarray(mod["y"], np.int32)
        d = oracle.cCGDecomposeCounts(O, s, z, y, p["K"], 10)
        ref = oracle.cCGCalcLL(p["K"], 10, d["mCPByS"], d["nTSByCP"], d["nByG"], d["nByTS"], d["nGByTS"], d["nS"],
                               counts.shape[0], 1.0, 1.0, 1.0, 1.0)
        assert abs(p["log_likelihood"] - ref) <= 1e-9 * abs(ref)


def test_recursiveSplitModule_counts_and_collapsed_forms(oracle, gold_cg):
    counts = gold_cg["counts"]
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    for tempK in (None, 20):
        res = recursive.recursiveSplitModule(A, initialL=4, maxL=7, tempK=tempK, seed=3, perplexity=False)
        assert [p["L"] for p in res["runParams"]] == [4, 5, 6, 7]
        for k, (mod, p) in enumerate(zip(res["resList"], res["runParams"])):
            y = np.asarray(mod["y"], np.int32)
            assert y.shape == (counts.shape[0],) and y.min() >= 1 and y.max() <= p["L"]
            if tempK is None or k > 0:  # the collapsed form's first model keeps the collapsed matrix's log-likelihood
                d = oracle.cGDecomposeCounts(O, y, p["L"])
                ref = oracle.cGCalcLL(d["nTSByC"], d["nByTS"], d["nByG"], d["nGByTS"], counts.shape[1], counts.shape[0],
                                      p["L"], 1.0, 1.0, 1.0)
                assert abs(p["log_likelihood"] - ref) <= 1e-9 * abs(ref)
