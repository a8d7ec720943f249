"""'split' initialisation (.initializeSplitZ / .initializeSplitY, R/initialize_clusters.R:60-400) on the device against
the same orchestration on the CPU oracle: identical labels from identical random draws."""
import numpy as np
import pytest

import celda_b200 as cb
from celda_b200 import model, split

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def obackend(oracle):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from split_oracle_backend import OracleBackend
    return OracleBackend(oracle)


@pytest.mark.parametrize("K,seed", [(5, 1), (8, 2), (3, 3)])
def test_initializeSplitZ_matches_oracle(gold_cg, obackend, K, seed):
    counts = cb.CSC.from_dense(gold_cg["counts"])
    zd = split.initializeSplitZ(counts, K, rng=np.random.default_rng(seed))
    zo = split.initializeSplitZ(counts, K, rng=np.random.default_rng(seed), backend=obackend)
    assert np.array_equal(zd, zo)
    assert zd.shape == (counts.ncol,) and set(np.unique(zd)) == set(range(1, K + 1))


@pytest.mark.parametrize("L,seed", [(10, 4), (6, 5)])
def test_initializeSplitY_matches_oracle(gold_cg, obackend, L, seed):
    counts = cb.CSC.from_dense(gold_cg["counts"])
    yd = split.initializeSplitY(counts, L, rng=np.random.default_rng(seed))
    yo = split.initializeSplitY(counts, L, rng=np.random.default_rng(seed), backend=obackend)
    assert np.array_equal(yd, yo)
    assert yd.shape == (counts.nrow,) and set(np.unique(yd)) == set(range(1, L + 1))


def test_split_initialisation_recovers_simulated_structure(gold_cg):
    """On celdaCGSim (K = 5, L = 10) the split initialisation alone puts most cells / genes with their true partners
    (the reference's point of using it as the default start): adjusted agreement well above a random start."""
    counts = cb.CSC.from_dense(gold_cg["counts"])
    z = split.initializeSplitZ(counts, 5, rng=np.random.default_rng(11))
    y = split.initializeSplitY(counts, 10, rng=np.random.default_rng(12))

    def purity(lab, truth):
        tot = 0
        for v in np.unique(lab):
            tot += np.bincount(truth[lab == v]).max()
        return tot / lab.size

    assert purity(z, gold_cg["sim_z"]) > 0.8
    assert purity(y, gold_cg["sim_y"]) > 0.6


def test_celda_CG_default_split_initialisation(gold_cg):
    """zInitialize = yInitialize = "split" is the reference default (R/celda_CG.R:85-110): the chain starts from
    the split labels and ends at least as likely as it started."""
    counts = cb.CSC.from_dense(gold_cg["counts"])
    res = model.celda_CG(counts, gold_cg["s"], 5, 10, algorithm="Gibbs", maxIter=5, nchains=1, seed=3,
                         zInitialize="split", yInitialize="split")
    ll = res["model"]["completeLogLik"]
    assert ll.size == 6 and res["model"]["finalLogLik"] >= ll[0]
    assert set(np.unique(res["model"]["z"])) <= set(range(1, 6))


def _draws(it, rng, nG, nM):
    return (rng.permutation(nG).astype(np.int32) + 1, rng.random(nG), rng.permutation(nM).astype(np.int32) + 1,
            rng.random(nM))


@pytest.mark.parametrize("start,seed", [("random", 21), ("merged", 22)])
def test_chain_loop_with_split_heuristic_matches_oracle(gold_cg, oracle, obackend, start, seed):
    """The .celda_CG loop with its split steps (.cCGSplitY / .cCGSplitZ every splitOnIter iterations,
    R/celda_CG.R:604-716, R/split_clusters.R:154-588) on the device and on the CPU oracle: same labels, same
    log-likelihood trace, same split decisions.  "merged": the truth with two populations and two modules merged and
    one label emptied, so a split is the right move."""
    from split_oracle_backend import OracleOpsCG
    counts, s, K, L = cb.CSC.from_dense(gold_cg["counts"]), gold_cg["s"], 5, 10
    rng0 = np.random.default_rng(seed)
    if start == "random":
        z = model.initialize_cluster(K, counts.ncol, None, rng0)
        y = model.initialize_cluster(L, counts.nrow, None, rng0)
    else:
        z, y = gold_cg["sim_z"].astype(np.int32).copy(), gold_cg["sim_y"].astype(np.int32).copy()
        z[z == 5] = 4
        z[:3] = 5            # population 5: three stray cells
        y[y == 10] = 9
        y[:3] = 10
    out = []
    for which in ("device", "oracle"):
        rng = np.random.default_rng(seed + 100)
        if which == "device":
            ch = cb.Chain(counts, s, K=K, L=L)
            ch.set_state(z, y)
            ops, be = split.DeviceOpsCG(ch), split.DeviceBackend()
        else:
            ops, be = OracleOpsCG(oracle, counts, s, K, L, z, y), obackend
        out.append(split.run_chain_CG_with_splits(ops, counts, K, L, _draws, rng, be, maxIter=9, stopIter=10, splitOnIter=3))
        if which == "device":
            ch.close()
    d, o = out
    assert [m[:2] for m in d["messages"]] == [m[:2] for m in o["messages"]] and len(d["messages"]) >= 2
    assert d["messages"] == o["messages"]
    assert np.array_equal(d["z"], o["z"]) and np.array_equal(d["y"], o["y"])
    assert np.allclose(d["completeLogLik"], o["completeLogLik"], rtol=1e-9, atol=0)
    assert d["finalLogLik"] >= d["completeLogLik"][0]
    if start == "merged":
        assert any("was split in two" in m[2] for m in d["messages"])


def test_cCSplitZ_and_cGSplitY_match_oracle(gold_c, gold_g, obackend):
    """.cCSplitZ (R/split_clusters.R:2-150) and .cGSplitY (:592-763) on the device and on the oracle from a state where
    one cluster is nearly empty and two true clusters share a label."""
    dev = split.DeviceBackend()
    # celda_C
    counts, K = cb.CSC.from_dense(gold_c["counts"]), 10
    z = gold_c["sim_z"].astype(np.int32).copy()
    z[z == 10] = 9
    z[:3] = 10
    res = []
    for be in (dev, obackend):
        ses = be.session_C(counts, 1.0, 1.0)
        new, msg = split.cCSplitZ(counts, z, K, ses.probs(z, K), lambda zz: ses.ll(zz, K), np.random.default_rng(5), be,
                                  maxClustersToTry=K)
        res.append((new, msg, ses.ll(new, K), ses.ll(z, K)))
        ses.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    assert "was split in two" in res[0][1] and res[0][2] > res[0][3]
    # celda_G
    counts, L = cb.CSC.from_dense(gold_g["counts"]), 10
    y = gold_g["y"].astype(np.int32).copy()
    y[y == 10] = 9
    y[:3] = 10
    res = []
    for be in (dev, obackend):
        ses = be.session_G(counts, 1.0, 1.0, 1.0)
        new, msg = split.cGSplitY(counts, y, L, ses.probs(y, L), lambda yy: ses.ll(yy, L), np.random.default_rng(6), be)
        res.append((new, msg, ses.ll(new, L), ses.ll(y, L)))
        ses.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    assert res[0][2] >= res[0][3]


def test_model_drivers_with_reference_defaults(gold_cg, gold_c, gold_g):
    """celda_CG / celda_C / celda_G with the reference's defaults ('split' initialisation, split heuristic every 10
    iterations and at a stall): the chains run, report their split decisions and end at least as likely as they began."""
    A = cb.CSC.from_dense(gold_cg["counts"])
    r = model.celda_CG(A, gold_cg["s"], 5, 10, algorithm="Gibbs", maxIter=25, nchains=1, seed=5)["model"]
    assert r["finalLogLik"] >= r["completeLogLik"][0] and any(m[0] == 10 for m in r["messages"])
    assert abs(r["finalLogLik"] - (-1215541.0958389007)) < 0.02 * 1215541  # the neighbourhood of the reference's fit
    r = model.celda_C(cb.CSC.from_dense(gold_c["counts"]), gold_c["s_mod"], 10, maxIter=25, nchains=1, seed=5)["model"]
    assert r["finalLogLik"] >= r["completeLogLik"][0] and set(np.unique(r["z"])) <= set(range(1, 11))
    r = model.celda_G(cb.CSC.from_dense(gold_g["counts"]), 10, maxIter=12, nchains=1, seed=5)["model"]
    assert r["finalLogLik"] >= r["completeLogLik"][0] and any(m[0] == 10 for m in r["messages"])
