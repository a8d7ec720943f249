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
