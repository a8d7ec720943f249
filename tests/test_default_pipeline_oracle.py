"""The reference's DEFAULT pipelines -- 'split' initialisation, the chain loop, the split / merge heuristic every 10
iterations (R/celda_CG.R:432-790, R/celda_C.R:377-589, R/celda_G.R:352-520) -- as orchestrated by the product's host
code (celda_b200/split.py), run here on the CPU checker: the backend is the oracle, the Gibbs draws are the ones the
device generator makes (the host build of celda_b200/csrc/philox.cuh).  Pinned to the reference itself: on the
reference's simulated data sets the pipelines end at the partitions and log-likelihoods of the reference's own stored
model fits (celdaCGMod / celdaCMod / celdaGMod in /root/reference/data, extracted by tests/golden/make_golden.py).
The GPU suite runs the same orchestration on the device (tests/test_gpu_split_init.py: label for label equal to this
CPU route under injected draws)."""
import numpy as np
import pytest

import celda_b200 as cb
from celda_b200 import split


@pytest.fixture(scope="module")
def host_order(tmp_path_factory):
    from test_philox_order import build_host_order
    return build_host_order(str(tmp_path_factory.mktemp("philox_pipeline")))


@pytest.fixture(scope="module")
def obackend(oracle):
    from split_oracle_backend import OracleBackend
    return OracleBackend(oracle)


def same_partition(a, b):
    """Equal up to the numbering of the clusters."""
    pairs = set(zip(np.asarray(a).tolist(), np.asarray(b).tolist()))
    return len(pairs) == len(set(np.asarray(a).tolist())) == len(set(np.asarray(b).tolist()))


def _dense(counts):
    d = np.zeros((counts.nrow, counts.ncol))
    d[counts.i, np.repeat(np.arange(counts.ncol), np.diff(counts.p))] = counts.x
    return d


class OracleOpsG:
    """celda_b200.split.DeviceOpsG on the oracle: the celda_G y sweep runs on the raw counts."""

    def __init__(self, o, counts, L, y, hyper=(1.0, 1.0, 1.0)):
        self.o, self.O, self.L, self.h = o, (counts.nrow, counts.ncol, counts.p, counts.i, counts.x), L, hyper
        self.nG, self.nM, self.dense, self.py = counts.nrow, counts.ncol, _dense(counts), None
        self.set_state(y)

    def set_state(self, y):
        self.y = np.asarray(y, np.int32).copy()
        self.p = self.o.cGDecomposeCounts(self.O, self.y, self.L)

    def _ll(self, p):
        return self.o.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], self.nM, self.nG, self.L, *self.h)

    def ll(self):
        return self._ll(self.p)

    def ll_alt(self, y):
        return self._ll(self.o.cGDecomposeCounts(self.O, np.asarray(y, np.int32), self.L))

    def state(self):
        return self.y.copy()

    def probs(self):
        return self.py

    def iterate(self, ixy, uy, ixz, uz):
        p = self.p
        r = self.o.cGCalcGibbsProbY(self.dense, p["nTSByC"], p["nByTS"], p["nGByTS"], p["nByG"], self.y, self.L, self.nG,
                                    *self.h, ixy, uy)
        self.py = r["probs"]
        self.set_state(r["y"])
        return self.ll()


class OracleOpsC:
    """celda_b200.split.DeviceOpsC(algorithm = "EM") on the oracle."""

    def __init__(self, o, counts, s, K, z, hyper=(1.0, 1.0)):
        self.o, self.O, self.s, self.K, self.h = o, (counts.nrow, counts.ncol, counts.p, counts.i, counts.x), s, K, hyper
        self.pz = None
        self.set_state(z)

    def set_state(self, z):
        self.z = np.asarray(z, np.int32).copy()
        self.p = self.o.cCDecomposeCounts(self.O, self.s, self.z, self.K)

    def _ll(self, p):
        return self.o.cCCalcLL(p["mCPByS"], p["nGByCP"], self.K, p["nS"], p["nG"], *self.h)

    def ll(self):
        return self._ll(self.p)

    def ll_alt(self, z):
        return self._ll(self.o.cCDecomposeCounts(self.O, self.s, np.asarray(z, np.int32), self.K))

    def state(self):
        return self.z.copy()

    def probs(self):
        return self.pz

    def iterate(self, ixy, uy, ixz, uz):
        p = self.p
        r = self.o.cCCalcEMProbZ(self.O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], self.z, self.s, self.K, p["nG"],
                                 p["nM"], *self.h)
        self.pz = r["probs"]
        self.set_state(r["z"])
        return self.ll()


def test_celda_CG_default_pipeline_reaches_the_reference_fit(oracle, obackend, host_order, gold_cg):
    """celda_CG(K = 5, L = 10) with the defaults, Gibbs draws from the device generator (seed 5, stream 0: the y sweep of
    iteration `it` uses sweep id 2 (it - 1), the z sweep 2 (it - 1) + 1): what test_model_drivers_with_reference_defaults
    runs on the GPU."""
    from split_oracle_backend import OracleOpsCG
    counts, s, K, L, seed = cb.CSC.from_dense(gold_cg["counts"]), gold_cg["s"], 5, 10, 5
    rng = np.random.Generator(np.random.PCG64(seed))
    z = split.initializeSplitZ(counts, K, alpha=1.0, beta=1.0, rng=rng, backend=obackend)
    y = split.initializeSplitY(counts, L, beta=1.0, delta=1.0, gamma=1.0, rng=rng, backend=obackend)

    def draws(it, rng_, nG, nM):
        iy, uy = host_order(nG, seed=seed, sweep=2 * (it - 1), stream=0, uniforms=True)
        iz, uz = host_order(nM, seed=seed, sweep=2 * (it - 1) + 1, stream=0, uniforms=True)
        return iy[0], uy[0], iz[0], uz[0]

    r = split.run_chain_CG_with_splits(OracleOpsCG(oracle, counts, s, K, L, z, y), counts, K, L, draws, rng, obackend,
                                       maxIter=25, stopIter=10, splitOnIter=10, splitOnLast=True)
    ref = float(gold_cg["finalLogLik"])  # celdaCGMod, fitted by the reference
    assert abs(r["finalLogLik"] - ref) <= 1e-12 * abs(ref)
    assert same_partition(r["z"], gold_cg["z"]) and same_partition(r["y"], gold_cg["y"])
    assert r["finalLogLik"] >= r["completeLogLik"][0] and any(m[0] == 10 for m in r["messages"])


def test_celda_G_default_pipeline_reaches_the_reference_fit(oracle, obackend, host_order, gold_g):
    counts, L, seed = cb.CSC.from_dense(gold_g["counts"]), 10, 5
    rng = np.random.Generator(np.random.PCG64(seed))
    y = split.initializeSplitY(counts, L, beta=1.0, delta=1.0, gamma=1.0, rng=rng, backend=obackend)

    def draws(it, rng_, nG, nM):
        iy, uy = host_order(nG, seed=seed, sweep=it - 1, stream=0, uniforms=True)
        return iy[0], uy[0], None, None

    r = split.run_chain_one_axis_with_splits(OracleOpsG(oracle, counts, L, y), counts, L, "y", draws, rng, obackend,
                                             maxIter=12, stopIter=10, splitOnIter=10, splitOnLast=True)
    ref = float(gold_g["finalLogLik"])  # celdaGMod
    assert abs(r["finalLogLik"] - ref) <= 1e-12 * abs(ref)
    assert same_partition(r["y"], gold_g["y"])
    assert any(m[0] == 10 for m in r["messages"])


def test_celda_C_default_pipeline_reaches_the_reference_fit(oracle, obackend, gold_c):
    """celda_C(K = 10) with the defaults: 'split' initialisation, EM z steps (stopIter forced to 1), .cCSplitZ."""
    counts, s, K, seed = cb.CSC.from_dense(gold_c["counts"]), gold_c["s_mod"], 10, 5
    rng = np.random.Generator(np.random.PCG64(seed))
    z = split.initializeSplitZ(counts, K, alpha=1.0, beta=1.0, rng=rng, backend=obackend)
    r = split.run_chain_one_axis_with_splits(OracleOpsC(oracle, counts, s, K, z), counts, K, "z",
                                             lambda it, r_, G, M: (None,) * 4, rng, obackend, maxIter=25, stopIter=1,
                                             splitOnIter=10, splitOnLast=True)
    ref = float(gold_c["finalLogLik"])  # celdaCMod
    assert abs(r["finalLogLik"] - ref) <= 1e-12 * abs(ref)
    assert same_partition(r["z"], gold_c["z"])


def test_grid_of_default_pipelines_reaches_or_beats_the_reference_grid(oracle, obackend, host_order, gold_cg, gold_grid):
    """The reference's stored celdaGridSearch result on celdaCGSim (celdaCGGridSearchRes: K = 4..6 x L = 9..11, one
    model each): for every (K, L) the better of two default-pipeline chains (seeds 12345 and 5, device-generator draws)
    is at least as likely as the reference's model, and for the six points with K <= 5 it IS the reference's model
    (same log-likelihood to 1e-12)."""
    from split_oracle_backend import OracleOpsCG
    counts, s = cb.CSC.from_dense(gold_cg["counts"]), gold_cg["s"]

    def chain(K, L, seed):
        rng = np.random.Generator(np.random.PCG64(seed))
        z = split.initializeSplitZ(counts, K, alpha=1.0, beta=1.0, rng=rng, backend=obackend)
        y = split.initializeSplitY(counts, L, beta=1.0, delta=1.0, gamma=1.0, rng=rng, backend=obackend)

        def draws(it, rng_, nG, nM):
            iy, uy = host_order(nG, seed=seed, sweep=2 * (it - 1), stream=0, uniforms=True)
            iz, uz = host_order(nM, seed=seed, sweep=2 * (it - 1) + 1, stream=0, uniforms=True)
            return iy[0], uy[0], iz[0], uz[0]

        return split.run_chain_CG_with_splits(OracleOpsCG(oracle, counts, s, K, L, z, y), counts, K, L, draws, rng,
                                              obackend, maxIter=50, stopIter=10, splitOnIter=10, splitOnLast=True)

    exact = 0
    for K, L, ref in zip(gold_grid["K"], gold_grid["L"], gold_grid["finalLogLik"]):
        best = max(chain(int(K), int(L), seed)["finalLogLik"] for seed in (12345, 5))
        assert best >= ref - 1e-9 * abs(ref), (K, L, best, ref)
        same = abs(best - ref) <= 1e-12 * abs(ref)
        exact += same
        assert same or K == 6, (K, L, best, ref)
    assert exact == 6
