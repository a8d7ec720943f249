"""Host logic of recursiveSplitCell / recursiveSplitModule (celda_b200/recursive.py; R/recursiveSplit.R) with a scripted
runner: which clusters are tried, how the labels are rewritten, the K / L sequence, the fallbacks for degenerate chains
and the reference quirks (zInit ignored under zInitialize = "random", default priors in the sub-chains)."""
import numpy as np

import celda_b200.recursive as rs
from celda_b200 import CSC


class ScriptedRunner:
    """Deterministic stand-in for the device: a K = 2 sub-chain splits its columns (rows) in halves, a K-cluster chain
    returns `labels mod K`, log-likelihoods prefer balanced partitions."""

    def __init__(self, lose_population_at=None):
        self.calls = []
        self.lose = lose_population_at

    def _ll(self, lab, n):
        c = np.bincount(lab, minlength=n + 1)[1:]
        return -float((c.astype(np.float64) ** 2).sum())

    def celda_C(self, counts, s, K, **kw):
        self.calls.append(("celda_C", counts.ncol, K, dict(kw)))
        n = counts.ncol
        if K == 2:
            z = np.where(np.arange(n) < n // 2, 1, 2).astype(np.int32)
        else:
            z = (np.arange(n) % K + 1).astype(np.int32)
            if self.lose == K:
                z[z == K] = 1
        return dict(z=z, y=None, finalLogLik=self._ll(z, K), completeLogLik=np.zeros(1), params={})

    def celda_G(self, counts, L, **kw):
        self.calls.append(("celda_G", counts.nrow, L, dict(kw)))
        n = counts.nrow
        y = (np.where(np.arange(n) < n // 2, 1, 2) if L == 2 and kw.get("yInitialize") == "random"
             else (kw["yInit"] if kw.get("yInitialize") == "predefined" else np.arange(n) % L + 1)).astype(np.int32)
        return dict(z=None, y=y, finalLogLik=self._ll(y, L), completeLogLik=np.zeros(1), params={})

    def ll_C(self, counts, s, z, K, alpha, beta):
        self.calls.append(("ll_C", counts.ncol, K, dict(alpha=alpha, beta=beta)))
        return self._ll(np.asarray(z), K)

    def ll_G(self, counts, y, L, beta, delta, gamma):
        self.calls.append(("ll_G", counts.nrow, L, dict(beta=beta)))
        return self._ll(np.asarray(y), L)

    def perplexity(self, counts, s, mod):
        return 42.0


def _counts(nG=12, nM=40):
    rng = np.random.default_rng(0)
    return CSC.from_dense(rng.integers(1, 5, (nG, nM)).astype(np.int32))


def test_singleSplitZ_picks_best_candidate_and_relabels():
    counts = _counts()
    z = np.repeat([1, 2, 3], [20, 16, 4]).astype(np.int32)  # population 3 has 4 cells: > minCell = 3, so it is tried too
    r = ScriptedRunner()
    out = rs.singleSplitZ(r, counts, z, np.ones(40, np.int32), 4, minCell=3, alpha=2.0, beta=3.0,
                          rng=np.random.default_rng(1))
    subs = [c for c in r.calls if c[0] == "celda_C"]
    assert [c[1] for c in subs] == [20, 16, 4]                      # columns of each population
    assert all(c[2] == 2 and c[3]["zInitialize"] == "random" and c[3]["splitOnIter"] == -1 and
               c[3]["splitOnLast"] is False and "alpha" not in c[3] for c in subs)  # default priors in the sub-chain
    assert all(c[3] == dict(alpha=2.0, beta=3.0) for c in r.calls if c[0] == "ll_C")  # caller's priors for the score
    # splitting the largest population gives the most balanced partition: first half -> K, second half keeps label 1
    want = z.copy()
    want[:10] = 4
    assert np.array_equal(out["z"], want) and out["ll"] == -(10 ** 2 + 10 ** 2 + 16 ** 2 + 4 ** 2)


def test_singleSplitZ_nothing_to_split():
    z = np.repeat([1, 2], [3, 3]).astype(np.int32)
    out = rs.singleSplitZ(ScriptedRunner(), _counts(5, 6), z, np.ones(6, np.int32), 3, rng=np.random.default_rng(1))
    assert out["ll"] == -np.inf and np.array_equal(out["z"], z)


def test_recursiveSplitCell_sequence_and_zinit_quirk():
    counts = _counts()
    r = ScriptedRunner()
    res = rs.recursiveSplitCell(counts, initialK=2, maxK=5, runner=r, perplexity=True)
    assert [p["K"] for p in res["runParams"]] == [2, 3, 4, 5]
    assert [p["index"] for p in res["runParams"]] == [1, 2, 3, 4]
    assert res["perplexity"].shape == (4, 1) and res["runParams"][0]["mean_perplexity"] == 42.0
    full = [c for c in r.calls if c[0] == "celda_C" and c[1] == 40 and c[2] != 2 or (c[0] == "celda_C" and c[3].get("nchains") == 1)]
    assert full[0][3]["zInitialize"] == "split"
    for c in full[1:]:  # zInit is handed over but the chain is started at random, stopIter = 5, splitOnLast = FALSE
        assert c[3]["zInitialize"] == "random" and c[3]["zInit"] is not None and c[3]["stopIter"] == 5
        assert c[3]["splitOnLast"] is False
    for mod, p in zip(res["resList"], res["runParams"]):
        assert np.unique(mod["z"]).size == p["K"] and p["log_likelihood"] == mod["finalLogLik"]


def test_recursiveSplitCell_lost_population_falls_back_to_split_labels():
    counts = _counts()
    r = ScriptedRunner(lose_population_at=4)
    res = rs.recursiveSplitCell(counts, initialK=3, maxK=4, runner=r, perplexity=False)
    m = res["resList"][-1]
    assert m["params"]["K"] == 4 and np.unique(m["z"]).size == 4       # the split's labels, not the degenerate chain's
    assert m["finalLogLik"] == r._ll(m["z"], 4) and "perplexity" not in res


def test_recursiveSplitModule_counts_form():
    counts = _counts(nG=24, nM=10)
    r = ScriptedRunner()
    res = rs.recursiveSplitModule(counts, initialL=2, maxL=4, tempK=None, runner=r, perplexity=False)
    assert [p["L"] for p in res["runParams"]] == [2, 3, 4]
    first = next(c for c in r.calls if c[0] == "celda_G")
    assert first[3]["maxIter"] == 20 and first[3]["nchains"] == 1
    warm = [c for c in r.calls if c[0] == "celda_G" and c[3].get("yInitialize") == "predefined"]
    assert [c[2] for c in warm] == [3, 4] and all(c[3]["stopIter"] == 5 and c[3]["splitOnIter"] == -1 for c in warm)
    for mod, p in zip(res["resList"], res["runParams"]):
        assert np.unique(mod["y"]).size == p["L"]
