"""CPU tests of the host logic: chain loop bookkeeping, grid-search partitioning, cell sharding, and the N > 1
paths over a world_size-2 gloo group (no GPU: chains / table builders are injected from the oracle)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from celda_b200 import CSC, model, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expand_grid_and_seeds_match_reference_order():
    runs = model.expand_grid({"K": [4, 5, 6], "L": [9, 10]}, nchains=3)
    assert len(runs) == 18
    # base::expand.grid: chain fastest, then K, then L (R/celdaGridSearch.R:298-310)
    assert [(r["chain"], r["K"], r["L"]) for r in runs[:4]] == [(1, 4, 9), (2, 4, 9), (3, 4, 9), (1, 5, 9)]
    assert (runs[-1]["chain"], runs[-1]["K"], runs[-1]["L"]) == (3, 6, 10)
    # allSeeds[ifelse(i %% nchains == 0, nchains, i %% nchains)] (R/celdaGridSearch.R:381-383)
    assert [model.chain_seed(12345, r["index"], 3) for r in runs[:6]] == [12345, 12346, 12347, 12345, 12346, 12347]


def test_partition_runs_lpt():
    runs = model.expand_grid({"K": list(range(5, 55, 5)), "L": list(range(10, 210, 10))}, nchains=3)  # configs[4]
    assert len(runs) == 600
    parts = model.partition_runs(runs, 8)
    seen = sorted(r["index"] for p in parts for r in p)
    assert seen == list(range(1, 601))
    loads = [sum(r["K"] * r["L"] for r in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.01  # >= 7x of 8 GPUs needs balance within a few percent
    assert model.partition_runs(runs, 8) == parts  # deterministic


def test_initialize_cluster():
    rng = np.random.default_rng(0)
    z = model.initialize_cluster(7, 50, rng=rng)
    assert sorted(np.unique(z)) == list(range(1, 8)) and z.size == 50
    assert model.initialize_cluster(3, 5, initial=[3, 3, 1, 2, 1]).tolist() == [3, 3, 1, 2, 1]
    with pytest.raises(ValueError):
        model.initialize_cluster(3, 5, initial=[1, 1, 1, 2, 2])


def test_partition_columns_by_nnz_and_shard():
    rng = np.random.default_rng(1)
    dense = rng.poisson(0.3, (40, 200)).astype(np.int32)
    dense[:, 50:60] = 0
    A = CSC.from_dense(dense)
    for n in (1, 2, 3, 8):
        b = shard.partition_columns_by_nnz(A.p, n)
        assert b[0] == 0 and b[-1] == 200 and all(b[i] < b[i + 1] for i in range(n))
        nnz = [int(A.p[b[i + 1]] - A.p[b[i]]) for i in range(n)]
        assert max(nnz) - min(nnz) <= 2 * 40
        back = np.concatenate([_dense(shard.shard_csc(A, b[i], b[i + 1])) for i in range(n)], axis=1)
        assert np.array_equal(back, dense)


def _dense(c):
    out = np.zeros((c.nrow, c.ncol), dtype=np.int32)
    for j in range(c.ncol):
        out[c.i[c.p[j]:c.p[j + 1]], j] = c.x[c.p[j]:c.p[j + 1]]
    return out


def test_chain_loop_bookkeeping():
    """Best-so-far / stopping rules of R/celda_CG.R:734-741 on a scripted log-likelihood trace."""
    trace = iter([-100.0, -90.0, -95.0, -80.0, -85.0, -84.0, -83.0])

    class Fake:
        def __init__(self):
            self.t = 0

        def loglik(self):
            return next(trace)

        def get_state(self):
            return np.array([self.t]), np.array([self.t])

        def iterate_philox(self):
            self.t += 1
            return next(trace)

    r = model._run_chain(Fake(), "Gibbs", maxIter=6, stopIter=2, draws="philox", rng=None, nG=1, nM=1, model=model.MODEL_CG)
    # iteration 1 always becomes best (-90); -95 no; -80 best; -85, -84 no improvement -> stop after stopIter exceeded
    assert r["finalLogLik"] == -80.0 and r["z"].tolist() == [3]
    assert r["completeLogLik"].tolist() == [-100.0, -90.0, -95.0, -80.0, -85.0, -84.0, -83.0][:len(r["completeLogLik"])]
    assert len(r["completeLogLik"]) == 6  # noImp after each iteration: 1, 2, 1, 2, 3 -> 3 > stopIter ends the loop


@pytest.mark.timeout(300)
def test_world_size_2_gloo_paths():
    """Grid-search gather and cell-sharded table all-reduce over a 2-rank gloo group (CPU)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561",
                        os.path.join(ROOT, "tests", "gloo_worker.py")], capture_output=True, text=True, env=env,
                       timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("WORKER_OK") == 2, r.stdout[-2000:]


def test_grid_search_isolates_a_failed_run():
    """One run that raises is reported in its row and the other runs still complete (failure isolation)."""
    import math

    from celda_b200 import model

    def runner(run, seed):
        if run["K"] == 3 and run["chain"] == 2:
            raise RuntimeError("sampler failed")
        return dict(finalLogLik=-float(run["K"] * 10 + run["L"]), z=None, y=None)

    res = model.celdaGridSearch(None, None, dict(K=[2, 3], L=[4]), nchains=2, runner=runner)
    rows = res["runParams"]
    assert [r["index"] for r in rows] == [1, 2, 3, 4]
    bad = [r for r in rows if r["error"]]
    assert len(bad) == 1 and bad[0]["K"] == 3 and bad[0]["chain"] == 2 and math.isnan(bad[0]["logLikelihood"])
    assert "sampler failed" in bad[0]["error"] and res["resList"][bad[0]["index"] - 1] is None
    assert all(r["error"] is None and r["logLikelihood"] == -(r["K"] * 10 + r["L"]) for r in rows if r is not bad[0])
    import pytest
    with pytest.raises(RuntimeError):
        model.celdaGridSearch(None, None, dict(K=[2, 3], L=[4]), nchains=2, runner=runner, isolate_failures=False)


def test_selectBestModel_and_subsetCeldaList():
    """R/celdaGridSearch.R:546-581, 668-687 on a celdaGridSearch-shaped result."""
    from celda_b200 import model
    runs = model.expand_grid({"K": [4, 5], "L": [9, 10]}, nchains=3)
    lls = [-5.0, -3.0, -3.0, float("nan"), -7.0, -6.5, -1.0, -2.0, float("nan"), float("nan"), float("nan"), -9.0]
    rp = [dict(index=r["index"], chain=r["chain"], K=r["K"], L=r["L"], seed=12345 + r["index"] - 1, logLikelihood=ll,
               seconds=0.1, rank=0, error=None) for r, ll in zip(runs, lls)]
    res = dict(runParams=rp, resList=[{"id": r["index"]} for r in rp])
    best = model.selectBestModel(res, asList=True)
    assert [(r["K"], r["L"], r["chain"]) for r in best["runParams"]] == [(4, 9, 2), (5, 9, 3), (4, 10, 1), (5, 10, 3)]
    assert [m["id"] for m in best["resList"]] == [r["index"] for r in best["runParams"]]
    one = model.subsetCeldaList(best, {"K": 5, "L": 10})
    assert one == {"id": 12}                                  # a single match returns the model
    sub = model.subsetCeldaList(res, {"K": [4], "chain": [1, 2]})
    assert [r["index"] for r in sub["runParams"]] == [1, 2, 7, 8] and len(sub["resList"]) == 4
    assert model.selectBestModel(model.subsetCeldaList(res, {"K": 4, "L": 9})) == {"id": 2}
    with pytest.raises(ValueError, match="not columns in runParams"):
        model.subsetCeldaList(res, {"Q": 1})
    with pytest.raises(ValueError, match="No runs matched"):
        model.subsetCeldaList(res, {"K": 6})
