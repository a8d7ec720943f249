"""2-rank gloo worker for tests/test_host_logic.py (CPU): exercises the N > 1 host paths with the ORACLE standing in
for the device chain (the oracle is the checker here, never the product)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from celda_b200 import CSC, model, shard  # noqa: E402
from oracle.oracle import Oracle, dense_to_csc  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    o = Oracle()
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "celdaCG.npz")))
    counts, s = g["counts"], g["s"]
    A = CSC.from_dense(counts)

    # ---- grid search: runs partitioned over ranks, gathered with all_gather_object; fake runner = oracle LL of
    # a deterministic labelling that depends on (K, L, seed)
    def runner(run, seed):
        rng = np.random.default_rng(seed * 1000 + run["K"] * 10 + run["L"])
        z = model.initialize_cluster(run["K"], counts.shape[1], rng=rng)
        y = model.initialize_cluster(run["L"], counts.shape[0], rng=rng)
        p = o.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, run["K"], run["L"])
        ll = o.cCGCalcLL(run["K"], run["L"], p["mCPByS"], p["nTSByCP"], p["nByG"], p["nByTS"], p["nGByTS"], p["nS"],
                         p["nG"], 1.0, 1.0, 1.0, 1.0)
        return dict(finalLogLik=ll, z=z, y=y)

    def gather(local):
        out = [None] * world
        dist.all_gather_object(out, local)
        return out

    params = {"K": [4, 5, 6], "L": [9, 10, 11]}
    res = model.celdaGridSearch(A, s, params, nchains=2, seed=12345, rank=rank, nranks=world, runner=runner,
                                gather=gather)
    single = model.celdaGridSearch(A, s, params, nchains=2, seed=12345, rank=0, nranks=1, runner=runner)
    assert [r["index"] for r in res["runParams"]] == list(range(1, 19))
    assert [r["logLikelihood"] for r in res["runParams"]] == [r["logLikelihood"] for r in single["runParams"]]

    # ---- cell sharding: partial tables of the shards all-reduced == tables of the whole matrix; an EM step on
    # each shard from the global tables reproduces the unsharded step
    K = 6
    z = model.initialize_cluster(K, counts.shape[1], rng=np.random.default_rng(5))
    b = shard.partition_columns_by_nnz(A.p, world)
    lo, hi = b[rank], b[rank + 1]
    loc = shard.shard_csc(A, lo, hi)
    O_loc = (loc.nrow, loc.ncol, loc.p, loc.i, loc.x)
    nS = int(s.max())
    part_n = o.colSumByGroupSparse(O_loc, z[lo:hi], K)
    part_m = o.table_zs(z[lo:hi], s[lo:hi], K, nS).astype(np.int64)
    import torch
    tn, tm = torch.from_numpy(np.ascontiguousarray(part_n)), torch.from_numpy(np.ascontiguousarray(part_m))
    dist.all_reduce(tn)
    dist.all_reduce(tm)
    full = o.cCDecomposeCounts(dense_to_csc(counts), s, z, K)
    assert np.array_equal(tn.numpy(), np.ascontiguousarray(full["nGByCP"]))
    assert np.array_equal(tm.numpy(), np.ascontiguousarray(full["mCPByS"]))
    ref = o.cCCalcEMProbZ(dense_to_csc(counts), full["mCPByS"], full["nGByCP"], full["nByC"], full["nCP"], z, s, K,
                          full["nG"], full["nM"], 1.0, 1.0)
    mine = o.cCCalcEMProbZ(O_loc, full["mCPByS"], full["nGByCP"], None, full["nCP"], z[lo:hi], s[lo:hi], K, full["nG"],
                           hi - lo, 1.0, 1.0, doSample=False)
    assert np.array_equal(mine["probs"].argmax(0) + 1, ref["z"][lo:hi])
    dist.barrier()
    print("WORKER_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
