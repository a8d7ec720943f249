"""N-rank worker (one process per GPU, torchrun): cell-sharded celda_C EM through libcelda_cuda's NCCL all-reduce
against the single-GPU chain, and a grid search partitioned over the ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celda_b200 as cb  # noqa: E402
from celda_b200 import model, shard  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # host plumbing only (unique id, gathers); the data path is NCCL inside the library
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "celdaC.npz")))
    counts, s, K = g["counts"], g["s_mod"], 10
    A = cb.CSC.from_dense(counts)
    z0 = model.initialize_cluster(K, counts.shape[1], rng=np.random.default_rng(3))

    ids = [shard.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = shard.Comm(ids[0], world, rank, local)
    sh = shard.ShardedChainC.from_full(A, s, K, comm)
    sh.set_state(z0)
    ref = cb.Chain(A, s, K=K, model=cb.MODEL_C, device=local)
    ref.set_state(z=z0)
    for it in range(4):
        ll_s, ll_r = sh.iterate_em(), ref.iterate_em()
        assert ll_s == ll_r, (it, ll_s, ll_r)  # integer tables are bit-identical, so the LL is too
        zs = sh.local_z()
        zr, _ = ref.get_state()
        assert np.array_equal(zs, zr[sh.lo:sh.hi]), it
        ts, tr = sh.chain.tables(), ref.tables()
        for k in ("nGByCP", "mCPByS", "nCP"):
            assert np.array_equal(ts[k], tr[k]), (it, k)
    ps, pr = sh.chain.perplexity(), ref.perplexity()
    assert abs(ps - pr) <= 1e-9 * pr, (ps, pr)
    sh.close()
    ref.close()
    comm.close()

    # grid search partitioned over the ranks (no NCCL): same log-likelihoods as one rank doing everything
    gcg = dict(np.load(os.path.join(ROOT, "tests", "golden", "celdaCG.npz")))
    Acg = cb.CSC.from_dense(gcg["counts"])

    def gather(local_res):
        out = [None] * world
        dist.all_gather_object(out, local_res)
        return out

    params = {"K": [4, 5], "L": [9, 10]}
    res = model.celdaGridSearch(Acg, gcg["s"], params, maxIter=5, nchains=2, seed=7, rank=rank, nranks=world,
                                device=local, gather=gather)
    if rank == 0:
        single = model.celdaGridSearch(Acg, gcg["s"], params, maxIter=5, nchains=2, seed=7, device=local)
        assert [r["logLikelihood"] for r in res["runParams"]] == [r["logLikelihood"] for r in single["runParams"]]
        for a, b in zip(res["resList"], single["resList"]):
            assert np.array_equal(a["z"], b["z"]) and np.array_equal(a["y"], b["y"])
    dist.barrier()
    print("MULTI_GPU_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
