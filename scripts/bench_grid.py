#!/usr/bin/env python
"""configs[4]: celdaGridSearch(celda_CG) over a K x L grid on 50k cells x 20k genes, the (chain, K, L) runs
partitioned over the GPUs of one box (one chain per GPU at a time, longest-first by K*L, no data-path collective).
Reports wall time of the whole search and chains/s; run with 1 and with N ranks to get the scaling.

  python scripts/bench_grid.py
  python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_grid.py"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=50000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--Ks", default="10,20,30,40,50")
    ap.add_argument("--Ls", default="50,100,150,200")
    ap.add_argument("--nchains", type=int, default=1)
    ap.add_argument("--maxIter", type=int, default=6)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import celda_b200 as cb
    from bench_support import simulate_cells_cg
    from celda_b200 import model
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    d = simulate_cells_cg(12345, 10, a.cells // 10, a.genes, 30, 150, (1000, 10000))
    counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    params = {"K": [int(v) for v in a.Ks.split(",")], "L": [int(v) for v in a.Ls.split(",")]}

    def gather(local_res):
        if world == 1:
            return [local_res]
        out = [None] * world
        dist.all_gather_object(out, [{k: v for k, v in r.items() if k != "result"} | {"result": {"finalLogLik": r["logLikelihood"]}}
                                     for r in local_res])
        return out

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = cb.launch_count()
    t0 = time.perf_counter()
    res = model.celdaGridSearch(counts, d["s"], params, maxIter=a.maxIter, nchains=a.nchains, seed=12345,
                                algorithm="Gibbs", rank=rank, nranks=world, device=local, gather=gather,
                                zInitialize="random", yInitialize="random")
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    if rank == 0:
        n = len(res["runParams"])
        print(json.dumps({"metric": "celdaGridSearch_chains_per_sec", "value": n / wall, "unit": "chains/s",
                          "n_gpus": world, "wall_s": wall, "runs": n, "maxIter": a.maxIter,
                          "gpu_launches_rank0": int(cb.launch_count() - l0),
                          "config": {"workload": f"configs[4]: celdaGridSearch celda_CG Gibbs, {d['nM']} cells x {d['nG']} "
                                                 f"genes, K in {params['K']}, L in {params['L']}, nchains={a.nchains}, "
                                                 f"maxIter={a.maxIter}, random init, Philox draws"},
                          "logLikelihood": [r["logLikelihood"] for r in res["runParams"]]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
