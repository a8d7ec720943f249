#!/usr/bin/env python
"""configs[2]: celda_C EM-mode z update, 1M cells x 5k genes, K=50, cells sharded over the ranks (one process per
GPU), partial G x K / K x S tables combined by the in-library NCCL all-reduce.  Strong scaling: total work fixed.

  python scripts/bench_em.py                                    1 GPU
  python -m torch.distributed.run --nproc-per-node N ... scripts/bench_em.py   N GPUs
Prints one JSON line on rank 0 (iterations/s, cells*iterations/s, kernel shares, SpMM FP64 rate)."""
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
    ap.add_argument("--cells", type=int, default=1000000)
    ap.add_argument("--genes", type=int, default=5000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--S", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import celda_b200 as cb
    from bench_support import simulate_cells_c
    from celda_b200 import shard
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")  # host plumbing (unique id, timing max); data path = NCCL inside libcelda_cuda
    per = a.cells // world
    lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else a.cells
    t0 = time.time()
    d = simulate_cells_c(12345, a.S, a.cells // a.S, a.genes, a.K, (500, 3000), cell_range=(lo, hi))
    gen_s = time.time() - t0
    counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    rng = np.random.Generator(np.random.PCG64(999))
    z = d["z"].copy()
    flip = rng.random(z.size) < 0.1
    z[flip] = rng.integers(1, a.K + 1, int(flip.sum()))
    comm = None
    if world > 1:
        ids = [shard.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = shard.Comm(ids[0], world, rank, local)
        sh = shard.ShardedChainC(counts, d["s"][lo:hi], a.K, comm, nS=a.S, lo=lo)
        ch = sh.chain
    else:
        ch = cb.Chain(counts, d["s"], K=a.K, model=cb.MODEL_C, device=local)
    ch.set_state(z=z[lo:hi])
    lls = [ch.iterate_em() for _ in range(a.warmup)]
    ch.profile(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = cb.launch_count()
    ch.timer_start()
    for _ in range(a.steps):
        lls.append(ch.iterate_em())
    ms = ch.timer_stop()
    launches = cb.launch_count() - l0
    prof = ch.profile_get()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nnz = torch.tensor([counts.nnz], dtype=torch.float64)
        dist.all_reduce(nnz)
        Z = float(nnz.item())
    else:
        Z = float(counts.nnz)
    if rank == 0:
        it_s = a.steps / (ms / 1e3)
        em_ms = prof["sweep_z"]["ms"] / max(1, prof["sweep_z"]["launches"])
        print(json.dumps({
            "metric": "celda_C_EM_iterations_per_sec", "value": it_s, "unit": "iter/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "scaling": "strong",
            "cells_iters_per_sec": it_s * a.cells, "gpu_launches": int(launches),
            "config": {"workload": f"configs[2]: celda_C EM z step, {a.cells} cells x {a.genes} genes, K={a.K}, nnz={int(Z)}, "
                                   f"cells sharded over {world} GPU(s), NCCL all-reduce of G x K + K x S int32 tables"},
            "em_step_ms_rank0": em_ms, "loglik_ms_rank0": prof["loglik"]["ms"] / max(1, prof["loglik"]["launches"]),
            "spmm_fp64_tflops": 2.0 * Z * a.K / (em_ms * 1e-3) / 1e12 / world * world,
            "logLik_first_last": [lls[0], lls[-1]], "gen_s": gen_s}))
    ch.close()
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
