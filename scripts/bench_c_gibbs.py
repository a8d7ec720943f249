#!/usr/bin/env python
"""celda_C with algorithm = "Gibbs" at the headline matrix (100k cells x 20k genes, K=30): one step = one iteration
of the .celda_C loop body (R/celda_C.R:427-470): z sweep on the raw counts + .cCCalcLL.  Not a BASELINE config of
its own: it is the celda_C form of SURVEY.md 8 row a8, measured on the configs[1] counts."""
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
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--K", type=int, default=30)
    ap.add_argument("--L", type=int, default=150, help="gene modules used only to shape the synthetic counts")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--start", default="truth10", choices=["truth10", "random"])
    a = ap.parse_args()
    import celda_b200 as cb
    from bench_support import simulate_cells_cg
    t0 = time.time()
    d = simulate_cells_cg(12345, 10, a.cells // 10, a.genes, a.K, a.L, (500, 3000))
    gen_s = time.time() - t0
    counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    nG, nM = d["nG"], d["nM"]
    rng = np.random.Generator(np.random.PCG64(4))
    z = d["z"].copy()
    if a.start == "random":
        z = rng.integers(1, a.K + 1, nM).astype(np.int32)
    else:
        flip = rng.random(nM) < 0.1
        z[flip] = rng.integers(1, a.K + 1, int(flip.sum()))
    t0 = time.time()
    ch = cb.Chain(counts, d["s"], K=a.K, model=cb.MODEL_C)
    ch.set_state(z=z)
    setup_s = time.time() - t0
    ch.seed(12345)
    lls = [ch.iterate(None, None, None, None) for _ in range(a.warmup)]
    ch.profile(True)
    l0 = cb.launch_count()
    ch.timer_start()
    for _ in range(a.steps):
        lls.append(ch.iterate(None, None, None, None))
    ms = ch.timer_stop()
    prof = ch.profile_get()
    Z = float(counts.nnz)
    sweep_ms = prof["sweep_z"]["ms"] / max(1, prof["sweep_z"]["launches"])
    print(json.dumps({
        "metric": "celda_C_gibbs_iterations_per_sec", "value": a.steps / (ms / 1e3), "unit": "iter/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "gpu_launches": int(cb.launch_count() - l0),
        "config": {"workload": f"celda_C Gibbs z sweep, {nG} genes x {nM} cells, K={a.K}, nnz={int(Z)}, "
                               f"start={a.start}, Philox draws, exact full-column scan"},
        "sweep_z_ms": sweep_ms, "movers_per_sweep": prof["sweep_z"]["movers"] / a.steps,
        "rounds_per_sweep": prof["sweep_z"]["rounds"] / a.steps, "tasks_per_sweep": prof["sweep_z"]["units"] / a.steps,
        "reference_work": {"table_lookups_per_sweep": 2.0 * nG * nM * a.K, "fresh_lgamma_per_sweep": Z * a.K},
        "dadd_rate_tflops": 1.0 * nG * nM * a.K / (sweep_ms * 1e-3) / 1e12,
        "logLik_first_last": [lls[0], lls[-1]], "gen_s": gen_s, "setup_s": setup_s}))
    ch.close()


if __name__ == "__main__":
    main()
