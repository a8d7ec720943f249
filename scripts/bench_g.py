#!/usr/bin/env python
"""configs[3]: celda_G module-only y sweep, 20k genes x 500k cells, L=200 (one GPU; a Gibbs chain does not shard).
One step = one iteration of the .celda_G loop body (R/celda_G.R:400-424): y sweep on the raw counts + .cGCalcLL."""
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
    ap.add_argument("--cells", type=int, default=500000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--L", type=int, default=200)
    ap.add_argument("--K", type=int, default=50, help="cell populations used only to shape the synthetic counts")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scan-mode", type=int, default=0, help="1 = opt-in scan of the count columns only (reports the margin)")
    ap.add_argument("--flip", type=float, default=0.1, help="fraction of the true module labels re-drawn for the start state")
    a = ap.parse_args()
    import celda_b200 as cb
    from bench_support import simulate_cells_cg
    t0 = time.time()
    d = simulate_cells_cg(12345, 10, a.cells // 10, a.genes, a.K, a.L, (500, 3000))
    gen_s = time.time() - t0
    counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    nG = d["nG"]
    rng = np.random.Generator(np.random.PCG64(4))
    y = d["y"].copy()
    flip = rng.random(nG) < a.flip
    y[flip] = rng.integers(1, a.L + 1, int(flip.sum()))
    t0 = time.time()
    ch = cb.Chain(counts, None, L=a.L, model=cb.MODEL_G)
    ch.set_state(y=y)
    setup_s = time.time() - t0
    ch.set_scan_mode(a.scan_mode)
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
    sweep_ms = prof["sweep_y"]["ms"] / max(1, prof["sweep_y"]["launches"])
    print(json.dumps({
        "metric": "celda_G_gibbs_iterations_per_sec", "value": a.steps / (ms / 1e3), "unit": "iter/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "gpu_launches": int(cb.launch_count() - l0),
        "config": {"workload": f"configs[3]: celda_G y sweep, {nG} genes x {d['nM']} cells, L={a.L}, nnz={int(Z)}, "
                               f"Philox draws, " + ("count columns only (scan mode 1)" if a.scan_mode else "exact full-column scan")},
        "sweep_y_ms": sweep_ms, "movers_per_sweep": prof["sweep_y"]["movers"] / a.steps,
        "rounds_per_sweep": prof["sweep_y"]["rounds"] / a.steps,
        "cta0_mclk_per_sweep": {q: prof["sweep_y"]["clk_" + q] / a.steps / 1e6 for q in ("commit", "units", "draw", "barrier")},
        "tasks_per_sweep": prof["sweep_y"]["units"] / a.steps,
        "reference_work": {"table_lookups_per_sweep": 2.0 * nG * d["nM"] * a.L, "fresh_lgamma_per_sweep": Z * a.L},
        "dadd_rate_tflops": 2.0 * nG * d["nM"] * a.L / (sweep_ms * 1e-3) / 1e12,
        "scan_mode": a.scan_mode, "scan_min_margin_over_bound_and_unproven_draws": list(ch.scan_margin()) if a.scan_mode else None,
        "logLik_first_last": [lls[0], lls[-1]], "gen_s": gen_s, "setup_s": setup_s}))
    ch.close()


if __name__ == "__main__":
    main()
