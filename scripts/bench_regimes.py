#!/usr/bin/env python
"""Per-iteration device time of the celda_CG chain from a given start (churn regimes), with sweep statistics.

  python scripts/bench_regimes.py --start random --iters 6      [CELDA_CUDA_MC=1 for the one-mover-per-round schedule]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import celda_b200 as cb  # noqa: E402


def run(d, a, start, iters, seed=12345):
    a.start = start
    z0, y0, rng = bench.start_state(d, a, seed)
    csc = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    ch = cb.Chain(csc, d["s"], K=a.K, L=a.L, device=0)
    ch.set_state(z0, y0)
    ixy, uy, ixz, uz = bench.draw_inputs(rng, iters, d["nG"], d["nM"])
    ch.stage_draws(ixy, uy, ixz, uz)
    rows = []
    for it in range(iters):
        ch.profile(True)
        ch.timer_start()
        ll = ch.iterate_staged(it)
        ms = ch.timer_stop()
        pr = ch.profile_get()
        ch.profile(False)
        rows.append({"iter": it, "ms": ms, "logLik": ll,
                     **{f"{k}_{q}": pr[k][q] for k in ("sweep_y", "sweep_z")
                        for q in ("ms", "movers", "rounds", "redraws", "discards", "units", "clk_commit", "clk_units",
                                  "clk_draw", "clk_barrier")}})
    zf, yf = ch.get_state()
    ch.close()
    return rows, zf, yf


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--start", default="random")
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--K", type=int, default=30)
    ap.add_argument("--L", type=int, default=150)
    ap.add_argument("--S", type=int, default=10)
    ap.add_argument("--nmin", type=int, default=1000)
    ap.add_argument("--nmax", type=int, default=10000)
    ap.add_argument("--mc", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    d = bench.make_data(a)
    res = {}
    ref = None
    for mc in ([m for m in a.mc.split(",") if m] or [os.environ.get("CELDA_CUDA_MC", "16")]):
        os.environ["CELDA_CUDA_MC"] = mc
        rows, zf, yf = run(d, a, a.start, a.iters)
        res["mc" + mc] = rows
        if ref is None:
            ref = (zf, yf)
        else:
            res["labels_equal_mc%s" % mc] = bool(np.array_equal(ref[0], zf) and np.array_equal(ref[1], yf))
        print("MC", mc, [(round(r["ms"], 2), r["sweep_y_movers"], r["sweep_y_rounds"], r["sweep_z_movers"], r["sweep_z_rounds"]) for r in rows], flush=True)
        for k in ("sweep_y", "sweep_z"):
            r = rows[0]
            n = max(1, r[k + "_rounds"])
            print("   ", k, "iter0 per-round kclk: commit %.1f units %.1f draw %.1f barrier %.1f; units/round %.0f" % (
                r[k + "_clk_commit"] / n / 1e3, r[k + "_clk_units"] / n / 1e3, r[k + "_clk_draw"] / n / 1e3,
                r[k + "_clk_barrier"] / n / 1e3, r[k + "_units"] / n), flush=True)
    txt = json.dumps({"start": a.start, "cells": a.cells, "genes": a.genes, "K": a.K, "L": a.L, "result": res})
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")
