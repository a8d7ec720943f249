#!/usr/bin/env python
"""Wall time of celda_CG with the reference's defaults at configs[1] size: 'split' initialisation of z and y, then Gibbs
iterations with the split heuristic every 10 iterations (nchains = 1).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import celda_b200 as cb  # noqa: E402
from celda_b200 import model, split  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--genes", type=int, default=20000)
ap.add_argument("--K", type=int, default=30)
ap.add_argument("--L", type=int, default=150)
ap.add_argument("--maxIter", type=int, default=30)
a = ap.parse_args()
a.S, a.nmin, a.nmax = 10, 1000, 10000
d = bench.make_data(a)
counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
rng = np.random.default_rng(1)
t0 = time.perf_counter()
z = split.initializeSplitZ(counts, a.K, rng=rng)
t1 = time.perf_counter()
times_z = dict(split.TIMES)
split.TIMES.clear()
y = split.initializeSplitY(counts, a.L, rng=rng)
t2 = time.perf_counter()
times_y = dict(split.TIMES)
split.TIMES.clear()


def purity(lab, truth):
    return float(sum(np.bincount(truth[lab == v]).max() for v in np.unique(lab)) / lab.size)


ch = cb.Chain(counts, d["s"], K=a.K, L=a.L)
ch.set_state(z, y)
ch.seed(12345, 0)
t3 = time.perf_counter()
res = split.run_chain_CG_with_splits(split.DeviceOpsCG(ch), counts, a.K, a.L, lambda it, r, G, M: (None,) * 4, rng,
                                     split.DeviceBackend(), maxIter=a.maxIter, stopIter=10, splitOnIter=10)
t4 = time.perf_counter()
ch.close()
print(json.dumps({"cells": d["nM"], "genes": d["nG"], "K": a.K, "L": a.L, "initializeSplitZ_s": t1 - t0,
                  "initializeSplitY_s": t2 - t1, "z_purity_after_init": purity(z, d["z"]), "y_purity_after_init": purity(y, d["y"]),
                  "iterations": int(res["completeLogLik"].size - 1), "chain_with_splits_s": t4 - t3,
                  "splits": [(m[0], m[1], m[2][:60]) for m in res["messages"]],
                  "logLik_first_last": [float(res["completeLogLik"][0]), float(res["completeLogLik"][-1])],
                  "z_purity_final": purity(res["z"], d["z"]), "y_purity_final": purity(res["y"], d["y"]),
                  "seconds_by_operation": {"initializeSplitZ": times_z, "initializeSplitY": times_y,
                                           "chain_with_splits": dict(split.TIMES)}}))
