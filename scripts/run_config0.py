#!/usr/bin/env python
"""configs[0]: celda_CG on the reference's celdaCGSim fixture (100 genes x 425 cells, S=5, K=5, L=10), Gibbs,
maxIter = 200 -- the case that runs on the CPU today.  Split heuristics / 'split' initialisation are next rows
(SURVEY 8f), so chains start from random labels; reports wall time, iterations run and the best log-likelihood next
to the reference model's stored finalLogLik."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celda_b200 as cb  # noqa: E402
from celda_b200 import model  # noqa: E402

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "celdaCG.npz")))
A = cb.CSC.from_dense(g["counts"])
cb.Chain(A, g["s"], K=5, L=10).close()  # context + kernels warm
t0 = time.perf_counter()
res = model.celda_CG(A, g["s"], K=5, L=10, algorithm="Gibbs", maxIter=200, stopIter=10, nchains=3, seed=12345)  # default "split" initialisation
wall = time.perf_counter() - t0
iters = [len(c["completeLogLik"]) - 1 for c in res["chains"]]
best = max(c["finalLogLik"] for c in res["chains"])
ari_z = None
print(json.dumps({"config": "configs[0]: celda_CG(celdaCGSim, K=5, L=10, algorithm='Gibbs', maxIter=200, nchains=3), "
                            "random init, no split heuristics, Philox draws",
                  "wall_s": wall, "iterations_per_chain": iters, "ms_per_iteration": 1e3 * wall / max(1, sum(iters)),
                  "finalLogLik_chains": [c["finalLogLik"] for c in res["chains"]], "best": best,
                  "reference_finalLogLik": float(g["finalLogLik"]),
                  "truth_logLik": None, "gpu_launches": cb.launch_count()}))
