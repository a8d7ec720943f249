#!/usr/bin/env python
"""Debug aid: which sampler paths the sweeps take and how long the general path's segments run (device clocks).
Needs the instrumented build: nvcc ... -DCELDA_DEBUG_STATS -shared -o celda_b200/libcelda_cuda_dbg.so celda_cuda.cu
(see celda_b200/csrc/Makefile for the flags); the product library carries no counters."""
import ctypes as C, sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celda_b200._lib as L
L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), "libcelda_cuda_dbg.so")
import celda_b200 as cb
from bench_support import simulate_cells_cg
d = simulate_cells_cg(12345, 10, 10000, 20000, 30, 150, (1000, 10000))
counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
rng = np.random.Generator(np.random.PCG64(4))
z, y = d["z"].copy(), d["y"].copy()
fz, fy = rng.random(d["nM"]) < 0.1, rng.random(d["nG"]) < 0.1
z[fz] = rng.integers(1, 31, int(fz.sum())); y[fy] = rng.integers(1, 151, int(fy.sum()))
ch = cb.Chain(counts, d["s"], K=30, L=150)
ch.set_state(z, y); ch.seed(1)
lib = L.lib()
out = (C.c_longlong * 8)()
for it in range(6):
    ch.sweep_y_philox()
    lib.celda_cuda_debug_counters(out, 1); ycnt = list(out)
    lib.celda_cuda_debug_segments(out, 1); print("   y general-path segment clocks per draw:", [round(v / max(1, ycnt[2])) for v in list(out)[:6]], "rank sort clocks per rank-sorted draw", round(out[6] / max(1, ycnt[3])))
    ch.sweep_z_gibbs_philox()
    lib.celda_cuda_debug_segments(out, 1)
    lib.celda_cuda_debug_counters(out, 1); zcnt = list(out)
    print(it, "Y draws/compact/general/ranksort/exact", ycnt[:5], "maxclk,sumclk,n>20k", ycnt[5:], "Z", zcnt[:5], flush=True)
