#!/usr/bin/env python
"""Debug aid (instrumented build -DCELDA_DEBUG_STATS -> celda_b200/libcelda_cuda_dbgs.so): sampler path counts and the
general path's segment clocks for the first y / z sweeps from uniform-random labels."""
import ctypes as C, sys, os, numpy as np, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celda_b200._lib as L
L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), "libcelda_cuda_dbgs.so")
import celda_b200 as cb, bench
a = argparse.Namespace(S=10, cells=100000, genes=20000, K=30, L=150, nmin=1000, nmax=10000, start="random")
d = bench.make_data(a)
z0, y0, rng = bench.start_state(d, a, 12345)
ch = cb.Chain(cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"]), d["s"], K=a.K, L=a.L)
ch.set_state(z0, y0); ch.seed(1)
lib = L.lib()
out = (C.c_longlong * 8)()
for it in range(2):
    ch.sweep_y_philox()
    lib.celda_cuda_debug_counters(out, 1); ycnt = list(out)
    lib.celda_cuda_debug_segments(out, 1); seg = list(out)
    print(it, "Y draws/compact/general/ranksort/exact", ycnt[:5], "segment clocks per general draw", [round(v / max(1, ycnt[2])) for v in seg[:7]], flush=True)
    ch.sweep_z_gibbs_philox()
    lib.celda_cuda_debug_counters(out, 1); zcnt = list(out)
    lib.celda_cuda_debug_segments(out, 1); seg = list(out)
    print(it, "Z draws/compact/general/ranksort/exact", zcnt[:5], "segment clocks per general draw", [round(v / max(1, zcnt[2])) for v in seg[:7]], flush=True)
