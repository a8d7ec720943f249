#!/usr/bin/env python
"""Debug aid: sub-phase clocks of the sweeps' settle/commit phase (CTA 0).  Needs the instrumented build
   nvcc ... -DCELDA_COMMIT_PROF -shared -o celda_b200/libcelda_cuda_dbg.so celda_cuda.cu  (flags as in csrc/Makefile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celda_b200._lib as L
L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), "libcelda_cuda_dbg.so")
import argparse
import bench, celda_b200 as cb
ap = argparse.ArgumentParser(); ap.add_argument("--start", default="random"); a0 = ap.parse_args()
a = argparse.Namespace(S=10, cells=100000, genes=20000, K=30, L=150, nmin=1000, nmax=10000, start=a0.start)
d = bench.make_data(a)
z0, y0, rng = bench.start_state(d, a, 12345)
ch = cb.Chain(cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"]), d["s"], K=a.K, L=a.L)
ch.set_state(z0, y0)
ixy, uy, ixz, uz = bench.draw_inputs(rng, 1, d["nG"], d["nM"])
ch.profile(True)
ch.iterate(ixy[0], uy[0], ixz[0], uz[0])
pr = ch.profile_get()
names = ["first_invalid", "standing+dcarry", "select", "records+CL", "post-images"]
for sw, sub in (("sweep_y", "rowSumByGroupChange"), ("sweep_z", "colSumByGroupChange")):
    n = max(1, pr[sw]["rounds"])
    v = [pr[sub][k] for k in ("rounds", "units", "movers", "clk_commit", "clk_units")]
    print(sw, "rounds", n, "ms", round(pr[sw]["ms"], 1), "commit kclk/round", round(pr[sw]["clk_commit"] / n / 1e3, 1),
          {nm: round(x / n / 1e3, 2) for nm, x in zip(names, v)})
