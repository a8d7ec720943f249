#!/usr/bin/env python
"""bench.py -- celda_CG collapsed-Gibbs iterations/s on synthetic 10x-shaped counts (BASELINE.json metric).

One "step" = one chain iteration without split heuristics (R/celda_CG.R:543-602): y sweep -> rowSumByGroupChange
-> z sweep (Gibbs) -> colSumByGroupChange -> cCGCalcLL, on configs[1]: 100k cells x 20k genes, K=30, L=150.
N GPUs = N independent chains (celdaGridSearch partitioning, R/celdaGridSearch.R:290-391: one chain per GPU, no
data-path collective) => weak scaling; value = chains * steps / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm  (libcelda_cuda through the C ABI)
  python bench.py --impl reference ...                             CPU arm  (oracle port on the host cores)

Besides the timed steady-state steps the line carries `regimes` (the same chain under churn: first iteration from
the 10 %-perturbed truth, the first iterations and a whole 50-iteration chain from uniform-random labels) and, under
torchrun (N > 1), two more legs that exercise the multi-GPU paths of SURVEY.md 8(e): `em_sharded` (configs[2],
cells sharded over the ranks, NCCL all-reduce of the tables inside libcelda_cuda, checked against the unsharded
chain) and `grid_search` (configs[4] subset, runs pulled from a shared queue, one resident chain per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# FP64 operations per lgamma(x > 10) in the shared algorithm (celda_math.cuh / oracle o_lgamma), a division
# counted as one: log 28 (1 div) + Stirling correction 23 (2 div) + assembly 5.
FLOP_PER_LGAMMA = 56


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--K", type=int, default=30)
    ap.add_argument("--L", type=int, default=150)
    ap.add_argument("--S", type=int, default=10)
    ap.add_argument("--nmin", type=int, default=1000)
    ap.add_argument("--nmax", type=int, default=10000)
    ap.add_argument("--start", default="truth10", choices=["truth10", "random"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the em_sharded / grid_search legs under torchrun")
    ap.add_argument("--em-cells", type=int, default=1000000)
    ap.add_argument("--grid-cells", type=int, default=50000)
    ap.add_argument("--g-cells", type=int, default=500000, help="cells of the configs[3] celda_G leg (N = 1 only)")
    ap.add_argument("--grid-iters", type=int, default=8)
    return ap.parse_args()


def make_data(a):
    from bench_support import simulate_cells_cg
    t = time.time()
    d = simulate_cells_cg(12345, a.S, a.cells // a.S, a.genes, a.K, a.L, (a.nmin, a.nmax))
    d["gen_s"] = time.time() - t
    return d


def start_state(d, a, chain_seed):
    """Labels from the generator's truth with 10 % of the entries re-drawn (steady-state regime), or uniform random."""
    rng = np.random.Generator(np.random.PCG64(chain_seed))
    z, y = d["z"].copy(), d["y"].copy()
    if a.start == "random":
        z = rng.integers(1, a.K + 1, z.size).astype(np.int32)
        y = rng.integers(1, a.L + 1, y.size).astype(np.int32)
    else:
        fz, fy = rng.random(z.size) < 0.1, rng.random(y.size) < 0.1
        z[fz] = rng.integers(1, a.K + 1, int(fz.sum()))
        y[fy] = rng.integers(1, a.L + 1, int(fy.sum()))
    return z, y, rng


def draw_inputs(rng, n, nG, nM):
    ixy = np.stack([rng.permutation(nG).astype(np.int32) + 1 for _ in range(n)])
    ixz = np.stack([rng.permutation(nM).astype(np.int32) + 1 for _ in range(n)])
    return ixy, rng.random((n, nG)), ixz, rng.random((n, nM))


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).
    NVML is polled from a thread every ~2 ms (the timed region is tens of milliseconds, too short for
    `nvidia-smi -lms`); if NVML cannot be loaded, `nvidia-smi -lms 100` is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.h, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.mask, self.max_mhz = [], 0, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch
                pr = torch.cuda.get_device_properties(index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.stop_flag, self.sm, self.mask = False, [], 0
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join()
            nv = self.nv
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            reasons = [n for n in self.NAMES if self.mask & int(bits[n])]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for k, n in enumerate(self.NAMES) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def cpu_iteration_sample(d, a, z, y, ixy, uy, ixz, uz, sample_cells, sample_genes, tables=None):
    """The oracle (CPU port of the reference loops) on one iteration, or on a bounded sample of it; returns seconds
    per FULL iteration (sweeps scale linearly in visited items when a sample is extrapolated), the sample
    description and whether anything was extrapolated."""
    from oracle.oracle import Oracle
    o = Oracle()
    O = (d["nG"], d["nM"], d["p"], d["i"], d["x"])
    K, L, nG, nM = a.K, a.L, d["nG"], d["nM"]
    t0 = time.perf_counter()
    p = o.cCGDecomposeCounts(O, d["s"], z, y, K, L) if tables is None else tables
    t_dec = time.perf_counter() - t0
    gs, cs = min(sample_genes, nG), min(sample_cells, nM)
    t0 = time.perf_counter()
    oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0, 1.0, ixy,
                            uy, count=gs, want_probs=False)
    t_y = time.perf_counter() - t0
    t0 = time.perf_counter()
    nTSByC = o.rowSumByGroupChangeSparse(O, p["nTSByC"], oy["y"], y, L)
    t_rc = time.perf_counter() - t0
    t0 = time.perf_counter()
    oz = o.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, d["s"], K, L, nM, 1.0, 1.0, ixz,
                            uz, count=cs, want_probs=False)
    t_z = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.colSumByGroupChangeSparse(O, p["nGByCP"], oz["z"], z, K)
    o.cCGCalcLL(K, L, oz["mCPByS"], oz["nGByCP"], p["nByG"], oy["nByTS"], oy["nGByTS"], p["nS"], nG, 1.0, 1.0, 1.0, 1.0)
    t_cc = time.perf_counter() - t0
    full = t_y * nG / gs + t_rc + t_z * nM / cs + t_cc
    extrapolated = gs < nG or cs < nM
    desc = (f"y sweep on {gs}/{nG} genes ({t_y:.2f}s) + full rowSumByGroupChange ({t_rc:.2f}s) + z sweep on {cs}/{nM} "
            f"cells ({t_z:.2f}s) + colSumByGroupChange+LL ({t_cc:.2f}s); "
            + ("sweeps extrapolated linearly; " if extrapolated else "one FULL iteration, nothing extrapolated; ")
            + f"decompose {t_dec:.1f}s untimed; oracle -O2, 1 thread/chain, direct lgamma instead of the per-iteration "
              "lgbeta table; optimistic for the reference (no R interpreter)")
    return full, desc, extrapolated


def config_of(a, d):
    return {"workload": f"configs[1]: celda_CG Gibbs, {d['nM']} cells x {d['nG']} genes, K={a.K}, L={a.L}, S={a.S}, "
                        f"simulateCells-shaped counts NRange=({a.nmin},{a.nmax}), nnz={int(d['x'].size)}, "
                        f"start={a.start}", "cells": int(d["nM"]), "genes": int(d["nG"]), "K": a.K, "L": a.L,
            "nnz": int(d["x"].size), "l2": "inputs_exceed_l2 (CSC + nTSByC = %.1f GB per step)" %
            ((8 * d["x"].size + 4 * a.L * d["nM"]) / 1e9), "parallelism": "one independent chain per GPU (grid search)"}


def run_reference(a):
    """The reference's CPU path (oracle port: R is absent here) on the host cores, FULL iterations of the
    same workload, one chain per requested GPU (a Gibbs chain is sequential: one core each).  If W + K full
    iterations would not fit the budget (about five minutes) every step becomes a stated fraction of an iteration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = make_data(a)
    z, y, rng = start_state(d, a, 12345)
    from concurrent.futures import ThreadPoolExecutor
    chains = max(1, min(a.gpus, os.cpu_count() or 1))
    per_step = []
    from oracle.oracle import Oracle
    o = Oracle()
    tables = o.cCGDecomposeCounts((d["nG"], d["nM"], d["p"], d["i"], d["x"]), d["s"], z, y, a.K, a.L)
    # ~1.3e-4 s per cell and ~6e-5 s per gene-module row at K=30, L=150 on this class of host: budget 300 s
    est_full = 1.35e-4 * d["nM"] * (a.K * a.L / 4500.0) + 1.5
    frac = min(1.0, 300.0 / max(1e-9, est_full * (a.warmup + a.steps)))
    cs, gs = max(1000, int(d["nM"] * frac)), max(200, int(d["nG"] * frac))
    desc, extrap = "", False
    for step in range(a.warmup + a.steps):
        ixy, uy, ixz, uz = draw_inputs(rng, 1, d["nG"], d["nM"])

        def one(_):
            t = {k: (v.copy(order="F") if isinstance(v, np.ndarray) else v) for k, v in tables.items()}
            return cpu_iteration_sample(d, a, z, y, ixy[0], uy[0], ixz[0], uz[0], cs, gs, t)

        with ThreadPoolExecutor(chains) as ex:
            res = list(ex.map(one, range(chains)))
        full = max(r[0] for r in res)
        desc, extrap = res[0][1], res[0][2]
        if step >= a.warmup:
            per_step.append(full)
    sec = float(np.mean(per_step))
    value = chains / sec
    natives = reference_natives(d, a, z, y, tables)
    print(json.dumps({
        "impl": "reference", "metric": "celda_CG_gibbs_iterations_per_sec", "value": value, "unit": "iter/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(a, d),
        "cells_iters_per_sec": value * d["nM"],
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": chains, "kind": "port", "sample": desc,
                         "extrapolated": extrap},
        "reference_natives": natives,
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def reference_natives(d, a, z, y, tables):
    """The reference's OWN compiled sparse aggregation natives (oracle/_ref/libcelda_ref_cpp.so = src/matrixSumsSparse.cpp
    built unmodified, oracle/Makefile target `ref`) on the full workload: seconds per call on one host core, and whether
    their tables equal the port's.  None when the library was not built (no /root/reference at build time)."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libcelda_ref_cpp.so")
    if not os.path.exists(so):
        return None
    try:
        L = C.CDLL(so)
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        pp, ii, xx = (np.ascontiguousarray(d["p"], np.int32), np.ascontiguousarray(d["i"], np.int32),
                      np.ascontiguousarray(d["x"], np.float64))
        zz, yy = np.ascontiguousarray(z, np.int32), np.ascontiguousarray(y, np.int32)
        nG, nM = int(d["nG"]), int(d["nM"])
        rs, cs = np.zeros((a.L, nM), order="F"), np.zeros((nG, a.K), order="F")
        t0 = time.perf_counter()
        rc1 = L.ref_rowSumByGroupSparse(nG, nM, pp.ctypes.data_as(ip), ii.ctypes.data_as(ip), xx.ctypes.data_as(dp),
                                        yy.ctypes.data_as(ip), C.c_long(nG), a.L, rs.ctypes.data_as(dp))
        t1 = time.perf_counter()
        rc2 = L.ref_colSumByGroupSparse(nG, nM, pp.ctypes.data_as(ip), ii.ctypes.data_as(ip), xx.ctypes.data_as(dp),
                                        zz.ctypes.data_as(ip), C.c_long(nM), a.K, cs.ctypes.data_as(dp))
        t2 = time.perf_counter()
        if rc1 or rc2:
            return {"error": "status %d / %d" % (rc1, rc2)}
        # the y sweep's native part as the reference runs it (R/celda_G.R:602-660): the lgamma tables are rebuilt by every
        # call of the closure, then one call of the compiled cG_CalcGibbsProbY per gene (labels not updated here)
        from oracle.oracle import Oracle
        o = Oracle()
        t3 = time.perf_counter()
        lgb = o.lgamma(np.arange(0, int(tables["nCP"].max()) + 1, dtype=np.float64) + 1.0)
        lgg = o.lgamma(np.arange(0, nG + a.L + 1, dtype=np.float64) + 1.0)
        lgd = np.concatenate([[np.nan], o.lgamma(np.arange(1, nG + a.L + 1, dtype=np.float64) * 1.0)])
        t4 = time.perf_counter()
        cnt = np.ascontiguousarray(tables["nGByCP"], np.float64)  # row g = counts[g, ] as R passes it
        tt = np.asfortranarray(tables["nTSByCP"], np.float64)
        ts, gt = np.ascontiguousarray(tables["nByTS"], np.float64), np.ascontiguousarray(tables["nGByTS"], np.int32)
        bg, probs = np.ascontiguousarray(tables["nByG"], np.float64), np.zeros(a.L)
        f = L.ref_cG_CalcGibbsProbY
        fixed = (tt.ctypes.data_as(dp), ts.ctypes.data_as(dp), gt.ctypes.data_as(ip), bg.ctypes.data_as(dp),
                 yy.ctypes.data_as(ip), a.L, nG, lgb.ctypes.data_as(dp), C.c_long(lgb.size), lgg.ctypes.data_as(dp),
                 C.c_long(lgg.size), lgd.ctypes.data_as(dp), C.c_long(lgd.size), C.c_double(1.0), probs.ctypes.data_as(dp))
        base, stride = cnt.ctypes.data, cnt.strides[0]
        t5 = time.perf_counter()
        for g in range(nG):
            f(g + 1, C.cast(base + g * stride, dp), a.K, *fixed)
        t6 = time.perf_counter()
        return {"kind": "reference", "what": "src/matrixSumsSparse.cpp and src/cG_calcGibbsProbY.cpp compiled unmodified (oracle/_ref), one core, full workload",
                "y_sweep_native_part_s": {"lgamma_tables": t4 - t3, "cG_CalcGibbsProbY_calls": t6 - t5, "genes": nG,
                                          "note": "tables by the port's lgammafn; the R loop, sampling and table updates "
                                                  "around the calls are not included"},
                "rowSumByGroupSparse_s": t1 - t0, "colSumByGroupSparse_s": t2 - t1,
                "GB_per_s": [8.0 * xx.size / 1e9 / (t1 - t0), 8.0 * xx.size / 1e9 / (t2 - t1)],
                "tables_equal_port": bool(np.array_equal(rs, tables["nTSByC"]) and np.array_equal(cs, tables["nGByCP"]))}
    except Exception as e:  # a reported extra, never fatal for the arm
        return {"error": str(e)}


def run_regimes(cb, csc, d, a, device):
    """The same chain under churn (BENCH `regimes`): per-iteration device time, movers and rounds of (1) the first
    iterations from the truth with 10 % of the labels re-drawn, (2) the first iterations from uniform-random labels and
    (3) a whole 50-iteration chain from uniform-random labels (total)."""
    out = {}
    for name, start, iters in (("truth10", "truth10", 3), ("random", "random", 50)):
        b = argparse.Namespace(**vars(a))
        b.start = start
        z0, y0, rng = start_state(d, b, 424242)
        ch = cb.Chain(csc, d["s"], K=a.K, L=a.L, device=device)
        ch.set_state(z0, y0)
        ixy, uy, ixz, uz = draw_inputs(rng, iters, d["nG"], d["nM"])
        ch.stage_draws(ixy, uy, ixz, uz)
        rows = []
        for it in range(iters):
            ch.profile(True)
            ch.timer_start()
            ll = ch.iterate_staged(it)
            ms = ch.timer_stop()
            pr = ch.profile_get()
            ch.profile(False)
            rows.append({"ms": ms, "logLik": ll, "y_movers": pr["sweep_y"]["movers"], "z_movers": pr["sweep_z"]["movers"],
                         "y_rounds": pr["sweep_y"]["rounds"], "z_rounds": pr["sweep_z"]["rounds"],
                         "y_ms": pr["sweep_y"]["ms"], "z_ms": pr["sweep_z"]["ms"],
                         "discarded_commits": pr["sweep_y"]["discards"] + pr["sweep_z"]["discards"],
                         "carried_redraws": pr["sweep_y"]["redraws"] + pr["sweep_z"]["redraws"],
                         "cta0_mclk": {k: {q: pr[k]["clk_" + q] / 1e6 for q in ("commit", "units", "draw", "barrier")}
                                       for k in ("sweep_y", "sweep_z")}})
        ch.close()
        if name == "truth10":
            out["truth10_first_iteration_ms"] = rows[0]["ms"]
            out["truth10_iterations"] = rows
        else:
            out["random_start_ms_per_iter"] = [r["ms"] for r in rows[:4]]
            out["random_start_iterations"] = rows[:6]
            out["random_start_50_iterations_total_ms"] = float(sum(r["ms"] for r in rows))
            out["random_start_50_iterations_iter_per_s"] = 50.0 / (sum(r["ms"] for r in rows) / 1e3)
            out["random_start_50_iterations_movers"] = [int(r["y_movers"] + r["z_movers"]) for r in rows]
            mv = sum(r["y_movers"] + r["z_movers"] for r in rows[:4])
            out["random_start_us_per_mover_first4"] = 1e3 * sum(r["ms"] for r in rows[:4]) / max(1, mv)
    out["movers_per_step"] = {"truth10_first": int(out["truth10_iterations"][0]["y_movers"] + out["truth10_iterations"][0]["z_movers"]),
                              "random_first4": [int(r["y_movers"] + r["z_movers"]) for r in out["random_start_iterations"][:4]]}
    return out


def leg_em_sharded(cb, a, rank, world, local, dist, torch):
    """configs[2] on the N ranks of this launch: celda_C EM z step, cells sharded, one NCCL all-reduce of the G x K and
    K x S tables per iteration inside libcelda_cuda; rank 0 also runs the unsharded chain and compares."""
    from bench_support import simulate_cells_c
    from celda_b200 import shard
    cells, G, K, S, iters, warm = a.em_cells, 5000, 50, 10, 6, 2
    per = cells // world
    lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else cells
    dd = simulate_cells_c(12345, S, cells // S, G, K, (500, 3000), cell_range=(lo, hi))
    counts = cb.CSC(dd["nG"], dd["nM"], dd["p"], dd["i"], dd["x"])
    rng = np.random.Generator(np.random.PCG64(999))
    z = dd["z"].copy()  # full-length labels: the 10 % re-draw is made on the whole vector so that shards agree
    flip = rng.random(z.size) < 0.1
    z[flip] = rng.integers(1, K + 1, int(flip.sum()))
    comm = None
    if world > 1:
        ids = [shard.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = shard.Comm(ids[0], world, rank, local)
        ch = shard.ShardedChainC(counts, dd["s"][lo:hi], K, comm, nS=S, lo=lo).chain
    else:
        ch = cb.Chain(counts, dd["s"], K=K, model=cb.MODEL_C, device=local)
    ch.set_state(z=z[lo:hi])
    lls = [ch.iterate_em() for _ in range(warm)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ch.profile(True)
    ch.timer_start()
    for _ in range(iters):
        lls.append(ch.iterate_em())
    ms = ch.timer_stop()
    prof = ch.profile_get()
    ch.profile(False)
    Z = float(counts.nnz)
    if world > 1:
        t = torch.tensor([ms, Z], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        ms, Z = float(tm[0].item()), float(t[1].item())
    tab = ch.tables()["nGByCP"] if rank == 0 else None
    ch.close()
    if comm is not None:
        comm.close()
    out = None
    if rank == 0:
        em_ms = prof["sweep_z"]["ms"] / max(1, prof["sweep_z"]["launches"])
        out = {"workload": f"configs[2]: celda_C EM z step, {cells} cells x {G} genes, K={K}, nnz={int(Z)}, cells sharded "
                           f"over {world} GPU(s), one NCCL all-reduce of the G x K and K x S int32 tables per iteration",
               "ms_per_step": ms / iters, "iter_per_s": iters / (ms / 1e3), "em_z_step_ms_rank0": em_ms,
               "logLik_last": lls[-1], "scaling": "strong"}
        if world > 1:
            full = simulate_cells_c(12345, S, cells // S, G, K, (500, 3000))
            one = cb.Chain(cb.CSC(full["nG"], full["nM"], full["p"], full["i"], full["x"]), full["s"], K=K,
                           model=cb.MODEL_C, device=local)
            one.set_state(z=z)
            ref = [one.iterate_em() for _ in range(warm + iters)]
            reft = one.tables()["nGByCP"]
            one.timer_start()
            one.iterate_em()
            out["one_gpu_ms_per_step"] = one.timer_stop()
            one.close()
            out["ll_equal"] = bool(all(x == y for x, y in zip(lls, ref)))
            out["nGByCP_equal"] = bool(np.array_equal(tab, reft))
    if world > 1:
        dist.barrier()
    return out


def leg_grid_search(cb, a, rank, world, local, dist, torch):
    """configs[4] (declared subset): celdaGridSearch celda_CG Gibbs over K x L x chains on 50k cells x 20k genes; every
    rank keeps one resident chain and pulls the next run from a shared counter (the process group's store)."""
    from bench_support import simulate_cells_cg
    from celda_b200 import model
    dd = simulate_cells_cg(12345, 10, a.grid_cells // 10, a.genes, 30, 150, (1000, 10000))
    counts = cb.CSC(dd["nG"], dd["nM"], dd["p"], dd["i"], dd["x"])
    params = {"K": [5, 10, 15, 20, 25, 30, 35, 40, 45, 50], "L": [50, 100, 150, 200]}  # x 3 chains = 120 runs
    counter = [0]
    if world > 1:
        store = dist.distributed_c10d._get_default_store()

        def queue():
            return int(store.add("celda_grid_queue", 1)) - 1
    else:
        def queue():
            counter[0] += 1
            return counter[0] - 1

    def gather(local_res):
        slim = [{k: v for k, v in r.items() if k != "result"} | {"result": {"finalLogLik": r["logLikelihood"]}}
                for r in local_res]
        if world == 1:
            return [slim]
        out = [None] * world
        dist.all_gather_object(out, slim)
        return out

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = model.celdaGridSearch(counts, dd["s"], params, maxIter=a.grid_iters, nchains=3, seed=12345, algorithm="Gibbs",
                                rank=rank, nranks=world, device=local, gather=gather, queue=queue,
                                zInitialize="random", yInitialize="random")
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    if rank != 0:
        return None
    rp = res["runParams"]
    serial = float(sum(r["seconds"] for r in rp))
    per_rank = [sum(1 for r in rp if r["rank"] == q) for q in range(world)]
    return {"workload": f"configs[4] subset: celdaGridSearch celda_CG Gibbs, {dd['nM']} cells x {dd['nG']} genes, K in "
                        f"{params['K']} x L in {params['L']} x 3 chains = {len(rp)} runs, maxIter={a.grid_iters}, random "
                        "init, Philox draws; counts resident per GPU (celda_chain_reconfigure), shared work queue",
            "runs": len(rp), "wall_s": wall, "sum_of_run_seconds": serial, "speedup_vs_serial": serial / wall,
            "runs_per_rank": per_rank, "chains_per_s": len(rp) / wall,
            "logLikelihood_checksum": float(np.sum([r["logLikelihood"] for r in rp])),
            "failed_runs": int(sum(1 for r in rp if r.get("error"))),
            "note": "sum_of_run_seconds is what one GPU spends on the same runs one after the other (the N=1 line of the "
                    "scaling run reports it as wall_s); log-likelihoods are independent of N (checksum)"}


def leg_celda_g(cb, a, device, sm_mhz):
    """configs[3]: celda_G module-only y sweep on the raw counts, 20k genes x 500k cells, L = 200 (one GPU: a Gibbs chain
    does not shard).  Two sweeps from the generator's labels, exact full-column scan; the roofline is the sweep's own
    bound, shared-memory wavefronts (one 8-byte read of the cached lgamma per two dependent adds)."""
    from bench_support import simulate_cells_cg
    d = simulate_cells_cg(12345, 10, a.g_cells // 10, a.genes, 50, 200, (500, 3000))
    counts = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    L = 200
    ch = cb.Chain(counts, None, L=L, model=cb.MODEL_G, device=device)
    ch.set_state(y=d["y"].copy())
    ch.seed(12345)
    ch.profile(True)
    ch.timer_start()
    lls = [ch.iterate(None, None, None, None) for _ in range(2)]
    ms = ch.timer_stop() / 2
    pr = ch.profile_get()["sweep_y"]
    ch.close()
    nchunk = (L + 31) // 32
    wavefronts = 2.0 * nchunk * d["nG"] * d["nM"]          # one 256-byte warp read per (gene, 32-module chunk, cell)
    peak = 148 * (sm_mhz or 1965.0) * 1e6                   # one wavefront per clock and SM
    return {"workload": "configs[3]: celda_G y sweep, %d genes x %d cells, L=%d, nnz=%d, Philox draws, exact full-column scan"
                        % (d["nG"], d["nM"], L, int(d["x"].size)),
            "ms_per_sweep": ms, "iter_per_s": 1e3 / ms, "movers_per_sweep": pr["movers"] / 2, "rounds_per_sweep": pr["rounds"] / 2,
            "dependent_adds_per_s": 2.0 * d["nG"] * d["nM"] * L / (ms * 1e-3),
            "roofline": {"bound": "shared", "achieved": wavefronts / (ms * 1e-3) / 1e9, "peak": peak / 1e9, "unit": "Gwavefront/s",
                         "frac": wavefronts / (ms * 1e-3) / peak,
                         "note": "algorithmic wavefronts of the count-free columns only; a mover adds a pass over all cells"},
            "logLik_last": lls[-1]}


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch

    import celda_b200 as cb
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d = make_data(a)
    csc = cb.CSC(d["nG"], d["nM"], d["p"], d["i"], d["x"])
    nG, nM = d["nG"], d["nM"]
    z0, y0, rng = start_state(d, a, 12345 + rank)  # chain seeds 12345 + chain - 1 (R/celdaGridSearch.R:294)
    W, K = a.warmup, a.steps
    ch = cb.Chain(csc, d["s"], K=a.K, L=a.L, device=local)
    ch.set_state(z0, y0)

    # ------------------------------------------------------------ leg 1: inputs resident in HBM (value)
    ixy, uy, ixz, uz = draw_inputs(rng, W + K, nG, nM)
    ch.stage_draws(ixy, uy, ixz, uz)
    lls = [ch.iterate_staged(it) for it in range(W)]
    ch.profile(True)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = cb.launch_count()
    ch.timer_start()
    for it in range(K):
        lls.append(ch.iterate_staged(W + it))
    ms = ch.timer_stop()
    barrier()
    clk = clocks.stop()
    launches = cb.launch_count() - l0
    prof = ch.profile_get()
    ch.profile(False)
    ms = max_over_ranks(ms)
    value = world * K / (ms / 1e3)

    # ------------------------------------------------------------ leg 2: end to end from pinned host buffers
    hx = draw_inputs(rng, W + K, nG, nM)
    pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in hx]
    pixy, puy, pixz, puz = [t.numpy() for t in pinned]
    for it in range(W):
        ch.iterate(pixy[it], puy[it], pixz[it], puz[it])
    barrier()
    t0 = time.perf_counter()
    ch.timer_start()
    for it in range(W, W + K):
        lls.append(ch.iterate(pixy[it], puy[it], pixz[it], puz[it]))  # H2D of this step's draws, D2H of its logLik
    ms_e2e = ch.timer_stop()
    wall_e2e = time.perf_counter() - t0
    barrier()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e * 1e3))
    e2e = world * K / (ms_e2e / 1e3)
    h2d = 12 * (nG + nM)

    # ------------------------------------------------------------ leg 3: the production generator (device Philox)
    # same chain, no injected draws: the visiting orders (a keyed permutation, philox.cuh) and the uniforms are made
    # on the device, one kernel per sweep
    ph_err, ms_ph = None, float("nan")
    try:
        ch.seed(12345 + rank, 0)
        for it in range(W):
            ch.iterate_philox()
    except Exception as e:  # reported in the line, never fatal for the headline legs
        ph_err = str(e)
    barrier()
    l0 = cb.launch_count()
    ch.timer_start()
    try:
        for it in range(K):
            if ph_err is None:
                ch.iterate_philox()
    except Exception as e:
        ph_err = str(e)
    ms_ph = max_over_ranks(ch.timer_stop())
    philox = ({"error": ph_err} if ph_err else
              {"ms_per_step": ms_ph / K, "iter_per_s": world * K / (ms_ph / 1e3),
               "launches_per_step": (cb.launch_count() - l0) / K,
               "note": "celda_chain_iterate with NULL draws: the visiting orders (keyed Feistel permutation over Philox4x32-10, no "
                       "sort) and the uniforms are computed on the device inside the timed region; 8 B D2H per step"})

    out = {"metric": "celda_CG_gibbs_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(a, d),
           "cells_iters_per_sec": value * nM, "clocks": clk, "gpu_launches": int(launches),
           "e2e": {"value": e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                   "ms_per_step": ms_e2e / K},
           "logLik_first_last": [lls[0], lls[-1]], "philox_mode": philox}
    if rank == 0:
        default_cfg = (a.cells, a.genes, a.K, a.L, a.start) == (100000, 20000, 30, 150, "truth10")
        out.update(rooflines(cb, ch, a, d, prof, clk, default_cfg))
        out["kernel_ms_per_step"] = {k: v["ms"] / K for k, v in prof.items()}
        out["sweep_stats_per_step"] = {k: {"rounds": prof[k]["rounds"] / K, "units": prof[k]["units"] / K,
                                           "movers": prof[k]["movers"] / K, "carried_redraws": prof[k]["redraws"] / K,
                                           "discarded_commits": prof[k]["discards"] / K,
                                           "cta0_mclk": {q: prof[k]["clk_" + q] / K / 1e6 for q in
                                                         ("commit", "units", "draw", "barrier")}}
                                       for k in ("sweep_y", "sweep_z")}
        out["regime_of_value"] = ("steady state: the timed steps follow %d warm-up iterations from start=%s; see `regimes` "
                                  "for the same chain under churn" % (W, a.start))
        if world == 1 and not a.no_cpu_baseline:
            zc, yc = ch.get_state()
            c_ixy, c_uy, c_ixz, c_uz = draw_inputs(rng, 1, nG, nM)
            sec, desc, extrap = cpu_iteration_sample(d, a, zc, yc, c_ixy[0], c_uy[0], c_ixz[0], c_uz[0], nM, nG)
            out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "iter/s", "cores": 1, "kind": "port", "sample": desc,
                                   "extrapolated": extrap}
    ch.close()
    if rank == 0 and not a.no_regimes:
        out["regimes"] = run_regimes(cb, csc, d, a, local)
    del csc
    if not a.no_legs:
        em = leg_em_sharded(cb, a, rank, world, local, dist, torch)
        gs = leg_grid_search(cb, a, rank, world, local, dist, torch)
        if rank == 0:
            out["em_sharded"], out["grid_search"] = em, gs
        if world == 1:
            out["celda_g"] = leg_celda_g(cb, a, local, (out.get("clocks") or {}).get("sm_mhz"))
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def rooflines(cb, ch, a, d, prof, clk, default_cfg):
    """Roofline objects of the line.  The sweeps are neither HBM- nor tensor-bound: their tables live in shared memory
    and L2 and the arithmetic is a dependent chain per candidate, so the executed-work measure is ISSUE SLOTS: warp
    instructions executed per launch (from the committed ncu capture of this same command, profiles/r2_sweep_ncu.json)
    over the live-measured launch time, against SMs x 4 schedulers x the SM clock sampled during the timed region.
    The dominant kernel is whichever sweep is longer; the equivalent-flop figure of round 1 stays as a side key."""
    nG, nM = d["nG"], d["nM"]
    out = {}
    ncu = {}
    path = os.path.join(ROOT, "profiles", "r2_sweep_ncu.json")
    if default_cfg and os.path.exists(path):
        ncu = json.load(open(path))
    sms = 148
    try:
        import torch
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        pass
    mhz = clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0
    issue_peak = sms * 4 * mhz * 1e6 / 1e9  # G warp-instructions / s
    fp64 = fp64_peak_tflops(cb)
    alg = {"k_sweep_z": FLOP_PER_LGAMMA * (float(nM) * a.K * a.L + 2.0 * nM * a.K),
           "k_sweep_y": FLOP_PER_LGAMMA * (float(nG) * a.L * a.K + 6.0 * nG * a.L)}
    objs = {}
    for kern, key in (("k_sweep_z", "sweep_z"), ("k_sweep_y", "sweep_y")):
        ms1 = prof[key]["ms"] / max(1, prof[key]["launches"])
        n = ncu.get(kern, {})
        inst = n.get("inst_executed")
        ach = inst / (ms1 * 1e-3) / 1e9 if inst else None
        objs[kern] = {
            "bound": "issue", "kernel": kern, "achieved": ach, "peak": issue_peak, "unit": "Gwarp-inst/s",
            "frac": ach / issue_peak if ach else None, "traffic": n.get("dram_bytes"), "ms_per_launch": ms1,
            "executed": {"warp_instructions_per_launch": inst, "fp64_pipe_pct": n.get("fp64_pipe_pct"),
                         "issue_active_pct_ncu": n.get("issue_active_pct"),
                         "shared_wavefronts_per_launch": n.get("shared_wavefronts"),
                         "shared_bank_conflict_wavefronts": n.get("shared_bank_conflicts"),
                         "source": "profiles/r2_sweep_ncu.json (ncu --set full of this command)" if n else None},
            "algorithmic": {"flop_per_launch": alg[kern], "tflops": alg[kern] / (ms1 * 1e-3) / 1e12, "fp64_fma_peak_tflops": fp64,
                            "note": "the reference's lgamma count at 56 flop each / device time: most terms are served from "
                                    "shared-memory / L2 value tables (bit-identical values), so this is NOT executed work"},
            "note": "peak = SMs x 4 warp schedulers x SM clock during the timed region; ms_per_launch of k_sweep_z includes "
                    "its two helper kernels (visiting-order transpose, table sizing)"}
    dom = max(objs, key=lambda k: objs[k]["ms_per_launch"])
    out["roofline"] = objs[dom]
    out["roofline_other_sweep"] = objs[[k for k in objs if k != dom][0]]
    # ---- HBM-bound aggregation kernels: the full decomposition (set_state), timed separately from the step
    zc, yc = ch.get_state()
    ch.profile(True)
    for _ in range(5):
        ch.set_state(zc, yc)
    agg = ch.profile_get()
    ch.profile(False)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    Z = float(d["x"].size)
    bytes_alg = {"rowSumByGroup": 8 * Z + 4 * (nM + 1) + 4 * nG + 4.0 * a.L * nM,
                 "colSumByGroup": 8 * Z + 4 * (nM + 1) + 4 * nM + 4.0 * nG * a.K}
    out["roofline_aggregation"] = {}
    for kname, kern in (("rowSumByGroup", "k_rowsum_csc_v4"), ("colSumByGroup", "k_colsum_twin_v5")):
        ms1 = agg[kname]["ms"] / max(1, agg[kname]["launches"])
        out["roofline_aggregation"][kname] = {
            "bound": "hbm", "kernel": kern, "achieved": bytes_alg[kname] / (ms1 * 1e-3) / 1e9, "peak": hbm,
            "unit": "GB/s", "frac": bytes_alg[kname] / (ms1 * 1e-3) / 1e9 / hbm,
            "traffic": ncu.get(kern.split("<")[0], {}).get("dram_bytes"),
            "peak_source": "measured" if peaks else "fallback", "ms_per_launch": ms1,
            "algorithmic_bytes": bytes_alg[kname], "where": "decomposition (set_state), outside the timed step"}
    return out


def fp64_peak_tflops(cb):
    import ctypes as C
    if not hasattr(cb.lib(), "celda_cuda_fp64_peak"):
        return None
    v = C.c_double(0)
    cb.check(cb.lib().celda_cuda_fp64_peak(C.byref(v)))
    return v.value


if __name__ == "__main__":
    main()
