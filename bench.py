#!/usr/bin/env python
"""bench.py -- celda_CG collapsed-Gibbs iterations/s on synthetic 10x-shaped counts (BASELINE.json metric).

One "step" = one chain iteration without split heuristics (R/celda_CG.R:543-602): y sweep -> rowSumByGroupChange
-> z sweep (Gibbs) -> colSumByGroupChange -> cCGCalcLL, on configs[1]: 100k cells x 20k genes, K=30, L=150.
N GPUs = N independent chains (celdaGridSearch partitioning, R/celdaGridSearch.R:290-391: one chain per GPU, no
data-path collective) => weak scaling; value = chains * steps / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm  (libcelda_cuda through the C ABI)
  python bench.py --impl reference ...                             CPU arm  (oracle port on the host cores)
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
    """The oracle (CPU port of the reference loops) on a bounded sample of one iteration; returns seconds per FULL
    iteration extrapolated from the sample (sweeps scale linearly in visited items) and the sample description."""
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
    desc = (f"y sweep on {gs}/{nG} genes ({t_y:.2f}s) + full rowSumByGroupChange ({t_rc:.2f}s) + z sweep on {cs}/{nM} "
            f"cells ({t_z:.2f}s) + colSumByGroupChange+LL ({t_cc:.2f}s); sweeps extrapolated linearly; decompose "
            f"{t_dec:.1f}s untimed; oracle -O2, 1 thread/chain, direct lgamma instead of the per-iteration lgbeta table")
    return full, desc


def config_of(a, d):
    return {"workload": f"configs[1]: celda_CG Gibbs, {d['nM']} cells x {d['nG']} genes, K={a.K}, L={a.L}, S={a.S}, "
                        f"simulateCells-shaped counts NRange=({a.nmin},{a.nmax}), nnz={int(d['x'].size)}, "
                        f"start={a.start}", "cells": int(d["nM"]), "genes": int(d["nG"]), "K": a.K, "L": a.L,
            "nnz": int(d["x"].size), "l2": "inputs_exceed_l2 (CSC + nTSByC = %.1f GB per step)" %
            ((8 * d["x"].size + 4 * a.L * d["nM"]) / 1e9), "parallelism": "one independent chain per GPU (grid search)"}


def run_reference(a):
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
    desc = ""
    for step in range(a.warmup + a.steps):
        ixy, uy, ixz, uz = draw_inputs(rng, 1, d["nG"], d["nM"])

        def one(_):
            t = {k: (v.copy(order="F") if isinstance(v, np.ndarray) else v) for k, v in tables.items()}
            return cpu_iteration_sample(d, a, z, y, ixy[0], uy[0], ixz[0], uz[0], 1000, 200, t)

        t0 = time.perf_counter()
        with ThreadPoolExecutor(chains) as ex:
            res = list(ex.map(one, range(chains)))
        full = max(r[0] for r in res)
        desc = res[0][1]
        if step >= a.warmup:
            per_step.append(full)
    sec = float(np.mean(per_step))
    value = chains / sec
    print(json.dumps({
        "impl": "reference", "metric": "celda_CG_gibbs_iterations_per_sec", "value": value, "unit": "iter/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(a, d),
        "cells_iters_per_sec": value * d["nM"],
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": chains, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


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

    out = {"metric": "celda_CG_gibbs_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(a, d),
           "cells_iters_per_sec": value * nM, "clocks": clk, "gpu_launches": int(launches),
           "e2e": {"value": e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                   "ms_per_step": ms_e2e / K},
           "logLik_first_last": [lls[0], lls[-1]]}
    if rank == 0:
        # ---- roofline of the dominant kernel (z sweep: FP64-pipe bound, tables in shared memory)
        pz, py = prof["sweep_z"], prof["sweep_y"]
        peak = fp64_peak_tflops(cb)
        z_ms = pz["ms"] / max(1, pz["launches"])
        z_flop = FLOP_PER_LGAMMA * (float(nM) * a.K * a.L + 2.0 * nM * a.K)  # algorithmic: C*K*L sums + 2 scalars/cand
        ach = z_flop / (z_ms * 1e-3) / 1e12
        # DRAM traffic per launch from the committed ncu --set full capture of this same command
        # (profiles/r1_ncu_full_sweeps_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum), default config only
        default_cfg = (a.cells, a.genes, a.K, a.L) == (100000, 20000, 30, 150)
        out["roofline"] = {
            "bound": "fp64", "kernel": "k_sweep_z", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak if peak else None,
            "traffic": (150.641920 + 75.385344) * 1e6 if default_cfg else None,
            "traffic_note": "bytes; algorithmic HBM bytes are 4*L*C (nTSByC) + 4*L*C (visiting-order copy, written and read) "
                            "+ 8*K*C (probs) = 144 MB; the rest is the per-(cell, population) partial sums and the round's "
                            "value tables spilling from L2",
            "note": "achieved = ALGORITHMIC flops (C*K*L + 2*C*K lgamma at 56 flop, the reference's work) / device time; "
                    "peak = DFMA rate measured live (2 flop/FMA; MEASURED_PEAKS.json has no FP64 entry).  In large rounds "
                    "the kernel serves most lgamma terms from shared-memory value tables (bit-identical values), so the "
                    "FP64 pipe itself is only ~10 % busy (ncu) and the limit is issue / shared-memory and L2 latency; "
                    "ms_per_launch includes the two small helper kernels that lay the counts out in visiting order",
            "ms_per_launch": z_ms}
        y_ms = py["ms"] / max(1, py["launches"])
        y_flop = FLOP_PER_LGAMMA * (float(nG) * a.L * a.K + 6.0 * nG * a.L)
        out["roofline_y"] = {"bound": "fp64", "kernel": "k_sweep_y", "achieved": y_flop / (y_ms * 1e-3) / 1e12, "peak": peak,
                             "unit": "TFLOP/s", "frac": y_flop / (y_ms * 1e-3) / 1e12 / peak if peak else None,
                             "traffic": (11.512064 + 4.040192) * 1e6 if default_cfg else None, "ms_per_launch": y_ms,
                             "note": "unfused bit-reproducible arithmetic: 0.5 of the FMA peak is the ceiling for evaluated "
                                     "terms; 3 of 4 columns read lgamma(k + beta) from a 32 MB table in L2 instead (same "
                                     "bits), ncu: fp64 pipe 21 % busy, barrier stalls dominate: ~11 movers per sweep each "
                                     "cost a grid-wide round whose length is one full draw (tens of microseconds of "
                                     "dependent FP64)"}
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
        for kname, kern in (("rowSumByGroup", "k_rowsum_csc<int>"), ("colSumByGroup", "k_colsum_twin")):
            ms1 = agg[kname]["ms"] / max(1, agg[kname]["launches"])
            out["roofline_aggregation"][kname] = {
                "bound": "hbm", "kernel": kern, "achieved": bytes_alg[kname] / (ms1 * 1e-3) / 1e9, "peak": hbm,
                "unit": "GB/s", "frac": bytes_alg[kname] / (ms1 * 1e-3) / 1e9 / hbm, "traffic": None,
                "peak_source": "measured" if peaks else "fallback", "ms_per_launch": ms1,
                "algorithmic_bytes": bytes_alg[kname], "where": "decomposition (set_state), outside the timed step"}
        out["kernel_ms_per_step"] = {k: v["ms"] / K for k, v in prof.items()}
        out["sweep_stats_per_step"] = {k: {"rounds": prof[k]["rounds"] / K, "units": prof[k]["units"] / K,
                                           "movers": prof[k]["movers"] / K, "carried_redraws": prof[k]["redraws"] / K,
                                           "cta0_mclk": {q: prof[k]["clk_" + q] / K / 1e6 for q in
                                                         ("commit", "units", "draw", "barrier")}}
                                       for k in ("sweep_y", "sweep_z")}
        if world == 1 and not a.no_cpu_baseline:
            c_ixy, c_uy, c_ixz, c_uz = draw_inputs(rng, 1, nG, nM)
            sec, desc = cpu_iteration_sample(d, a, zc, yc, c_ixy[0], c_uy[0], c_ixz[0], c_uz[0], 4000, 800)
            out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "iter/s", "cores": 1, "kind": "port", "sample": desc}
        print(json.dumps(out))
    ch.close()
    if dist is not None:
        dist.destroy_process_group()


def fp64_peak_tflops(cb):
    import ctypes as C
    if not hasattr(cb.lib(), "celda_cuda_fp64_peak"):
        return None
    v = C.c_double(0)
    cb.check(cb.lib().celda_cuda_fp64_peak(C.byref(v)))
    return v.value


if __name__ == "__main__":
    main()
