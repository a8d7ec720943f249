"""Extract golden vectors from the reference's own fixtures into small .npz files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Sources: data/celda{CG,C,G}Sim.rda (counts + labels),
data/celda{CG,C,G}Mod.rda (fitted z/y, finalLogLik, completeLogLik, params),
data/celdaCGGridSearchRes.rda (9 celda_CG models K 4-6 x L 9-11 with logLikelihood,
stored perplexity of resampled counts).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rda_reader import RObj, list_get, read_rda  # noqa: E402

REF = "/root/reference/data"


def _vec(ob, dtype):
    return np.asarray(ob.value, dtype=dtype)


def _matrix(ob, dtype):
    dim = ob.attr["dim"].value
    return np.asarray(ob.value, dtype=dtype).reshape(dim[1], dim[0]).T.copy()  # column-major -> [r, c]


def _factor_codes(ob):
    return np.asarray(ob.value, dtype=np.int32), [str(s) for s in ob.attr["levels"].value]


def _scalar(ob):
    return float(ob.value[0])


def _params(mod):
    p = mod.attr["params"]
    names = p.attr["names"].value
    out = {}
    for n, v in zip(names, p.value):
        if isinstance(v, RObj) and v.kind in ("real", "int") and len(v.value) == 1:
            out[n] = float(v.value[0])
    return out


def _sample_codes(sim_label, n):
    """sampleLabel in the Sim lists: factor or character vector -> 1-based int codes (as.integer(factor))."""
    if sim_label.kind == "int" and "levels" in sim_label.attr:
        return np.asarray(sim_label.value, dtype=np.int32)
    labels = [str(v) for v in sim_label.value]
    levels = sorted(set(labels))
    return np.asarray([levels.index(v) + 1 for v in labels], dtype=np.int32)


def main(outdir=HERE):
    """Writes the four .npz files into `outdir` (the committed ones live next to this script)."""
    out = {}
    # ---- celda_CG
    sim = read_rda(f"{REF}/celdaCGSim.rda")["celdaCGSim"]
    mod = read_rda(f"{REF}/celdaCGMod.rda")["celdaCGMod"]
    counts = _matrix(list_get(sim, "counts"), np.int32)
    cl = mod.attr["clusters"]
    s_mod = mod.attr["sampleLabel"]
    cg = dict(
        counts=counts,
        sim_z=_vec(list_get(sim, "z"), np.int32), sim_y=_vec(list_get(sim, "y"), np.int32),
        s=_sample_codes(list_get(sim, "sampleLabel"), counts.shape[1]),
        z=_vec(list_get(cl, "z"), np.int32), y=_vec(list_get(cl, "y"), np.int32),
        s_mod=_factor_codes(s_mod)[0] if "levels" in s_mod.attr else _sample_codes(s_mod, counts.shape[1]),
        finalLogLik=np.float64(_scalar(mod.attr["finalLogLik"])),
        completeLogLik=_vec(mod.attr["completeLogLik"], np.float64),
        **{f"param_{k}": np.float64(v) for k, v in _params(mod).items()},
    )
    np.savez_compressed(f"{outdir}/celdaCG.npz", **cg)
    out["celdaCG"] = cg
    # ---- celda_C
    sim = read_rda(f"{REF}/celdaCSim.rda")["celdaCSim"]
    mod = read_rda(f"{REF}/celdaCMod.rda")["celdaCMod"]
    counts = _matrix(list_get(sim, "counts"), np.int32)
    s_mod = mod.attr["sampleLabel"]
    c = dict(
        counts=counts, sim_z=_vec(list_get(sim, "z"), np.int32),
        s=_sample_codes(list_get(sim, "sampleLabel"), counts.shape[1]),
        z=_vec(list_get(mod.attr["clusters"], "z"), np.int32),
        s_mod=_factor_codes(s_mod)[0] if "levels" in s_mod.attr else _sample_codes(s_mod, counts.shape[1]),
        finalLogLik=np.float64(_scalar(mod.attr["finalLogLik"])),
        completeLogLik=_vec(mod.attr["completeLogLik"], np.float64),
        **{f"param_{k}": np.float64(v) for k, v in _params(mod).items()},
    )
    np.savez_compressed(f"{outdir}/celdaC.npz", **c)
    out["celdaC"] = c
    # ---- celda_G
    sim = read_rda(f"{REF}/celdaGSim.rda")["celdaGSim"]
    mod = read_rda(f"{REF}/celdaGMod.rda")["celdaGMod"]
    counts = _matrix(list_get(sim, "counts"), np.int32)
    g = dict(
        counts=counts, sim_y=_vec(list_get(sim, "y"), np.int32),
        y=_vec(list_get(mod.attr["clusters"], "y"), np.int32),
        finalLogLik=np.float64(_scalar(mod.attr["finalLogLik"])),
        completeLogLik=_vec(mod.attr["completeLogLik"], np.float64),
        **{f"param_{k}": np.float64(v) for k, v in _params(mod).items()},
    )
    np.savez_compressed(f"{outdir}/celdaG.npz", **g)
    out["celdaG"] = g
    # ---- grid search: 9 models on the celdaCGSim counts
    gs = read_rda(f"{REF}/celdaCGGridSearchRes.rda")["celdaCGGridSearchRes"]
    rp = gs.attr["runParams"]
    rp_names = rp.attr["names"].value
    cols = {n: np.asarray(v.value, dtype=np.float64) for n, v in zip(rp_names, rp.value)}
    res = gs.attr["resList"].value
    zs, ys, lls, Ks, Ls = [], [], [], [], []
    for m in res:
        zs.append(_vec(list_get(m.attr["clusters"], "z"), np.int32))
        ys.append(_vec(list_get(m.attr["clusters"], "y"), np.int32))
        lls.append(_scalar(m.attr["finalLogLik"]))
        pr = _params(m)
        Ks.append(pr["K"])
        Ls.append(pr["L"])
    perp = gs.attr["perplexity"]
    grid = dict(
        z=np.stack(zs), y=np.stack(ys), finalLogLik=np.asarray(lls), K=np.asarray(Ks, dtype=np.int32),
        L=np.asarray(Ls, dtype=np.int32), runParams_K=cols["K"], runParams_L=cols["L"],
        runParams_logLikelihood=cols["logLikelihood"], perplexity=_matrix(perp, np.float64),
        s=_factor_codes(res[0].attr["sampleLabel"])[0],
    )
    np.savez_compressed(f"{outdir}/celdaCGGrid.npz", **grid)
    out["grid"] = grid
    for k, v in out.items():
        print(k, {n: (a.shape if hasattr(a, "shape") and a.shape else a) for n, a in v.items()})


if __name__ == "__main__":
    main()
