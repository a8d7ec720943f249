"""recursiveSplitCell / recursiveSplitModule (R/recursiveSplit.R:1-150, 452-870, 1250-1650): warm-start orchestration
over the chain drivers of celda_b200.model -- a model with K (or L) clusters is initialised from the K-1 model by
splitting one cluster in two with a small K = 2 (L = 2) sub-chain, the split with the best log-likelihood wins.

The control flow below restates the reference's, quirks included, and is written against a small `runner`
interface (DeviceRunner: libcelda_cuda chains; the CPU test plugs in a scripted runner), so the bookkeeping --
which clusters are tried, how labels are rewritten, which K the next round uses, what a degenerate model is
replaced with -- is testable without a GPU.

Reference behaviours kept on purpose:
  * .singleSplitZ / .singleSplitY call the sub-chain with DEFAULT hyper-parameters (alpha/beta/... are not passed,
    R/recursiveSplit.R:13-20, 53-60) and score the candidate on the caller's matrix with the caller's priors;
  * the celda_C branches pass `zInit = tempSplit$z` together with `zInitialize = "random"` (:684-697, :786-799), and
    .celda_C ignores zInit unless zInitialize == "predefined" (R/celda_C.R:377-397): the K-cluster chain starts from
    random labels and the split only decides the fallback labels when the chain loses a population;
  * in the tempL branch the first model's log-likelihood is computed with K = (its populations + 1) (:663-670).
Not reproduced: `reorder = TRUE` (hierarchical re-numbering of the clusters; partitions are unaffected) and R's
session RNG -- sub-chain seeds come from a numpy generator seeded with `seed`.
"""
import numpy as np

from . import CSC, MODEL_C, MODEL_CG, MODEL_G, Chain
from . import model as _model
from . import split as _split


def _n_unique(v):
    return int(np.unique(np.asarray(v)).size)


class DeviceRunner:
    """The model fits and scores recursiveSplit* needs, on libcelda_cuda (no CPU fallback)."""

    def __init__(self, device=0):
        self.device = device

    def celda_C(self, counts, s, K, **kw):
        if s is None:
            s = np.ones(counts.ncol, dtype=np.int32)
        return _model.celda_C(counts, s, K, device=self.device, **kw)["model"]

    def celda_G(self, counts, L, **kw):
        return _model.celda_G(counts, L, device=self.device, **kw)["model"]

    def celda_CG(self, counts, s, K, L, **kw):
        if s is None:
            s = np.ones(counts.ncol, dtype=np.int32)
        return _model.celda_CG(counts, s, K, L, device=self.device, **kw)["model"]

    # .logLikelihoodcelda_C / _G / _CG: decompose from the labels, then the collapsed log-likelihood
    def ll_C(self, counts, s, z, K, alpha, beta):
        ch = Chain(counts, s, K=K, alpha=alpha, beta=beta, model=MODEL_C, device=self.device)
        try:
            ch.set_state(z=z)
            return ch.loglik()
        finally:
            ch.close()

    def ll_G(self, counts, y, L, beta, delta, gamma):
        ch = Chain(counts, None, L=L, beta=beta, delta=delta, gamma=gamma, model=MODEL_G, device=self.device)
        try:
            ch.set_state(y=y)
            return ch.loglik()
        finally:
            ch.close()

    def ll_CG(self, counts, s, z, y, K, L, alpha, beta, delta, gamma):
        ch = Chain(counts, s, K=K, L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma, model=MODEL_CG,
                   device=self.device)
        try:
            ch.set_state(z, y)
            return ch.loglik()
        finally:
            ch.close()

    def rowSumByGroup(self, counts, y, L):
        from . import rowSumByGroupSparse, set_device
        set_device(self.device)
        return CSC.from_dense(np.asarray(rowSumByGroupSparse(counts, y, L)))

    def colSumByGroup(self, counts, z, K):
        from . import colSumByGroupSparse, set_device
        set_device(self.device)
        return CSC.from_dense(np.asarray(colSumByGroupSparse(counts, z, K)))

    def initializeSplitY(self, counts, L, tempK, minFeature, rng):
        return _split.initializeSplitY(counts, L, tempK=tempK, minFeature=minFeature, rng=rng,
                                       backend=_split.DeviceBackend(self.device))

    def initializeSplitZ(self, counts, K, minCell, rng):
        return _split.initializeSplitZ(counts, K, minCell=minCell, rng=rng, backend=_split.DeviceBackend(self.device))

    def perplexity(self, counts, s, mod):
        """perplexity of one model on the counts it was summarised for (R/perplexity.R:449-498, doResampling = FALSE)."""
        kind = MODEL_CG if ("z" in mod and "y" in mod and mod.get("y") is not None and mod.get("z") is not None) else (
            MODEL_C if mod.get("z") is not None else MODEL_G)
        p = mod["params"]
        ch = Chain(counts, s if kind != MODEL_G else None, K=p.get("K", 1), L=p.get("L", 1), alpha=p.get("alpha", 1.0),
                   beta=p.get("beta", 1.0), delta=p.get("delta", 1.0), gamma=p.get("gamma", 1.0), model=kind,
                   device=self.device)
        try:
            ch.set_state(z=mod.get("z") if kind != MODEL_G else None, y=mod.get("y") if kind != MODEL_C else None)
            return ch.perplexity()
        finally:
            ch.close()


def _seed(rng):
    return int(rng.integers(1, 2**31 - 1))


def singleSplitZ(runner, counts, z, s, K, minCell=3, alpha=1.0, beta=1.0, rng=None):
    """.singleSplitZ (R/recursiveSplit.R:1-35): try to split every population with more than `minCell` cells in two
    with a K = 2 celda_C sub-chain on its columns; the new population gets label K.  Returns dict(ll, z) of the best
    candidate (ll = -inf and z unchanged when no split produced two groups)."""
    z = np.asarray(z, dtype=np.int32)
    zTa = np.bincount(z, minlength=K + 1)[1:K + 1]
    bestZ, bestLl = z, -np.inf
    for i in (np.flatnonzero(zTa > minCell) + 1):
        ix = z == i
        clust = runner.celda_C(_split.csc_columns(counts, ix), None, 2, zInitialize="random", splitOnIter=-1,
                               splitOnLast=False, seed=_seed(rng))
        cz = np.asarray(clust["z"])
        if _n_unique(cz) == 2:
            newZ = z.copy()
            newZ[ix] = np.where(cz == 2, i, K).astype(np.int32)
            ll = runner.ll_C(counts, s, newZ, K, alpha, beta)
            if ll > bestLl:
                bestZ, bestLl = newZ, ll
    return dict(ll=bestLl, z=bestZ)


def singleSplitY(runner, counts, y, L, minFeature=3, beta=1.0, delta=1.0, gamma=1.0, rng=None):
    """.singleSplitY (R/recursiveSplit.R:38-73): the same over modules with an L = 2 celda_G sub-chain (one chain) on the
    module's rows; the new module gets label L."""
    y = np.asarray(y, dtype=np.int32)
    yTa = np.bincount(y, minlength=L + 1)[1:L + 1]
    bestY, bestLl = y, -np.inf
    for i in (np.flatnonzero(yTa > minFeature) + 1):
        ix = y == i
        clust = runner.celda_G(_split.csc_rows(counts, ix), 2, yInitialize="random", splitOnIter=-1, splitOnLast=False,
                               nchains=1, seed=_seed(rng))
        cy = np.asarray(clust["y"])
        if _n_unique(cy) == 2:
            newY = y.copy()
            newY[ix] = np.where(cy == 2, i, L).astype(np.int32)
            ll = runner.ll_G(counts, newY, L, beta, delta, gamma)
            if ll > bestLl:
                bestY, bestLl = newY, ll
    return dict(ll=bestLl, y=bestY)


def _model_C(z, K, alpha, beta, ll):
    return dict(z=np.asarray(z, np.int32), y=None, finalLogLik=float(ll), completeLogLik=np.asarray([ll]),
                params=dict(K=int(K), alpha=alpha, beta=beta))


def _with_params(mod, **params):
    mod = dict(mod)
    mod["params"] = dict(mod.get("params", {}), **params)
    return mod


def _finish(runner, counts, s, resList, keys, perplexity, doResampling, numResample, seed):
    runParams = []
    for k, mod in enumerate(resList):
        row = dict(index=k + 1, **{q: int(mod["params"][q]) for q in keys})
        row["log_likelihood"] = float(mod["finalLogLik"])
        runParams.append(row)
    out = dict(runParams=runParams, resList=resList)
    if perplexity:  # resamplePerplexity(counts, celdaRes, doResampling, numResample) (R/recursiveSplit.R:854-859)
        if doResampling:
            raise NotImplementedError("recursiveSplit*: doResampling = TRUE -- call model.resamplePerplexity on a chain of the "
                                      "model of interest")
        perp = [runner.perplexity(counts, s, m) for m in resList]
        out["perplexity"] = np.asarray(perp).reshape(-1, 1)
        for row, p in zip(runParams, perp):
            row["mean_perplexity"] = float(p)
    return out


def recursiveSplitCell(counts, sampleLabel=None, initialK=5, maxK=25, tempL=None, yInit=None, alpha=1.0, beta=1.0, delta=1.0,
                       gamma=1.0, minCell=3, seed=12345, perplexity=True, doResampling=False, numResample=5, device=0,
                       runner=None):
    """.recursiveSplitCell (R/recursiveSplit.R:452-870) on a CSC count matrix.  Three forms, as in the reference:
    `yInit` given -> celda_CG models on the module labels (cells are split on the collapsed module x cell matrix);
    `tempL` given -> celda_C models on counts collapsed to tempL temporary modules (.initializeSplitY);
    neither -> celda_C models on the counts themselves.  Returns dict(runParams, resList[, perplexity])."""
    runner = runner or DeviceRunner(device)
    rng = np.random.Generator(np.random.PCG64(seed))
    nM = counts.ncol
    s = np.ones(nM, dtype=np.int32) if sampleLabel is None else _split.as_factor_int(sampleLabel)
    if yInit is not None:
        L = _n_unique(yInit)
        overallY = _model.initialize_cluster(L, counts.nrow, yInit)
        countsY = runner.rowSumByGroup(counts, overallY, L)
        first = runner.celda_CG(counts, s, int(initialK), L, zInitialize="split", yInitialize="predefined", nchains=1,
                                yInit=overallY, alpha=alpha, beta=beta, gamma=gamma, delta=delta, seed=_seed(rng))
        first = _with_params(first, K=int(initialK), L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma)
        currentK = _n_unique(first["z"]) + 1
        overallZ = np.asarray(first["z"], np.int32)
        resList = [first]
        while currentK <= maxK:
            tempSplit = singleSplitZ(runner, countsY, overallZ, s, currentK, minCell=3, alpha=alpha, beta=beta, rng=rng)
            temp = runner.celda_CG(counts, s, int(currentK), L, yInit=overallY, zInit=tempSplit["z"], nchains=1,
                                   zInitialize="predefined", yInitialize="predefined", splitOnLast=False, stopIter=5,
                                   alpha=alpha, beta=beta, gamma=gamma, delta=delta, seed=_seed(rng))
            temp = _with_params(temp, K=int(currentK), L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma)
            if _n_unique(temp["z"]) == currentK:
                overallZ = np.asarray(temp["z"], np.int32)
            else:  # the chain lost a population: keep the split's labels with the chain's modules (:572-603)
                overallZ = np.asarray(tempSplit["z"], np.int32)
                ty = np.asarray(temp["y"], np.int32)
                ll = runner.ll_CG(counts, s, overallZ, ty, currentK, L, alpha, beta, delta, gamma)
                temp = dict(z=overallZ, y=ty, finalLogLik=float(ll), completeLogLik=np.asarray([ll]),
                            params=dict(K=int(currentK), L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma))
            resList.append(temp)
            currentK = _n_unique(overallZ) + 1
        return _finish(runner, counts, s, resList, ("L", "K"), perplexity, doResampling, numResample, seed)
    if tempL is not None:
        tempY = _split.as_factor_int(runner.initializeSplitY(counts, int(tempL), max(100, maxK), 3, rng))
        L = _n_unique(tempY)  # some modules may be empty (:635-636)
        work = runner.rowSumByGroup(counts, tempY, L)
    else:
        work = counts
    first = runner.celda_C(work, s, int(initialK), zInitialize="split", nchains=1, alpha=alpha, beta=beta, seed=_seed(rng))
    first = _with_params(first, K=int(initialK), alpha=alpha, beta=beta)
    currentK = _n_unique(first["z"]) + 1
    overallZ = np.asarray(first["z"], np.int32)
    if tempL is not None:  # log-likelihood on the full counts, with K = currentK as the reference writes it (:663-672)
        ll = runner.ll_C(counts, s, overallZ, currentK, alpha, beta)
        first = dict(first, finalLogLik=float(ll), completeLogLik=np.asarray([ll]))
    resList = [first]
    while currentK <= maxK:
        tempSplit = singleSplitZ(runner, work, overallZ, s, currentK, minCell=3, alpha=alpha, beta=beta, rng=rng)
        # zInitialize = "random": .celda_C ignores zInit (see the module docstring)
        temp = runner.celda_C(work, s, int(currentK), nchains=1, zInitialize="random", alpha=alpha, beta=beta, stopIter=5,
                              splitOnLast=False, zInit=tempSplit["z"], seed=_seed(rng))
        temp = _with_params(temp, K=int(currentK), alpha=alpha, beta=beta)
        lost = _n_unique(temp["z"]) != currentK
        overallZ = np.asarray(tempSplit["z"] if lost else temp["z"], np.int32)
        if tempL is not None or lost:  # (:700-733, :801-823)
            temp = _model_C(overallZ, currentK, alpha, beta, runner.ll_C(counts, s, overallZ, currentK, alpha, beta))
        resList.append(temp)
        currentK = _n_unique(overallZ) + 1
    return _finish(runner, counts, s, resList, ("K",), perplexity, doResampling, numResample, seed)


def recursiveSplitModule(counts, initialL=10, maxL=100, tempK=100, zInit=None, sampleLabel=None, alpha=1.0, beta=1.0,
                         delta=1.0, gamma=1.0, minFeature=3, seed=12345, perplexity=True, doResampling=False, numResample=5,
                         device=0, runner=None):
    """.recursiveSplitModule (R/recursiveSplit.R:1250-1650).  `zInit` given -> celda_CG models with the cell labels fixed
    as initial values (modules are split on the gene x population matrix); `tempK` given (the default, 100) -> celda_G
    models on counts collapsed to tempK temporary populations (.initializeSplitZ), log-likelihood re-computed on the
    full counts; tempK = None -> celda_G models on the counts themselves (first model maxIter = 20)."""
    runner = runner or DeviceRunner(device)
    rng = np.random.Generator(np.random.PCG64(seed))
    nM = counts.ncol
    s = np.ones(nM, dtype=np.int32) if sampleLabel is None else _split.as_factor_int(sampleLabel)
    if zInit is not None:
        K = _n_unique(zInit)
        overallZ = _model.initialize_cluster(K, nM, zInit)
        countsZ = runner.colSumByGroup(counts, overallZ, K)
        first = runner.celda_CG(counts, s, K, int(initialL), zInitialize="predefined", yInitialize="split", nchains=1,
                                zInit=overallZ, alpha=alpha, beta=beta, gamma=gamma, delta=delta, seed=_seed(rng))
        first = _with_params(first, K=K, L=int(initialL), alpha=alpha, beta=beta, delta=delta, gamma=gamma)
        currentL = _n_unique(first["y"]) + 1
        overallY = np.asarray(first["y"], np.int32)
        resList = [first]
        while currentL <= maxL:
            tempSplit = singleSplitY(runner, countsZ, overallY, currentL, minFeature=3, beta=beta, delta=delta, gamma=gamma,
                                     rng=rng)
            # hyper-parameters are NOT forwarded here (:1336-1349): the K x L chain runs with the defaults
            temp = runner.celda_CG(counts, None, K, int(currentL), stopIter=5, splitOnIter=-1, splitOnLast=False, nchains=1,
                                   yInitialize="predefined", zInitialize="predefined", yInit=tempSplit["y"], zInit=overallZ,
                                   seed=_seed(rng))
            temp = _with_params(temp, K=K, L=int(currentL), alpha=1.0, beta=1.0, delta=1.0, gamma=1.0)
            overallY = np.asarray(temp["y"], np.int32)
            resList.append(temp)
            currentL += 1
        return _finish(runner, counts, s, resList, ("L", "K"), perplexity, doResampling, numResample, seed)
    if tempK is not None:
        z = _split.as_factor_int(runner.initializeSplitZ(counts, int(tempK), 3, rng))
        work = runner.colSumByGroup(counts, z, _n_unique(z))
        first = runner.celda_G(work, int(initialL), nchains=1, seed=_seed(rng))
    else:
        work = counts
        first = runner.celda_G(work, int(initialL), maxIter=20, nchains=1, seed=_seed(rng))
    first = _with_params(first, L=int(initialL), beta=1.0, delta=1.0, gamma=1.0)
    overallY = np.asarray(first["y"], np.int32)
    currentL = _n_unique(overallY) + 1
    resList = [first]
    while currentL <= maxL:
        tempSplit = singleSplitY(runner, work, overallY, currentL, minFeature=3, beta=beta, delta=delta, gamma=gamma, rng=rng)
        temp = runner.celda_G(work, int(currentL), stopIter=5, splitOnIter=-1, splitOnLast=False, nchains=1,
                              yInitialize="predefined", yInit=tempSplit["y"], seed=_seed(rng))
        temp = _with_params(temp, L=int(currentL), beta=1.0, delta=1.0, gamma=1.0)
        overallY = np.asarray(temp["y"], np.int32)
        if tempK is not None:
            # the reference keeps nTSByC of the FULL counts current with .cGReDecomposeCounts and evaluates .cGCalcLL on
            # it (:1430-1462); decomposing from the labels gives the same integer tables
            ll = runner.ll_G(counts, overallY, currentL, beta, delta, gamma)
            temp = dict(temp, finalLogLik=float(ll), completeLogLik=np.asarray([ll]))
        resList.append(temp)
        currentL += 1
    return _finish(runner, counts, s, resList, ("L",), perplexity, doResampling, numResample, seed)
