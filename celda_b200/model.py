"""Host-side chain drivers over the device-resident chain handle: the reference's `while` loops
(.celda_CG R/celda_CG.R:432-835, .celda_C R/celda_C.R:370-616) without the split heuristics, and the chain
partitioning of celdaGridSearch (R/celdaGridSearch.R:290-391) over GPUs: one chain per GPU at a time, no
collective on the data path.

Randomness: `draws="philox"` uses the device generator seeded with (seed, chain); `draws=callable` injects
(ix_y, u_y, ix_z, u_z) per iteration, which is how a run is reproduced seed for seed from R.
Label initialisation: "predefined" (zInit / yInit) follows the reference exactly; "random" restates
.initializeCluster (R/initialize_clusters.R:1-57) and "split" (the reference default) .initializeSplitZ / Y
(celda_b200/split.py) with numpy's generator (R's RNG stream is not available).
"""
import itertools

import numpy as np

from . import MODEL_C, MODEL_CG, MODEL_G, Chain


def initialize_cluster(N, length, initial=None, rng=None):
    """.initializeCluster(N, len, initial = ...) without `fixed` (R/initialize_clusters.R:1-57)."""
    if initial is not None:
        initial = np.asarray(initial)
        vals = np.unique(initial)
        if vals.size != N or initial.size != length or not np.all(np.isin(vals, np.arange(1, N + 1))):
            raise ValueError("'initial' needs to be a vector of length 'len' containing N unique values.")
        return (np.searchsorted(vals, initial) + 1).astype(np.int32)
    z = rng.integers(1, N + 1, length).astype(np.int32)
    for i in np.setdiff1d(np.arange(1, N + 1), z):
        cnt = np.bincount(z, minlength=N + 1)
        top = int(cnt.argmax())
        if cnt[top] == 1:
            raise ValueError("'len' is not long enough to accomodate 'N' unique values")
        ix = np.flatnonzero(z == top)
        z[rng.choice(ix)] = i
    return z


def best_chain(lls):
    """Index of the chain to return: the first one with the highest final log-likelihood
    (`result$finalLogLik > bestResult$finalLogLik`, R/celda_C.R:577-589).  Chains that converge to the same partition
    have the same log-likelihood up to the rounding of the reduction, so values within 1e-10 relative count as equal:
    the earliest of them is returned whatever the summation order."""
    top = max(lls)
    tol = 1e-10 * abs(top)
    return next(i for i, v in enumerate(lls) if v >= top - tol)


def _run_chain(ch, algorithm, maxIter, stopIter, draws, rng, nG, nM, model):
    """The iteration / best-so-far / stopping logic of R/celda_CG.R:521-755 (splitOnIter = -1, splitOnLast = FALSE)."""
    ll = [ch.loglik()]
    z, y = ch.get_state()
    zBest, yBest, llBest = z, y, ll[0]
    it, noImp = 1, 0
    while it <= maxIter and noImp <= stopIter:
        if draws == "philox":
            if algorithm == "Gibbs":  # celda_CG: y + z sweeps; celda_C: z sweep; celda_G: y sweep; then the LL
                temp = ch.iterate_philox()
            else:
                if model == MODEL_CG:
                    ch.sweep_y_philox()
                ch.sweep_z_em()
                temp = ch.loglik()
        else:
            ixy, uy, ixz, uz = draws(it, rng, nG, nM)
            temp = ch.iterate(ixy, uy, ixz, uz) if algorithm == "Gibbs" else ch.iterate_em(ixy, uy)
        if all(temp > v for v in ll) or it == 1:  # R/celda_CG.R:734
            zBest, yBest = ch.get_state()
            llBest = temp
            noImp = 1
        else:
            noImp += 1
        ll.append(temp)
        it += 1
    return dict(z=zBest, y=yBest, completeLogLik=np.asarray(ll), finalLogLik=llBest)


def _celda_CG_chain(ch, c, K, L, seed, algorithm, maxIter, stopIter, zInitialize, yInitialize, zInit, yInit, draws,
                    splitOnIter=-1, splitOnLast=False):
    """Chain number c (1-based) of .celda_CG on a resident handle already sized for (K, L) (R/celda_CG.R:432-790)."""
    nG, nM = ch.nG, ch.nM
    rng = np.random.Generator(np.random.PCG64(seed + c - 1))
    a, b, d, g = ch.hyper
    if zInitialize == "split":  # the reference default (R/celda_CG.R:463-469)
        from . import split
        z = split.initializeSplitZ(ch.counts, K, alpha=a, beta=b, rng=rng, backend=split.DeviceBackend(ch.device))
    else:
        z = initialize_cluster(K, nM, zInit if zInitialize == "predefined" else None, rng)
    if yInitialize == "split":  # R/celda_CG.R:487-493
        from . import split
        y = split.initializeSplitY(ch.counts, L, beta=b, delta=d, gamma=g, rng=rng,
                                   backend=split.DeviceBackend(ch.device))
    else:
        y = initialize_cluster(L, nG, yInit if yInitialize == "predefined" else None, rng)
    ch.set_state(z, y)
    ch.seed(seed, stream=c - 1)
    if splitOnIter > 0 or splitOnLast:  # the split / merge heuristic of R/celda_CG.R:604-716
        from . import split
        dr = (lambda it, r, G, M: (None, None, None, None)) if draws == "philox" else draws
        return split.run_chain_CG_with_splits(split.DeviceOpsCG(ch, algorithm), ch.counts, K, L, dr, rng,
                                              split.DeviceBackend(ch.device), maxIter=maxIter, stopIter=stopIter,
                                              splitOnIter=splitOnIter, splitOnLast=splitOnLast)
    return _run_chain(ch, algorithm, maxIter, stopIter, draws, rng, nG, nM, MODEL_CG)


def celda_CG(counts, sampleLabel, K, L, alpha=1.0, beta=1.0, delta=1.0, gamma=1.0, algorithm="EM", stopIter=10,
             maxIter=200, seed=12345, nchains=3, zInitialize="split", yInitialize="split", zInit=None, yInit=None,
             draws="philox", device=0, chain=None, splitOnIter=10, splitOnLast=True):
    """celda_CG(x, sampleLabel, K, L, ...) (R/celda_CG.R:85-261) on a CSC count matrix, with the reference's defaults:
    'split' initialisation and the split / merge heuristic every `splitOnIter` iterations and at a stall
    (celda_b200/split.py); splitOnIter = -1 and splitOnLast = False switch the heuristic off.
    Returns the per-chain results and `model`, built like the reference from the LAST chain (quirk Q1,
    R/celda_CG.R:796-811).  `chain`: a resident handle over the same counts and hyper-parameters to re-use
    (celdaGridSearch); it is re-sized to (K, L) and left open."""
    if algorithm not in ("EM", "Gibbs"):
        raise ValueError("'algorithm' must be 'EM' or 'Gibbs'")
    chains = []
    if chain is None:
        ch = Chain(counts, sampleLabel, K=K, L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma, model=MODEL_CG,
                   device=device)
    else:
        ch = chain
        if (ch.K, ch.L) != (int(K), int(L)):
            ch.reconfigure(K, L)
    try:
        for c in range(1, nchains + 1):
            chains.append(_celda_CG_chain(ch, c, K, L, seed, algorithm, maxIter, stopIter, zInitialize, yInitialize,
                                          zInit, yInit, draws, splitOnIter, splitOnLast))
    finally:
        if chain is None:
            ch.close()
    best = best_chain([c["finalLogLik"] for c in chains])
    return dict(chains=chains, model=chains[-1], bestChain=best + 1,
                params=dict(K=K, L=L, alpha=alpha, beta=beta, delta=delta, gamma=gamma, algorithm=algorithm, seed=seed))


def celda_C(counts, sampleLabel, K, alpha=1.0, beta=1.0, algorithm="EM", stopIter=10, maxIter=200, seed=12345,
            nchains=3, zInitialize="split", zInit=None, device=0, splitOnIter=10, splitOnLast=True):
    """celda_C (R/celda_C.R:306-616), algorithm = "EM" (forces stopIter <- 1, :351-353) or "Gibbs" (the z sweep on
    the raw counts, :427-440).  The best chain is returned (R/celda_C.R:577-589 does this correctly)."""
    if algorithm not in ("EM", "Gibbs"):
        raise ValueError("'algorithm' must be 'EM' or 'Gibbs'")
    if algorithm == "EM":
        stopIter = 1
    chains = []
    ch = Chain(counts, sampleLabel, K=K, alpha=alpha, beta=beta, model=MODEL_C, device=device)
    try:
        for c in range(1, nchains + 1):
            rng = np.random.Generator(np.random.PCG64(seed + c - 1))
            if zInitialize == "split":
                from . import split
                z = split.initializeSplitZ(counts, K, alpha=alpha, beta=beta, rng=rng, backend=split.DeviceBackend(device))
            else:
                z = initialize_cluster(K, counts.ncol, zInit if zInitialize == "predefined" else None, rng)
            ch.set_state(z=z)
            ch.seed(seed, stream=c - 1)
            if splitOnIter > 0 or splitOnLast:  # .cCSplitZ steps (R/celda_C.R:459-508)
                from . import split
                chains.append(split.run_chain_one_axis_with_splits(
                    split.DeviceOpsC(ch, algorithm), counts, K, "z", lambda it, r, G, M: (None, None, None, None), rng,
                    split.DeviceBackend(device), maxIter=maxIter, stopIter=stopIter, splitOnIter=splitOnIter,
                    splitOnLast=splitOnLast))
                continue
            chains.append(_run_chain(ch, algorithm, maxIter, stopIter, "philox", rng, counts.nrow, counts.ncol, MODEL_C))
    finally:
        ch.close()
    best = best_chain([c["finalLogLik"] for c in chains])
    return dict(chains=chains, model=chains[best], bestChain=best + 1,
                params=dict(K=K, alpha=alpha, beta=beta, algorithm=algorithm, seed=seed))


def celda_G(counts, L, beta=1.0, delta=1.0, gamma=1.0, stopIter=10, maxIter=200, seed=12345, nchains=3,
            yInitialize="split", yInit=None, device=0, draws="philox", splitOnIter=10, splitOnLast=True):
    """celda_G (R/celda_G.R:268-520): Gibbs y sweeps on the raw counts, split heuristics off; best chain returned
    (R/celda_G.R:483-494).  `draws`: "philox" or a callable (it, rng, nG, nM) -> (ix_y, u_y, None, None)."""
    chains = []
    ch = Chain(counts, None, L=L, beta=beta, delta=delta, gamma=gamma, model=MODEL_G, device=device)
    try:
        for c in range(1, nchains + 1):
            rng = np.random.Generator(np.random.PCG64(seed + c - 1))
            if yInitialize == "split":
                from . import split
                y = split.initializeSplitY(counts, L, beta=beta, delta=delta, gamma=gamma, rng=rng,
                                           backend=split.DeviceBackend(device))
            else:
                y = initialize_cluster(L, counts.nrow, yInit if yInitialize == "predefined" else None, rng)
            ch.set_state(y=y)
            ch.seed(seed, stream=c - 1)
            if splitOnIter > 0 or splitOnLast:  # .cGSplitY steps (R/celda_G.R:437-470)
                from . import split
                dr = (lambda it, r, G, M: (None, None, None, None)) if draws == "philox" else draws
                chains.append(split.run_chain_one_axis_with_splits(
                    split.DeviceOpsG(ch), counts, L, "y", dr, rng, split.DeviceBackend(device), maxIter=maxIter,
                    stopIter=stopIter, splitOnIter=splitOnIter, splitOnLast=splitOnLast))
                continue
            chains.append(_run_chain(ch, "Gibbs", maxIter, stopIter, draws, rng, counts.nrow, counts.ncol, MODEL_G))
    finally:
        ch.close()
    best = best_chain([c["finalLogLik"] for c in chains])
    return dict(chains=chains, model=chains[best], bestChain=best + 1,
                params=dict(L=L, beta=beta, delta=delta, gamma=gamma, seed=seed))


# ------------------------------------------------------------------ celdaGridSearch chain partitioning
def expand_grid(paramsTest, nchains):
    """base::expand.grid(chain = seq_len(nchains), paramsTest...) (R/celdaGridSearch.R:298-310): the first factor
    (chain) varies fastest.  Returns a list of dicts with a 1-based 'index'."""
    names = list(paramsTest)
    rows = []
    for combo in itertools.product(*[paramsTest[n] for n in reversed(names)]):
        vals = dict(zip(reversed(names), combo))
        for chain in range(1, nchains + 1):
            rows.append(dict(chain=chain, **{n: vals[n] for n in names}))
    for i, r in enumerate(rows):
        r["index"] = i + 1
    return rows


def chain_seed(seed, index, nchains):
    """allSeeds[ifelse(i %% nchains == 0, nchains, i %% nchains)] with allSeeds = seed + 0..nchains-1
    (R/celdaGridSearch.R:291-295,381-383)."""
    k = index % nchains
    return seed + (nchains if k == 0 else k) - 1


def partition_runs(runs, nranks, cost=lambda r: r.get("K", 1) * r.get("L", 1)):
    """Longest-processing-time-first assignment of grid-search runs to ranks (one chain per GPU at a time).
    Deterministic; returns a list (per rank) of run lists.  Cost defaults to K * L (work per sweep)."""
    order = sorted(range(len(runs)), key=lambda i: (-cost(runs[i]), i))
    load = [0.0] * nranks
    out = [[] for _ in range(nranks)]
    for i in order:
        r = min(range(nranks), key=lambda q: (load[q], q))
        out[r].append(runs[i])
        load[r] += cost(runs[i])
    return out


def run_order(runs, cost=lambda r: r.get("K", 1) * r.get("L", 1)):
    """Longest-first order of the grid-search runs (indices into `runs`), the order a shared work queue hands them out."""
    return sorted(range(len(runs)), key=lambda i: (-cost(runs[i]), i))


def celdaGridSearch(counts, sampleLabel, paramsTest, maxIter=200, nchains=3, seed=12345, algorithm="Gibbs",
                    rank=0, nranks=1, device=0, runner=None, gather=None, queue=None, isolate_failures=True,
                    bestOnly=False, **fixed):
    """celdaGridSearch(x, model = "celda_CG", paramsTest, nchains, ...) (R/celdaGridSearch.R:214-439): every
    (chain, K, L) run is an independent single-chain celda_CG on the SAME counts (:290-391).  Each rank keeps one
    resident chain handle (counts uploaded and gene-major twin built once) and re-sizes it per run
    (celda_chain_reconfigure).  Work distribution over `nranks` GPUs: `queue` -- a callable returning the next
    position (0-based, shared atomic counter, e.g. a torch.distributed store `add`) in the longest-first order, so a
    GPU pulls the next run when it finishes one -- or, without a queue, the static longest-processing-time partition.
    `runner(run, seed)` -> result dict can be injected (tests); `gather` all-gathers picklable objects across ranks.
    A run that raises is reported in its runParams row (`error`, logLikelihood = NaN) and the remaining runs go on with a
    fresh resident chain (`isolate_failures=False` re-raises, like a failed PSOCK worker fails the reference's search).
    `bestOnly=True` (the reference's default, R/celdaGridSearch.R:419-425) keeps the best chain of every parameter
    combination (selectBestModel); the default here returns every run, which is what the benchmark legs count."""
    import time
    runs = expand_grid(paramsTest, nchains)
    resident = {"ch": None}

    def default_runner(run, sd):
        if resident["ch"] is None:
            hyper = {k: fixed[k] for k in ("alpha", "beta", "delta", "gamma") if k in fixed}
            resident["ch"] = Chain(counts, sampleLabel, K=run["K"], L=run["L"], model=MODEL_CG, device=device, **hyper)
        res = celda_CG(counts, sampleLabel, run["K"], run["L"], maxIter=maxIter, nchains=1, seed=sd,
                       algorithm=algorithm, device=device, chain=resident["ch"], **fixed)
        return res["model"]

    runner = runner or default_runner
    if queue is None:
        mine = iter(partition_runs(runs, nranks)[rank])
    else:
        order = run_order(runs)

        def pull():
            while True:
                k = queue()
                if k >= len(order):
                    return
                yield runs[order[k]]

        mine = pull()
    local = []
    try:
        for run in mine:
            sd = chain_seed(seed, run["index"], nchains)
            t0 = time.perf_counter()
            try:
                res, err = runner(run, sd), None
            except Exception as e:  # one failed run does not take the grid down: it is reported, the others go on
                if not isolate_failures:
                    raise
                res, err = None, "%s: %s" % (type(e).__name__, e)
                if resident["ch"] is not None:  # the chain state behind a failed sweep is unspecified: start afresh
                    resident["ch"].close()
                    resident["ch"] = None
            local.append(dict(index=run["index"], chain=run["chain"], K=run.get("K"), L=run.get("L"), seed=sd,
                              logLikelihood=float(res["finalLogLik"]) if res is not None else float("nan"),
                              seconds=time.perf_counter() - t0, rank=rank, result=res, error=err))
    finally:
        if resident["ch"] is not None:
            resident["ch"].close()
    everything = [local] if gather is None else gather(local)
    flat = sorted((r for part in everything for r in part), key=lambda r: r["index"])
    out = dict(runParams=[{k: r[k] for k in ("index", "chain", "K", "L", "seed", "logLikelihood", "seconds", "rank", "error")}
                          for r in flat],
               resList=[r["result"] for r in flat])
    return selectBestModel(out, asList=True) if bestOnly else out


_NOT_PARAMS = ("index", "chain", "logLikelihood", "mean_perplexity", "seed", "seconds", "rank", "error")


def subsetCeldaList(res, params):
    """subsetCeldaList(x, params) (R/celdaGridSearch.R:546-581) on a celdaGridSearch result: keep the runs whose
    runParams match every entry of `params` (a value or a list of values per column); a single match returns that
    model, otherwise a result of the same shape."""
    rp = res["runParams"]
    cols = set().union(*(r.keys() for r in rp)) if rp else set()
    bad = [k for k in params if k not in cols]
    if bad:
        raise ValueError("The following elements in 'params' are not columns in runParams (x) " + ",".join(bad))
    keep = list(range(len(rp)))
    for k, want in params.items():
        want = set(np.atleast_1d(want).tolist())
        keep = [i for i in keep if rp[i][k] in want]
        if not keep:
            raise ValueError("No runs matched the criteria given in 'params'. Check 'runParams(x)' for complete list of "
                             "parameters used to generate 'x'.")
    if len(keep) == 1:
        return res["resList"][keep[0]]
    return dict(runParams=[rp[i] for i in keep], resList=[res["resList"][i] for i in keep])


def selectBestModel(res, asList=False):
    """selectBestModel(x, asList) (R/celdaGridSearch.R:668-687): for every combination of the tested parameters the
    chain with the highest log-likelihood (`which.max`: the first one on ties, failed runs (NaN) never win), groups in
    the order they first appear.  One combination and asList = False returns the model itself.  This is what
    celdaGridSearch(bestOnly = TRUE), the reference's default, returns."""
    rp = res["runParams"]
    groups = {}
    for i, r in enumerate(rp):
        key = tuple((k, r[k]) for k in r if k not in _NOT_PARAMS)
        ll = r["logLikelihood"]
        cur = groups.get(key)
        if cur is None:
            groups[key] = i
        elif not np.isnan(ll) and (np.isnan(rp[cur]["logLikelihood"]) or ll > rp[cur]["logLikelihood"]):
            groups[key] = i
    ix = list(groups.values())
    if len(ix) == 1 and not asList:
        return res["resList"][ix[0]]
    return dict(runParams=[rp[i] for i in ix], resList=[res["resList"][i] for i in ix])


# ------------------------------------------------------------------ cold callers of the hot path
def clusterProbability(chain, log=False):
    """clusterProbability(x) (R/clusterProbability.R:117-252): the sweeps with doSample = FALSE, then each row of
    probabilities normalised with the log-sum-exp trick (.normalizeLogProbs).  Returns dict(zProbability [nM, K],
    yProbability [nG, L]) for the chain's model; the chain state is left unchanged."""
    from . import MODEL_C as _C, MODEL_G as _G
    out = {}
    if chain.model != _G:
        ix = np.arange(1, chain.nM + 1, dtype=np.int32)
        if chain.model == _C:
            chain.sweep_z_em(doSample=False)
        else:
            chain.sweep_z_gibbs(ix, np.zeros(chain.nM), doSample=False)
    if chain.model != _C:
        chain.sweep_y(np.arange(1, chain.nG + 1, dtype=np.int32), np.zeros(chain.nG), doSample=False)
    pz, py = chain.probs()

    def norm(p):
        p = p.T  # items x candidates
        mx = p.max(axis=1, keepdims=True)
        lse = mx + np.log(np.exp(p - mx).sum(axis=1, keepdims=True))
        return p - lse if log else np.exp(p - lse)

    if pz is not None:
        out["zProbability"] = norm(pz)
    if py is not None:
        out["yProbability"] = norm(py)
    return out


def resamplePerplexity(chain, doResampling=False, numResample=5, seed=12345):
    """.resamplePerplexity for one model (R/perplexity.R:449-498): the default (doResampling = FALSE) is the
    perplexity on the counts themselves (one column); with resampling, `numResample` columns from multinomial
    resamples of the counts.  Returns (perplexities, their mean) like the celdaList's perplexity row and
    runParams$mean_perplexity."""
    if not isinstance(doResampling, bool):
        raise ValueError("The 'doResampling' parameter needs to be logical (TRUE/FALSE).")
    if doResampling and (not isinstance(numResample, (int, np.integer)) or numResample < 1):
        raise ValueError("The 'numResample' parameter needs to be an integer greater than 0.")
    perp = chain.perplexity_resampled(seed, numResample) if doResampling else np.array([chain.perplexity()])
    return perp, float(perp.mean())


def factorizeMatrix(chain, type=("counts", "proportion", "posterior")):
    """.factorizeMatrixCG (R/factorizeMatrix.R:193-299) from the chain's device-resident tables: counts /
    proportions / posterior of sample (K x S), cellPopulation (L x K), cell (L x C), module (G x L),
    geneDistribution (L, the +1 pseudo-gene included exactly like the reference)."""
    t = chain.tables()
    z, y = chain.get_state()
    a, b, d, g = chain.hyper
    G, L = chain.nG, chain.L
    GByTS = np.zeros((G, L))
    GByTS[np.arange(G), y - 1] = t["nByG"]
    nGByTS = t["nGByTS"].astype(np.float64)
    nGByTS[nGByTS == 0] = 1.0  # :221
    res = {}

    def colnorm(m):  # normalizeCounts(m, normalize = "proportion")
        return m / m.sum(axis=0, keepdims=True)

    if "counts" in type:
        res["counts"] = dict(sample=t["mCPByS"].astype(np.float64), cellPopulation=t["nTSByCP"], cell=t["nTSByC"],
                             module=GByTS, geneDistribution=nGByTS)
    if "proportion" in type:
        uz, uy = np.unique(z) - 1, np.unique(y) - 1  # only populated populations / modules are normalised (:255-266)
        cp, mod = t["nTSByCP"].copy(), GByTS.copy()
        cp[:, uz] = colnorm(cp[:, uz])
        mod[:, uy] = colnorm(mod[:, uy])
        res["proportions"] = dict(sample=colnorm(t["mCPByS"].astype(np.float64)), cellPopulation=cp,
                                  cell=colnorm(t["nTSByC"]), module=mod, geneDistribution=nGByTS / nGByTS.sum())
    if "posterior" in type:
        gs = GByTS.copy()
        gs[np.arange(G), y - 1] += d
        res["posterior"] = dict(sample=colnorm(t["mCPByS"] + a), cellPopulation=colnorm(t["nTSByCP"] + b),
                                module=colnorm(gs), geneDistribution=(nGByTS + g) / (nGByTS + g).sum())
    return res
