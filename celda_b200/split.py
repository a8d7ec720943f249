"""'split' initialisation of the cluster labels -- the reference's default zInitialize / yInitialize
(.initializeSplitZ, .initializeSplitY: R/initialize_clusters.R:60-400; call sites R/celda_CG.R:463-493,
R/celda_C.R, R/celda_G.R).

The functions below restate the reference's control flow and are written against a small backend interface, so that
the same orchestration runs on the device (DeviceBackend: resident chain handles, EM / Gibbs kernels, the
decomposition and log-likelihood kernels) and, in the parity tests, on a CPU checker -- labels must agree exactly.

Randomness: the reference draws from R's session generator (random sub-chain initialisations, `sample()` of the
clusters to split, Gibbs draws); that stream cannot be replayed here, so a numpy generator supplies the same draws
in the same places.  Sub-matrices (`counts[, z == i]`, `counts[y == i, ]`) are copied into chains of their own.
"""
import numpy as np

from . import CSC, MODEL_C, MODEL_G, Chain
from . import model as _model


# seconds spent per kind of operation (diagnostics: scripts/run_default_pipeline.py prints them)
TIMES = {}


def _timed(kind):
    import functools
    import time

    def deco(fn):
        @functools.wraps(fn)
        def wrap(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                TIMES[kind] = TIMES.get(kind, 0.0) + time.perf_counter() - t0
                TIMES[kind + "_calls"] = TIMES.get(kind + "_calls", 0) + 1
        return wrap
    return deco


# ------------------------------------------------------------------ small helpers on the CSC container
@_timed("csc_columns")
def csc_columns(counts, keep):
    """counts[, keep, drop = FALSE] for a boolean mask (touches only the selected columns' entries)."""
    cols = np.flatnonzero(np.asarray(keep, dtype=bool))
    starts = counts.p[cols].astype(np.int64)
    lens = counts.p[cols + 1].astype(np.int64) - starts
    p = np.zeros(cols.size + 1, dtype=np.int64)
    np.cumsum(lens, out=p[1:])
    # positions of the selected entries: for column k the run starts[k] .. starts[k] + lens[k] - 1
    idx = np.arange(p[-1], dtype=np.int64) + np.repeat(starts - p[:-1], lens)
    return CSC(counts.nrow, cols.size, p.astype(np.int32), counts.i[idx], counts.x[idx])


@_timed("csc_rows")
def csc_rows(counts, keep, drop_empty_columns=False):
    """counts[keep, , drop = FALSE]; optionally also the columns whose sum over the kept rows is 0."""
    keep = np.asarray(keep, dtype=bool)
    remap = np.cumsum(keep) - 1
    entry = keep[counts.i]
    col_of = np.repeat(np.arange(counts.ncol), np.diff(counts.p))[entry]
    lens = np.bincount(col_of, minlength=counts.ncol)
    if drop_empty_columns:
        lens = lens[lens > 0]
    p = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=p[1:])
    return CSC(int(keep.sum()), lens.size, p.astype(np.int32), remap[counts.i[entry]].astype(np.int32), counts.x[entry])


def as_factor_int(v):
    """as.integer(as.factor(v)): relabel to 1..n in the order of the sorted distinct values."""
    return (np.unique(np.asarray(v), return_inverse=True)[1] + 1).astype(np.int32)


def _second_best(probs, labels):
    """zProb[cbind(seq(n), z)] <- NA; apply(zProb, 1, which.max): probs is [candidates, items]."""
    pr = np.array(probs, dtype=np.float64, copy=True)
    pr[labels - 1, np.arange(pr.shape[1])] = -np.inf
    return (pr.argmax(axis=0) + 1).astype(np.int32)  # first maximum on ties, like which.max


def _first_max(ll):
    """which.max(llShuffle).  Removing population i or j gives the same partition (and the same log-likelihood up to
    the rounding of the reduction) when each is the other's second choice, so values within 1e-10 relative count as
    equal and the lowest index wins whatever the summation order."""
    top = float(np.max(ll))
    return int(np.flatnonzero(ll >= top - 1e-10 * abs(top))[0])


def _draws_G(it, rng, nG, nM):
    return rng.permutation(nG).astype(np.int32) + 1, rng.random(nG), None, None


# ------------------------------------------------------------------ the device backend
class DeviceBackend:
    """The operations of the split initialisation on libcelda_cuda (no CPU fallback)."""

    def __init__(self, device=0):
        from . import set_device
        self.device = device
        set_device(device)  # the stateless entry points (colSumByGroup) run on the library's current device

    @_timed("sub_celda_C")
    def celda_C_labels(self, counts, K, alpha, beta, maxIter, seed):
        """.celda_C(counts, K, maxIter, zInitialize = "random", algorithm = "EM", nchains = 3): best chain's z."""
        s = np.ones(counts.ncol, dtype=np.int32)
        res = _model.celda_C(counts, s, K, alpha=alpha, beta=beta, algorithm="EM", maxIter=maxIter, seed=seed, nchains=3,
                             zInitialize="random", device=self.device, splitOnIter=-1, splitOnLast=False)
        return res["model"]["z"]

    @_timed("sub_celda_G")
    def celda_G_labels(self, counts, L, beta, delta, gamma, maxIter, seed):
        """.celda_G(counts, L, maxIter, yInitialize = "random", nchains = 3): best chain's y (injected Gibbs draws)."""
        res = _model.celda_G(counts, L, beta=beta, delta=delta, gamma=gamma, maxIter=maxIter, seed=seed, nchains=3,
                             yInitialize="random", device=self.device, draws=_draws_G, splitOnIter=-1, splitOnLast=False)
        return res["model"]["y"]

    def session_C(self, counts, alpha, beta):
        return _SessionC(counts, alpha, beta, self.device)

    def session_G(self, counts, beta, delta, gamma):
        return _SessionG(counts, beta, delta, gamma, self.device)

    @_timed("colSumByGroup")
    def colSumByGroup(self, counts, z, K):
        from . import colSumByGroupSparse
        return colSumByGroupSparse(counts, z, K)


class _SessionC:
    """One resident celda_C chain over `counts`: EM probabilities (doSample = FALSE) and log-likelihoods for label
    vectors with varying K (celda_chain_reconfigure)."""

    def __init__(self, counts, alpha, beta, device):
        self.counts, self.s = counts, np.ones(counts.ncol, dtype=np.int32)
        self.ch = Chain(counts, self.s, K=1, alpha=alpha, beta=beta, model=MODEL_C, device=device)

    def _state(self, z, K):
        if self.ch.K != K:
            self.ch.reconfigure(K=K)
        self.ch.set_state(z=z)

    def probs(self, z, K):
        self._state(z, K)
        self.ch.sweep_z_em(doSample=False)
        return self.ch.probs()[0]

    def ll(self, z, K):
        self._state(z, K)
        return self.ch.loglik()

    def close(self):
        self.ch.close()


class _SessionG:
    def __init__(self, counts, beta, delta, gamma, device):
        self.counts = counts
        self.ch = Chain(counts, None, L=1, beta=beta, delta=delta, gamma=gamma, model=MODEL_G, device=device)

    def _state(self, y, L):
        if self.ch.L != L:
            self.ch.reconfigure(L=L)
        self.ch.set_state(y=y)

    def probs(self, y, L):
        self._state(y, L)
        nG = self.counts.nrow
        self.ch.sweep_y(np.arange(1, nG + 1, dtype=np.int32), np.zeros(nG), doSample=False)
        return self.ch.probs()[1]

    def ll(self, y, L):
        self._state(y, L)
        return self.ch.loglik()

    def close(self):
        self.ch.close()


# ------------------------------------------------------------------ .initializeSplitZ
def initializeSplitZ(counts, K, KSubcluster=None, alpha=1.0, beta=1.0, minCell=3, rng=None, backend=None):
    """.initializeSplitZ (R/initialize_clusters.R:60-222): cluster into ceiling(sqrt(K)) populations, split every
    population that is large enough again, then remove populations one by one -- each time the one whose cells,
    moved to their second-best population, leave the highest log-likelihood -- until K remain."""
    rng = np.random.default_rng(12345) if rng is None else rng
    be = DeviceBackend() if backend is None else backend
    nM = counts.ncol
    KSub = int(np.ceil(np.sqrt(K))) if KSubcluster is None else int(KSubcluster)

    def seed():
        return int(rng.integers(1, 2**31 - 4))

    overallZ = as_factor_int(be.celda_C_labels(counts, min(KSub, nM), alpha, beta, 20, seed()))
    currentK = int(overallZ.max())
    counter = 0
    while currentK < K and counter < 25:
        KPerCluster = min(int(np.ceil(K / currentK)), KSub)
        KToUse = 2 if KPerCluster < 2 else KPerCluster
        zTa = np.bincount(overallZ, minlength=currentK + 1)[1:]
        zToSplit = np.flatnonzero((zTa > minCell) & (zTa > KToUse)) + 1
        if zToSplit.size > 1:
            zToSplit = rng.permutation(zToSplit)  # sample(zToSplit)
        elif zToSplit.size == 0:
            break
        for i in zToSplit:
            ix = overallZ == i
            tempZ = as_factor_int(be.celda_C_labels(csc_columns(counts, ix), KToUse, alpha, beta, 20, seed()))
            splitIx = tempZ > 1
            newZ = overallZ[ix]
            newZ[splitIx] = currentK + tempZ[splitIx] - 1
            overallZ[ix] = newZ
            currentK = int(overallZ.max())
            if currentK > K + 10:
                break
        counter += 1
    if currentK > K:
        ses = be.session_C(counts, alpha, beta)
        try:
            while currentK > K:
                zSecond = _second_best(ses.probs(overallZ, currentK), overallZ)
                zTa = np.bincount(overallZ, minlength=currentK + 1)[1:]
                llShuffle = np.full(currentK, -np.inf)
                for i in np.flatnonzero(zTa > 0) + 1:
                    newZ = overallZ.copy()
                    ix = overallZ == i
                    newZ[ix] = zSecond[ix]
                    llShuffle[i - 1] = ses.ll(newZ, currentK)
                zToRemove = _first_max(llShuffle) + 1
                ix = overallZ == zToRemove
                overallZ[ix] = zSecond[ix]
                overallZ = as_factor_int(overallZ)
                currentK -= 1
        finally:
            ses.close()
    return overallZ


# ------------------------------------------------------------------ .initializeSplitY
def initializeSplitY(counts, L, LSubcluster=None, tempK=100, beta=1.0, delta=1.0, gamma=1.0, minFeature=3, rng=None,
                     backend=None):
    """.initializeSplitY (R/initialize_clusters.R:225-400): collapse the cells to tempK pseudo-cells, cluster the
    genes into ceiling(sqrt(L)) modules with celda_G, split the large modules again, then remove modules one by one
    (second-best module by the Gibbs probabilities, highest log-likelihood after removal) until L remain."""
    rng = np.random.default_rng(12345) if rng is None else rng
    be = DeviceBackend() if backend is None else backend
    LSub = int(np.ceil(np.sqrt(L))) if LSubcluster is None else int(LSubcluster)

    def seed():
        return int(rng.integers(1, 2**31 - 4))

    if tempK is not None and counts.ncol > tempK:
        z = initializeSplitZ(counts, K=tempK, rng=rng, backend=be)
        collapsed = be.colSumByGroup(counts, z, int(np.unique(z).size))
        counts = CSC.from_dense(collapsed)
    nG = counts.nrow
    overallY = as_factor_int(be.celda_G_labels(counts, LSub, beta, delta, gamma, 10, seed()))
    currentL = int(overallY.max())
    counter = 0
    while currentL < L and counter < 25:
        yTa = np.bincount(overallY, minlength=currentL + 1)[1:]
        yToSplit = np.flatnonzero((yTa > minFeature) & (yTa > LSub)) + 1
        if yToSplit.size == 0:
            break
        yToSplit = rng.permutation(yToSplit)  # sample(which(...))
        for i in yToSplit:
            ix = overallY == i
            countsY = csc_rows(counts, ix, drop_empty_columns=True)
            if countsY.ncol == 0:
                continue
            tempY = as_factor_int(be.celda_G_labels(countsY, min(LSub, countsY.nrow), beta, delta, gamma, 20, seed()))
            splitIx = tempY > 1
            newY = overallY[ix]
            newY[splitIx] = currentL + tempY[splitIx] - 1
            overallY[ix] = newY
            currentL = int(overallY.max())
            if currentL > L + 10:
                break
        counter += 1
    if currentL > L:
        ses = be.session_G(counts, beta, delta, gamma)
        try:
            while currentL > L:
                ySecond = _second_best(ses.probs(overallY, currentL), overallY)
                yTa = np.bincount(overallY, minlength=currentL + 1)[1:]
                llShuffle = np.full(currentL, -np.inf)
                for i in np.flatnonzero(yTa > 0) + 1:
                    newY = overallY.copy()
                    ix = overallY == i
                    newY[ix] = ySecond[ix]
                    llShuffle[i - 1] = ses.ll(newY, currentL)
                yToRemove = _first_max(llShuffle) + 1
                ix = overallY == yToRemove
                overallY[ix] = ySecond[ix]
                overallY = as_factor_int(overallY)
                currentL -= 1
        finally:
            ses.close()
    assert overallY.size == nG
    return overallY


# ================================================================== split / merge heuristic (R/split_clusters.R)
def _order_decreasing(ll, n):
    """utils::head(order(ll, decreasing = TRUE, na.last = NA), n) with NA = -inf: indices (0-based) by decreasing value,
    the lower index first among values within 1e-10 relative (equal up to the rounding of the reduction)."""
    ll = np.array(ll, dtype=np.float64, copy=True)
    out = []
    while len(out) < n and np.isfinite(ll).any():
        k = _first_max(np.where(np.isfinite(ll), ll, -np.inf))
        out.append(k)
        ll[k] = -np.inf
    return out


def _split_heuristic(lab, N, min_size, sub_split, second, ll_of, max_try):
    """The body shared by .cCSplitZ, .cCGSplitZ, .cCGSplitY and .cGSplitY (R/split_clusters.R:2-763): every cluster with
    at least `min_size` members is split in two by a short sub-chain (`sub_split(i)` -> 1/2 labels of its members);
    the clusters whose members, moved to their second-best cluster, cost the least log-likelihood are candidates for
    removal; for each candidate i and each splittable j != i the proposal "dissolve i, split j, give j's second half
    the label i" is scored by the log-likelihood, and the best proposal (or the current labels) is returned."""
    ta = np.bincount(lab, minlength=N + 1)[1:]
    to_split = np.flatnonzero(ta >= min_size) + 1
    non_empty = np.flatnonzero(ta > 0) + 1
    if to_split.size == 0:
        return lab, "Cluster sizes too small. No additional splitting was performed."
    clust = {int(i): np.asarray(sub_split(int(i)), dtype=np.int32) for i in to_split}
    cands, lls = [lab], [ll_of(lab)]
    ll_shuffle = np.full(N, -np.inf)
    for i in non_empty:
        new = lab.copy()
        ix = lab == i
        new[ix] = second[ix]
        ll_shuffle[i - 1] = ll_of(new)
    pairs = [None]
    for i0 in _order_decreasing(ll_shuffle, max_try):
        i = i0 + 1
        for j in to_split:
            if j == i:
                continue
            new = lab.copy()
            move = lab == i
            new[move] = second[move]
            sp = lab == j
            new[sp] = np.where(clust[int(j)] == 1, j, i)
            cands.append(new)
            lls.append(ll_of(new))
            pairs.append((i, int(j)))
    sel = _first_max(np.asarray(lls))
    if sel == 0:
        return lab, "No additional splitting was performed."
    return cands[sel], "Cluster %d was reassigned and cluster %d was split in two." % pairs[sel]


def _sub_seed(rng):
    return int(rng.integers(1, 2**31 - 4))


def cCSplitZ(counts, z, K, zProb, ll_of, rng, backend, maxClustersToTry=10, minCell=3):
    """.cCSplitZ (R/split_clusters.R:2-150).  zProb: [K, nM] log-probabilities of the last z update; ll_of(z) -> .cCCalcLL."""
    second = _second_best(zProb, z)
    return _split_heuristic(
        z, K, minCell, lambda i: backend.celda_C_labels(csc_columns(counts, z == i), 2, 1.0, 1.0, 5, _sub_seed(rng)),
        second, ll_of, int(maxClustersToTry))


def cCGSplitZ(counts, z, K, zProb, ll_of, rng, backend, maxClustersToTry=10, minCell=3):
    """.cCGSplitZ (R/split_clusters.R:154-347): as .cCSplitZ, scored with .cCGCalcLL at fixed y."""
    return cCSplitZ(counts, z, K, zProb, ll_of, rng, backend, maxClustersToTry, minCell)


def cCGSplitY(counts, y, z, K, L, yProb, ll_of, rng, backend, maxClustersToTry=10, KSubclusters=10, minCell=3):
    """.cCGSplitY (R/split_clusters.R:350-588): the cells are first reduced to at most KSubclusters pseudo-cells per
    population (celda_C sub-chains), the module splits are found by celda_G on those pseudo-cells."""
    zTa = np.bincount(z, minlength=K + 1)[1:]
    tempZ = np.zeros(z.size, dtype=np.int32)
    top = 0
    for i in np.flatnonzero(zTa > 0) + 1:
        ix = z == i
        if zTa[i - 1] <= KSubclusters:
            tempZ[ix] = np.arange(top + 1, top + zTa[i - 1] + 1)
        else:
            tempZ[ix] = backend.celda_C_labels(csc_columns(counts, ix), KSubclusters, 1.0, 1.0, 5, _sub_seed(rng)) + top
        top = int(tempZ.max())
    pseudo = np.asarray(backend.colSumByGroup(counts, tempZ, top))  # G x top, dense
    second = _second_best(yProb, y)

    def sub(i):
        rows = CSC.from_dense(pseudo[y == i, :])
        return backend.celda_G_labels(rows, 2, 1.0, 1.0, 1.0, 5, _sub_seed(rng))

    return _split_heuristic(y, L, minCell, sub, second, ll_of, int(maxClustersToTry))


def cGSplitY(counts, y, L, yProb, ll_of, rng, backend, minFeature=3, maxClustersToTry=10):
    """.cGSplitY (R/split_clusters.R:592-763)."""
    second = _second_best(yProb, y)

    def sub(i):
        return backend.celda_G_labels(csc_rows(counts, y == i), 2, 1.0, 1.0, 1.0, 5, _sub_seed(rng))

    return _split_heuristic(y, L, minFeature, sub, second, ll_of, int(maxClustersToTry))


# ------------------------------------------------------------------ chain loop with the split steps
class DeviceOpsCG:
    """The chain operations the loop needs, on a resident celda_CG handle."""

    def __init__(self, ch, algorithm="Gibbs"):
        self.ch, self.algorithm = ch, algorithm

    def ll(self):
        return self.ch.loglik()

    def iterate(self, ixy, uy, ixz, uz):
        if self.algorithm == "Gibbs":
            return self.ch.iterate(ixy, uy, ixz, uz)  # all None = device Philox draws
        if ixy is None:
            self.ch.sweep_y_philox()
            self.ch.sweep_z_em()
            return self.ch.loglik()
        return self.ch.iterate_em(ixy, uy)

    def state(self):
        return self.ch.get_state()

    def set_state(self, z, y):
        self.ch.set_state(z, y)

    def probs(self):
        return self.ch.probs()

    @_timed("ll_alt_z")
    def ll_alt_z(self, z):
        """.cCGCalcLL with other population labels, the chain's modules and state unchanged (celda_chain_loglik_alt)."""
        return self.ch.loglik_alt(z=z)

    @_timed("ll_alt_y")
    def ll_alt_y(self, y):
        return self.ch.loglik_alt(y=y)


def run_chain_CG_with_splits(ops, counts, K, L, draws, rng, backend, maxIter=200, stopIter=10, splitOnIter=10,
                             splitOnLast=True):
    """The while loop of .celda_CG (R/celda_CG.R:536-755) with the split steps: Gibbs y sweep, z sweep, log-likelihood,
    module split (.cCGSplitY) and population split (.cCGSplitZ) every `splitOnIter` iterations or at a stall, the
    best-so-far bookkeeping and the stopping rule.  draws(it, rng, nG, nM) supplies the visiting orders and uniforms."""
    nG, nM = counts.nrow, counts.ncol
    ll = [ops.ll()]
    zBest, yBest = ops.state()
    llBest = ll[0]
    it, noImp, doCell, doGene = 1, 0, True, True
    messages = []
    while it <= maxIter and noImp <= stopIter:
        temp = ops.iterate(*draws(it, rng, nG, nM))
        z, y = ops.state()
        pz, py = ops.probs()
        if L > 2 and it != maxIter and ((noImp == stopIter and not all(temp >= v for v in ll) and splitOnLast) or
                                        (splitOnIter > 0 and it % splitOnIter == 0 and doGene)):
            newY, msg = cCGSplitY(counts, y, z, K, L, py, ops.ll_alt_y, rng, backend, maxClustersToTry=max(L / 2, 10),
                                  minCell=3)
            messages.append((it, "y", msg))
            if not np.array_equal(newY, y):
                noImp, doGene = 1, True
                y = newY
                ops.set_state(z, y)  # nTSByC <- rowSumByGroup(counts, y) etc. (R/celda_CG.R:654-659)
            else:
                doGene = False
        if K > 2 and it != maxIter and ((noImp == stopIter and not all(temp > v for v in ll) and splitOnLast) or
                                        (splitOnIter > 0 and it % splitOnIter == 0 and doCell)):
            newZ, msg = cCGSplitZ(counts, z, K, pz, ops.ll_alt_z, rng, backend, maxClustersToTry=K, minCell=3)
            messages.append((it, "z", msg))
            if not np.array_equal(newZ, z):
                noImp, doCell = 0, True
                z = newZ
                ops.set_state(z, y)
            else:
                doCell = False
        temp = ops.ll()
        if all(temp > v for v in ll) or it == 1:
            zBest, yBest, llBest, noImp = z.copy(), y.copy(), temp, 1
        else:
            noImp += 1
        ll.append(temp)
        it += 1
    return dict(z=zBest, y=yBest, completeLogLik=np.asarray(ll), finalLogLik=llBest, messages=messages)


class DeviceOpsC:
    """Chain operations of the celda_C loop on a resident handle (algorithm "EM" or "Gibbs")."""

    def __init__(self, ch, algorithm="EM"):
        self.ch, self.algorithm = ch, algorithm

    def ll(self):
        return self.ch.loglik()

    def iterate(self, ixy, uy, ixz, uz):
        if self.algorithm == "EM":
            return self.ch.iterate_em()
        return self.ch.iterate(None, None, ixz, uz)  # ixz None = device Philox draws

    def state(self):
        return self.ch.get_state()[0]

    def set_state(self, z):
        self.ch.set_state(z=z)

    def probs(self):
        return self.ch.probs()[0]

    def ll_alt(self, z):
        self.ch.set_state(z=z)
        return self.ch.loglik()


class DeviceOpsG:
    """Chain operations of the celda_G loop on a resident handle."""

    def __init__(self, ch):
        self.ch = ch

    def ll(self):
        return self.ch.loglik()

    def iterate(self, ixy, uy, ixz, uz):
        return self.ch.iterate(ixy, uy, None, None)

    def state(self):
        return self.ch.get_state()[1]

    def set_state(self, y):
        self.ch.set_state(y=y)

    def probs(self):
        return self.ch.probs()[1]

    def ll_alt(self, y):
        self.ch.set_state(y=y)
        return self.ch.loglik()


def run_chain_one_axis_with_splits(ops, counts, N, axis, draws, rng, backend, maxIter=200, stopIter=10, splitOnIter=10,
                                   splitOnLast=True):
    """The while loops of .celda_C (axis = "z", R/celda_C.R:412-543, .cCSplitZ) and .celda_G (axis = "y",
    R/celda_G.R:384-517, .cGSplitY) with their split step."""
    nG, nM = counts.nrow, counts.ncol
    ll = [ops.ll()]
    best, llBest = ops.state(), ll[0]
    it, noImp, doSplit = 1, 0, True
    messages = []
    while it <= maxIter and noImp <= stopIter:
        temp = ops.iterate(*draws(it, rng, nG, nM))
        lab = ops.state()
        if N > 2 and it != maxIter and ((noImp == stopIter and not all(temp >= v for v in ll) and splitOnLast) or
                                        (splitOnIter > 0 and it % splitOnIter == 0 and doSplit)):
            if axis == "z":
                new, msg = cCSplitZ(counts, lab, N, ops.probs(), ops.ll_alt, rng, backend, maxClustersToTry=N, minCell=3)
            else:
                new, msg = cGSplitY(counts, lab, N, ops.probs(), ops.ll_alt, rng, backend, minFeature=3,
                                    maxClustersToTry=max(N / 2, 10))
            messages.append((it, axis, msg))
            if not np.array_equal(new, lab):
                noImp, doSplit = (0 if axis == "z" else 1), True
            else:
                doSplit = False
            lab = new
            ops.set_state(lab)
        temp = ops.ll()
        if all(temp > v for v in ll) or it == 1:
            best, llBest, noImp = lab.copy(), temp, 1
        else:
            noImp += 1
        ll.append(temp)
        it += 1
    return dict(**{axis: best}, completeLogLik=np.asarray(ll), finalLogLik=llBest, messages=messages)
