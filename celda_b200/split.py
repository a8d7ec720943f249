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


# ------------------------------------------------------------------ small helpers on the CSC container
def csc_columns(counts, keep):
    """counts[, keep, drop = FALSE] for a boolean mask."""
    keep = np.asarray(keep, dtype=bool)
    lens = np.diff(counts.p)[keep]
    p = np.zeros(int(keep.sum()) + 1, dtype=np.int64)
    np.cumsum(lens, out=p[1:])
    entry = np.repeat(keep, np.diff(counts.p))
    return CSC(counts.nrow, int(keep.sum()), p.astype(np.int32), counts.i[entry], counts.x[entry])


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
        self.device = device

    def celda_C_labels(self, counts, K, alpha, beta, maxIter, seed):
        """.celda_C(counts, K, maxIter, zInitialize = "random", algorithm = "EM", nchains = 3): best chain's z."""
        s = np.ones(counts.ncol, dtype=np.int32)
        res = _model.celda_C(counts, s, K, alpha=alpha, beta=beta, algorithm="EM", maxIter=maxIter, seed=seed, nchains=3,
                             zInitialize="random", device=self.device)
        return res["model"]["z"]

    def celda_G_labels(self, counts, L, beta, delta, gamma, maxIter, seed):
        """.celda_G(counts, L, maxIter, yInitialize = "random", nchains = 3): best chain's y (injected Gibbs draws)."""
        res = _model.celda_G(counts, L, beta=beta, delta=delta, gamma=gamma, maxIter=maxIter, seed=seed, nchains=3,
                             yInitialize="random", device=self.device, draws=_draws_G)
        return res["model"]["y"]

    def session_C(self, counts, alpha, beta):
        return _SessionC(counts, alpha, beta, self.device)

    def session_G(self, counts, beta, delta, gamma):
        return _SessionG(counts, beta, delta, gamma, self.device)

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
