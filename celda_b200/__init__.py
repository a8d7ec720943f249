"""celda_b200: host-side mirror (Python + ctypes) of celda's hot-path interface over libcelda_cuda.so.

The product is the CUDA library behind include/celda_cuda.h; this package only marshals numpy arrays across the
C ABI with the reference's names, argument meaning and error behaviour (R/RcppExports.R, R/matrixSums.R,
R/celda_C.R, R/celda_G.R, R/celda_CG.R).  Matrices are column-major like R (numpy order="F"); labels are 1-based.
There is no CPU fallback: every call raises CeldaCudaError if the CUDA library or a device is missing.
"""
import ctypes as C

import numpy as np

from ._lib import CeldaCudaError, build, check, declared_functions, lib  # noqa: F401

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

MODEL_CG, MODEL_C, MODEL_G = 0, 1, 2


def _d(a):
    return np.asfortranarray(a, dtype=np.float64)


def _i(a):
    return np.asfortranarray(a, dtype=np.int32)


def _pd(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_ip)


class CSC:
    """A dgCMatrix: Dim, p (int32[ncol+1]), i (int32[nnz], 0-based), x (float64[nnz])  (SURVEY.md L0)."""

    def __init__(self, nrow, ncol, p, i, x):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.p, self.i, self.x = np.ascontiguousarray(p, np.int32), np.ascontiguousarray(i, np.int32), \
            np.ascontiguousarray(x, np.float64)
        if self.p.size != self.ncol + 1 or self.i.size != self.x.size or self.p[-1] != self.x.size:
            raise ValueError("inconsistent dgCMatrix slots")

    @classmethod
    def from_dense(cls, mat):
        """methods::as(x, "CsparseMatrix") (R/celda_CG.R:225)."""
        mat = np.asarray(mat)
        nr, nc = mat.shape
        cols, rows = np.nonzero(mat.T)
        p = np.zeros(nc + 1, dtype=np.int32)
        np.cumsum(np.bincount(cols, minlength=nc), out=p[1:])
        return cls(nr, nc, p, rows.astype(np.int32), mat.T[cols, rows].astype(np.float64))

    @property
    def nnz(self):
        return int(self.x.size)


def device_count():
    n = C.c_int(0)
    check(lib().celda_cuda_device_count(C.byref(n)))
    return n.value


def set_device(d):
    check(lib().celda_cuda_set_device(int(d)))


def launch_count():
    return int(lib().celda_cuda_launch_count())


def _vec(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    check(fn(_pd(x), x.size, _pd(out)))
    return out


def philox4x32(counters, key):
    """Raw Philox4x32-10 blocks (uint32 words) for known-answer tests."""
    c = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 4)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros_like(c)
    check(lib().celda_cuda_philox4x32(c.view(np.int32).ctypes.data_as(_ip), k.view(np.int32).ctypes.data_as(_ip),
                                      c.shape[0], out.view(np.int32).ctypes.data_as(_ip)))
    return out


def lgamma(x):
    """lgamma table builder (R/celda_CG.R:428-429,543): device arithmetic, value-identical to the sweeps'."""
    return _vec(lib().celda_cuda_lgamma, x)


def log(x):
    return _vec(lib().celda_cuda_log, x)


def exp(x):
    return _vec(lib().celda_cuda_exp, x)


# ------------------------------------------------------------------ src/matrixSumsSparse.cpp
def colSumByGroupSparse(counts, group, K):
    group = _i(group)
    out = np.zeros((counts.nrow, K), order="F")
    check(lib().celda_colSumByGroupSparse(counts.nrow, counts.ncol, _pi(counts.p), _pi(counts.i), _pd(counts.x),
                                          _pi(group), group.size, K, _pd(out)))
    return out


def rowSumByGroupSparse(counts, group, L):
    group = _i(group)
    out = np.zeros((L, counts.ncol), order="F")
    check(lib().celda_rowSumByGroupSparse(counts.nrow, counts.ncol, _pi(counts.p), _pi(counts.i), _pd(counts.x),
                                          _pi(group), group.size, L, _pd(out)))
    return out


def colSumByGroupChangeSparse(counts, px, group, pgroup, K):
    """px is updated IN PLACE and returned (reference quirk Q3)."""
    group, pgroup = _i(group), _i(pgroup)
    if not (px.flags.f_contiguous and px.dtype == np.float64):
        raise TypeError("px must be a column-major float64 matrix (it is updated in place)")
    check(lib().celda_colSumByGroupChangeSparse(counts.nrow, counts.ncol, _pi(counts.p), _pi(counts.i), _pd(counts.x),
                                                _pd(px), px.shape[0], _pi(group), group.size, _pi(pgroup),
                                                pgroup.size, K))
    return px


def rowSumByGroupChangeSparse(counts, px, group, pgroup, L):
    group, pgroup = _i(group), _i(pgroup)
    if not (px.flags.f_contiguous and px.dtype == np.float64):
        raise TypeError("px must be a column-major float64 matrix (it is updated in place)")
    check(lib().celda_rowSumByGroupChangeSparse(counts.nrow, counts.ncol, _pi(counts.p), _pi(counts.i), _pd(counts.x),
                                                _pd(px), px.shape[1], _pi(group), group.size, _pi(pgroup),
                                                pgroup.size, L))
    return px


# ------------------------------------------------------------------ src/matrixSums.c (dense twins)
def _dense_kind(x):
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.integer):
        return "int", np.int32, _pi
    return "numeric", np.float64, _pd


def _dense_sum(name, x, group, nlevels):
    kind, dt, ptr = _dense_kind(x)
    x, group = np.asfortranarray(x, dtype=dt), _i(group)
    nr, nc = x.shape
    nl = max(int(nlevels), 0)
    out = np.zeros((nl, nc) if name == "rowSumByGroup" else (nr, nl), dtype=dt, order="F")
    check(getattr(lib(), f"celda_{name}_{kind}")(ptr(x), nr, nc, _pi(group), group.size, int(nlevels), ptr(out)))
    return out


def rowSumByGroup_dense(x, group, nlevels):
    """_rowSumByGroup / _rowSumByGroup_numeric (src/matrixSums.c:5-48,198-236); nlevels < 0 == not a factor."""
    return _dense_sum("rowSumByGroup", x, group, nlevels)


def colSumByGroup_dense(x, group, nlevels):
    """_colSumByGroup / _colSumByGroup_numeric (src/matrixSums.c:50-93,240-278)."""
    return _dense_sum("colSumByGroup", x, group, nlevels)


def _dense_change(name, x, px, group, pgroup, nlevels, pnlevels):
    kind, dt, ptr = _dense_kind(x)
    x, group, pgroup = np.asfortranarray(x, dtype=dt), _i(group), _i(pgroup)
    if not (px.flags.f_contiguous and px.dtype == dt):
        raise TypeError("px must be a column-major matrix of the same type as x (it is updated in place)")
    pnlevels = nlevels if pnlevels is None else pnlevels
    check(getattr(lib(), f"celda_{name}_{kind}")(ptr(x), x.shape[0], x.shape[1], ptr(px), px.shape[0], px.shape[1],
                                                 _pi(group), group.size, int(nlevels), _pi(pgroup), pgroup.size,
                                                 int(pnlevels)))
    return px


def rowSumByGroupChange_dense(x, px, group, pgroup, nlevels, pnlevels=None):
    """_rowSumByGroupChange[_numeric] (src/matrixSums.c:97-143,282-327); px updated in place."""
    return _dense_change("rowSumByGroupChange", x, px, group, pgroup, nlevels, pnlevels)


def colSumByGroupChange_dense(x, px, group, pgroup, nlevels, pnlevels=None):
    """_colSumByGroupChange[_numeric] (src/matrixSums.c:147-194,331-379); px updated in place."""
    return _dense_change("colSumByGroupChange", x, px, group, pgroup, nlevels, pnlevels)


# ------------------------------------------------------------------ R/matrixSums.R:1-51 type dispatchers
def rowSumByGroup(counts, group, L):
    """.rowSumByGroup: int matrix / numeric matrix -> factor(group, levels = seq(L)) + dense native; dgCMatrix -> sparse."""
    return rowSumByGroupSparse(counts, group, L) if isinstance(counts, CSC) else rowSumByGroup_dense(counts, group, L)


def colSumByGroup(counts, group, K):
    return colSumByGroupSparse(counts, group, K) if isinstance(counts, CSC) else colSumByGroup_dense(counts, group, K)


def rowSumByGroupChange(counts, pcounts, group, pgroup, L):
    if isinstance(counts, CSC):
        return rowSumByGroupChangeSparse(counts, pcounts, group, pgroup, L)
    return rowSumByGroupChange_dense(counts, pcounts, group, pgroup, L)


def colSumByGroupChange(counts, pcounts, group, pgroup, K):
    if isinstance(counts, CSC):
        return colSumByGroupChangeSparse(counts, pcounts, group, pgroup, K)
    return colSumByGroupChange_dense(counts, pcounts, group, pgroup, K)


# ------------------------------------------------------------------ R/RcppExports.R stubs
def cG_CalcGibbsProbY(index, counts, nTSbyC, nbyTS, nGbyTS, nbyG, y, L, nG, beta, gamma, delta):
    """cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:187-249); the lgamma tables are evaluated on the device."""
    c, t, ts, gt, g, y = _d(counts), _d(nTSbyC), _d(nbyTS), _i(nGbyTS), _d(nbyG), _i(y)
    out = np.zeros(L)
    check(lib().celda_cG_CalcGibbsProbY(int(index), _pd(c), c.size, _pd(t), _pd(ts), _pi(gt), _pd(g), _pi(y), L, nG,
                                        beta, gamma, delta, _pd(out)))
    return out


def fastNormPropLog(counts, alpha):
    x = _d(counts)
    out = np.empty_like(x, order="F")
    check(lib().celda_fastNormPropLog(_pd(x), x.shape[0], x.shape[1], alpha, _pd(out)))
    return out


def eigenMatMultInt(A, B):
    A, B = _d(A), _i(B)
    out = np.empty((A.shape[1], B.shape[1]), order="F")
    check(lib().celda_eigenMatMultInt(_pd(A), A.shape[0], A.shape[1], _pi(B), B.shape[1], _pd(out)))
    return out


def eigenMatMultNumeric(A, B):
    A, B = _d(A), _d(B)
    out = np.empty((A.shape[1], B.shape[1]), order="F")
    check(lib().celda_eigenMatMultNumeric(_pd(A), A.shape[0], A.shape[1], _pd(B), B.shape[1], _pd(out)))
    return out


def perplexityG(x, phi, psi, group, nlevels):
    """_perplexityG (src/perplexity.c:5-62): sum x log(phi[g(i), j] psi[i, g(i)])."""
    x, phi, psi, group = _i(x), _d(phi), _d(psi), _i(group)
    ans = C.c_double(0)
    check(lib().celda_perplexityG(_pi(x), x.shape[0], x.shape[1], _pd(phi), phi.shape[0], phi.shape[1], _pd(psi),
                                  psi.shape[0], psi.shape[1], _pi(group), group.size, int(nlevels), C.byref(ans)))
    return ans.value


# ------------------------------------------------------------------ R closures at sweep granularity
def cCCalcGibbsProbZ(counts, mCPByS, nGByCP, nByC, nCP, z, s, K, nG, nM, alpha, beta, ix, u, doSample=True):
    """.cCCalcGibbsProbZ (R/celda_C.R:620-714) on a dense nG x nM matrix (nTSByC in celda_CG) with the visiting order
    `ix` and uniforms `u` injected.  Returns list(mCPByS, nGByCP, nCP, z, probs) like the reference (inputs copied)."""
    counts = _d(counts)
    m, n, nbc, ncp, z, s = _i(mCPByS).copy(order="F"), _d(nGByCP).copy(order="F"), _d(nByC), _d(nCP).copy(), \
        _i(z).copy(), _i(s)
    ix, u = _i(ix), _d(u)
    probs = np.full((K, nM), np.nan, order="F")
    check(lib().celda_cCCalcGibbsProbZ(_pd(counts), _pi(m), m.shape[1], _pd(n), _pd(nbc), _pd(ncp), _pi(z), _pi(s), K, nG,
                                       nM, alpha, beta, int(doSample), _pi(ix), _pd(u), _pd(probs)))
    return dict(mCPByS=m, nGByCP=n, nCP=ncp, z=z, probs=probs)


def cCCalcEMProbZ(counts, mCPByS, nGByCP, nByC, nCP, z, s, K, nG, nM, alpha, beta, doSample=True):
    """.cCCalcEMProbZ (R/celda_C.R:717-756); counts: CSC (dgCMatrix) or a dense matrix."""
    m, n, ncp, z, s = _i(mCPByS).copy(order="F"), _d(nGByCP).copy(order="F"), _d(nCP).copy(), _i(z).copy(), _i(s)
    probs = np.zeros((K, nM), order="F")
    if isinstance(counts, CSC):
        args = (_pi(counts.p), _pi(counts.i), _pd(counts.x))
    else:
        dense = _d(counts)
        args = (None, None, _pd(dense))
    check(lib().celda_cCCalcEMProbZ(nG, nM, *args, _pi(m), m.shape[1], _pd(n), _pd(ncp), _pi(z), _pi(s), K, alpha, beta,
                                    int(doSample), _pd(probs)))
    return dict(mCPByS=m, nGByCP=n, nCP=ncp, z=z, probs=probs)


def cGCalcGibbsProbY(counts, nTSByC, nByTS, nGByTS, nByG, y, L, nG, beta, delta, gamma, ix, u, doSample=True):
    """.cGCalcGibbsProbY (R/celda_G.R:602-660) on a dense nG x ncol matrix (nGByCP in celda_CG).  The lgbeta /
    lggamma / lgdelta tables of the reference are evaluated on the device (value-identical)."""
    counts = _d(counts)
    t, ts, gt, g, y = _d(nTSByC).copy(order="F"), _d(nByTS).copy(), _i(nGByTS).copy(), _d(nByG), _i(y).copy()
    ix, u = _i(ix), _d(u)
    probs = np.full((L, nG), np.nan, order="F")
    check(lib().celda_cGCalcGibbsProbY(_pd(counts), counts.shape[1], _pd(t), _pd(ts), _pi(gt), _pd(g), _pi(y), L, nG,
                                       beta, delta, gamma, int(doSample), _pi(ix), _pd(u), _pd(probs)))
    return dict(nTSByC=t, nGByTS=gt, nByTS=ts, y=y, probs=probs)


def cCGCalcLL(K, L, mCPByS, nTSByCP, nByG, nByTS, nGByTS, nS, nG, alpha, beta, delta, gamma):
    """.cCGCalcLL (R/celda_CG.R:839-888)."""
    m, t, g, ts, gt = _i(mCPByS), _d(nTSByCP), _d(nByG), _d(nByTS), _i(nGByTS)
    ll = C.c_double(0)
    check(lib().celda_cCGCalcLL(K, L, _pi(m), _pd(t), _pd(g), g.size, _pd(ts), _pi(gt), nS, alpha, beta, delta, gamma,
                                C.byref(ll)))
    return ll.value


def cCCalcLL(mCPByS, nGByCP, s, z, K, nS, nG, alpha, beta):
    """.cCCalcLL (R/celda_C.R:760-788); s and z are unused by the reference formula, kept for the signature."""
    m, n = _i(mCPByS), _d(nGByCP)
    ll = C.c_double(0)
    check(lib().celda_cCCalcLL(_pi(m), _pd(n), K, nS, nG, alpha, beta, C.byref(ll)))
    return ll.value


def cGCalcLL(nTSByC, nByTS, nByG, nGByTS, nM, nG, L, beta, delta, gamma):
    """.cGCalcLL (R/celda_G.R:664-702)."""
    t, ts, g, gt = _d(nTSByC), _d(nByTS), _d(nByG), _i(nGByTS)
    ll = C.c_double(0)
    check(lib().celda_cGCalcLL(_pd(t), _pd(ts), _pd(g), g.size, _pi(gt), nM, L, beta, delta, gamma, C.byref(ll)))
    return ll.value


# ------------------------------------------------------------------ device-resident chain
class Chain:
    """One chain of .celda_CG / .celda_C / .celda_G with counts and tables resident in HBM (celda_chain_*)."""

    def __init__(self, counts, sampleLabel=None, K=1, L=1, alpha=1.0, beta=1.0, delta=1.0, gamma=1.0,
                 model=MODEL_CG, device=0):
        self._h = C.c_void_p()
        self.counts, self.model, self.K, self.L = counts, model, int(K), int(L)
        self.device = int(device)
        self.nG, self.nM = counts.nrow, counts.ncol
        self.hyper = (float(alpha), float(beta), float(delta), float(gamma))
        s = None
        self.nS = 1
        if model != MODEL_G:
            s = _i(sampleLabel)
            self.nS = int(s.max())
        self.s = s
        check(lib().celda_chain_create(C.byref(self._h), device, model, counts.nrow, counts.ncol, _pi(counts.p),
                                       _pi(counts.i), _pd(counts.x), _pi(s), self.nS, self.K, self.L, alpha, beta,
                                       delta, gamma))

    def reconfigure(self, K=None, L=None):
        """Another (K, L) of a grid search on the same resident counts (R/celdaGridSearch.R:290-391)."""
        K = self.K if K is None else int(K)
        L = self.L if L is None else int(L)
        check(lib().celda_chain_reconfigure(self._h, K, L))
        self.K, self.L = (1 if self.model == MODEL_G else K), (1 if self.model == MODEL_C else L)

    def close(self):
        if self._h:
            lib().celda_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, z=None, y=None):
        z = None if z is None else _i(z)
        y = None if y is None else _i(y)
        check(lib().celda_chain_set_state(self._h, _pi(z), _pi(y)))

    def get_state(self):
        z = np.zeros(self.nM, np.int32) if self.model != MODEL_G else None
        y = np.zeros(self.nG, np.int32) if self.model != MODEL_C else None
        check(lib().celda_chain_get_state(self._h, _pi(z), _pi(y)))
        return z, y

    def sweep_y(self, ix, u, doSample=True):
        ix, u = _i(ix), _d(u)
        check(lib().celda_chain_sweep_y(self._h, _pi(ix), _pd(u), int(doSample)))

    def sweep_z_gibbs(self, ix, u, doSample=True):
        ix, u = _i(ix), _d(u)
        check(lib().celda_chain_sweep_z_gibbs(self._h, _pi(ix), _pd(u), int(doSample)))

    def seed(self, seed, stream=0):
        """Philox mode: key = seed, stream = chain index (R/celdaGridSearch.R:294 uses seed + chain - 1)."""
        check(lib().celda_chain_seed(self._h, int(seed), int(stream)))

    def sweep_y_philox(self, doSample=True):
        check(lib().celda_chain_sweep_y_philox(self._h, int(doSample)))

    def sweep_z_gibbs_philox(self, doSample=True):
        check(lib().celda_chain_sweep_z_gibbs_philox(self._h, int(doSample)))

    def iterate_philox(self):
        ll = C.c_double(0)
        check(lib().celda_chain_iterate(self._h, None, None, None, None, C.byref(ll)))
        return ll.value

    def get_draws(self, which):
        n = self.nG if which in (0, "y") else self.nM
        ix, u = np.zeros(n, np.int32), np.zeros(n)
        check(lib().celda_chain_get_draws(self._h, 0 if which in (0, "y") else 1, _pi(ix), _pd(u)))
        return ix, u

    def sweep_z_em(self, doSample=True):
        """.cCCalcEMProbZ (R/celda_C.R:717-756) + re-decomposition."""
        check(lib().celda_chain_sweep_z_em(self._h, int(doSample)))

    def iterate_em(self, ix_y=None, u_y=None):
        """One iteration with algorithm = "EM" (celda_C: EM z + LL; celda_CG: y Gibbs sweep + EM z + LL)."""
        ll = C.c_double(0)
        ix_y = None if ix_y is None else _i(ix_y)
        u_y = None if u_y is None else _d(u_y)
        check(lib().celda_chain_iterate_em(self._h, _pi(ix_y), _pd(u_y), C.byref(ll)))
        return ll.value

    def iterate(self, ix_y, u_y, ix_z, u_z):
        """One iteration of R/celda_CG.R:543-602 (no split heuristics); returns the log-likelihood.  On a celda_C
        handle pass (None, None, ix_z, u_z), on a celda_G handle (ix_y, u_y, None, None); all None = Philox."""
        ll = C.c_double(0)
        # int32 / float64 arrays pass through untouched (pinned buffers stay pinned); anything else is converted
        ix_y, ix_z = (None if v is None else _i(v) for v in (ix_y, ix_z))
        u_y, u_z = (None if v is None else _d(v) for v in (u_y, u_z))
        check(lib().celda_chain_iterate(self._h, _pi(ix_y), _pd(u_y), _pi(ix_z), _pd(u_z), C.byref(ll)))
        return ll.value

    def stage_draws(self, ix_y, u_y, ix_z, u_z):
        """ix_y/u_y: [n_iter, nG]; ix_z/u_z: [n_iter, nM] (C-contiguous)."""
        ix_y, u_y = np.ascontiguousarray(ix_y, np.int32), np.ascontiguousarray(u_y, np.float64)
        ix_z, u_z = np.ascontiguousarray(ix_z, np.int32), np.ascontiguousarray(u_z, np.float64)
        check(lib().celda_chain_stage_draws(self._h, ix_y.shape[0], _pi(ix_y), _pd(u_y), _pi(ix_z), _pd(u_z)))

    def iterate_staged(self, it):
        ll = C.c_double(0)
        check(lib().celda_chain_iterate_staged(self._h, int(it), C.byref(ll)))
        return ll.value

    def loglik(self):
        ll = C.c_double(0)
        check(lib().celda_chain_loglik(self._h, C.byref(ll)))
        return ll.value

    def loglik_alt(self, z=None, y=None):
        """.cCGCalcLL with one label vector replaced (celda_CG), the chain's state untouched (split heuristics)."""
        z = None if z is None else _i(z)
        y = None if y is None else _i(y)
        ll = C.c_double(0)
        check(lib().celda_chain_loglik_alt(self._h, _pi(z), _pi(y), C.byref(ll)))
        return ll.value

    def perplexity(self):
        """perplexity(x) on the model's own counts (R/perplexity.R:233-325)."""
        v = C.c_double(0)
        check(lib().celda_chain_perplexity(self._h, C.byref(v)))
        return v.value

    def perplexity_new(self, newCounts):
        """perplexity(x, celdaMod, newCounts = ...) (R/perplexity.R:233-325): `newCounts` is a CSC of the same shape."""
        if (newCounts.nrow, newCounts.ncol) != (self.nG, self.nM):
            raise ValueError("newCounts should have the same number of rows and columns as counts")
        v = C.c_double(0)
        check(lib().celda_chain_perplexity_new(self._h, _pi(newCounts.p), _pi(newCounts.i), _pd(newCounts.x), C.byref(v)))
        return v.value

    def perplexity_resampled(self, seed=12345, numResample=5, return_counts=False):
        """.resamplePerplexity(doResampling = TRUE) for this model (R/perplexity.R:449-476): `numResample` multinomial
        resamples of the counts (.resampleCountMatrix, :764-775; device Philox) and the perplexity of each."""
        out = np.zeros(int(numResample))
        check(lib().celda_chain_perplexity_resampled(self._h, int(seed), int(numResample), _pd(out)))
        if not return_counts:
            return out
        x = np.zeros(self.counts.nnz, np.int32)
        check(lib().celda_chain_get_resampled(self._h, _pi(x)))
        return out, x

    def tables(self):
        """The sufficient statistics as R holds them (.cCGDecomposeCounts, R/celda_CG.R:902-934)."""
        K, L, nG, nM, nS = self.K, self.L, self.nG, self.nM, self.nS
        t = dict(mCPByS=np.zeros((K, nS), np.int32, order="F"), nTSByC=np.zeros((L, nM), order="F"),
                 nTSByCP=np.zeros((L, K), order="F"), nGByCP=np.zeros((nG, K), order="F"), nCP=np.zeros(K),
                 nByC=np.zeros(nM), nByG=np.zeros(nG), nByTS=np.zeros(L), nGByTS=np.zeros(L, np.int32))
        check(lib().celda_chain_get_tables(self._h, _pi(t["mCPByS"]), _pd(t["nTSByC"]), _pd(t["nTSByCP"]),
                                           _pd(t["nGByCP"]), _pd(t["nCP"]), _pd(t["nByC"]), _pd(t["nByG"]),
                                           _pd(t["nByTS"]), _pi(t["nGByTS"])))
        return t

    def probs(self):
        pz = np.zeros((self.K, self.nM), order="F") if self.model != MODEL_G else None
        py = np.zeros((self.L, self.nG), order="F") if self.model != MODEL_C else None
        check(lib().celda_chain_get_probs(self._h, _pd(pz), _pd(py)))
        return pz, py

    def timer_start(self):
        check(lib().celda_chain_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        check(lib().celda_chain_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_scan_mode(self, mode):
        """celda_G y sweep: 0 = exact full-column scan (default), 1 = only the cells where the gene has a count (opt-in)."""
        check(lib().celda_chain_set_scan_mode(self._h, int(mode)))

    def scan_margin(self):
        """(smallest margin / bound over the draws of the last mode-1 sweep, number of draws whose margin does not exceed
        their bound); no such draw proves that the sweep's labels equal the exact scan's (include/celda_cuda.h)."""
        m = (C.c_double * 2)()
        check(lib().celda_chain_scan_margin(self._h, m))
        return float(m[0]), int(m[1])

    def profile(self, enable=True):
        check(lib().celda_chain_profile(self._h, int(enable)))

    def profile_get(self):
        names = ["sweep_y", "rowSumByGroupChange", "sweep_z", "colSumByGroupChange", "loglik", "rowSumByGroup",
                 "colSumByGroup"]
        out = {}
        for w, n in enumerate(names):
            ms, nl = C.c_double(0), C.c_longlong(0)
            st = (C.c_longlong * 8)()
            check(lib().celda_chain_profile_get(self._h, w, C.byref(ms), C.byref(nl), st))
            out[n] = dict(ms=ms.value, launches=nl.value, rounds=st[0], units=st[1], movers=st[2],
                          clk_commit=st[3], clk_units=st[4], clk_draw=st[5], clk_barrier=st[6],
                          redraws=st[7] & ((1 << 40) - 1), discards=st[7] >> 40)
        return out
