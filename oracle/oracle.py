"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
All matrices follow R: column-major (numpy order="F"), labels 1-based int32.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "libcelda_oracle.so")
    src = os.path.join(_HERE, "celda_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def fd(a):
    return np.asfortranarray(a, dtype=np.float64)


def fi(a):
    return np.asfortranarray(a, dtype=np.int32)


class OracleError(RuntimeError):
    pass


class Oracle:
    def __init__(self, long_double=False, lib_path=None):
        """lib_path: another build of celda_oracle.c (the sanitizer build, oracle/Makefile target `asan`)."""
        build()
        self.lib = C.CDLL(lib_path or os.path.join(_HERE, "libcelda_oracle_ld.so" if long_double else "libcelda_oracle.so"))
        L = self.lib
        L.o_last_error.restype = C.c_char_p
        for n in ("o_log_vec", "o_exp_vec", "o_lgamma_vec"):
            getattr(L, n).argtypes = [_dp, C.c_long, _dp]
        for n in ("o_cCGCalcLL", "o_cCCalcLL", "o_cGCalcLL", "o_perplexity_C", "o_perplexity_CG", "o_perplexity_G"):
            getattr(L, n).restype = C.c_double
        L.o_cCGCalcLL.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, _dp, _ip, C.c_int] + [C.c_double] * 4
        L.o_cCCalcLL.argtypes = [_ip, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.o_cGCalcLL.argtypes = [_dp, _dp, _dp, C.c_int, _ip, C.c_long, C.c_int] + [C.c_double] * 3
        L.o_sample_ll.argtypes = [_dp, C.c_int, C.c_double, _dp, _ip, _dp]
        L.o_cCCalcGibbsProbZ.argtypes = [_dp, C.c_int, C.c_long, _ip, C.c_int, _dp, _dp, _dp, _ip, _ip, C.c_int,
                                         C.c_double, C.c_double, C.c_int, _ip, _dp, C.c_long, C.c_long, _dp, _dp]
        L.o_cGCalcGibbsProbY.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, _ip, _dp, _ip, C.c_int, C.c_double,
                                         C.c_double, C.c_double, _dp, _dp, _dp, C.c_int, _ip, _dp, C.c_long, C.c_long,
                                         _dp, _dp]
        L.o_cCCalcEMProbZ.argtypes = [C.c_int, C.c_int, C.c_long, _ip, _ip, _dp, _ip, C.c_int, _dp, _dp, _ip, _ip,
                                      C.c_int, C.c_double, C.c_double, C.c_int, _dp]
        L.o_cG_CalcGibbsProbY.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, _ip, _dp, _ip, C.c_int, _dp, _dp, _dp,
                                          C.c_double, C.c_double, C.c_double, _dp]
        L.o_cG_calcGibbsProbY_Simple.argtypes = [_dp, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _ip, C.c_int, C.c_int,
                                                 C.c_double, C.c_double, C.c_double, _dp]
        L.o_fastNormPropLog.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp]
        L.o_matMultT.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.o_table_zs.argtypes = [_ip, _ip, C.c_long, C.c_int, C.c_int, _ip]
        L.o_perplexity_C.argtypes = [C.c_int, C.c_long, _ip, _ip, _dp, _ip, C.c_int, _dp, C.c_int, _ip, C.c_double,
                                     C.c_double]
        L.o_perplexity_CG.argtypes = [C.c_int, C.c_long, _ip, _ip, _dp, _ip, C.c_int, _dp, _dp, _ip, C.c_int, C.c_int,
                                      _ip, C.c_double, C.c_double, C.c_double]
        L.o_perplexity_G.argtypes = [C.c_int, C.c_long, _ip, _ip, _dp, _dp, _dp, _ip, C.c_int, C.c_double, C.c_double]

    def _chk(self, rc):
        if rc:
            raise OracleError(self.lib.o_last_error().decode())

    # ---------------------------------------------------------------- math
    def _vec(self, fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        fn(_p(x, _dp), x.size, _p(out, _dp))
        return out

    def log(self, x):
        return self._vec(self.lib.o_log_vec, x)

    def exp(self, x):
        return self._vec(self.lib.o_exp_vec, x)

    def lgamma(self, x):
        return self._vec(self.lib.o_lgamma_vec, x)

    # ---------------------------------------------------------------- aggregation (sparse)
    def colSumByGroupSparse(self, csc, group, K):
        nr, nc, p, i, x = csc
        group = fi(group)
        out = np.zeros((nr, K), order="F")
        self._chk(self.lib.o_colSumByGroupSparse(C.c_int(nr), C.c_int(nc), _p(p, _ip), _p(i, _ip), _p(x, _dp),
                                                 _p(group, _ip), C.c_long(group.size), C.c_int(K), _p(out, _dp)))
        return out

    def rowSumByGroupSparse(self, csc, group, L):
        nr, nc, p, i, x = csc
        group = fi(group)
        out = np.zeros((L, nc), order="F")
        self._chk(self.lib.o_rowSumByGroupSparse(C.c_int(nr), C.c_int(nc), _p(p, _ip), _p(i, _ip), _p(x, _dp),
                                                 _p(group, _ip), C.c_long(group.size), C.c_int(L), _p(out, _dp)))
        return out

    def colSumByGroupChangeSparse(self, csc, px, group, pgroup, K):
        """px (float64, order F) is updated IN PLACE and returned (quirk Q3)."""
        nr, nc, p, i, x = csc
        group, pgroup = fi(group), fi(pgroup)
        assert px.flags.f_contiguous and px.dtype == np.float64
        self._chk(self.lib.o_colSumByGroupChangeSparse(
            C.c_int(nr), C.c_int(nc), _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(px, _dp), C.c_int(px.shape[0]),
            _p(group, _ip), C.c_long(group.size), _p(pgroup, _ip), C.c_long(pgroup.size), C.c_int(K)))
        return px

    def rowSumByGroupChangeSparse(self, csc, px, group, pgroup, L):
        nr, nc, p, i, x = csc
        group, pgroup = fi(group), fi(pgroup)
        assert px.flags.f_contiguous and px.dtype == np.float64
        self._chk(self.lib.o_rowSumByGroupChangeSparse(
            C.c_int(nr), C.c_int(nc), _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(px, _dp), C.c_int(px.shape[1]),
            _p(group, _ip), C.c_long(group.size), _p(pgroup, _ip), C.c_long(pgroup.size), C.c_int(L)))
        return px

    # ---------------------------------------------------------------- aggregation (dense; nl<0 == "not a factor")
    def _dense(self, name, x, group, nl):
        kind, ct, dt = ("int", _ip, np.int32) if np.issubdtype(np.asarray(x).dtype, np.integer) else \
            ("numeric", _dp, np.float64)
        x = np.asfortranarray(x, dtype=dt)
        group = fi(group)
        nr, nc = x.shape
        row = name.startswith("row")
        out = np.zeros((max(nl, 0), nc) if row else (nr, max(nl, 0)), dtype=dt, order="F")
        fn = getattr(self.lib, f"o_{name}_{kind}")
        self._chk(fn(_p(x, ct), C.c_int(nr), C.c_int(nc), _p(group, _ip), C.c_long(group.size), C.c_int(nl),
                     _p(out, ct)))
        return out

    def rowSumByGroup(self, x, group, nl):
        return self._dense("rowSumByGroup", x, group, nl)

    def colSumByGroup(self, x, group, nl):
        return self._dense("colSumByGroup", x, group, nl)

    def _dense_change(self, name, x, px, group, pgroup, nl, nlp):
        kind, ct, dt = ("int", _ip, np.int32) if np.issubdtype(np.asarray(x).dtype, np.integer) else \
            ("numeric", _dp, np.float64)
        x = np.asfortranarray(x, dtype=dt)
        assert px.flags.f_contiguous and px.dtype == dt
        group, pgroup = fi(group), fi(pgroup)
        fn = getattr(self.lib, f"o_{name}_{kind}")
        self._chk(fn(_p(x, ct), C.c_int(x.shape[0]), C.c_int(x.shape[1]), _p(px, ct), C.c_int(px.shape[0]),
                     C.c_int(px.shape[1]), _p(group, _ip), C.c_long(group.size), C.c_int(nl), _p(pgroup, _ip),
                     C.c_long(pgroup.size), C.c_int(nlp)))
        return px

    def rowSumByGroupChange(self, x, px, group, pgroup, nl, nlp=None):
        return self._dense_change("rowSumByGroupChange", x, px, group, pgroup, nl, nl if nlp is None else nlp)

    def colSumByGroupChange(self, x, px, group, pgroup, nl, nlp=None):
        return self._dense_change("colSumByGroupChange", x, px, group, pgroup, nl, nl if nlp is None else nlp)

    def table_zs(self, z, s, K, nS):
        z, s = fi(z), fi(s)
        m = np.zeros((K, nS), dtype=np.int32, order="F")
        self.lib.o_table_zs(_p(z, _ip), _p(s, _ip), z.size, K, nS, _p(m, _ip))
        return m

    # ---------------------------------------------------------------- decompose (R/celda_CG.R:902-934 etc.)
    def cCGDecomposeCounts(self, csc, s, z, y, K, L):
        nr, nc, p, i, x = csc
        nS = len(np.unique(s))
        mCPByS = self.table_zs(z, s, K, nS)
        nTSByC = self.rowSumByGroupSparse(csc, y, L)
        nGByCP = self.colSumByGroupSparse(csc, z, K)
        nTSByCP = self.colSumByGroup(nTSByC, z, K)
        nByC = np.add.reduceat(np.append(x, 0.0), p[:-1]) * (np.diff(p) > 0) if x.size else np.zeros(nc)
        nByG = np.bincount(i, weights=x, minlength=nr).astype(np.float64)
        nByTS = self.rowSumByGroup(nByG.reshape(-1, 1), y, L).ravel()
        nCP = nTSByCP.sum(axis=0)
        nGByTS = (np.bincount(np.asarray(y) - 1, minlength=L) + 1).astype(np.int32)
        return dict(mCPByS=mCPByS, nTSByC=nTSByC, nTSByCP=nTSByCP, nCP=nCP, nByG=nByG, nByC=nByC.astype(np.float64),
                    nByTS=nByTS, nGByTS=nGByTS, nGByCP=nGByCP, nM=nc, nG=nr, nS=nS)

    def cCDecomposeCounts(self, csc, s, z, K):
        nr, nc, p, i, x = csc
        nS = len(np.unique(s))
        nGByCP = self.colSumByGroupSparse(csc, z, K)
        nByC = np.add.reduceat(np.append(x, 0.0), p[:-1]) * (np.diff(p) > 0)
        return dict(mCPByS=self.table_zs(z, s, K, nS), nGByCP=nGByCP, nCP=nGByCP.sum(axis=0),
                    nByC=nByC.astype(np.float64), nS=nS, nG=nr, nM=nc)

    def cGDecomposeCounts(self, csc, y, L):
        nr, nc, p, i, x = csc
        nTSByC = self.rowSumByGroupSparse(csc, y, L)
        nByG = np.bincount(i, weights=x, minlength=nr).astype(np.float64)
        nByTS = self.rowSumByGroup(nByG.reshape(-1, 1), y, L).ravel()
        nGByTS = (np.bincount(np.asarray(y) - 1, minlength=L) + 1).astype(np.int32)
        return dict(nTSByC=nTSByC, nByG=nByG, nByTS=nByTS, nGByTS=nGByTS, nM=nc, nG=nr)

    # ---------------------------------------------------------------- log-likelihoods
    def cCGCalcLL(self, K, L, mCPByS, nTSByCP, nByG, nByTS, nGByTS, nS, nG, alpha, beta, delta, gamma):
        m, t, g, ts, gt = fi(mCPByS), fd(nTSByCP), fd(nByG), fd(nByTS), fi(nGByTS)
        return self.lib.o_cCGCalcLL(K, L, _p(m, _ip), _p(t, _dp), _p(g, _dp), g.size, _p(ts, _dp), _p(gt, _ip), nS,
                                    alpha, beta, delta, gamma)

    def cCCalcLL(self, mCPByS, nGByCP, K, nS, nG, alpha, beta):
        m, n = fi(mCPByS), fd(nGByCP)
        return self.lib.o_cCCalcLL(_p(m, _ip), _p(n, _dp), K, nS, nG, alpha, beta)

    def cGCalcLL(self, nTSByC, nByTS, nByG, nGByTS, nM, nG, L, beta, delta, gamma):
        t, ts, g, gt = fd(nTSByC), fd(nByTS), fd(nByG), fi(nGByTS)
        return self.lib.o_cGCalcLL(_p(t, _dp), _p(ts, _dp), _p(g, _dp), g.size, _p(gt, _ip), nM, L, beta, delta, gamma)

    # ---------------------------------------------------------------- sampler / sweeps
    def sampleLl(self, ll, u):
        ll = fd(ll)
        wp, wperm, mg = np.empty(ll.size), np.empty(ll.size, dtype=np.int32), np.zeros(1)
        r = self.lib.o_sample_ll(_p(ll, _dp), ll.size, u, _p(wp, _dp), _p(wperm, _ip), _p(mg, _dp))
        return r, float(mg[0])

    def cCCalcGibbsProbZ(self, counts, mCPByS, nGByCP, nByC, nCP, z, s, K, nG, nM, alpha, beta, ix, u,
                         doSample=True, first=0, count=None, want_probs=True, want_margin=False):
        """Returns dict(mCPByS, nGByCP, nCP, z, probs[, margin]); inputs are copied, like R."""
        counts = fd(counts)
        m, n, nbc, ncp, z, s = fi(mCPByS).copy(order="F"), fd(nGByCP).copy(order="F"), fd(nByC), fd(nCP).copy(), \
            fi(z).copy(), fi(s)
        ix, u = fi(ix), fd(u)
        count = nM if count is None else count
        probs = np.full((K, nM), np.nan, order="F") if want_probs else None
        margin = np.full(nM, np.nan) if want_margin else None
        assert counts.shape == (nG, nM)
        self._chk(self.lib.o_cCCalcGibbsProbZ(_p(counts, _dp), nG, nM, _p(m, _ip), m.shape[1], _p(n, _dp), _p(nbc, _dp),
                                              _p(ncp, _dp), _p(z, _ip), _p(s, _ip), K, alpha, beta, int(doSample),
                                              _p(ix, _ip), _p(u, _dp), first, count, _p(probs, _dp), _p(margin, _dp)))
        out = dict(mCPByS=m, nGByCP=n, nCP=ncp, z=z, probs=probs)
        if want_margin:
            out["margin"] = margin
        return out

    def cGCalcGibbsProbY(self, counts, nTSByC, nByTS, nGByTS, nByG, y, L, nG, beta, delta, gamma, ix, u,
                         lgbeta=None, lggamma=None, lgdelta=None, doSample=True, first=0, count=None,
                         want_probs=True, want_margin=False):
        counts = fd(counts)
        t, ts, gt, g, y = fd(nTSByC).copy(order="F"), fd(nByTS).copy(), fi(nGByTS).copy(), fd(nByG), fi(y).copy()
        ix, u = fi(ix), fd(u)
        ncol = counts.shape[1]
        count = nG if count is None else count
        probs = np.full((L, nG), np.nan, order="F") if want_probs else None
        margin = np.full(nG, np.nan) if want_margin else None
        lb, lg_, ld = (None if a is None else fd(a) for a in (lgbeta, lggamma, lgdelta))
        self._chk(self.lib.o_cGCalcGibbsProbY(_p(counts, _dp), nG, ncol, _p(t, _dp), _p(ts, _dp), _p(gt, _ip),
                                              _p(g, _dp), _p(y, _ip), L, beta, delta, gamma, _p(lb, _dp), _p(lg_, _dp),
                                              _p(ld, _dp), int(doSample), _p(ix, _ip), _p(u, _dp), first, count,
                                              _p(probs, _dp), _p(margin, _dp)))
        out = dict(nTSByC=t, nGByTS=gt, nByTS=ts, y=y, probs=probs)
        if want_margin:
            out["margin"] = margin
        return out

    def cG_CalcGibbsProbY(self, index, counts_row, nTSbyC, nbyTS, nGbyTS, nbyG, y, L, beta, gamma, delta,
                          lg_beta=None, lg_gamma=None, lg_delta=None):
        c, t, ts, gt, g, y = fd(counts_row), fd(nTSbyC), fd(nbyTS), fi(nGbyTS), fd(nbyG), fi(y)
        lb, lg_, ld = (None if a is None else fd(a) for a in (lg_beta, lg_gamma, lg_delta))
        out = np.zeros(L)
        self.lib.o_cG_CalcGibbsProbY(index, _p(c, _dp), c.size, _p(t, _dp), _p(ts, _dp), _p(gt, _ip), _p(g, _dp),
                                     _p(y, _ip), L, _p(lb, _dp), _p(lg_, _dp), _p(ld, _dp), beta, gamma, delta,
                                     _p(out, _dp))
        return out

    def cG_calcGibbsProbY_Simple(self, counts, nGbyTS, nTSbyC, nbyTS, nbyG, y, L, index, gamma, beta, delta):
        c, gt, t, ts, g, y = fd(counts), fi(nGbyTS).copy(), fd(nTSbyC).copy(order="F"), fd(nbyTS).copy(), fd(nbyG), fi(y)
        out = np.zeros(L)
        self.lib.o_cG_calcGibbsProbY_Simple(_p(c, _dp), c.shape[0], c.shape[1], _p(gt, _ip), _p(t, _dp), _p(ts, _dp),
                                            _p(g, _dp), _p(y, _ip), L, index, gamma, beta, delta, _p(out, _dp))
        return out

    def cCCalcEMProbZ(self, counts, mCPByS, nGByCP, nByC, nCP, z, s, K, nG, nM, alpha, beta, doSample=True):
        """counts: CSC tuple (nr, nc, p, i, x) or a dense matrix."""
        m, n, ncp, z, s = fi(mCPByS).copy(order="F"), fd(nGByCP).copy(order="F"), fd(nCP).copy(), fi(z).copy(), fi(s)
        probs = np.zeros((K, nM), order="F")
        if isinstance(counts, tuple):
            nr, nc, p, i, x = counts
            rc = self.lib.o_cCCalcEMProbZ(1, nr, nM, _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(m, _ip), m.shape[1],
                                          _p(n, _dp), _p(ncp, _dp), _p(z, _ip), _p(s, _ip), K, alpha, beta,
                                          int(doSample), _p(probs, _dp))
        else:
            x = fd(counts)
            rc = self.lib.o_cCCalcEMProbZ(0, nG, nM, None, None, _p(x, _dp), _p(m, _ip), m.shape[1], _p(n, _dp),
                                          _p(ncp, _dp), _p(z, _ip), _p(s, _ip), K, alpha, beta, int(doSample),
                                          _p(probs, _dp))
        self._chk(rc)
        return dict(mCPByS=m, nGByCP=n, nCP=ncp, z=z, probs=probs)

    def fastNormPropLog(self, x, alpha):
        x = fd(x)
        out = np.empty_like(x, order="F")
        self._chk(self.lib.o_fastNormPropLog(_p(x, _dp), x.shape[0], x.shape[1], alpha, _p(out, _dp)))
        return out

    def matMultT(self, A, B):
        A, B = fd(A), fd(B)
        out = np.empty((A.shape[1], B.shape[1]), order="F")
        self.lib.o_matMultT(_p(A, _dp), A.shape[0], A.shape[1], _p(B, _dp), B.shape[1], _p(out, _dp))
        return out

    # ---------------------------------------------------------------- perplexity
    def perplexity_CG(self, csc, mCPByS, nTSByCP, nByG, y, K, L, s, alpha, beta, delta):
        nr, nc, p, i, x = csc
        m, t, g, y, s = fi(mCPByS), fd(nTSByCP), fd(nByG), fi(y), fi(s)
        return self.lib.o_perplexity_CG(nr, nc, _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(m, _ip), m.shape[1], _p(t, _dp),
                                        _p(g, _dp), _p(y, _ip), K, L, _p(s, _ip), alpha, beta, delta)

    def perplexity_C(self, csc, mCPByS, nGByCP, K, s, alpha, beta):
        nr, nc, p, i, x = csc
        m, n, s = fi(mCPByS), fd(nGByCP), fi(s)
        return self.lib.o_perplexity_C(nr, nc, _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(m, _ip), m.shape[1], _p(n, _dp),
                                       K, _p(s, _ip), alpha, beta)

    def perplexity_G(self, csc, nTSByC, nByG, y, L, beta, delta):
        nr, nc, p, i, x = csc
        t, g, y = fd(nTSByC), fd(nByG), fi(y)
        return self.lib.o_perplexity_G(nr, nc, _p(p, _ip), _p(i, _ip), _p(x, _dp), _p(t, _dp), _p(g, _dp), _p(y, _ip),
                                       L, beta, delta)

    def perplexityG_dense(self, x, phi, psi, group, nl):
        x, phi, psi, group = fi(x), fd(phi), fd(psi), fi(group)
        ans = np.zeros(1)
        self._chk(self.lib.o_perplexityG(_p(x, _ip), x.shape[0], x.shape[1], _p(phi, _dp), phi.shape[0], phi.shape[1],
                                         _p(psi, _dp), psi.shape[0], psi.shape[1], _p(group, _ip),
                                         C.c_long(group.size), nl, _p(ans, _dp)))
        return float(ans[0])


def dense_to_csc(mat):
    """Dense [nr, nc] -> (nr, nc, p, i, x) like as(x, "CsparseMatrix") (R/celda_CG.R:225): int32 p/i, double x."""
    mat = np.asarray(mat)
    nr, nc = mat.shape
    cols, rows = np.nonzero(mat.T)
    x = mat.T[cols, rows].astype(np.float64)
    p = np.zeros(nc + 1, dtype=np.int32)
    np.cumsum(np.bincount(cols, minlength=nc), out=p[1:])
    return nr, nc, p, rows.astype(np.int32), x
