"""CPU backend of celda_b200.split for the parity tests: the same operations as DeviceBackend, computed by the oracle
(test infrastructure; the product never imports this)."""
import numpy as np

from celda_b200.model import best_chain, initialize_cluster


def _csc(counts):
    return (counts.nrow, counts.ncol, counts.p, counts.i, counts.x)


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    # .celda_C(algorithm = "EM", nchains = 3, zInitialize = "random"): R/celda_C.R:412-589 without the split steps
    def celda_C_labels(self, counts, K, alpha, beta, maxIter, seed):
        o, O = self.o, _csc(counts)
        s = np.ones(counts.ncol, dtype=np.int32)
        res = []
        for c in range(1, 4):
            rng = np.random.Generator(np.random.PCG64(seed + c - 1))
            z = initialize_cluster(K, counts.ncol, None, rng)
            p = o.cCDecomposeCounts(O, s, z, K)
            m, n, ncp = p["mCPByS"], p["nGByCP"], p["nCP"]
            ll = [o.cCCalcLL(m, n, K, 1, p["nG"], alpha, beta)]
            zBest, llBest, it, noImp = z, ll[0], 1, 0
            while it <= maxIter and noImp <= 1:  # EM forces stopIter <- 1 (R/celda_C.R:351-353)
                r = o.cCCalcEMProbZ(O, m, n, p["nByC"], ncp, z, s, K, p["nG"], p["nM"], alpha, beta)
                z, m, n, ncp = r["z"], r["mCPByS"], r["nGByCP"], r["nCP"]
                temp = o.cCCalcLL(m, n, K, 1, p["nG"], alpha, beta)
                if all(temp > v for v in ll) or it == 1:
                    zBest, llBest, noImp = z.copy(), temp, 1
                else:
                    noImp += 1
                ll.append(temp)
                it += 1
            res.append((llBest, zBest))
        return res[best_chain([r[0] for r in res])][1]

    # .celda_G(nchains = 3, yInitialize = "random"): R/celda_G.R:384-517 without the split steps
    def celda_G_labels(self, counts, L, beta, delta, gamma, maxIter, seed):
        o, O = self.o, _csc(counts)
        nG, nM = counts.nrow, counts.ncol
        dense = np.zeros((nG, nM))
        cols = np.repeat(np.arange(nM), np.diff(counts.p))
        dense[counts.i, cols] = counts.x
        res = []
        for c in range(1, 4):
            rng = np.random.Generator(np.random.PCG64(seed + c - 1))
            y = initialize_cluster(L, nG, None, rng)
            p = o.cGDecomposeCounts(O, y, L)
            t, ts, gt = p["nTSByC"], p["nByTS"], p["nGByTS"]
            ll = [o.cGCalcLL(t, ts, p["nByG"], gt, nM, nG, L, beta, delta, gamma)]
            yBest, llBest, it, noImp = y, ll[0], 1, 0
            while it <= maxIter and noImp <= 10:
                ix, u = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
                r = o.cGCalcGibbsProbY(dense, t, ts, gt, p["nByG"], y, L, nG, beta, delta, gamma, ix, u)
                y, t, ts, gt = r["y"], r["nTSByC"], r["nByTS"], r["nGByTS"]
                temp = o.cGCalcLL(t, ts, p["nByG"], gt, nM, nG, L, beta, delta, gamma)
                if all(temp > v for v in ll) or it == 1:
                    yBest, llBest, noImp = y.copy(), temp, 1
                else:
                    noImp += 1
                ll.append(temp)
                it += 1
            res.append((llBest, yBest))
        return res[best_chain([r[0] for r in res])][1]

    def session_C(self, counts, alpha, beta):
        return _SesC(self.o, counts, alpha, beta)

    def session_G(self, counts, beta, delta, gamma):
        return _SesG(self.o, counts, beta, delta, gamma)

    def colSumByGroup(self, counts, z, K):
        return self.o.colSumByGroupSparse(_csc(counts), z, K)


class _SesC:
    def __init__(self, o, counts, alpha, beta):
        self.o, self.O, self.a, self.b = o, _csc(counts), alpha, beta
        self.s = np.ones(counts.ncol, dtype=np.int32)

    def probs(self, z, K):
        p = self.o.cCDecomposeCounts(self.O, self.s, z, K)
        return self.o.cCCalcEMProbZ(self.O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, self.s, K, p["nG"], p["nM"],
                                    self.a, self.b, doSample=False)["probs"]

    def ll(self, z, K):
        p = self.o.cCDecomposeCounts(self.O, self.s, z, K)
        return self.o.cCCalcLL(p["mCPByS"], p["nGByCP"], K, 1, p["nG"], self.a, self.b)

    def close(self):
        pass


class _SesG:
    def __init__(self, o, counts, beta, delta, gamma):
        self.o, self.O, self.h = o, _csc(counts), (beta, delta, gamma)
        nG, nM = counts.nrow, counts.ncol
        self.dense = np.zeros((nG, nM))
        self.dense[counts.i, np.repeat(np.arange(nM), np.diff(counts.p))] = counts.x
        self.nG, self.nM = nG, nM

    def probs(self, y, L):
        b, d, g = self.h
        p = self.o.cGDecomposeCounts(self.O, y, L)
        ix = np.arange(1, self.nG + 1, dtype=np.int32)
        return self.o.cGCalcGibbsProbY(self.dense, p["nTSByC"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, self.nG, b, d, g,
                                       ix, np.zeros(self.nG), doSample=False)["probs"]

    def ll(self, y, L):
        b, d, g = self.h
        p = self.o.cGDecomposeCounts(self.O, y, L)
        return self.o.cGCalcLL(p["nTSByC"], p["nByTS"], p["nByG"], p["nGByTS"], self.nM, self.nG, L, b, d, g)

    def close(self):
        pass


class OracleOpsCG:
    """The chain operations of celda_b200.split.run_chain_CG_with_splits on the CPU oracle (celda_CG, Gibbs)."""

    def __init__(self, oracle, counts, s, K, L, z, y, hyper=(1.0, 1.0, 1.0, 1.0)):
        self.o, self.O, self.s, self.K, self.L, self.h = oracle, _csc(counts), np.asarray(s, np.int32), K, L, hyper
        self.nG, self.nM = counts.nrow, counts.ncol
        self.pz = self.py = None
        self.set_state(z, y)

    def set_state(self, z, y):
        self.z, self.y = np.asarray(z, np.int32).copy(), np.asarray(y, np.int32).copy()
        self.p = self.o.cCGDecomposeCounts(self.O, self.s, self.z, self.y, self.K, self.L)
        # the probabilities of the last sweeps survive a label change (zProb = t(nextZ$probs), R/celda_CG.R:674-694)

    def _ll(self, p):
        a, b, d, g = self.h
        return self.o.cCGCalcLL(self.K, self.L, p["mCPByS"], p["nTSByCP"], p["nByG"], p["nByTS"], p["nGByTS"], p["nS"],
                                self.nG, a, b, d, g)

    def ll(self):
        return self._ll(self.p)

    def ll_alt_z(self, z):
        return self._ll(self.o.cCGDecomposeCounts(self.O, self.s, np.asarray(z, np.int32), self.y, self.K, self.L))

    def ll_alt_y(self, y):
        return self._ll(self.o.cCGDecomposeCounts(self.O, self.s, self.z, np.asarray(y, np.int32), self.K, self.L))

    def state(self):
        return self.z.copy(), self.y.copy()

    def probs(self):
        return self.pz, self.py

    def iterate(self, ixy, uy, ixz, uz):
        o, p, (a, b, d, g) = self.o, self.p, self.h
        oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], self.y, self.L, self.nG, b,
                                d, g, ixy, uy)
        nTSByC = o.rowSumByGroupChangeSparse(self.O, p["nTSByC"].copy(order="F"), oy["y"], self.y, self.L)
        oz = o.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], self.z, self.s, self.K, self.L,
                                self.nM, a, b, ixz, uz)
        py, pz = oy["probs"], oz["probs"]
        self.set_state(oz["z"], oy["y"])
        self.pz, self.py = pz, py
        return self.ll()
