import sys, os, hashlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import celda_b200 as cb
from celda_b200 import split
from oracle.oracle import Oracle
from split_oracle_backend import OracleBackend
g = dict(np.load('/root/repo/tests/golden/celdaCG.npz'))
counts = cb.CSC.from_dense(g["counts"])

def h(a):
    a = np.ascontiguousarray(a)
    return hashlib.md5(a.tobytes()).hexdigest()[:8]

class Log:
    def __init__(self, be): self.be, self.log = be, []
    def celda_C_labels(self, counts, K, alpha, beta, maxIter, seed):
        z = self.be.celda_C_labels(counts, K, alpha, beta, maxIter, seed)
        self.log.append(("C", counts.ncol, counts.nnz, K, seed, h(z), np.bincount(z).tolist())); return z
    def celda_G_labels(self, counts, L, beta, delta, gamma, maxIter, seed):
        y = self.be.celda_G_labels(counts, L, beta, delta, gamma, maxIter, seed)
        self.log.append(("G", counts.nrow, counts.ncol, counts.nnz, L, seed, h(y), np.bincount(y).tolist())); return y
    def session_C(self, counts, alpha, beta):
        s = self.be.session_C(counts, alpha, beta); lg = self.log
        class S:
            def probs(_, z, K):
                p = s.probs(z, K); lg.append(("probsC", K, h(z), h(p))); return p
            def ll(_, z, K):
                v = s.ll(z, K); lg.append(("llC", K, h(z), repr(v))); return v
            def close(_): s.close()
        return S()
    def session_G(self, counts, beta, delta, gamma):
        s = self.be.session_G(counts, beta, delta, gamma); lg = self.log
        class S:
            def probs(_, y, L):
                p = s.probs(y, L); lg.append(("probsG", L, h(y), h(p))); return p
            def ll(_, y, L):
                v = s.ll(y, L); lg.append(("llG", L, h(y), repr(v))); return v
            def close(_): s.close()
        return S()
    def colSumByGroup(self, counts, z, K):
        m = self.be.colSumByGroup(counts, z, K); self.log.append(("colsum", K, h(z), h(m))); return m

for name, fn, arg, seed in (("Z", split.initializeSplitZ, 5, 1), ("Y", split.initializeSplitY, 10, 4)):
    d, o = Log(split.DeviceBackend()), Log(OracleBackend(Oracle()))
    rd = fn(counts, arg, rng=np.random.default_rng(seed), backend=d)
    ro = fn(counts, arg, rng=np.random.default_rng(seed), backend=o)
    print(name, "equal", np.array_equal(rd, ro), len(d.log), len(o.log))
    for k, (a, b) in enumerate(zip(d.log, o.log)):
        same = a == b or (a[0] in ('llC', 'llG') and a[:3] == b[:3] and abs(float(a[3]) - float(b[3])) <= 1e-9 * abs(float(b[3])))
        if not same:
            print("  first difference at call", k); print("   dev", a); print("   ora", b); break
    else:
        print("  logs identical up to", min(len(d.log), len(o.log)))
