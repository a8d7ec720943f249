"""Logic prototype of the multi-commit speculative round (host-side model of k_sweep_z's scheduling).

Not product code and not used by tests of the CUDA path: it checks, on small random problems, that the round
schedule (tentative commits as extra column states, stale-column sets, validity flags, discard of commits behind the
first invalid item) reproduces a strictly sequential sweep.  The arithmetic is a stand-in (scipy lgamma); only the
bookkeeping is under test.
"""
import numpy as np
from scipy.special import gammaln


def unit(col, x, sgn, beta):
    return float(np.sum(gammaln(col + sgn * x + beta)))


def colsum(col, beta):
    return float(np.sum(gammaln(col + beta)))


def draw(p, u):
    e = np.exp(p - p.max())
    q = e / e.sum()
    order = np.argsort(-q, kind="stable")
    c = np.cumsum(q[order])
    k = int(np.searchsorted(c, u, side="left"))
    return int(order[min(k, len(p) - 1)])


def probs_for(n, nCP, m, x, N, zi, si, alpha, beta):
    R, K = n.shape
    p = np.empty(K)
    for j in range(K):
        sgn = -1 if j == zi else 1
        S = unit(n[:, j], x, sgn, beta)
        A = gammaln(nCP[j] + sgn * N + R * beta)
        cs = colsum(n[:, j], beta)
        lc = gammaln(nCP[j] + R * beta)
        lm = np.log(m[j, si] - (j == zi) + alpha)
        p[j] = lm + S - A - cs + lc if j != zi else lm + cs - lc - S + A
    return p


def sweep_seq(X, n, nCP, m, z, s, ix, u, alpha, beta):
    n, nCP, m, z = n.copy(), nCP.copy(), m.copy(), z.copy()
    for t, i in enumerate(ix):
        zi = z[i]
        p = probs_for(n, nCP, m, X[:, i], X[:, i].sum(), zi, s[i], alpha, beta)
        zn = draw(p, u[t])
        if zn != zi:
            n[:, zi] -= X[:, i]; n[:, zn] += X[:, i]
            nCP[zi] -= X[:, i].sum(); nCP[zn] += X[:, i].sum()
            m[zi, s[i]] -= 1; m[zn, s[i]] += 1
            z[i] = zn
    return z, n


def sweep_multi(X, n, nCP, m, z, s, ix, u, alpha, beta, W=12, MC=4, stats=None):
    R, K = n.shape
    nM = len(ix)
    n, nCP, m, z = n.copy(), nCP.copy(), m.copy(), z.copy()
    # per item ring state
    S = {}      # t -> array K of (unit S, A) pairs
    spec = {}   # t -> (old,new) or None
    inval = {}  # t -> bool
    pos, hi, prev_end = 0, 0, 0
    commits = []  # last round's tentative commits: dict(t,i,a,b,si,N, colA, colB, nA, nB)
    Dcarry = set()
    rounds = 0
    have = False
    while True:
        if have:
            f = next((t for t in range(pos, prev_end) if inval[t]), None)
            qf = len(commits) if f is None else sum(1 for c in commits if c["t"] < f)
            for c in commits[:qf]:  # standing: advance the base
                n[:, c["a"]] = c["colA"]; n[:, c["b"]] = c["colB"]
                nCP[c["a"]] = c["nA"]; nCP[c["b"]] = c["nB"]
                m[c["a"], c["si"]] -= 1; m[c["b"], c["si"]] += 1
                z[c["i"]] = c["b"]
            Dcarry = set()
            for c in commits[qf:]:
                Dcarry |= {c["a"], c["b"]}
            if qf < len(commits):
                pos = f  # no unselected mover can lie before a discarded selected one
            elif commits:
                pos = commits[-1]["t"] + 1  # unselected movers (beyond MC) may follow: the next selection finds them
            hi = min(prev_end, pos + W)
            if stats is not None:
                stats.append((len(commits), qf))
        if pos >= nM:
            break
        rounds += 1
        # select tentative commits among carried items
        commits = []
        if have:
            for t in range(pos, hi):
                if spec[t] is not None:
                    i = ix[t]
                    commits.append(dict(t=t, i=i, a=spec[t][0], b=spec[t][1], si=s[i], N=X[:, i].sum()))
                    if len(commits) == MC:
                        break
            if not commits and not Dcarry:  # nothing moves and nothing is stale: the carried items are settled
                pos = hi
                if pos >= nM:
                    break
        # post-images
        cur = {j: (n[:, j], nCP[j]) for j in range(K)}
        for c in commits:
            x = X[:, c["i"]]
            ca, na = cur[c["a"]]; cb, nb = cur[c["b"]]
            c["colA"], c["nA"] = ca - x, na - c["N"]
            c["colB"], c["nB"] = cb + x, nb + c["N"]
            cur[c["a"]] = (c["colA"], c["nA"]); cur[c["b"]] = (c["colB"], c["nB"])

        def state(q, j):  # column j after the first q tentative commits
            for c in reversed(commits[:q]):
                if c["a"] == j: return c["colA"], c["nA"]
                if c["b"] == j: return c["colB"], c["nB"]
            return n[:, j], nCP[j]

        def mstate(q, j, si):
            v = m[j, si]
            for c in commits[:q]:
                if c["si"] == si:
                    v += int(c["b"] == j) - int(c["a"] == j)
            return v

        def qof(t):
            return sum(1 for c in commits if c["t"] < t)

        def colsQ(q):
            out = set(Dcarry)
            for c in commits[:q]:
                out |= {c["a"], c["b"]}
            return out

        topup = (not have) or not commits and not Dcarry or (hi - pos) * 2 < W
        end = max(hi, min(pos + W, nM)) if topup else hi
        M = len(commits)

        def evalunit(t, j, q):
            i = ix[t]; zi = z[i]
            col, ncp = state(q, j)
            sgn = -1 if j == zi else 1
            return unit(col, X[:, i], sgn, beta), gammaln(ncp + sgn * X[:, i].sum() + R * beta)

        def vector(t, q):
            i = ix[t]; zi = z[i]; si = s[i]
            p = np.empty(K)
            for j in range(K):
                col, ncp = state(q, j)
                cs, lc = colsum(col, beta), gammaln(ncp + R * beta)
                lm = np.log(mstate(q, j, si) - (j == zi) + alpha)
                Sv, Av = S[t][j]
                p[j] = lm + Sv - Av - cs + lc if j != zi else lm + cs - lc - Sv + Av
            return p

        # units
        for t in range(pos, hi):
            q = qof(t)
            for j in colsQ(q):
                S[t][j] = evalunit(t, j, q)
        for t in range(hi, end):
            S[t] = [evalunit(t, j, M) for j in range(K)]
        # draws / validation
        for t in range(pos, end):
            fresh = t >= hi
            q = M if fresh else qof(t)
            if not fresh and not colsQ(q):
                inval[t] = False
                continue
            i = ix[t]; zi = z[i]
            zn = draw(vector(t, q), u[t])
            new = None if zn == zi else (zi, zn)
            inval[t] = (not fresh) and (new != spec[t])
            spec[t] = new
        prev_end = end
        have = True
    return z, n, rounds


def main():
    rng = np.random.default_rng(0)
    tot = 0
    for trial in range(300):
        R, K, nM, nS = rng.integers(3, 8), rng.integers(2, 6), rng.integers(5, 60), rng.integers(1, 4)
        scale = rng.choice([1, 3, 30])
        X = rng.poisson(scale, size=(R, nM)).astype(np.int64) + (rng.random((R, nM)) < 0.3)
        z = rng.integers(0, K, nM); s = rng.integers(0, nS, nM)
        n = np.zeros((R, K), np.int64); m = np.zeros((K, nS), np.int64)
        for i in range(nM):
            n[:, z[i]] += X[:, i]; m[z[i], s[i]] += 1
        nCP = n.sum(0)
        ix = rng.permutation(nM); u = rng.random(nM)
        z0, n0 = sweep_seq(X, n, nCP, m, z, s, ix, u, 1.0, 1.0)
        st = []
        z1, n1, rounds = sweep_multi(X, n, nCP, m, z, s, ix, u, 1.0, 1.0, W=int(rng.integers(2, 20)), MC=int(rng.integers(1, 6)), stats=st)
        assert np.array_equal(z0, z1) and np.array_equal(n0, n1), trial
        tot += sum(1 for a, b in st if b < a)
    print("ok; rounds with discarded commits:", tot)


if __name__ == "__main__":
    main()
