"""Parity at BASELINE.json's full size (configs[1]: 100k cells x 20k genes, K=30, L=150) through properties that do
not need a full CPU run: the oracle replays the FIRST visited items of each sweep from the same state (a sweep is
sequential, so a prefix depends on nothing after it), incrementally updated tables must equal a decomposition from
scratch, and the chain must be deterministic."""
import numpy as np
import pytest

import celda_b200 as cb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from bench_support import simulate_cells_cg
    d = simulate_cells_cg(12345, 10, 10000, 20000, 30, 150, (1000, 10000))
    rng = np.random.default_rng(5)
    z, y = d["z"].copy(), d["y"].copy()
    fz, fy = rng.random(z.size) < 0.02, rng.random(y.size) < 0.02
    z[fz] = rng.integers(1, 31, int(fz.sum()))
    y[fy] = rng.integers(1, 151, int(fy.sum()))
    return d, z, y


@pytest.mark.timeout(900)
def test_full_size_prefix_parity_and_invariants(oracle, big):
    d, z0, y0 = big
    K, L, nG, nM = 30, 150, d["nG"], d["nM"]
    A = cb.CSC(nG, nM, d["p"], d["i"], d["x"])
    rng = np.random.default_rng(17)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    ch = cb.Chain(A, d["s"], K=K, L=L)
    ch.set_state(z0, y0)
    t0 = ch.tables()
    # ---- y sweep: the first 150 visited genes replayed by the oracle on the full-size tables
    ch.sweep_y(ixy, uy)
    _, y1 = ch.get_state()
    _, py = ch.probs()
    n_y = 150
    ref = oracle.cGCalcGibbsProbY(t0["nGByCP"], t0["nTSByCP"], t0["nByTS"], t0["nGByTS"], t0["nByG"], y0, L, nG, 1.0, 1.0,
                                  1.0, ixy, uy, count=n_y)
    head = ixy[:n_y] - 1
    assert np.array_equal(y1[head], ref["y"][head])
    assert np.array_equal(py[:, head], ref["probs"][:, head])
    t1 = ch.tables()
    # ---- z sweep: the first 300 visited cells
    ch.sweep_z_gibbs(ixz, uz)
    z1, _ = ch.get_state()
    pz, _ = ch.probs()
    n_z = 300
    ref = oracle.cCCalcGibbsProbZ(t1["nTSByC"], t1["mCPByS"], t1["nTSByCP"], t1["nByC"], t1["nCP"], z0, d["s"], K, L, nM,
                                  1.0, 1.0, ixz, uz, count=n_z)
    head = ixz[:n_z] - 1
    assert np.array_equal(z1[head], ref["z"][head])
    assert np.array_equal(pz[:, head], ref["probs"][:, head])
    # ---- incrementally updated tables == decomposition from scratch (all integer tables, bit-exact)
    t2 = ch.tables()
    fresh = cb.Chain(A, d["s"], K=K, L=L)
    fresh.set_state(z1, y1)
    tf = fresh.tables()
    for k in ("mCPByS", "nTSByC", "nTSByCP", "nGByCP", "nCP", "nByTS", "nGByTS", "nByC", "nByG"):
        assert np.array_equal(t2[k], tf[k]), k
    assert int(t2["nTSByC"].sum()) == int(d["x"].sum()) == int(t2["nGByCP"].sum())
    ll_a, ll_b = ch.loglik(), fresh.loglik()
    assert ll_a == ll_b
    ll_s = cb.cCGCalcLL(K, L, t2["mCPByS"], t2["nTSByCP"], t2["nByG"], t2["nByTS"], t2["nGByTS"], 10, nG, 1.0, 1.0, 1.0, 1.0)
    assert abs(ll_a - ll_s) <= 1e-12 * abs(ll_s)
    ll_o = oracle.cCGCalcLL(K, L, t2["mCPByS"], t2["nTSByCP"], t2["nByG"], t2["nByTS"], t2["nGByTS"], 10, nG, 1.0, 1.0,
                            1.0, 1.0)
    assert abs(ll_a - ll_o) <= 1e-9 * abs(ll_o)
    # ---- determinism: a second chain fed the same draws ends in the same state; doSample = FALSE changes nothing
    again = cb.Chain(A, d["s"], K=K, L=L)
    again.set_state(z0, y0)
    ll2 = again.iterate(ixy, uy, ixz, uz)
    z2, y2 = again.get_state()
    assert np.array_equal(z2, z1) and np.array_equal(y2, y1) and ll2 == ll_a
    again.sweep_z_gibbs(ixz, uz, doSample=False)
    z3, _ = again.get_state()
    assert np.array_equal(z3, z2)
    assert (z1 != z0).sum() > 500 and (y1 != y0).sum() > 100  # the sweeps did move labels
    for c in (ch, fresh, again):
        c.close()
