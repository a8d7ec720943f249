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


@pytest.mark.timeout(1200)
def test_full_size_whole_sweeps_bit_exact(oracle, big):
    """One FULL y sweep and one FULL z sweep at configs[1] against the oracle: every label, every table and both
    probability matrices, array_equal (R/celda_CG.R:543-585)."""
    d, z0, y0 = big
    K, L, nG, nM = 30, 150, d["nG"], d["nM"]
    A = cb.CSC(nG, nM, d["p"], d["i"], d["x"])
    O = (nG, nM, d["p"], d["i"], d["x"])
    rng = np.random.default_rng(23)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    ch = cb.Chain(A, d["s"], K=K, L=L)
    ch.set_state(z0, y0)
    p = oracle.cCGDecomposeCounts(O, d["s"], z0, y0, K, L)
    t0 = ch.tables()
    for k in ("mCPByS", "nTSByC", "nTSByCP", "nGByCP", "nCP", "nByTS", "nGByTS", "nByC", "nByG"):
        assert np.array_equal(t0[k], p[k]), k
    # ---- y sweep
    ch.sweep_y(ixy, uy)
    _, y1 = ch.get_state()
    _, py = ch.probs()
    oy = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y0, L, nG, 1.0, 1.0, 1.0,
                                 ixy, uy)
    assert np.array_equal(y1, oy["y"]), int((y1 != oy["y"]).sum())
    assert np.array_equal(py, oy["probs"])
    t1 = ch.tables()
    assert np.array_equal(t1["nTSByCP"], oy["nTSByC"]) and np.array_equal(t1["nByTS"], oy["nByTS"])
    assert np.array_equal(t1["nGByTS"], oy["nGByTS"])
    nTSByC = oracle.rowSumByGroupChangeSparse(O, p["nTSByC"], oy["y"], y0, L)
    assert np.array_equal(t1["nTSByC"], nTSByC)
    # ---- z sweep
    ch.sweep_z_gibbs(ixz, uz)
    z1, _ = ch.get_state()
    pz, _ = ch.probs()
    oz = oracle.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z0, d["s"], K, L, nM, 1.0, 1.0,
                                 ixz, uz)
    assert np.array_equal(z1, oz["z"]), int((z1 != oz["z"]).sum())
    assert np.array_equal(pz, oz["probs"])
    t2 = ch.tables()
    assert np.array_equal(t2["nTSByCP"], oz["nGByCP"]) and np.array_equal(t2["mCPByS"], oz["mCPByS"])
    assert np.array_equal(t2["nCP"], oz["nCP"])
    nGByCP = oracle.colSumByGroupChangeSparse(O, p["nGByCP"], oz["z"], z0, K)
    assert np.array_equal(t2["nGByCP"], nGByCP)
    assert (z1 != z0).sum() > 1000 and (y1 != y0).sum() > 200
    ch.close()


def test_commits_per_round_do_not_change_results(gold_cg):
    """The number of tentative commits per round (CELDA_CUDA_MC) is a schedule, not a result: 1 (one mover per round),
    4 and the default give identical labels, tables and probabilities from a churning start."""
    import os
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(91)
    z = rng.integers(1, K + 1, nM).astype(np.int32)
    y = rng.integers(1, L + 1, nG).astype(np.int32)
    draws = [(rng.permutation(nG).astype(np.int32) + 1, rng.random(nG), rng.permutation(nM).astype(np.int32) + 1,
              rng.random(nM)) for _ in range(3)]
    res = []
    old = os.environ.get("CELDA_CUDA_MC")
    try:
        for mc in ("1", "4", None):
            if mc is None:
                os.environ.pop("CELDA_CUDA_MC", None)
            else:
                os.environ["CELDA_CUDA_MC"] = mc
            ch = cb.Chain(cb.CSC.from_dense(counts), s, K=K, L=L)
            ch.set_state(z, y)
            lls = [ch.iterate(*dr) for dr in draws]
            res.append((ch.get_state(), ch.probs(), ch.tables(), lls))
            ch.close()
    finally:
        if old is None:
            os.environ.pop("CELDA_CUDA_MC", None)
        else:
            os.environ["CELDA_CUDA_MC"] = old
    for r in res[1:]:
        assert np.array_equal(r[0][0], res[0][0][0]) and np.array_equal(r[0][1], res[0][0][1])
        assert np.array_equal(r[1][0], res[0][1][0]) and np.array_equal(r[1][1], res[0][1][1])
        assert all(np.array_equal(r[2][k], res[0][2][k]) for k in r[2])
        assert r[3] == res[0][3]
