"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Bit-exact for labels, integer tables and sweep probabilities; 1e-9 relative (north_star) for log-likelihood."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _csc(counts):
    return cb.CSC.from_dense(counts), dense_to_csc(counts)


def simulate_cg(rng, G, C, K, L, S, nrange=(500, 1000)):
    """Small restatement of .simulateCellscelda_CG (R/simulateCells.R:302-435) for test inputs."""
    s = rng.integers(1, S + 1, C).astype(np.int32)
    s[:S] = np.arange(1, S + 1)
    theta = rng.dirichlet(np.ones(K), S)
    z = np.array([rng.choice(K, p=theta[si - 1]) for si in s]) + 1
    phi = rng.dirichlet(np.ones(L), K)
    y = rng.integers(1, L + 1, G)
    y[:L] = np.arange(1, L + 1)
    counts = np.zeros((G, C), dtype=np.int32)
    psi = np.zeros((L, G))
    for l in range(L):
        idx = np.where(y == l + 1)[0]
        psi[l, idx] = rng.dirichlet(np.ones(idx.size))
    for c in range(C):
        n = rng.integers(nrange[0], nrange[1] + 1)
        mod = rng.multinomial(n, phi[z[c] - 1])
        for l in range(L):
            if mod[l]:
                counts[:, c] += rng.multinomial(mod[l], psi[l])
    keep = counts.sum(1) > 0
    counts[~keep, 0] = 1  # no all-zero rows (.validateCounts, R/celda_functions.R:477-487)
    counts[0, counts.sum(0) == 0] = 1
    return counts, s, z.astype(np.int32), y.astype(np.int32)


# ------------------------------------------------------------------ arithmetic
def test_device_math_bit_identical_to_oracle(oracle):
    rng = np.random.default_rng(3)
    x = np.concatenate([np.arange(0, 30000) + 1.0, np.arange(0, 5000) + 0.5, rng.uniform(1e-3, 12, 50000),
                        rng.uniform(10, 1e6, 200000), 10.0 ** rng.uniform(1, 18, 50000),
                        rng.integers(4_000_000, 3_000_000_000, 100000) + 1.0,  # table totals around lgammafn's cut-off
                        [1e-310, 4934720.0, 4934721.0, 94906265.0, 94906266.0, 1e17, 2e17]])
    assert np.array_equal(cb.lgamma(x), oracle.lgamma(x))
    assert np.array_equal(cb.log(x), oracle.log(x))
    e = np.concatenate([rng.uniform(-760, 5, 300000), -10.0 ** rng.uniform(-20, 2.9, 100000), [0.0, -0.0, -745.2]])
    assert np.array_equal(cb.exp(e), oracle.exp(e))


# ------------------------------------------------------------------ aggregation (a2-a5)
def test_sparse_aggregation_bit_exact(oracle, gold_cg):
    counts = gold_cg["counts"]
    A, O = _csc(counts)
    z, y = gold_cg["z"], gold_cg["y"]
    assert np.array_equal(cb.colSumByGroupSparse(A, z, 5), oracle.colSumByGroupSparse(O, z, 5))
    assert np.array_equal(cb.rowSumByGroupSparse(A, y, 10), oracle.rowSumByGroupSparse(O, y, 10))
    rng = np.random.default_rng(0)
    z2 = rng.integers(1, 6, z.size).astype(np.int32)
    y2 = rng.integers(1, 11, y.size).astype(np.int32)
    px = cb.colSumByGroupSparse(A, z, 5)
    r = cb.colSumByGroupChangeSparse(A, px, z2, z, 5)
    assert r is px and np.array_equal(px, oracle.colSumByGroupSparse(O, z2, 5))
    px = cb.rowSumByGroupSparse(A, y, 10)
    cb.rowSumByGroupChangeSparse(A, px, y2, y, 10)
    assert np.array_equal(px, oracle.rowSumByGroupSparse(O, y2, 10))


def test_sparse_aggregation_error_contract(gold_cg):
    A, _ = _csc(gold_cg["counts"])
    z, y = gold_cg["z"], gold_cg["y"]
    for bad in (lambda: cb.colSumByGroupSparse(A, z[:-1], 5), lambda: cb.colSumByGroupSparse(A, z, 4),
                lambda: cb.colSumByGroupSparse(A, z, 1000), lambda: cb.rowSumByGroupSparse(A, y, 101),
                lambda: cb.rowSumByGroupSparse(A, y * 0, 10), lambda: cb.rowSumByGroupSparse(A, y[:-1], 10),
                lambda: cb.colSumByGroupChangeSparse(A, np.zeros((99, 5), order="F"), z, z, 5),
                lambda: cb.rowSumByGroupChangeSparse(A, np.zeros((10, 424), order="F"), y, y, 10),
                lambda: cb.colSumByGroupChangeSparse(A, np.zeros((100, 5), order="F"), z, z[:-1], 5)):
        with pytest.raises(cb.CeldaCudaError):
            bad()


def test_aggregation_ragged_and_empty_columns(oracle):
    rng = np.random.default_rng(5)
    counts = rng.poisson(0.05, (300, 700)).astype(np.int32)
    counts[:, 10:20] = 0  # empty columns
    counts[5, :] = 0  # empty row
    counts[:, 33] = rng.integers(0, 50, 300)  # one dense column
    A, O = _csc(counts)
    z = rng.integers(1, 8, 700).astype(np.int32)
    y = rng.integers(1, 41, 300).astype(np.int32)
    assert np.array_equal(cb.colSumByGroupSparse(A, z, 7), oracle.colSumByGroupSparse(O, z, 7))
    assert np.array_equal(cb.rowSumByGroupSparse(A, y, 40), oracle.rowSumByGroupSparse(O, y, 40))


# ------------------------------------------------------------------ chain: decompose, LL
def test_chain_decompose_and_loglik_golden(oracle, gold_cg):
    counts, s, z, y = gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"]
    A, O = _csc(counts)
    ch = cb.Chain(A, s, K=5, L=10)
    ch.set_state(z, y)
    t = ch.tables()
    p = oracle.cCGDecomposeCounts(O, s, z, y, 5, 10)
    for k in ("mCPByS", "nTSByC", "nTSByCP", "nGByCP", "nCP", "nByC", "nByG", "nByTS", "nGByTS"):
        assert np.array_equal(t[k], p[k]), k
    ll = ch.loglik()
    assert abs(ll - float(gold_cg["finalLogLik"])) / abs(ll) < 1e-12  # data/celdaCGMod.rda
    ref = oracle.cCGCalcLL(5, 10, p["mCPByS"], p["nTSByCP"], p["nByG"], p["nByTS"], p["nGByTS"], p["nS"], p["nG"],
                           1.0, 1.0, 1.0, 1.0)
    assert abs(ll - ref) / abs(ref) < 1e-12
    ch.close()


def test_chain_loglik_grid_models(gold_cg, gold_grid):
    A, _ = _csc(gold_cg["counts"])
    for m in range(9):
        K, L = int(gold_grid["K"][m]), int(gold_grid["L"][m])
        ch = cb.Chain(A, gold_grid["s"], K=K, L=L)
        ch.set_state(gold_grid["z"][m], gold_grid["y"][m])
        ll = ch.loglik()
        assert abs(ll - gold_grid["finalLogLik"][m]) / abs(ll) < 1e-12
        ch.close()


def test_chain_rejects_bad_input(gold_cg):
    counts, s = gold_cg["counts"], gold_cg["s_mod"]
    A, _ = _csc(counts)
    bad = cb.CSC(A.nrow, A.ncol, A.p, A.i, A.x + 0.5)
    with pytest.raises(cb.CeldaCudaError):
        cb.Chain(bad, s, K=5, L=10)
    ch = cb.Chain(A, s, K=5, L=10)
    with pytest.raises(cb.CeldaCudaError, match="greater than the total number"):
        ch.set_state(gold_cg["z"] + 1, gold_cg["y"])  # R/celda_CG.R "greater than the total number" errors
    with pytest.raises(cb.CeldaCudaError, match="greater than the total number"):
        ch.set_state(gold_cg["z"], gold_cg["y"] + 5)
    with pytest.raises(cb.CeldaCudaError):
        ch.sweep_y(np.arange(1, 101), np.zeros(100))  # no state yet
    ch.set_state(gold_cg["z"], gold_cg["y"])
    with pytest.raises(cb.CeldaCudaError):
        ch.sweep_y(np.ones(100, np.int32), np.zeros(100))  # not a permutation
    ch.close()


# ------------------------------------------------------------------ sweeps, bit-exact under injected uniforms
def _oracle_iteration(o, O, s, z, y, K, L, ixy, uy, ixz, uz, hyper=(1.0, 1.0, 1.0, 1.0)):
    a, b, d, g = hyper
    nG, nM = O[0], O[1]
    p = o.cCGDecomposeCounts(O, s, z, y, K, L)
    oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, b, d, g, ixy, uy)
    nTSByC = o.rowSumByGroupChangeSparse(O, p["nTSByC"].copy(order="F"), oy["y"], y, L)
    oz = o.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, s, K, L, nM, a, b, ixz, uz)
    nGByCP = o.colSumByGroupChangeSparse(O, p["nGByCP"].copy(order="F"), oz["z"], z, K)
    ll = o.cCGCalcLL(K, L, oz["mCPByS"], oz["nGByCP"], p["nByG"], oy["nByTS"], oy["nGByTS"], p["nS"], nG, a, b, d, g)
    return dict(y=oy["y"], z=oz["z"], probsY=oy["probs"], probsZ=oz["probs"], nTSByC=nTSByC, nTSByCP=oz["nGByCP"],
                nGByCP=nGByCP, nCP=oz["nCP"], mCPByS=oz["mCPByS"], nByTS=oy["nByTS"], nGByTS=oy["nGByTS"], ll=ll)


def _check_iteration(ch, ref, ll):
    z, y = ch.get_state()
    assert np.array_equal(y, ref["y"]), f"{(y != ref['y']).sum()} y labels differ"
    assert np.array_equal(z, ref["z"]), f"{(z != ref['z']).sum()} z labels differ"
    t = ch.tables()
    for k in ("mCPByS", "nTSByC", "nTSByCP", "nGByCP", "nCP", "nByTS", "nGByTS"):
        assert np.array_equal(t[k], ref[k]), k
    pz, py = ch.probs()
    assert np.array_equal(py, ref["probsY"])
    assert np.array_equal(pz, ref["probsZ"])
    assert abs(ll - ref["ll"]) / abs(ref["ll"]) < 1e-9


@pytest.mark.parametrize("start", ["random", "fitted", "truth10"])
def test_cg_iteration_bit_exact_fixture(oracle, gold_cg, start):
    """configs[0]: celda_CG on celdaCGSim (K=5, L=10), Gibbs; three start states (worst-case churn, converged,
    10 % perturbed)."""
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(11)
    if start == "random":
        z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    elif start == "fitted":
        z, y = gold_cg["z"].copy(), gold_cg["y"].copy()
    else:
        z, y = gold_cg["sim_z"].copy(), gold_cg["sim_y"].copy()
        fz, fy = rng.random(nM) < 0.1, rng.random(nG) < 0.1
        z[fz] = rng.integers(1, K + 1, fz.sum())
        y[fy] = rng.integers(1, L + 1, fy.sum())
    A, O = _csc(counts)
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    for it in range(3):
        ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
        ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
        ref = _oracle_iteration(oracle, O, s, z, y, K, L, ixy, uy, ixz, uz)
        ll = ch.iterate(ixy, uy, ixz, uz)
        _check_iteration(ch, ref, ll)
        z, y = ref["z"], ref["y"]
    ch.close()


def test_cg_separate_sweeps_and_no_sample(oracle, gold_cg):
    """sweep_y / sweep_z entry points; doSample=FALSE computes probabilities only (clusterProbability use)."""
    counts, s, K, L = gold_cg["counts"], gold_cg["s_mod"], 5, 10
    z, y = gold_cg["z"], gold_cg["y"]
    nG, nM = counts.shape
    A, O = _csc(counts)
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
    ixz, ixy = np.arange(1, nM + 1, dtype=np.int32), np.arange(1, nG + 1, dtype=np.int32)
    ch.sweep_z_gibbs(ixz, np.zeros(nM), doSample=False)
    ch.sweep_y(ixy, np.zeros(nG), doSample=False)
    pz, py = ch.probs()
    oz = oracle.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0,
                                 ixz, np.zeros(nM), doSample=False)
    oy = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                 1.0, ixy, np.zeros(nG), doSample=False)
    assert np.array_equal(pz, oz["probs"]) and np.array_equal(py, oy["probs"])
    z2, y2 = ch.get_state()
    assert np.array_equal(z2, z) and np.array_equal(y2, y)
    ch.close()


@pytest.mark.parametrize("hyper", [(1.0, 1.0, 1.0, 1.0), (0.5, 0.1, 2.0, 5.0)])
def test_cg_iteration_bit_exact_simulated(oracle, hyper):
    """Mid-size simulated data, K=12 L=40, non-default concentration parameters included."""
    rng = np.random.default_rng(21)
    G, C, K, L, S = 400, 1500, 12, 40, 4
    counts, s, z_true, y_true = simulate_cg(rng, G, C, K, L, S)
    A, O = _csc(counts)
    z, y = z_true.copy(), y_true.copy()
    fz, fy = rng.random(C) < 0.1, rng.random(G) < 0.1
    z[fz] = rng.integers(1, K + 1, fz.sum())
    y[fy] = rng.integers(1, L + 1, fy.sum())
    ch = cb.Chain(A, s, K=K, L=L, alpha=hyper[0], beta=hyper[1], delta=hyper[2], gamma=hyper[3])
    ch.set_state(z, y)
    for it in range(2):
        ixy, uy = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
        ixz, uz = rng.permutation(C).astype(np.int32) + 1, rng.random(C)
        ref = _oracle_iteration(oracle, O, s, z, y, K, L, ixy, uy, ixz, uz, hyper)
        ll = ch.iterate(ixy, uy, ixz, uz)
        _check_iteration(ch, ref, ll)
        z, y = ref["z"], ref["y"]
    ch.close()


def test_sampler_ties_and_empty_clusters(oracle, gold_cg):
    """Two empty populations give exactly tied probabilities: the chosen label must follow R's heapsort order."""
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 8, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(4)
    z = rng.integers(1, 6, nM).astype(np.int32)  # populations 6, 7, 8 start empty
    y = gold_cg["sim_y"].copy()
    A, O = _csc(counts)
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    for it in range(2):
        ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
        ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
        uz[::7] = 1.0 - 1e-17  # beyond the total mass: falls through to the last sorted position
        ref = _oracle_iteration(oracle, O, s, z, y, K, L, ixy, uy, ixz, uz)
        ll = ch.iterate(ixy, uy, ixz, uz)
        _check_iteration(ch, ref, ll)
        z, y = ref["z"], ref["y"]
    ch.close()


def test_staged_iterations_match_host_fed(oracle, gold_cg):
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(8)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    n_it = 3
    ixy = np.stack([rng.permutation(nG).astype(np.int32) + 1 for _ in range(n_it)])
    ixz = np.stack([rng.permutation(nM).astype(np.int32) + 1 for _ in range(n_it)])
    uy, uz = rng.random((n_it, nG)), rng.random((n_it, nM))
    A, _ = _csc(counts)
    a, b = cb.Chain(A, s, K=K, L=L), cb.Chain(A, s, K=K, L=L)
    a.set_state(z, y)
    b.set_state(z, y)
    b.stage_draws(ixy, uy, ixz, uz)
    for it in range(n_it):
        la = a.iterate(ixy[it], uy[it], ixz[it], uz[it])
        lb = b.iterate_staged(it)
        assert la == lb
    za, ya = a.get_state()
    zb, yb = b.get_state()
    assert np.array_equal(za, zb) and np.array_equal(ya, yb)
    a.close()
    b.close()


def test_cg_iteration_more_than_32_populations(oracle):
    """K = 40 > 32 and L = 70: several candidates per lane in the draw and in the warp-mode units."""
    rng = np.random.default_rng(77)
    G, C, K, L, S = 300, 900, 40, 70, 3
    counts, s, z_true, y_true = simulate_cg(rng, G, C, K, L, S, nrange=(200, 400))
    A, O = _csc(counts)
    z, y = z_true.copy(), y_true.copy()
    fz, fy = rng.random(C) < 0.3, rng.random(G) < 0.3
    z[fz] = rng.integers(1, K + 1, fz.sum())
    y[fy] = rng.integers(1, L + 1, fy.sum())
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    for it in range(2):
        ixy, uy = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
        ixz, uz = rng.permutation(C).astype(np.int32) + 1, rng.random(C)
        ref = _oracle_iteration(oracle, O, s, z, y, K, L, ixy, uy, ixz, uz)
        ll = ch.iterate(ixy, uy, ixz, uz)
        _check_iteration(ch, ref, ll)
        z, y = ref["z"], ref["y"]
    ch.close()


@pytest.mark.parametrize("start", ["truth10", "random", "truth10-truncated"])
def test_cg_iteration_value_tables_large_windows(oracle, start):
    """>= 4096 cells: the z sweep's lane-per-unit rounds run on per-candidate value tables (built once per round by the
    grid, staged in shared memory one row group at a time) and must still equal the sequential oracle bit for bit.
    A few large counts stretch some rows' tables; in the 'truncated' case one module count exceeds the shared-memory
    capacity, so its row's table is cut short and look-ups beyond it are evaluated directly."""
    rng = np.random.default_rng(91)
    G, C, K, L, S = 240, 6000, 9, 24, 3
    counts, s, z_true, y_true = simulate_cg(rng, G, C, K, L, S, nrange=(300, 900))
    counts[:8, rng.integers(0, C, 40)] += rng.integers(500, 4000, (8, 40)).astype(np.int32)
    if start.endswith("truncated"):
        counts[3, rng.integers(0, C, 5)] += rng.integers(40000, 90000, 5).astype(np.int32)
        start = "truth10"
    A, O = _csc(counts)
    if start == "random":
        z, y = rng.integers(1, K + 1, C).astype(np.int32), rng.integers(1, L + 1, G).astype(np.int32)
    else:
        z, y = z_true.copy(), y_true.copy()
        fz, fy = rng.random(C) < 0.1, rng.random(G) < 0.1
        z[fz] = rng.integers(1, K + 1, fz.sum())
        y[fy] = rng.integers(1, L + 1, fy.sum())
    ch = cb.Chain(A, s, K=K, L=L)
    ch.set_state(z, y)
    for it in range(3):
        ixy, uy = rng.permutation(G).astype(np.int32) + 1, rng.random(G)
        ixz, uz = rng.permutation(C).astype(np.int32) + 1, rng.random(C)
        ref = _oracle_iteration(oracle, O, s, z, y, K, L, ixy, uy, ixz, uz)
        ll = ch.iterate(ixy, uy, ixz, uz)
        _check_iteration(ch, ref, ll)
        z, y = ref["z"], ref["y"]
    ch.close()
