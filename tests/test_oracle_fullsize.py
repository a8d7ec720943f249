"""Oracle accumulation modes at configs[1] magnitudes (|logLik| ~ 4.9e9, table entries ~1e5): the long-double sum()
of x86 R and the plain-double sum() the device implements must draw the same labels."""
import numpy as np
import pytest


@pytest.mark.timeout(900)
def test_ld_and_double_modes_draw_identically_at_c2_magnitudes(oracle, oracle_ld):
    from bench_support import simulate_cells_cg
    d = simulate_cells_cg(12345, 10, 10000, 20000, 30, 150, (1000, 10000))
    K, L, nG, nM = 30, 150, d["nG"], d["nM"]
    rng = np.random.default_rng(5)
    z, y = d["z"].copy(), d["y"].copy()
    fz, fy = rng.random(z.size) < 0.05, rng.random(y.size) < 0.05
    z[fz] = rng.integers(1, K + 1, int(fz.sum()))
    y[fy] = rng.integers(1, L + 1, int(fy.sum()))
    O = (nG, nM, d["p"], d["i"], d["x"])
    p = oracle.cCGDecomposeCounts(O, d["s"], z, y, K, L)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    ny, nz = 1500, 6000
    out = []
    for o in (oracle, oracle_ld):
        oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0, 1.0, ixy,
                                uy, count=ny)
        oz = o.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, d["s"], K, L, nM, 1.0, 1.0,
                                ixz, uz, count=nz)
        out.append((oy["y"], oz["z"], oy["probs"][:, ixy[:ny] - 1], oz["probs"][:, ixz[:nz] - 1]))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert (out[0][0] != y).sum() > 30 and (out[0][1] != z).sum() > 100
    # the probabilities themselves differ in the last bits only (sums of ~1e6-magnitude terms)
    assert np.allclose(out[0][2], out[1][2], rtol=1e-9, atol=1e-6) and np.allclose(out[0][3], out[1][3], rtol=1e-9, atol=1e-6)
