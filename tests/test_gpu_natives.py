"""GPU parity for the stand-alone native entry points: dense aggregation twins, the
single-gene probability vector, normaliser, A^T B products and _perplexityG."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _rowsum(mat, lab, n):
    out = np.zeros((n, mat.shape[1]), dtype=mat.dtype)
    for i, g in enumerate(lab):
        out[g - 1] += mat[i]
    return out


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_matrixSums_known_answers(dtype):
    """tests/testthat/test-matrixSums.R:5-76 through the C ABI."""
    mat = np.asfortranarray(np.tile(np.arange(1, 6), 20).reshape(10, 10, order="F").astype(dtype))
    lab = np.repeat([1, 2], 5).astype(np.int32)
    want_r, want_c = _rowsum(mat, lab, 2), _rowsum(mat.T, lab, 2).T
    assert np.array_equal(cb.rowSumByGroup_dense(mat, lab, 2), want_r)
    assert np.array_equal(cb.colSumByGroup_dense(mat, lab, 2), want_c)
    lab2 = np.array([1, 2, 1, 2, 1, 2, 1, 2, 1, 2], dtype=np.int32)
    px = np.asfortranarray(want_r.copy())
    assert cb.rowSumByGroupChange_dense(mat, px, lab2, lab, 2) is px
    assert np.array_equal(px, _rowsum(mat, lab2, 2))
    px = np.asfortranarray(want_c.copy())
    cb.colSumByGroupChange_dense(mat, px, lab2, lab, 2)
    assert np.array_equal(px, _rowsum(mat.T, lab2, 2).T)
    with pytest.raises(cb.CeldaCudaError, match="must be a factor"):
        cb.rowSumByGroup_dense(mat, lab, -1)
    with pytest.raises(cb.CeldaCudaError, match="must match the number of rows"):
        cb.rowSumByGroup_dense(mat, lab[:9], 2)
    with pytest.raises(cb.CeldaCudaError, match="must match the number of columns"):
        cb.colSumByGroup_dense(mat, lab[:9], 2)
    with pytest.raises(cb.CeldaCudaError, match="same number of levels"):
        cb.rowSumByGroupChange_dense(mat, np.asfortranarray(want_r.copy()), lab2, lab, 2, 3)
    with pytest.raises(cb.CeldaCudaError):
        cb.colSumByGroupChange_dense(mat, np.asfortranarray(np.zeros((3, 2), dtype=dtype)), lab2, lab, 2)
    na = lab.copy()
    na[3] = np.iinfo(np.int32).min
    with pytest.raises(cb.CeldaCudaError, match="must not be NA"):
        cb.rowSumByGroup_dense(mat, na, 2)


def test_dense_twins_bit_exact_fractional_and_wrap(oracle):
    """Fractional doubles (DecontX-style input): the gather order equals the reference loop order, so
    results are bit-identical; int32 sums wrap like the reference's int arithmetic."""
    rng = np.random.default_rng(9)
    x = np.asfortranarray(rng.gamma(0.3, 2.0, (257, 331)))
    gr, gc = rng.integers(1, 8, 257).astype(np.int32), rng.integers(1, 12, 331).astype(np.int32)
    assert np.array_equal(cb.rowSumByGroup_dense(x, gr, 7), oracle.rowSumByGroup(x, gr, 7))
    assert np.array_equal(cb.colSumByGroup_dense(x, gc, 11), oracle.colSumByGroup(x, gc, 11))
    gr2, gc2 = rng.integers(1, 8, 257).astype(np.int32), rng.integers(1, 12, 331).astype(np.int32)
    px, po = cb.rowSumByGroup_dense(x, gr, 7), oracle.rowSumByGroup(x, gr, 7)
    assert np.array_equal(cb.rowSumByGroupChange_dense(x, px, gr2, gr, 7), oracle.rowSumByGroupChange(x, po, gr2, gr, 7))
    px, po = cb.colSumByGroup_dense(x, gc, 11), oracle.colSumByGroup(x, gc, 11)
    assert np.array_equal(cb.colSumByGroupChange_dense(x, px, gc2, gc, 11), oracle.colSumByGroupChange(x, po, gc2, gc, 11))
    big = np.asfortranarray(np.full((4, 3), 2 ** 30, dtype=np.int32))
    g = np.ones(4, np.int32)
    assert np.array_equal(cb.rowSumByGroup_dense(big, g, 1), oracle.rowSumByGroup(big, g, 1))  # wraps to 0
    # dispatchers (R/matrixSums.R:1-51)
    counts = rng.poisson(1.0, (60, 90)).astype(np.int32)
    A = cb.CSC.from_dense(counts)
    y, z = rng.integers(1, 6, 60).astype(np.int32), rng.integers(1, 5, 90).astype(np.int32)
    assert np.array_equal(cb.rowSumByGroup(A, y, 5), cb.rowSumByGroup(counts, y, 5))
    assert np.array_equal(cb.colSumByGroup(A, z, 4), cb.colSumByGroup(counts.astype(np.float64), z, 4))


def test_cG_CalcGibbsProbY_single_gene(oracle, gold_cg):
    counts, s, z, y = gold_cg["counts"], gold_cg["s_mod"], gold_cg["z"], gold_cg["y"]
    K, L, nG = 5, 10, counts.shape[0]
    p = oracle.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, K, L)
    for gene in (1, 14, 58, 100):
        got = cb.cG_CalcGibbsProbY(gene, p["nGByCP"][gene - 1], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L,
                                   nG, 1.0, 1.0, 1.0)
        want = oracle.cG_CalcGibbsProbY(gene, p["nGByCP"][gene - 1], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"],
                                        y, L, 1.0, 1.0, 1.0)
        assert np.array_equal(got, want)
    # celda_G form: the raw count row against the L x C table
    pg = oracle.cGDecomposeCounts(dense_to_csc(counts), y, L)
    got = cb.cG_CalcGibbsProbY(7, counts[6].astype(np.float64), pg["nTSByC"], pg["nByTS"], pg["nGByTS"], pg["nByG"], y,
                               L, nG, 0.5, 2.0, 1.5)
    want = oracle.cG_CalcGibbsProbY(7, counts[6].astype(np.float64), pg["nTSByC"], pg["nByTS"], pg["nGByTS"],
                                    pg["nByG"], y, L, 0.5, 2.0, 1.5)
    assert np.array_equal(got, want)


def test_norm_matmult_perplexityG(oracle):
    rng = np.random.default_rng(12)
    x = rng.poisson(3.0, (40, 17)).astype(np.float64)
    assert np.array_equal(cb.fastNormPropLog(x, 0.7), oracle.fastNormPropLog(x, 0.7))
    with pytest.raises(cb.CeldaCudaError, match="Division by 0"):
        cb.fastNormPropLog(np.zeros((3, 2)), 0.0)
    A = rng.normal(size=(40, 6))
    Bi = rng.poisson(2.0, (40, 23)).astype(np.int32)
    want = oracle.matMultT(A, Bi.astype(np.float64))
    assert np.array_equal(cb.eigenMatMultInt(A, Bi), want)
    assert np.array_equal(cb.eigenMatMultNumeric(A, Bi.astype(np.float64)), want)
    assert np.allclose(want, A.T @ Bi, rtol=1e-12, atol=1e-12)  # Eigen's blocked order is unspecified: 1e-12
    xg = rng.poisson(1.5, (30, 25)).astype(np.int32)
    grp = rng.integers(1, 5, 30).astype(np.int32)
    phi = np.asfortranarray(rng.dirichlet(np.ones(4), 25).T)
    psi = np.asfortranarray(rng.random((30, 4)) + 0.01)
    got = cb.perplexityG(xg, phi, psi, grp, 4)
    want = oracle.perplexityG_dense(xg, phi, psi, grp, 4)
    assert abs(got - want) <= 1e-12 * abs(want)
    with pytest.raises(cb.CeldaCudaError):
        cb.perplexityG(xg, phi[:3], psi, grp, 4)
