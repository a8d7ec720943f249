"""The device against the REFERENCE'S OWN compiled natives (oracle/_ref, built from /root/reference/src by
`make -C oracle ref` in the build container; the built files travel to the GPU box): sparse and dense aggregation bit
for bit, the single-gene y kernel within the rounding the reference itself has across platforms (its two libm lgamma
terms per module), same first maximum.  tests/test_reference_natives.py holds the CPU oracle against the same natives;
this file closes the triangle on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc
from test_reference_natives import REF_SO, SPARSE_SO, RefNatives, _RefSparse, _dp, _ip, _tables, _y_state

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF_SO) and os.path.exists(SPARSE_SO)), reason="oracle/_ref not built")]


def test_device_sparse_aggregation_equals_the_reference_natives(gold_cg):
    counts = gold_cg["counts"]
    nG, nM = counts.shape
    rng = np.random.default_rng(8)
    z, z2 = rng.integers(1, 6, nM).astype(np.int32), rng.integers(1, 6, nM).astype(np.int32)
    y, y2 = rng.integers(1, 11, nG).astype(np.int32), rng.integers(1, 11, nG).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    ref = _RefSparse(C.CDLL(SPARSE_SO))
    ref.L.ref_last_error.restype = C.c_char_p
    cs, rs = ref.sum("col", O, z, 5), ref.sum("row", O, y, 10)
    assert np.array_equal(cb.colSumByGroupSparse(A, z, 5), cs)
    assert np.array_equal(cb.rowSumByGroupSparse(A, y, 10), rs)
    a, b = cs.copy(order="F"), cs.copy(order="F")
    ref.change("col", O, a, z2, z, 5)
    assert np.array_equal(cb.colSumByGroupChangeSparse(A, b, z2, z, 5), a)
    a, b = rs.copy(order="F"), rs.copy(order="F")
    ref.change("row", O, a, y2, y, 10)
    assert np.array_equal(cb.rowSumByGroupChangeSparse(A, b, y2, y, 10), a)


@pytest.mark.parametrize("kind", ["int", "numeric"])
def test_device_dense_aggregation_equals_the_reference_natives(kind):
    ref = RefNatives(REF_SO)
    sfx = "" if kind == "int" else "_numeric"
    rng = np.random.default_rng(4)
    x = rng.integers(0, 50, (37, 23)).astype(np.int32) if kind == "int" else rng.random((37, 23)) * 10 - 3
    gr, gr2 = rng.integers(1, 8, 37).astype(np.int32), rng.integers(1, 8, 37).astype(np.int32)
    gc = rng.integers(1, 8, 23).astype(np.int32)
    want_r = np.asarray(ref.call("_rowSumByGroup" + sfx, ref.matrix(x), ref.factor(gr, 7))).reshape((7, 23), order="F")
    want_c = np.asarray(ref.call("_colSumByGroup" + sfx, ref.matrix(x), ref.factor(gc, 7))).reshape((37, 7), order="F")
    assert np.array_equal(cb.rowSumByGroup_dense(x, gr, 7), want_r)
    assert np.array_equal(cb.colSumByGroup_dense(x, gc, 7), want_c)
    px = ref.matrix(want_r)
    ref.call("_rowSumByGroupChange" + sfx, ref.matrix(x), px, ref.factor(gr2, 7), ref.factor(gr, 7))
    mine = np.asfortranarray(want_r.copy())
    cb.rowSumByGroupChange_dense(x, mine, gr2, gr, 7)
    assert np.array_equal(mine, np.asarray(ref.value(px)).reshape((7, 23), order="F"))
    ref.L.rstub_reset()


def test_device_y_kernel_against_the_reference_native(oracle, gold_g):
    L, (beta, gamma, delta) = 10, (1.0, 1.0, 1.0)
    counts, y, p = _y_state(oracle, gold_g, L, 17)
    nG, nM = counts.shape
    lgb, lgg, lgd = _tables(oracle, p, nG, L, beta, gamma, delta)
    t, ts, gt, g = np.asfortranarray(p["nTSByC"], dtype=np.float64), p["nByTS"].astype(np.float64), \
        p["nGByTS"].astype(np.int32), p["nByG"].astype(np.float64)
    lib = C.CDLL(SPARSE_SO)
    ulp_top = float(np.spacing(oracle.lgamma(np.array([ts.max() + g.max() + (nG + 1) * delta]))[0]))
    for index in (1, 2, 17, 50, 99, nG):
        row = np.ascontiguousarray(counts[index - 1], dtype=np.float64)
        want = np.zeros(L)
        rc = lib.ref_cG_CalcGibbsProbY(index, _dp(row), nM, _dp(t), _dp(ts), _ip(gt), _dp(g), _ip(y), L, nG, _dp(lgb),
                                       C.c_long(lgb.size), _dp(lgg), C.c_long(lgg.size), _dp(lgd), C.c_long(lgd.size),
                                       C.c_double(delta), _dp(want))
        assert rc == 0
        got = cb.cG_CalcGibbsProbY(index, row, t, ts, gt, g, y, L, nG, beta, gamma, delta)
        ok = ~(np.isnan(got) & np.isnan(want))
        assert (np.isnan(got) == np.isnan(want)).all()
        assert np.abs(got[ok] - want[ok]).max(initial=0.0) <= 4 * ulp_top
        if ok.all():
            assert int(np.argmax(got)) == int(np.argmax(want))
