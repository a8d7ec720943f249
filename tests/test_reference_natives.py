"""The oracle against the REFERENCE'S OWN compiled C natives.  oracle/_ref/libcelda_ref_natives.so is
/root/reference/src/matrixSums.c + src/perplexity.c compiled in place (oracle/Makefile target `ref`, run by
__graft_entry__.build() where /root/reference exists) against the R C-API miniature of tests/r_stub; the built file
travels to the GPU box.  Every dense aggregation native (int and numeric, plain and *Change) and _perplexityG is driven
through `.Call`-style calls on random matrices, ragged shapes, wrapping integers and the error cases, and must agree
with the oracle's restatement bit for bit (values) and word for word (messages).  The CUDA entry points are compared
with the oracle in tests/test_gpu_natives.py, so this closes the chain reference C -> oracle -> device."""
import ctypes as C
import os

import numpy as np
import pytest

from test_r_shim import RStub

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.environ.get("CELDA_REF_DIR") or os.path.join(ROOT, "oracle", "_ref")  # a sanitizer build, see the last test
REF_SO = os.path.join(REF_DIR, "libcelda_ref_natives.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")


class RefNatives(RStub):
    """The reference natives are plain exported functions (celda registers them through Rcpp's table, which is not part
    of these two files): call them by address."""

    def __init__(self, so):
        L = self.L = C.CDLL(so)
        vp = C.c_void_p
        for name, res, args in [
            ("rstub_nil", vp, []), ("rstub_int", vp, [C.c_ssize_t, vp]), ("rstub_real", vp, [C.c_ssize_t, vp]),
            ("rstub_set_dim", None, [vp, C.c_int, C.c_int]), ("rstub_set_levels", None, [vp, C.c_int]),
            ("rstub_type", C.c_int, [vp]), ("rstub_length", C.c_longlong, [vp]), ("rstub_data", vp, [vp]),
            ("rstub_nrow", C.c_int, [vp]), ("rstub_ncol", C.c_int, [vp]), ("rstub_vector_elt", vp, [vp, C.c_longlong]),
            ("rstub_name", C.c_char_p, [vp, C.c_longlong]), ("rstub_protect_balance", C.c_int, []),
            ("rstub_call_ptr", vp, [vp, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]), ("rstub_reset", C.c_int, []),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    def call(self, name, *args, raw=False):
        fn = C.cast(getattr(self.L, name), C.c_void_p)
        arr = (C.c_void_p * max(len(args), 1))(*args)
        err = C.create_string_buffer(1024)
        out = self.L.rstub_call_ptr(fn, len(args), arr, err, len(err))
        if not out:
            raise RuntimeError(err.value.decode())
        return out if raw else self.value(out)


@pytest.fixture(scope="module")
def ref():
    r = RefNatives(REF_SO)
    yield r
    r.L.rstub_reset()


def _mat(rng, kind, nr, nc):
    if kind == "int":
        return rng.integers(0, 50, (nr, nc)).astype(np.int32)
    return rng.random((nr, nc)) * 10 - 3


@pytest.mark.parametrize("kind", ["int", "numeric"])
@pytest.mark.parametrize("nr,nc,nl", [(10, 10, 2), (37, 5, 7), (1, 9, 3), (64, 1, 4), (23, 41, 23)])
def test_dense_aggregation_natives_equal_the_oracle(ref, oracle, kind, nr, nc, nl):
    sfx = "" if kind == "int" else "_numeric"
    rng = np.random.default_rng(nr * 1000 + nc)
    x = _mat(rng, kind, nr, nc)
    gr, gr2 = rng.integers(1, nl + 1, nr), rng.integers(1, nl + 1, nr)
    gc, gc2 = rng.integers(1, nl + 1, nc), rng.integers(1, nl + 1, nc)
    got = ref.call("_rowSumByGroup" + sfx, ref.matrix(x), ref.factor(gr, nl))
    want = oracle.rowSumByGroup(x, gr, nl)
    assert got.dtype == want.dtype and np.array_equal(np.asarray(got).reshape(want.shape, order="F"), want)
    got = ref.call("_colSumByGroup" + sfx, ref.matrix(x), ref.factor(gc, nl))
    want_c = oracle.colSumByGroup(x, gc, nl)
    assert np.array_equal(np.asarray(got).reshape(want_c.shape, order="F"), want_c)
    # the incremental variants update px in place and return it (src/matrixSums.c:136-142)
    px = ref.matrix(want)
    out = ref.call("_rowSumByGroupChange" + sfx, ref.matrix(x), px, ref.factor(gr2, nl), ref.factor(gr, nl), raw=True)
    assert out == px
    po = oracle.rowSumByGroupChange(x, want.copy(order="F"), gr2, gr, nl)
    assert np.array_equal(np.asarray(ref.value(px)).reshape(po.shape, order="F"), po)
    px = ref.matrix(want_c)
    ref.call("_colSumByGroupChange" + sfx, ref.matrix(x), px, ref.factor(gc2, nl), ref.factor(gc, nl))
    po = oracle.colSumByGroupChange(x, want_c.copy(order="F"), gc2, gc, nl)
    assert np.array_equal(np.asarray(ref.value(px)).reshape(po.shape, order="F"), po)


def test_integer_sums_wrap_like_the_reference(ref, oracle):
    big = np.full((4, 3), 2 ** 30, dtype=np.int32)
    g = np.array([1, 1, 1, 1])
    got = ref.call("_rowSumByGroup", ref.matrix(big), ref.factor(g, 1))
    assert np.array_equal(np.asarray(got).ravel(), oracle.rowSumByGroup(big, g, 1).ravel())  # 4 * 2^30 wraps to 0


@pytest.mark.parametrize("kind", ["int", "numeric"])
def test_error_messages_equal_the_reference(ref, oracle, kind):
    sfx = "" if kind == "int" else "_numeric"
    rng = np.random.default_rng(1)
    x = _mat(rng, kind, 6, 4)
    gr, gc = np.array([1, 2, 1, 2, 1, 2]), np.array([1, 2, 2, 1])
    na = gr.copy()
    na[2] = np.iinfo(np.int32).min
    px_r, px_c = oracle.rowSumByGroup(x, gr, 2), oracle.colSumByGroup(x, gc, 2)
    cases = [
        ("_rowSumByGroup", (x, ("v", gr)), lambda: oracle.rowSumByGroup(x, gr, -1)),                   # not a factor
        ("_rowSumByGroup", (x, ("f", gr[:5], 2)), lambda: oracle.rowSumByGroup(x, gr[:5], 2)),         # length
        ("_colSumByGroup", (x, ("f", gc[:3], 2)), lambda: oracle.colSumByGroup(x, gc[:3], 2)),
        ("_rowSumByGroupChange", (x, px_r, ("f", gr, 2), ("f", gr, 3)),
         lambda: oracle.rowSumByGroupChange(x, px_r.copy(order="F"), gr, gr, 2, 3)),                   # levels differ
        ("_colSumByGroupChange", (x, px_r, ("f", gc, 2), ("f", gc, 2)),
         lambda: oracle.colSumByGroupChange(x, px_r.copy(order="F"), gc, gc, 2)),                      # px shape
        ("_rowSumByGroupChange", (x, px_r, ("f", na, 2), ("f", gr, 2)),
         lambda: oracle.rowSumByGroupChange(x, px_r.copy(order="F"), na, gr, 2)),                      # NA label
    ]
    if kind == "int":  # the numeric natives do not check NA in the plain sums (src/matrixSums.c:198-278)
        cases.append(("_rowSumByGroup", (x, ("f", na, 2)), lambda: oracle.rowSumByGroup(x, na, 2)))
    for name, args, ofn in cases:
        rargs = []
        for a in args:
            if isinstance(a, tuple):
                rargs.append(ref.integer(a[1]) if a[0] == "v" else ref.factor(a[1], a[2]))
            else:
                rargs.append(ref.matrix(a))
        with pytest.raises(RuntimeError) as e_ref:
            ref.call(name + sfx, *rargs)
        with pytest.raises(Exception) as e_or:
            ofn()
        assert str(e_ref.value) == str(e_or.value), name


def test_perplexityG_equals_the_oracle(ref, oracle):
    rng = np.random.default_rng(9)
    nr, nc, nl = 30, 12, 5
    x = rng.integers(0, 20, (nr, nc)).astype(np.int32)
    group = rng.integers(1, nl + 1, nr)
    phi = rng.random((nl, nc)) + 0.01
    psi = rng.random((nr, nl)) + 0.01
    got = ref.call("_perplexityG", ref.matrix(x), ref.matrix(phi), ref.matrix(psi), ref.factor(group, nl))[0]
    want = oracle.perplexityG_dense(x, phi, psi, group, nl)
    assert got == want or abs(got - want) <= 1e-13 * abs(want)  # the native calls libm's log, the oracle its own
    with pytest.raises(RuntimeError) as e_ref:
        ref.call("_perplexityG", ref.matrix(x), ref.matrix(phi[:, :5]), ref.matrix(psi), ref.factor(group, nl))
    with pytest.raises(Exception) as e_or:
        oracle.perplexityG_dense(x, phi[:, :5], psi, group, nl)
    assert str(e_ref.value) == str(e_or.value)


# ------------------------------------------------------------------ src/matrixSumsSparse.cpp (rows a2-a5, sparse)
SPARSE_SO = os.path.join(REF_DIR, "libcelda_ref_cpp.so")


@pytest.fixture(scope="module")
def refsp():
    if not os.path.exists(SPARSE_SO):
        pytest.skip("oracle/_ref/libcelda_ref_cpp.so not built")
    L = C.CDLL(SPARSE_SO)
    L.ref_last_error.restype = C.c_char_p
    return L


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _RefSparse:
    def __init__(self, L):
        self.L = L

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.ref_last_error().decode())

    def sum(self, which, O, group, n):
        nr, nc, p, i, x = O
        group = np.ascontiguousarray(group, np.int32)
        out = np.zeros((nr, n) if which == "col" else (n, nc), order="F")
        fn = getattr(self.L, f"ref_{which}SumByGroupSparse")
        self._chk(fn(nr, nc, _ip(p), _ip(i), _dp(x), _ip(group), C.c_long(group.size), n, _dp(out)))
        return out

    def change(self, which, O, px, group, pgroup, n):
        nr, nc, p, i, x = O
        group, pgroup = np.ascontiguousarray(group, np.int32), np.ascontiguousarray(pgroup, np.int32)
        fn = getattr(self.L, f"ref_{which}SumByGroupChangeSparse")
        self._chk(fn(nr, nc, _ip(p), _ip(i), _dp(x), _dp(px), px.shape[0], px.shape[1], _ip(group), C.c_long(group.size),
                     _ip(pgroup), C.c_long(pgroup.size), n))
        return px


@pytest.mark.parametrize("nr,nc,K,L,density", [(100, 425, 5, 10, 0.3), (57, 33, 4, 6, 0.05), (8, 300, 7, 8, 0.9),
                                                (40, 40, 1, 1, 0.2)])
def test_sparse_aggregation_natives_equal_the_oracle(refsp, oracle, nr, nc, K, L, density):
    """The reference's compiled colSumByGroupSparse / rowSumByGroupSparse / *ChangeSparse against the restatement:
    identical doubles on random sparse counts with empty columns and rows, px updated in place."""
    from oracle.oracle import dense_to_csc
    rng = np.random.default_rng(nr + nc)
    dense = rng.integers(1, 30, (nr, nc)) * (rng.random((nr, nc)) < density)
    dense[:, rng.integers(0, nc, 3)] = 0   # empty columns
    dense[rng.integers(0, nr, 2), :] = 0   # empty rows
    O = dense_to_csc(dense)
    ref = _RefSparse(refsp)
    z, z2 = rng.integers(1, K + 1, nc).astype(np.int32), rng.integers(1, K + 1, nc).astype(np.int32)
    y, y2 = rng.integers(1, L + 1, nr).astype(np.int32), rng.integers(1, L + 1, nr).astype(np.int32)
    cs, rs = ref.sum("col", O, z, K), ref.sum("row", O, y, L)
    assert np.array_equal(cs, oracle.colSumByGroupSparse(O, z, K))
    assert np.array_equal(rs, oracle.rowSumByGroupSparse(O, y, L))
    assert np.array_equal(cs, np.stack([dense[:, z == k].sum(1) for k in range(1, K + 1)], axis=1))
    a, b = cs.copy(order="F"), cs.copy(order="F")
    assert ref.change("col", O, a, z2, z, K) is a
    assert np.array_equal(a, oracle.colSumByGroupChangeSparse(O, b, z2, z, K))
    assert np.array_equal(a, ref.sum("col", O, z2, K))
    a, b = rs.copy(order="F"), rs.copy(order="F")
    ref.change("row", O, a, y2, y, L)
    assert np.array_equal(a, oracle.rowSumByGroupChangeSparse(O, b, y2, y, L))
    assert np.array_equal(a, ref.sum("row", O, y2, L))


def test_sparse_aggregation_error_messages_equal_the_reference(refsp, oracle):
    from oracle.oracle import dense_to_csc
    rng = np.random.default_rng(3)
    dense = rng.integers(0, 5, (12, 9))
    O = dense_to_csc(dense)
    ref = _RefSparse(refsp)
    z, y = rng.integers(1, 4, 9).astype(np.int32), rng.integers(1, 5, 12).astype(np.int32)
    px_c, px_r = oracle.colSumByGroupSparse(O, z, 3), oracle.rowSumByGroupSparse(O, y, 4)
    bad_z, bad_y = z.copy(), y.copy()
    bad_z[0], bad_y[0] = 4, 0
    cases = [
        (lambda o: o.colSumByGroupSparse(O, z[:8], 3), lambda: ref.sum("col", O, z[:8], 3)),           # length
        (lambda o: o.rowSumByGroupSparse(O, y[:11], 4), lambda: ref.sum("row", O, y[:11], 4)),
        (lambda o: o.colSumByGroupSparse(O, bad_z, 3), lambda: ref.sum("col", O, bad_z, 3)),           # label range
        (lambda o: o.rowSumByGroupSparse(O, bad_y, 4), lambda: ref.sum("row", O, bad_y, 4)),
        (lambda o: o.colSumByGroupSparse(O, np.ones(9, np.int32), 10), lambda: ref.sum("col", O, np.ones(9, np.int32), 10)),
        (lambda o: o.rowSumByGroupSparse(O, np.ones(12, np.int32), 13), lambda: ref.sum("row", O, np.ones(12, np.int32), 13)),
        (lambda o: o.colSumByGroupChangeSparse(O, px_c.copy(order="F"), z, z[:8], 3),
         lambda: ref.change("col", O, px_c.copy(order="F"), z, z[:8], 3)),                             # group vs pgroup
        (lambda o: o.colSumByGroupChangeSparse(O, px_c.copy(order="F"), z, bad_z, 3),
         lambda: ref.change("col", O, px_c.copy(order="F"), z, bad_z, 3)),                             # pgroup range
        (lambda o: o.rowSumByGroupChangeSparse(O, px_r.copy(order="F"), bad_y, y, 4),
         lambda: ref.change("row", O, px_r.copy(order="F"), bad_y, y, 4)),
        (lambda o: o.colSumByGroupChangeSparse(O, px_c[:11].copy(order="F"), z, z, 3),
         lambda: ref.change("col", O, px_c[:11].copy(order="F"), z, z, 3)),                            # px shape
        (lambda o: o.rowSumByGroupChangeSparse(O, px_r[:, :8].copy(order="F"), y, y, 4),
         lambda: ref.change("row", O, px_r[:, :8].copy(order="F"), y, y, 4)),
    ]
    for k, (ofn, rfn) in enumerate(cases):
        with pytest.raises(RuntimeError) as e_ref:
            rfn()
        with pytest.raises(Exception) as e_or:
            ofn(oracle)
        assert str(e_ref.value) == str(e_or.value), (k, str(e_ref.value), str(e_or.value))


# ------------------------------------------------------------------ src/cG_calcGibbsProbY.cpp (row a10: the y kernel)
def _y_state(oracle, gold_g, L, seed):
    from oracle.oracle import dense_to_csc
    counts = gold_g["counts"]
    nG, nM = counts.shape
    y = np.random.default_rng(seed).integers(1, L + 1, nG).astype(np.int32)
    p = oracle.cGDecomposeCounts(dense_to_csc(counts), y, L)
    return counts, y, p


def _tables(oracle, p, nG, L, beta, gamma, delta):
    """lgbeta / lggamma / lgdelta as .celda_G builds them (R/celda_G.R:380-383), by the oracle's lgamma."""
    top = int(p["nTSByC"].sum(0).max())
    lgb = oracle.lgamma(np.arange(0, top + 1) + beta)
    lgg = oracle.lgamma(np.arange(0, nG + L + 1) + gamma)
    lgd = np.concatenate([[np.nan], oracle.lgamma(np.arange(1, nG + L + 1) * delta)])
    return lgb, lgg, lgd


@pytest.mark.parametrize("L,hyper", [(10, (1.0, 1.0, 1.0)), (4, (0.5, 2.0, 1.5))])
def test_y_kernel_native_equals_the_oracle(refsp, oracle, gold_g, L, hyper):
    """The reference's compiled cG_CalcGibbsProbY against the restatement, gene by gene on a random celda_G state.  The
    native takes lgamma of the two nByTS terms per module from the C library while the oracle (and the device) evaluate
    R's lgammafn (which is what builds the tables in R): each entry is a difference of lgamma values near 1e5, so the
    vectors agree to a few units in the last place OF THOSE TERMS (4 ulp allowed, 2-3 measured), not of the result --
    which is why DESIGN.md section 3 fixes one arithmetic for oracle and device.  Same first maximum and the same
    sampled module for the same uniform."""
    beta, gamma, delta = hyper
    counts, y, p = _y_state(oracle, gold_g, L, 17)
    nG, nM = counts.shape
    lgb, lgg, lgd = _tables(oracle, p, nG, L, beta, gamma, delta)
    t, ts, gt, g = np.asfortranarray(p["nTSByC"], dtype=np.float64), p["nByTS"].astype(np.float64), \
        p["nGByTS"].astype(np.int32), p["nByG"].astype(np.float64)
    rng = np.random.default_rng(1)
    ulp_top = float(np.spacing(oracle.lgamma(np.array([ts.max() + g.max() + (nG + 1) * delta]))[0]))  # the largest term
    genes = rng.choice(nG, 40, replace=False) + 1
    for index in genes:
        row = np.ascontiguousarray(counts[index - 1], dtype=np.float64)
        got = np.zeros(L)
        rc = refsp.ref_cG_CalcGibbsProbY(int(index), _dp(row), nM, _dp(t), _dp(ts), _ip(gt), _dp(g), _ip(y), L, nG, _dp(lgb),
                                         C.c_long(lgb.size), _dp(lgg), C.c_long(lgg.size), _dp(lgd), C.c_long(lgd.size),
                                         C.c_double(delta), _dp(got))
        assert rc == 0, refsp.ref_last_error()
        want = oracle.cG_CalcGibbsProbY(int(index), row, t, ts, gt, g, y, L, beta, gamma, delta, lgb, lgg, lgd)
        both_nan = np.isnan(got) & np.isnan(want)  # lg_delta[0] is NA: a gene alone in its module (R aborts there)
        assert (np.isnan(got) == np.isnan(want)).all()
        assert np.abs(got[~both_nan] - want[~both_nan]).max(initial=0.0) <= 4 * ulp_top  # measured: 2-3 ulp of that term
        if not both_nan.any():
            assert int(np.argmax(got)) == int(np.argmax(want))
            u = float(rng.random())
            assert oracle.sampleLl(got, u)[0] == oracle.sampleLl(want, u)[0]


def test_reference_cross_check_variants_agree_with_its_production_kernel(refsp, oracle, gold_g):
    """The reference ships three more forms of the kernel 'as a sanity check when modifying the faster function'
    (src/cG_calcGibbsProbY.cpp:4-6).  Run here for the same purpose: _fastRow and _ori reproduce the production
    vector, _Simple (full expression per candidate) reproduces it up to its constant -- and so does the oracle."""
    L, (beta, gamma, delta) = 6, (1.0, 1.0, 1.0)
    counts, y, p = _y_state(oracle, gold_g, L, 23)
    nG, nM = counts.shape
    lgb, lgg, lgd = _tables(oracle, p, nG, L, beta, gamma, delta)
    ci = np.asfortranarray(counts, dtype=np.int32)
    ti, tsi, gti, gi = np.asfortranarray(p["nTSByC"], dtype=np.int32), p["nByTS"].astype(np.int32), \
        p["nGByTS"].astype(np.int32), p["nByG"].astype(np.int32)
    t, ts, g = ti.astype(np.float64, order="F"), tsi.astype(np.float64), gi.astype(np.float64)
    for index in (1, 7, 50, nG):
        row = np.ascontiguousarray(counts[index - 1], dtype=np.float64)
        want = oracle.cG_CalcGibbsProbY(index, row, t, ts, gti, g, y, L, beta, gamma, delta, lgb, lgg, lgd)
        for which in (0, 1):
            got = np.zeros(L)
            rc = refsp.ref_cG_CalcGibbsProbY_int(which, index, _ip(ci), nM, _ip(ti), _ip(tsi), _ip(gti), _ip(gi), _ip(y), L, nG,
                                                 _dp(lgb), C.c_long(lgb.size), _dp(lgg), C.c_long(lgg.size), _dp(lgd),
                                                 C.c_long(lgd.size), C.c_double(delta), _dp(got))
            assert rc == 0
            if which == 0:
                assert np.allclose(got, want, rtol=1e-12, atol=0)
            else:  # _ori sums every module's terms for every candidate: equal up to a constant
                assert np.allclose(got - got[0], want - want[0], rtol=0, atol=1e-7 * np.abs(got).max())
        simple = np.zeros(L)
        before = (ti.copy(), tsi.copy(), gti.copy())
        rc = refsp.ref_cG_calcGibbsProbY_Simple(_ip(ci), nM, _ip(gti), _ip(ti), _ip(tsi), _ip(gi), _ip(y), L, nG, index,
                                                C.c_double(gamma), C.c_double(beta), C.c_double(delta), _dp(simple))
        assert rc == 0
        assert all(np.array_equal(a, b) for a, b in zip(before, (ti, tsi, gti)))  # it restores the tables it edits
        assert np.allclose(simple - simple[0], want - want[0], rtol=0, atol=1e-7 * np.abs(simple).max())


# ------------------------------------------------------------------ the product's error contract, on the CPU
def test_library_error_messages_equal_the_reference_natives(ref, refsp, gold_cg):
    """libcelda_cuda validates its arguments on the host before the first CUDA call, so its error contract can be held
    against the reference's compiled natives without a device: for every precondition the reference checks (lengths,
    label range, K / L against the matrix, px shape, factor-ness, NA labels) the library returns the SAME message."""
    import celda_b200 as cb
    from oracle.oracle import dense_to_csc
    rng = np.random.default_rng(3)
    dense = rng.integers(0, 5, (12, 9))
    O = dense_to_csc(dense)
    A = cb.CSC.from_dense(dense)
    rs = _RefSparse(refsp)
    z, y = rng.integers(1, 4, 9).astype(np.int32), rng.integers(1, 5, 12).astype(np.int32)
    bad_z, bad_y = z.copy(), y.copy()
    bad_z[0], bad_y[0] = 4, 0
    pc, pr = np.zeros((12, 3), order="F"), np.zeros((4, 9), order="F")
    sparse_cases = [
        (lambda: cb.colSumByGroupSparse(A, z[:8], 3), lambda: rs.sum("col", O, z[:8], 3)),
        (lambda: cb.rowSumByGroupSparse(A, y[:11], 4), lambda: rs.sum("row", O, y[:11], 4)),
        (lambda: cb.colSumByGroupSparse(A, bad_z, 3), lambda: rs.sum("col", O, bad_z, 3)),
        (lambda: cb.rowSumByGroupSparse(A, bad_y, 4), lambda: rs.sum("row", O, bad_y, 4)),
        (lambda: cb.colSumByGroupSparse(A, np.ones(9, np.int32), 10), lambda: rs.sum("col", O, np.ones(9, np.int32), 10)),
        (lambda: cb.rowSumByGroupSparse(A, np.ones(12, np.int32), 13), lambda: rs.sum("row", O, np.ones(12, np.int32), 13)),
        (lambda: cb.colSumByGroupChangeSparse(A, pc.copy(order="F"), z, z[:8], 3),
         lambda: rs.change("col", O, pc.copy(order="F"), z, z[:8], 3)),
        (lambda: cb.colSumByGroupChangeSparse(A, pc.copy(order="F"), z, bad_z, 3),
         lambda: rs.change("col", O, pc.copy(order="F"), z, bad_z, 3)),
        (lambda: cb.rowSumByGroupChangeSparse(A, pr.copy(order="F"), bad_y, y, 4),
         lambda: rs.change("row", O, pr.copy(order="F"), bad_y, y, 4)),
        (lambda: cb.colSumByGroupChangeSparse(A, pc[:11].copy(order="F"), z, z, 3),
         lambda: rs.change("col", O, pc[:11].copy(order="F"), z, z, 3)),
        (lambda: cb.rowSumByGroupChangeSparse(A, pr[:, :8].copy(order="F"), y, y, 4),
         lambda: rs.change("row", O, pr[:, :8].copy(order="F"), y, y, 4)),
    ]
    for k, (ours, theirs) in enumerate(sparse_cases):
        with pytest.raises(cb.CeldaCudaError) as e_ours:
            ours()
        with pytest.raises(RuntimeError) as e_ref:
            theirs()
        assert str(e_ours.value) == str(e_ref.value), (k, str(e_ours.value), str(e_ref.value))

    # dense natives (src/matrixSums.c): factor-ness, lengths, levels, px shape, NA
    for dtype, sfx in ((np.int32, ""), (np.float64, "_numeric")):
        x = np.asfortranarray(rng.integers(0, 9, (6, 4)).astype(dtype))
        gr, gc = np.array([1, 2, 1, 2, 1, 2], np.int32), np.array([1, 2, 2, 1], np.int32)
        na = gr.copy()
        na[2] = np.iinfo(np.int32).min
        px_r, px_c = np.zeros((2, 4), dtype, order="F"), np.zeros((6, 2), dtype, order="F")
        dense_cases = [
            (lambda: cb.rowSumByGroup_dense(x, gr, -1), ("_rowSumByGroup", [x, ("v", gr)])),
            (lambda: cb.colSumByGroup_dense(x, gc, -1), ("_colSumByGroup", [x, ("v", gc)])),
            (lambda: cb.rowSumByGroup_dense(x, gr[:5], 2), ("_rowSumByGroup", [x, ("f", gr[:5], 2)])),
            (lambda: cb.colSumByGroup_dense(x, gc[:3], 2), ("_colSumByGroup", [x, ("f", gc[:3], 2)])),
            (lambda: cb.rowSumByGroupChange_dense(x, px_r.copy(order="F"), gr, gr, 2, 3),
             ("_rowSumByGroupChange", [x, px_r, ("f", gr, 2), ("f", gr, 3)])),
            (lambda: cb.colSumByGroupChange_dense(x, px_r.copy(order="F"), gc, gc, 2),
             ("_colSumByGroupChange", [x, px_r, ("f", gc, 2), ("f", gc, 2)])),
            (lambda: cb.rowSumByGroupChange_dense(x, px_c.copy(order="F"), gr, gr, 2),
             ("_rowSumByGroupChange", [x, px_c, ("f", gr, 2), ("f", gr, 2)])),
            (lambda: cb.rowSumByGroupChange_dense(x, px_r.copy(order="F"), gr[:5], gr, 2),
             ("_rowSumByGroupChange", [x, px_r, ("f", gr[:5], 2), ("f", gr, 2)])),
            (lambda: cb.rowSumByGroupChange_dense(x, px_r.copy(order="F"), na, gr, 2),
             ("_rowSumByGroupChange", [x, px_r, ("f", na, 2), ("f", gr, 2)])),
            (lambda: cb.colSumByGroupChange_dense(x, px_c.copy(order="F"), gc, gc[:3], 2),
             ("_colSumByGroupChange", [x, px_c, ("f", gc, 2), ("f", gc[:3], 2)])),
        ]
        if dtype == np.int32:
            dense_cases.append((lambda: cb.rowSumByGroup_dense(x, na, 2), ("_rowSumByGroup", [x, ("f", na, 2)])))
        for k, (ours, (name, args)) in enumerate(dense_cases):
            rargs = [(ref.integer(a[1]) if a[0] == "v" else ref.factor(a[1], a[2])) if isinstance(a, tuple) else ref.matrix(a)
                     for a in args]
            with pytest.raises(cb.CeldaCudaError) as e_ours:
                ours()
            with pytest.raises(RuntimeError) as e_ref:
                ref.call(name + sfx, *rargs)
            assert str(e_ours.value) == str(e_ref.value), (sfx, k, name, str(e_ours.value), str(e_ref.value))

    # _perplexityG (src/perplexity.c:13-43).  The reference's label check is `< 0 || > nrow(x)` and would index out of
    # bounds for 0 or for a label above the number of levels; the library rejects those too (same message).
    nr, nc, nl = 8, 5, 3
    x = rng.integers(0, 9, (nr, nc)).astype(np.int32)
    phi, psi, grp = rng.random((nl, nc)) + 0.1, rng.random((nr, nl)) + 0.1, rng.integers(1, nl + 1, nr).astype(np.int32)
    na = grp.copy()
    na[1] = np.iinfo(np.int32).min
    for k, (xx, ph, ps, g, lev) in enumerate([(x, phi, psi, grp, -1), (x, phi, psi, grp[:7], nl), (x, phi[:, :4], psi, grp, nl),
                                              (x, phi[:2], psi, grp, nl), (x, phi, psi[:7], grp, nl),
                                              (x, phi, psi[:, :2], grp, nl), (x, phi, psi, na, nl)]):
        with pytest.raises(cb.CeldaCudaError) as e_ours:
            cb.perplexityG(xx, ph, ps, g, lev)
        with pytest.raises(RuntimeError) as e_ref:
            ref.call("_perplexityG", ref.matrix(xx), ref.matrix(np.asfortranarray(ph)), ref.matrix(np.asfortranarray(ps)),
                     ref.integer(g) if lev < 0 else ref.factor(g, lev))
        assert str(e_ours.value) == str(e_ref.value), (k, str(e_ours.value), str(e_ref.value))


def test_reference_natives_and_miniatures_clean_under_asan_and_ubsan(tmp_path):
    """This file once more in a subprocess against a second build of oracle/_ref with AddressSanitizer +
    UndefinedBehaviorSanitizer (`make -C oracle ref REFDIR=... SAN=...`, libasan preloaded): the R / Rcpp / Eigen
    miniatures hand the reference's code exactly the memory it indexes, and the reference's natives themselves run clean
    on these inputs."""
    import subprocess
    import sys
    if os.environ.get("CELDA_REF_DIR"):
        pytest.skip("already inside the sanitizer run")
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs the reference sources to build the instrumented copy")
    try:
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    except Exception:
        asan = ""
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("gcc's libasan is not available")
    san = "-g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer"
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "CC=gcc", "CXX=g++", f"REFDIR={tmp_path}",
                           f"SAN={san}"])
    # libstdc++ is preloaded next to libasan: python itself does not link it, and ASan resolves __cxa_throw when it starts
    # (the reference's stop() is a C++ exception thrown inside a dlopen'ed library)
    stdcpp = subprocess.run(["g++", "-print-file-name=libstdc++.so.6"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=f"{asan} {stdcpp}".strip(), CELDA_REF_DIR=str(tmp_path),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    # the wrap test is left out: there UBSan stops at the reference's own `ans[...] += x[...]` on int
    # (src/matrixSums.c:42, signed overflow is undefined in C; it wraps on every build we know, which is what the
    # oracle and the device reproduce) -- the one finding, and it is the reference's
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "not gpu",
                        "-k", "not wrap", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
