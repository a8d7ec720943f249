"""r/celda_cuda_shim.c -- the .Call binding INTEGRATION.md asks a celda maintainer to add -- compiled and driven
without an R installation: tests/r_stub/ is a miniature of the R C API (vectors with dim / levels / S4 slots,
external pointers with finalizers, Rf_error as a non-local exit, the R_registerRoutines table and an arity-checking
.Call).  The CPU tests cover what needs no device: the shim compiles against include/celda_cuda.h, links against
libcelda_cuda.so, registers every routine with the arity its definition has, and turns the library's status +
celda_cuda_last_error() into an R error.  The gpu tests run the testthat vectors of the reference
(tests/testthat/test-matrixSums.R:5-76) and one celda_CG chain iteration through `.Call`, against the CPU oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "r", "celda_cuda_shim.c")
STUB = os.path.join(ROOT, "tests", "r_stub")

INTSXP, REALSXP, VECSXP, EXTPTRSXP = 13, 14, 19, 22


class RStub:
    """`.Call` over the stub runtime; numpy arrays in, numpy arrays out (column-major like R)."""

    def __init__(self, so):
        L = self.L = C.CDLL(so)
        vp = C.c_void_p
        for name, res, args in [
            ("rstub_nil", vp, []), ("rstub_int", vp, [C.c_ssize_t, vp]), ("rstub_real", vp, [C.c_ssize_t, vp]),
            ("rstub_set_dim", None, [vp, C.c_int, C.c_int]), ("rstub_set_levels", None, [vp, C.c_int]),
            ("rstub_s4", vp, []), ("rstub_set_slot", None, [vp, C.c_char_p, vp]), ("rstub_type", C.c_int, [vp]),
            ("rstub_length", C.c_longlong, [vp]), ("rstub_data", vp, [vp]), ("rstub_nrow", C.c_int, [vp]),
            ("rstub_ncol", C.c_int, [vp]), ("rstub_vector_elt", vp, [vp, C.c_longlong]),
            ("rstub_protect_balance", C.c_int, []), ("rstub_n_routines", C.c_int, []),
            ("rstub_routine_name", C.c_char_p, [C.c_int]), ("rstub_routine_nargs", C.c_int, [C.c_int]),
            ("rstub_call", vp, [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]), ("rstub_reset", C.c_int, []),
            ("celda_cuda_register", None, [vp]), ("rstub_logical", vp, [C.c_int]),
            ("rstub_name", C.c_char_p, [vp, C.c_longlong]),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L.celda_cuda_register(None)

    # ---- R objects
    def integer(self, v, dim=None, levels=None):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.int32).ravel(order="F"))
        s = self.L.rstub_int(a.size, a.ctypes.data)
        if dim is not None:
            self.L.rstub_set_dim(s, *dim)
        if levels is not None:
            self.L.rstub_set_levels(s, levels)
        return s

    def numeric(self, v, dim=None):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel(order="F"))
        s = self.L.rstub_real(a.size, a.ctypes.data)
        if dim is not None:
            self.L.rstub_set_dim(s, *dim)
        return s

    def matrix(self, m):
        m = np.asarray(m)
        return (self.integer if m.dtype.kind == "i" else self.numeric)(m, dim=m.shape)

    def factor(self, codes, nlevels):
        return self.integer(codes, levels=nlevels)

    def dgCMatrix(self, dense):
        """The Matrix package's column-compressed S4 object: slots Dim, p, i (0-based), x (double)."""
        dense = np.asarray(dense)
        p, i, x = [0], [], []
        for c in range(dense.shape[1]):
            nz = np.nonzero(dense[:, c])[0]
            i += nz.tolist()
            x += dense[nz, c].tolist()
            p.append(len(i))
        s = self.L.rstub_s4()
        for name, v in (("Dim", self.integer(dense.shape)), ("p", self.integer(p)), ("i", self.integer(i)),
                        ("x", self.numeric(x))):
            self.L.rstub_set_slot(s, name.encode(), v)
        return s

    def logical(self, v):
        return self.L.rstub_logical(int(bool(v)))

    @property
    def NULL(self):
        return self.L.rstub_nil()

    def value(self, s):
        t, n = self.L.rstub_type(s), self.L.rstub_length(s)
        if t == VECSXP:
            vals = [self.value(self.L.rstub_vector_elt(s, k)) for k in range(n)]
            names = [self.L.rstub_name(s, k).decode() for k in range(n)]
            return dict(zip(names, vals)) if all(names) else vals
        if t not in (INTSXP, REALSXP):
            return s  # NULL, external pointer
        ct = C.c_int if t == INTSXP else C.c_double
        a = np.ctypeslib.as_array(C.cast(self.L.rstub_data(s), C.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
        nr, nc = self.L.rstub_nrow(s), self.L.rstub_ncol(s)
        return a.reshape((nr, nc), order="F") if nr * nc == n and nc != 1 else a

    # ---- .Call
    def call(self, name, *args, raw=False):
        arr = (C.c_void_p * max(len(args), 1))(*args)
        err = C.create_string_buffer(1024)
        out = self.L.rstub_call(name.encode(), len(args), arr, err, len(err))
        if not out:
            raise RuntimeError(err.value.decode())
        assert self.L.rstub_protect_balance() == 0, f"{name}: PROTECT / UNPROTECT imbalance"
        return out if raw else self.value(out)

    def routines(self):
        return {self.L.rstub_routine_name(k).decode(): self.L.rstub_routine_nargs(k)
                for k in range(self.L.rstub_n_routines())}


def _build_shim(tmpdir, mock=False):
    """The shim + the R API miniature as one shared object on top of libcelda_cuda.so; with mock=True the closure-level
    entry points resolve to tests/r_stub/mock_celda_cuda.c (the CPU oracle) linked in front of the product library."""
    from celda_b200 import _lib
    so_dir = os.path.dirname(_lib.build())
    warn = ["-O1", "-g", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-shared", "-fPIC"]
    if os.environ.get("CELDA_SHIM_SANITIZE") == "1":  # test_shim_and_stub_clean_under_asan_and_ubsan re-runs this file so
        warn += ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
    libs = ["-L" + so_dir, "-l:libcelda_cuda.so", "-Wl,-rpath," + so_dir]
    if mock:
        odir = os.path.join(ROOT, "oracle")
        subprocess.check_call(["make", "-s", "-C", odir, "all"])
        mock_so = os.path.join(tmpdir, "libmock_celda_cuda.so")
        r = subprocess.run(["gcc"] + warn + ["-I" + os.path.join(ROOT, "include"), "-o", mock_so,
                            os.path.join(STUB, "mock_celda_cuda.c"), "-L" + odir, "-l:libcelda_oracle.so",
                            "-Wl,-rpath," + odir], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        libs = ["-L" + tmpdir, "-l:libmock_celda_cuda.so", "-Wl,-rpath," + tmpdir] + libs
    out = os.path.join(tmpdir, "celda_shim_test.so")
    r = subprocess.run(["gcc"] + warn + ["-I" + os.path.join(STUB, "include"), "-I" + os.path.join(ROOT, "include"), "-o", out,
                        SHIM, os.path.join(STUB, "rstub.c")] + libs, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return RStub(out)


@pytest.fixture(scope="module")
def R(tmp_path_factory):
    stub = _build_shim(str(tmp_path_factory.mktemp("rshim")))
    yield stub
    stub.L.rstub_reset()


@pytest.fixture(scope="module")
def Rmock(tmp_path_factory):
    stub = _build_shim(str(tmp_path_factory.mktemp("rshim_mock")), mock=True)
    stub.L.mock_celda_cuda_calls.restype = C.c_int
    yield stub
    stub.L.rstub_reset()


def _shim_definitions():
    src = re.sub(r"/\*.*?\*/", "", open(SHIM).read(), flags=re.S)
    return {m.group(1): len([a for a in m.group(2).split(",") if a.strip() not in ("", "void")])
            for m in re.finditer(r"^SEXP\s+(_\w+)\s*\(([^)]*)\)\s*\{", src, flags=re.M)}


def test_shim_compiles_and_registers_every_routine(R):
    """-Wall -Wextra -Werror against include/celda_cuda.h, and the R_CallMethodDef table agrees with the definitions."""
    defs, table = _shim_definitions(), R.routines()
    assert len(defs) >= 44
    assert table == defs  # same names, and numArgs equals each definition's parameter count
    # the dense natives keep the .Call names of the reference (NAMESPACE useDynLib / src/RcppExports.cpp:308-337)
    for name in ("_rowSumByGroup", "_colSumByGroup", "_rowSumByGroupChange", "_colSumByGroupChange", "_perplexityG",
                 "_rowSumByGroup_numeric", "_colSumByGroup_numeric", "_rowSumByGroupChange_numeric",
                 "_colSumByGroupChange_numeric"):
        assert name in table


def test_registered_names_and_arities_match_the_reference_table(R):
    """Against the reference's own registration table (src/RcppExports.cpp:308-337): the nine plain-C natives keep their
    `.Call` names and arities, the Rcpp exports on the path map `_celda_<name>` -> `_celda_cuda_<name>` with the same
    arity (cG_CalcGibbsProbY takes beta / gamma where the reference takes the three lgamma tables: 12 for 13)."""
    path = "/root/reference/src/RcppExports.cpp"
    if not os.path.exists(path):
        pytest.skip("the reference is not present on this box")
    ref = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(_\w+)",\s*\(DL_FUNC\)\s*&\w+,\s*(\d+)\}', open(path).read())}
    mine = R.routines()
    plain = [n for n in ref if not n.startswith("_celda_")]
    assert len(plain) == 9
    for n in plain:
        assert mine[n] == ref[n], n
    for n in ("colSumByGroupSparse", "rowSumByGroupSparse", "colSumByGroupChangeSparse", "rowSumByGroupChangeSparse",
              "eigenMatMultInt", "eigenMatMultNumeric", "fastNormPropLog"):
        assert mine["_celda_cuda_" + n] == ref["_celda_" + n], n
    assert mine["_celda_cuda_cG_CalcGibbsProbY"] == ref["_celda_cG_CalcGibbsProbY"] - 1


def test_shim_binds_only_declared_entry_points():
    from celda_b200 import _lib
    used = set(re.findall(r"\b(celda_\w+)\s*\(", open(SHIM).read())) - {"celda_cuda_register"}
    decl = set(_lib.declared_functions())
    assert used and used <= decl, used - decl


def test_call_checks_arity_and_name(R):
    with pytest.raises(RuntimeError, match="Incorrect number of arguments"):
        R.call("_rowSumByGroup", R.NULL)
    with pytest.raises(RuntimeError, match="not available for .Call"):
        R.call("_celda_cuda_no_such_routine")


def test_library_status_becomes_an_R_error_without_a_device(R):
    """Argument validation precedes any CUDA call, so the error path of CHECK() runs on a CPU-only box:
    the messages are the reference's (src/matrixSums.c:13-22, 151-172)."""
    mat = np.asfortranarray(np.tile(np.arange(1, 6), 20).reshape(10, 10, order="F").astype(np.int32))
    lab = np.repeat([1, 2], 5)
    with pytest.raises(RuntimeError, match="must be a factor"):
        R.call("_rowSumByGroup", R.matrix(mat), R.integer(lab))  # an integer vector, not a factor
    with pytest.raises(RuntimeError, match="must match the number of rows"):
        R.call("_rowSumByGroup", R.matrix(mat), R.factor(lab[:9], 2))
    with pytest.raises(RuntimeError, match="must match the number of columns"):
        R.call("_colSumByGroup_numeric", R.matrix(mat.astype(np.float64)), R.factor(lab[:9], 2))
    with pytest.raises(RuntimeError, match="must not be NA"):
        na = lab.copy()
        na[3] = np.iinfo(np.int32).min
        R.call("_rowSumByGroup", R.matrix(mat), R.factor(na, 2))
    with pytest.raises(RuntimeError, match='no slot of name "Dim"'):
        R.call("_celda_cuda_colSumByGroupSparse", R.matrix(mat), R.integer(lab), R.integer([2]))
    with pytest.raises(RuntimeError, match="wrong SEXP type"):
        R.call("_celda_cuda_chain_loglik", R.integer([1]))  # not an external pointer


# ---------------------------------------------------------- closures through .Call, answered by the CPU oracle (mock)
def _same(got, want, names):
    assert list(got) == names  # the reference's list, in the reference's order
    for k in names:
        assert np.array_equal(got[k], want[k]), k


def test_gibbs_closures_through_dotCall_marshal_like_the_reference(Rmock, oracle, gold_cg):
    """.cCCalcGibbsProbZ / .cGCalcGibbsProbY as celda_CG calls them (R/celda_CG.R:566-585, 544-563): same arguments, same
    named list, arguments left untouched.  The library side is the mock (CPU oracle), so this pins the shim only."""
    n0 = Rmock.L.mock_celda_cuda_calls()
    gibbs_closures_roundtrip(Rmock, oracle, gold_cg)
    assert Rmock.L.mock_celda_cuda_calls() == n0 + 4  # the closure entry points resolved to the mock, not the GPU library


def gibbs_closures_roundtrip(Rmock, oracle, gold_cg):
    """Both Gibbs closures through `.Call` against the oracle called directly (shared with the GPU run of
    tests/test_zz_gpu_r_shim_closures.py, where the library behind the shim is the real one)."""
    from oracle.oracle import dense_to_csc
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(11)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    p = oracle.cCGDecomposeCounts(dense_to_csc(counts), s, z, y, K, L)
    ix, u = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    for doSample in (True, False):
        want = oracle.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0,
                                       0.7, ix, u, doSample=doSample)
        args = [Rmock.matrix(p["nTSByC"]), Rmock.matrix(p["mCPByS"]), Rmock.matrix(p["nTSByCP"]), Rmock.numeric(p["nByC"]),
                Rmock.numeric(p["nCP"]), Rmock.integer(z), Rmock.integer(s), Rmock.integer([K]), Rmock.integer([L]),
                Rmock.integer([nM]), Rmock.numeric([1.0]), Rmock.numeric([0.7]), Rmock.logical(doSample),
                Rmock.integer(ix), Rmock.numeric(u)]
        got = Rmock.call("_celda_cuda_cCCalcGibbsProbZ", *args)
        _same(got, want, ["mCPByS", "nGByCP", "nCP", "z", "probs"])
        assert got["probs"].shape == (K, nM) and got["mCPByS"].dtype == np.int32
        assert (got["z"] != z).any() == doSample
        # R semantics: the caller's objects are not modified
        assert np.array_equal(Rmock.value(args[1]), p["mCPByS"]) and np.array_equal(Rmock.value(args[5]), z)
        assert np.array_equal(Rmock.value(args[2]), p["nTSByCP"]) and np.array_equal(Rmock.value(args[4]), p["nCP"])
    # integer-typed tables (colSumByGroup of an integer matrix) are accepted and coerced
    got = Rmock.call("_celda_cuda_cCCalcGibbsProbZ", args[0], args[1], Rmock.matrix(p["nTSByCP"].astype(np.int32)),
                     Rmock.integer(p["nByC"]), Rmock.integer(p["nCP"]), *args[5:12], Rmock.logical(True), *args[13:])
    want = oracle.cCCalcGibbsProbZ(p["nTSByC"], p["mCPByS"], p["nTSByCP"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 0.7, ix, u)
    _same(got, want, ["mCPByS", "nGByCP", "nCP", "z", "probs"])

    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    want = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 0.9, 1.1, 1.3,
                                   ixy, uy)
    args = [Rmock.matrix(p["nGByCP"]), Rmock.matrix(p["nTSByCP"]), Rmock.numeric(p["nByTS"]), Rmock.numeric(p["nGByTS"]),
            Rmock.numeric(p["nByG"]), Rmock.integer(y), Rmock.integer([L]), Rmock.integer([nG]), Rmock.numeric([0.9]),
            Rmock.numeric([1.1]), Rmock.numeric([1.3]), Rmock.logical(True), Rmock.integer(ixy), Rmock.numeric(uy)]
    got = Rmock.call("_celda_cuda_cGCalcGibbsProbY", *args)
    _same(got, want, ["nTSByC", "nGByTS", "nByTS", "y", "probs"])
    assert got["probs"].shape == (L, nG) and got["nGByTS"].dtype == np.int32 and (got["y"] != y).any()
    assert np.array_equal(Rmock.value(args[1]), p["nTSByCP"]) and np.array_equal(Rmock.value(args[5]), y)


def test_em_closure_and_loglik_shims_marshal_like_the_reference(Rmock, oracle, gold_c, gold_g):
    from oracle.oracle import dense_to_csc
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    nG, nM = counts.shape
    z = np.random.default_rng(4).integers(1, K + 1, nM).astype(np.int32)
    O = dense_to_csc(counts)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    tail = [Rmock.integer(z), Rmock.integer(s), Rmock.integer([K]), Rmock.integer([nG]), Rmock.integer([nM]),
            Rmock.numeric([1.0]), Rmock.numeric([1.0]), Rmock.logical(True)]
    head = [Rmock.matrix(p["mCPByS"]), Rmock.matrix(p["nGByCP"]), Rmock.numeric(p["nByC"]), Rmock.numeric(p["nCP"])]
    want = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, nG, nM, 1.0, 1.0)
    for cnt in (Rmock.dgCMatrix(counts), Rmock.matrix(counts.astype(np.float64)), Rmock.matrix(counts.astype(np.int32))):
        got = Rmock.call("_celda_cuda_cCCalcEMProbZ", cnt, *head, *tail)
        assert list(got) == ["mCPByS", "nGByCP", "nCP", "z", "probs"]
        for k in ("mCPByS", "nGByCP", "nCP", "z"):
            assert np.array_equal(got[k], want[k]), k
        assert np.allclose(got["probs"], want["probs"], rtol=1e-12)  # sparse and dense products add in different orders
    assert np.array_equal(Rmock.value(tail[0]), z)
    ll = Rmock.call("_celda_cuda_cCCalcLL", head[0], head[1], tail[1], tail[0], Rmock.integer([K]),
                    Rmock.integer([p["nS"]]), Rmock.integer([nG]), Rmock.numeric([1.0]), Rmock.numeric([1.0]))[0]
    assert ll == oracle.cCCalcLL(p["mCPByS"], p["nGByCP"], K, p["nS"], nG, 1.0, 1.0)

    counts, L = gold_g["counts"], 10
    nG, nM = counts.shape
    y = np.random.default_rng(5).integers(1, L + 1, nG).astype(np.int32)
    q = oracle.cGDecomposeCounts(dense_to_csc(counts), y, L)
    ll = Rmock.call("_celda_cuda_cGCalcLL", Rmock.matrix(q["nTSByC"]), Rmock.numeric(q["nByTS"]), Rmock.numeric(q["nByG"]),
                    Rmock.numeric(q["nGByTS"]), Rmock.integer([nM]), Rmock.integer([nG]), Rmock.integer([L]),
                    Rmock.numeric([1.0]), Rmock.numeric([1.0]), Rmock.numeric([1.0]))[0]
    assert ll == oracle.cGCalcLL(q["nTSByC"], q["nByTS"], q["nByG"], q["nGByTS"], nM, nG, L, 1.0, 1.0, 1.0)


def test_closure_errors_surface_as_R_errors(Rmock, oracle, gold_cg):
    """A status from the library (here: the oracle's own argument check) becomes an R error after the PROTECT stack is
    balanced."""
    counts, s = gold_cg["counts"], gold_cg["s"]
    nG, nM = counts.shape
    bad = [Rmock.matrix(np.zeros((3, nM))), Rmock.matrix(np.zeros((2, int(s.max())), np.int32)), Rmock.matrix(np.zeros((3, 2))),
           Rmock.numeric(np.zeros(nM)), Rmock.numeric(np.zeros(2)), Rmock.integer(np.full(nM, 7)), Rmock.integer(s),
           Rmock.integer([2]), Rmock.integer([3]), Rmock.integer([nM]), Rmock.numeric([1.0]), Rmock.numeric([1.0]),
           Rmock.logical(True), Rmock.integer(np.arange(1, nM + 1)), Rmock.numeric(np.full(nM, 0.5))]
    with pytest.raises(RuntimeError):
        Rmock.call("_celda_cuda_cCCalcGibbsProbZ", *bad)  # z = 7 with K = 2
    assert Rmock.L.rstub_protect_balance() == 0


def test_shim_and_stub_clean_under_asan_and_ubsan():
    """The CPU tests of this file once more in a subprocess with the shim, the R API miniature and the mock built with
    AddressSanitizer + UndefinedBehaviorSanitizer (libasan preloaded): the marshalling of the closures (copies, dims,
    named lists), the PROTECT bookkeeping and the error exits touch no memory they do not own."""
    import sys
    if os.environ.get("CELDA_SHIM_SANITIZE") == "1":
        pytest.skip("already inside the sanitizer run")
    try:
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    except Exception:
        asan = ""
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("gcc's libasan is not available")
    env = dict(os.environ, LD_PRELOAD=asan, CELDA_SHIM_SANITIZE="1", ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
               UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "not gpu",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["int", "numeric"])
def test_matrixSums_known_answers_through_dotCall(R, kind):
    """tests/testthat/test-matrixSums.R:5-76 the way R would run it: `.Call` on matrices and factors."""
    dtype, sfx = (np.int32, "") if kind == "int" else (np.float64, "_numeric")
    mat = np.tile(np.arange(1, 6), 20).reshape(10, 10, order="F").astype(dtype)
    lab, lab2 = np.repeat([1, 2], 5), np.array([1, 2] * 5)
    want_r = np.stack([mat[lab == g].sum(0) for g in (1, 2)])
    want_c = np.stack([mat[:, lab == g].sum(1) for g in (1, 2)], axis=1)
    got_r = R.call("_rowSumByGroup" + sfx, R.matrix(mat), R.factor(lab, 2))
    got_c = R.call("_colSumByGroup" + sfx, R.matrix(mat), R.factor(lab, 2))
    assert got_r.dtype == dtype and np.array_equal(got_r, want_r) and np.array_equal(got_c, want_c)
    px = R.matrix(want_r)
    out = R.call("_rowSumByGroupChange" + sfx, R.matrix(mat), px, R.factor(lab2, 2), R.factor(lab, 2), raw=True)
    assert out == px  # modified in place and returned, like the reference (src/matrixSums.c:136-142)
    assert np.array_equal(R.value(px), np.stack([mat[lab2 == g].sum(0) for g in (1, 2)]))
    px = R.matrix(want_c)
    R.call("_colSumByGroupChange" + sfx, R.matrix(mat), px, R.factor(lab2, 2), R.factor(lab, 2))
    assert np.array_equal(R.value(px), np.stack([mat[:, lab2 == g].sum(1) for g in (1, 2)], axis=1))


@pytest.mark.gpu
def test_sparse_aggregation_and_natives_through_dotCall(R, oracle, gold_cg):
    from oracle.oracle import dense_to_csc
    counts = gold_cg["counts"]
    nG, nM = counts.shape
    rng = np.random.default_rng(3)
    z, z2 = rng.integers(1, 6, nM), rng.integers(1, 6, nM)
    y, y2 = rng.integers(1, 11, nG), rng.integers(1, 11, nG)
    O, A = dense_to_csc(counts), R.dgCMatrix(counts)
    cs = R.call("_celda_cuda_colSumByGroupSparse", A, R.integer(z), R.integer([5]))
    rs = R.call("_celda_cuda_rowSumByGroupSparse", A, R.integer(y), R.integer([10]))
    assert np.array_equal(cs, oracle.colSumByGroupSparse(O, z.astype(np.int32), 5))
    assert np.array_equal(rs, oracle.rowSumByGroupSparse(O, y.astype(np.int32), 10))
    px = R.matrix(cs)
    R.call("_celda_cuda_colSumByGroupChangeSparse", A, px, R.integer(z2), R.integer(z), R.integer([5]))
    assert np.array_equal(R.value(px), oracle.colSumByGroupSparse(O, z2.astype(np.int32), 5))
    px = R.matrix(rs)
    R.call("_celda_cuda_rowSumByGroupChangeSparse", A, px, R.integer(y2), R.integer(y), R.integer([10]))
    assert np.array_equal(R.value(px), oracle.rowSumByGroupSparse(O, y2.astype(np.int32), 10))
    # eigenMatMultInt(A, B) = t(A) %*% B (src/eigenMatMultInt.cpp:11-20), fastNormPropLog (src/matrixNorm.cpp:36-52)
    a, b = rng.random((40, 6)), rng.integers(0, 9, (40, 7)).astype(np.int32)
    assert np.allclose(R.call("_celda_cuda_eigenMatMultInt", R.matrix(a), R.matrix(b)), a.T @ b, rtol=1e-13)
    m = rng.integers(0, 50, (12, 9)).astype(np.float64)
    want = np.log((m + 0.5) / (m + 0.5).sum(0))
    assert np.allclose(R.call("_celda_cuda_fastNormPropLog", R.matrix(m), R.numeric([0.5])), want, rtol=1e-12)


@pytest.mark.gpu
def test_chain_iteration_through_dotCall_matches_oracle(R, oracle, gold_cg):
    """create -> set_state -> iterate -> get_state / loglik / perplexity on an external pointer, one celda_CG Gibbs
    iteration (R/celda_CG.R:543-602) with injected draws; labels bit-exact, logLik 1e-9; the finalizer frees it."""
    from oracle.oracle import dense_to_csc
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(0)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
    ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
    h = R.call("_celda_cuda_chain_create", R.integer([0]), R.dgCMatrix(counts), R.integer(s),
               R.integer([int(s.max())]), R.integer([K]), R.integer([L]), R.numeric([1, 1, 1, 1]), R.integer([0]),
               raw=True)
    assert R.L.rstub_type(h) == EXTPTRSXP
    R.call("_celda_cuda_chain_set_state", h, R.integer(z), R.integer(y))
    ll = R.call("_celda_cuda_chain_iterate", h, R.integer(ixy), R.numeric(uy), R.integer(ixz), R.numeric(uz))[0]
    zg, yg = R.call("_celda_cuda_chain_get_state", h, R.integer([nM]), R.integer([nG]))
    O = dense_to_csc(counts)
    p = oracle.cCGDecomposeCounts(O, s, z, y, K, L)
    oy = oracle.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0,
                                 1.0, ixy, uy)
    nTSByC = oracle.rowSumByGroupChangeSparse(O, p["nTSByC"].copy(order="F"), oy["y"], y, L)
    oz = oracle.cCCalcGibbsProbZ(nTSByC, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0,
                                 ixz, uz)
    ref = oracle.cCGCalcLL(K, L, oz["mCPByS"], oz["nGByCP"], p["nByG"], oy["nByTS"], oy["nGByTS"], p["nS"], nG, 1.0,
                           1.0, 1.0, 1.0)
    assert np.array_equal(zg, oz["z"]) and np.array_equal(yg, oy["y"])
    assert abs(ll - ref) <= 1e-9 * abs(ref)
    assert abs(R.call("_celda_cuda_chain_loglik", h)[0] - ref) <= 1e-9 * abs(ref)
    assert R.call("_celda_cuda_chain_perplexity", h)[0] > 1.0
    # an R-level error from inside the library (labels out of range) leaves the handle usable
    with pytest.raises(RuntimeError):
        R.call("_celda_cuda_chain_set_state", h, R.integer(np.full(nM, K + 1)), R.integer(y))
    R.call("_celda_cuda_chain_set_state", h, R.integer(zg), R.integer(yg))
    assert abs(R.call("_celda_cuda_chain_loglik", h)[0] - ref) <= 1e-9 * abs(ref)
    assert R.L.rstub_reset() == 1  # gc: the external pointer's finalizer ran celda_chain_destroy
    R.L.celda_cuda_register(None)
