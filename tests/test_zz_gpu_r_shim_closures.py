"""The R closures through `.Call` on the GPU: r/celda_cuda_shim.c over the real libcelda_cuda.so (tests/test_r_shim.py
runs the same round trip on the CPU with the oracle-backed mock behind the shim).  .cCCalcGibbsProbZ and
.cGCalcGibbsProbY with the reference's arguments and named result lists, every table, label and probability equal to
the oracle called directly; the EM closure on a dgCMatrix."""
import numpy as np
import pytest

from test_r_shim import _build_shim, gibbs_closures_roundtrip

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Rgpu(tmp_path_factory):
    stub = _build_shim(str(tmp_path_factory.mktemp("rshim_gpu")))
    yield stub
    stub.L.rstub_reset()


def test_gibbs_closures_through_dotCall_on_the_device(Rgpu, oracle, gold_cg):
    gibbs_closures_roundtrip(Rgpu, oracle, gold_cg)


def test_em_closure_through_dotCall_on_the_device(Rgpu, oracle, gold_c):
    from oracle.oracle import dense_to_csc
    counts, s, K = gold_c["counts"], gold_c["s_mod"], 10
    nG, nM = counts.shape
    z = np.random.default_rng(4).integers(1, K + 1, nM).astype(np.int32)
    O = dense_to_csc(counts)
    p = oracle.cCDecomposeCounts(O, s, z, K)
    want = oracle.cCCalcEMProbZ(O, p["mCPByS"], p["nGByCP"], p["nByC"], p["nCP"], z, s, K, nG, nM, 1.0, 1.0)
    got = Rgpu.call("_celda_cuda_cCCalcEMProbZ", Rgpu.dgCMatrix(counts), Rgpu.matrix(p["mCPByS"]), Rgpu.matrix(p["nGByCP"]),
                    Rgpu.numeric(p["nByC"]), Rgpu.numeric(p["nCP"]), Rgpu.integer(z), Rgpu.integer(s), Rgpu.integer([K]),
                    Rgpu.integer([nG]), Rgpu.integer([nM]), Rgpu.numeric([1.0]), Rgpu.numeric([1.0]), Rgpu.logical(True))
    assert list(got) == ["mCPByS", "nGByCP", "nCP", "z", "probs"]
    for k in ("mCPByS", "nGByCP", "nCP", "z"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["probs"], want["probs"], rtol=1e-12, atol=0)
