"""The CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer: the oracle is what every GPU result is compared
with, so a silent out-of-bounds read in it would void the parity claims.  The sanitizer build (oracle/Makefile target
`asan`) is loaded in a subprocess with libasan preloaded; any report aborts the subprocess."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from oracle.oracle import Oracle, dense_to_csc
o = Oracle(lib_path=os.path.join(sys.argv[1], "oracle", "libcelda_oracle_asan.so"))
g = dict(np.load(os.path.join(sys.argv[1], "tests", "golden", "celdaCG.npz")))
counts, s, K, L = g["counts"], g["s"], 5, 10
nG, nM = counts.shape
rng = np.random.default_rng(1)
O = dense_to_csc(counts)
z = rng.integers(1, K + 1, nM).astype(np.int32)
y = rng.integers(1, L + 1, nG).astype(np.int32)
p = o.cCGDecomposeCounts(O, s, z, y, K, L)
ixy, uy = rng.permutation(nG).astype(np.int32) + 1, rng.random(nG)
ixz, uz = rng.permutation(nM).astype(np.int32) + 1, rng.random(nM)
oy = o.cGCalcGibbsProbY(p["nGByCP"], p["nTSByCP"], p["nByTS"], p["nGByTS"], p["nByG"], y, L, nG, 1.0, 1.0, 1.0, ixy, uy)
t = o.rowSumByGroupChangeSparse(O, p["nTSByC"].copy(order="F"), oy["y"], y, L)
oz = o.cCCalcGibbsProbZ(t, p["mCPByS"], oy["nTSByC"], p["nByC"], p["nCP"], z, s, K, L, nM, 1.0, 1.0, ixz, uz)
ll = o.cCGCalcLL(K, L, oz["mCPByS"], oz["nGByCP"], p["nByG"], oy["nByTS"], oy["nGByTS"], p["nS"], nG, 1.0, 1.0, 1.0, 1.0)
assert np.isfinite(ll)
# celda_C: EM step + Gibbs on the raw counts; celda_G on the raw counts
pc = o.cCDecomposeCounts(O, s, z, K)
em = o.cCCalcEMProbZ(O, pc["mCPByS"], pc["nGByCP"], pc["nByC"], pc["nCP"], z, s, K, pc["nG"], pc["nM"], 1.0, 1.0)
assert em["z"].shape == (nM,)
pg = o.cGDecomposeCounts(O, y, L)
og = o.cGCalcGibbsProbY(counts.astype(np.float64), pg["nTSByC"], pg["nByTS"], pg["nGByTS"], pg["nByG"], y, L, nG, 1.0, 1.0,
                        1.0, ixy, uy)
assert og["y"].shape == (nG,)
# sampler: ties, fall-through, the Walker alias branch (> 200 likely categories)
for n in (3, 150, 260):
    llv = np.log(rng.dirichlet(np.full(n, 30.0)))
    for u in (1e-300, 0.3, 1 - 1e-16):
        lab, _ = o.sampleLl(llv, u)
        assert 1 <= lab <= n
lab, _ = o.sampleLl(np.zeros(7), 0.5)
assert 1 <= lab <= 7
# aggregation natives incl. the in-place Change variants
a = o.colSumByGroupSparse(O, z, K)
b = o.colSumByGroupChangeSparse(O, a.copy(order="F"), z, z[::-1].copy(), K)
assert a.shape == b.shape
print("SANITIZED_OK")
"""


def _libasan():
    try:
        path = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    except Exception:
        return None
    return path if os.path.isabs(path) and os.path.exists(path) else None


@pytest.mark.timeout(300)
def test_oracle_clean_under_asan_and_ubsan():
    asan = _libasan()
    if asan is None:
        pytest.skip("gcc's libasan is not available")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "asan"])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT], capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0 and "SANITIZED_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
