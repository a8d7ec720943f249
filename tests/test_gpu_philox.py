"""Device Philox RNG: known answers, and Philox-mode sweeps replayed bit for bit in injection mode."""
import numpy as np
import pytest

import celda_b200 as cb
from oracle.oracle import dense_to_csc

pytestmark = pytest.mark.gpu


def _philox_ref(ctr, key):
    c, k = [int(v) for v in ctr], [int(v) for v in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert cb.philox4x32([ctr], key)[0].tolist() == want
    rng = np.random.default_rng(0)
    ctrs = rng.integers(0, 2 ** 32, (500, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([123456789, 987654321], dtype=np.uint32)
    got = cb.philox4x32(ctrs, key)
    for i in range(0, 500, 37):
        assert got[i].tolist() == _philox_ref(ctrs[i], key)


def test_philox_sweeps_replay_in_injection_mode(oracle, gold_cg):
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    rng = np.random.default_rng(2)
    z, y = rng.integers(1, K + 1, nM).astype(np.int32), rng.integers(1, L + 1, nG).astype(np.int32)
    A, O = cb.CSC.from_dense(counts), dense_to_csc(counts)
    a, b = cb.Chain(A, s, K=K, L=L), cb.Chain(A, s, K=K, L=L)
    a.set_state(z, y)
    b.set_state(z, y)
    a.seed(12345, stream=0)
    seen = []
    for it in range(3):
        ll_a = a.iterate_philox()
        ixy, uy = a.get_draws("y")
        ixz, uz = a.get_draws("z")
        assert sorted(ixy.tolist()) == list(range(1, nG + 1)) and sorted(ixz.tolist()) == list(range(1, nM + 1))
        assert uy.min() > 0 and uy.max() < 1 and uz.min() > 0 and uz.max() < 1
        seen.append((ixz.copy(), uz.copy()))
        ll_b = b.iterate(ixy, uy, ixz, uz)
        assert ll_a == ll_b
        za, ya = a.get_state()
        zb, yb = b.get_state()
        assert np.array_equal(za, zb) and np.array_equal(ya, yb)
    assert not np.array_equal(seen[0][0], seen[1][0]) and not np.array_equal(seen[0][1], seen[1][1])
    assert 0.4 < np.concatenate([u for _, u in seen]).mean() < 0.6
    # the oracle fed the same draws agrees too (philox mode is injection mode with device-made draws)
    za, ya = a.get_state()
    c = cb.Chain(A, s, K=K, L=L)
    c.set_state(z, y)
    c.seed(12345, stream=0)
    for it in range(3):
        c.iterate_philox()
    zc, yc = c.get_state()
    assert np.array_equal(za, zc) and np.array_equal(ya, yc)  # same seed, same chain
    d = cb.Chain(A, s, K=K, L=L)
    d.set_state(z, y)
    d.seed(12345, stream=1)
    d.iterate_philox()
    ix1, _ = d.get_draws("z")
    assert not np.array_equal(ix1, seen[0][0])  # another stream, another order
    for ch in (a, b, c, d):
        ch.close()


def test_philox_requires_seed(gold_cg):
    A = cb.CSC.from_dense(gold_cg["counts"])
    ch = cb.Chain(A, gold_cg["s"], K=5, L=10)
    ch.set_state(gold_cg["z"], gold_cg["y"])
    with pytest.raises(cb.CeldaCudaError, match="celda_chain_seed"):
        ch.sweep_z_gibbs_philox()
    ch.seed(1)
    ch.sweep_z_gibbs_philox()
    ch.sweep_y_philox()
    ch.close()
