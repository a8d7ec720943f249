"""The visiting order of the Philox mode (celda_b200/csrc/philox.cuh: a keyed Feistel permutation with cycle walking,
no sort) checked on the CPU: the header is plain C++ under g++, so the very function the device kernel calls is
compiled here and tested for what a `sample(seq_len(n))` replacement must be -- a bijection of 1..n for every n,
a function of (seed, sweep, stream) only, and statistically uniform (every permutation of a small n equally often,
every item equally often at every position, no correlation between neighbours).  tests/test_gpu_philox.py checks
that the device produces exactly these orders."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "celda_b200", "csrc")

WRAPPER = r"""
#include "philox.cuh"
extern "C" {
// orders of `nsweep` consecutive sweeps: out[s * n + t] = 1-based item visited t-th in sweep sweep0 + s
void order_host(long long n, unsigned long long seed, unsigned sweep0, unsigned stream, int nsweep, int* out, double* u) {
    const unsigned bits = celda::philox_order_bits((unsigned long long)n);
    for (int s = 0; s < nsweep; ++s)
        for (long long t = 0; t < n; ++t) {
            out[(long long)s * n + t] = (int)celda::philox_order((unsigned long long)t, (unsigned long long)n, bits,
                                                                (unsigned)seed, (unsigned)(seed >> 32), sweep0 + s, stream) + 1;
            if (u) u[(long long)s * n + t] = celda::philox_uniform((unsigned long long)t, (unsigned)seed,
                                                                   (unsigned)(seed >> 32), sweep0 + s, stream);
        }
}
void philox_host(const unsigned* ctr, const unsigned* key, unsigned* out) {
    const celda::Philox4 r = celda::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int q = 0; q < 4; ++q) out[q] = r.x[q];
}
}
"""


def build_host_order(tmpdir):
    src, so = os.path.join(tmpdir, "order_host.cpp"), os.path.join(tmpdir, "liborder_host.so")
    open(src, "w").write(WRAPPER)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I" + CSRC, "-o", so,
                        src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    L.order_host.restype = None
    L.order_host.argtypes = [C.c_longlong, C.c_ulonglong, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_void_p]

    L.philox_host.restype = None
    L.philox_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

    def philox(ctr, key):
        c, k, out = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
        L.philox_host(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        return out.tolist()

    def order(n, seed=12345, sweep=0, stream=0, nsweep=1, uniforms=False):
        out = np.zeros((nsweep, n), dtype=np.int32)
        u = np.zeros((nsweep, n)) if uniforms else None
        L.order_host(n, seed, sweep, stream, nsweep, out.ctypes.data, u.ctypes.data if uniforms else None)
        return (out, u) if uniforms else out
    order.philox = philox
    return order


@pytest.fixture(scope="module")
def order(tmp_path_factory):
    return build_host_order(str(tmp_path_factory.mktemp("philox_order")))


def test_generator_known_answers_on_the_host_build(order):
    """Random123's kat_vectors for philox4x32-10 on the host build of the header the device compiles
    (tests/test_gpu_philox.py checks the same vectors on the GPU)."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert order.philox(ctr, key) == want


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 63, 64, 65, 83, 100, 127, 128, 129, 1000, 4097, 19999, 100000,
                               (1 << 20) + 1])
def test_order_is_a_permutation_for_every_n(order, n):
    ix = order(n)[0]
    assert np.array_equal(np.sort(ix), np.arange(1, n + 1, dtype=np.int32))


def test_order_is_a_permutation_for_all_small_n(order):
    """Every n up to 2100 (all the small bit widths and their boundaries), three sweeps each."""
    for n in range(1, 2101):
        ix = order(n, seed=n, sweep=n % 7, stream=n % 3, nsweep=3)
        want = np.arange(1, n + 1, dtype=np.int32)
        assert all(np.array_equal(np.sort(row), want) for row in ix), n


def test_order_depends_on_seed_sweep_and_stream_only(order):
    n = 5000
    base = order(n, seed=7, sweep=3, stream=1)[0]
    assert np.array_equal(base, order(n, seed=7, sweep=3, stream=1)[0])
    assert np.array_equal(base, order(n, seed=7, sweep=2, stream=1, nsweep=2)[1])  # sweep id, not call count
    for other in (order(n, seed=8, sweep=3, stream=1)[0], order(n, seed=7, sweep=4, stream=1)[0],
                  order(n, seed=7, sweep=3, stream=2)[0], order(n, seed=7 + (1 << 32), sweep=3, stream=1)[0]):
        assert (other == base).mean() < 0.01  # a fresh permutation agrees in about 1/n of the positions
    assert (base == np.arange(1, n + 1)).mean() < 0.01


def test_small_n_every_permutation_equally_often(order):
    """n = 4 lives in a 6-bit domain with heavy cycle walking: all 24 permutations, chi-square against uniform."""
    n, ns = 4, 48000
    ix = order(n, nsweep=ns)
    code = ((ix - 1) * (n ** np.arange(n))).sum(1)
    _, cnt = np.unique(code, return_counts=True)
    assert cnt.size == 24
    chi2 = ((cnt - ns / 24) ** 2 / (ns / 24)).sum()
    assert chi2 < 60.0, chi2  # 23 degrees of freedom: P(chi2 > 60) ~ 4e-5


@pytest.mark.parametrize("n", [7, 83, 257])
def test_every_item_equally_often_at_every_position(order, n):
    ns = 40000 if n < 100 else 20000
    ix = order(n, seed=99, nsweep=ns) - 1
    tab = np.zeros((n, n))
    np.add.at(tab, (np.broadcast_to(np.arange(n), ix.shape), ix), 1)
    # a sum of ns uniform permutation matrices has fixed margins: cell variance ns (1/n)(1 - 1/n), so the Pearson
    # statistic scaled by (n - 1) / n is chi-square with (n - 1)^2 degrees of freedom
    chi2 = ((tab - ns / n) ** 2 / (ns / n)).sum() * (n - 1) / n
    dof = (n - 1) ** 2
    assert abs(chi2 - dof) < 5.0 * np.sqrt(2.0 * dof), (chi2, dof)


def test_neighbours_in_the_order_are_uncorrelated(order):
    n = 20000
    ix = order(n, seed=5, nsweep=8).astype(np.float64)
    for row in ix:
        assert abs(np.corrcoef(row[:-1], row[1:])[0, 1]) < 4.0 / np.sqrt(n)
        assert abs(np.corrcoef(np.arange(n), row)[0, 1]) < 4.0 / np.sqrt(n)
    # consecutive sweeps are unrelated permutations
    assert abs(np.corrcoef(ix[0], ix[1])[0, 1]) < 4.0 / np.sqrt(n)


def test_uniforms_are_open_interval_and_uniform(order):
    _, u = order(100000, uniforms=True)
    u = u[0]
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 4.0 / np.sqrt(12 * u.size)
    hist = np.histogram(u, bins=64, range=(0, 1))[0]
    chi2 = ((hist - u.size / 64) ** 2 / (u.size / 64)).sum()
    assert chi2 < 130.0, chi2  # 63 degrees of freedom
