"""Device side of tests/test_philox_order.py: the orders and uniforms a Philox-mode iteration consumes on the GPU are
exactly those the host build of celda_b200/csrc/philox.cuh gives for (seed, sweep id, stream) -- sweep ids count the
draw sets since celda_chain_seed, the y sweep of iteration `it` is 2 it, its z sweep 2 it + 1."""
import numpy as np
import pytest

import celda_b200 as cb

pytestmark = pytest.mark.gpu


def test_device_orders_equal_the_host_build_of_philox_cuh(gold_cg, tmp_path):
    from test_philox_order import build_host_order
    host_order = build_host_order(str(tmp_path))
    counts, s, K, L = gold_cg["counts"], gold_cg["s"], 5, 10
    nG, nM = counts.shape
    ch = cb.Chain(cb.CSC.from_dense(counts), s, K=K, L=L)
    ch.set_state(gold_cg["z"], gold_cg["y"])
    for seed, stream in ((12345, 0), (2 ** 40 + 17, 3)):
        ch.seed(seed, stream=stream)
        for it in range(2):
            ch.iterate_philox()
            ixy, uy = ch.get_draws("y")
            ixz, uz = ch.get_draws("z")
            hy, hu = host_order(nG, seed=seed, sweep=2 * it, stream=stream, uniforms=True)
            hz, hv = host_order(nM, seed=seed, sweep=2 * it + 1, stream=stream, uniforms=True)
            assert np.array_equal(ixy, hy[0]) and np.array_equal(uy, hu[0])
            assert np.array_equal(ixz, hz[0]) and np.array_equal(uz, hv[0])
    ch.close()
