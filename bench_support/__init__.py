"""Benchmark / large-test support: synthetic celda_CG counts shaped like simulateCells output."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    so = os.path.join(_HERE, "libsimcells.so")
    src = os.path.join(_HERE, "simulate_cells.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def _threads():
    """Generator threads: all host cores divided among the ranks of a torchrun launch."""
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    return max(1, (os.cpu_count() or 1) // max(1, world))


def simulate_cells_cg(seed, S, cells_per_sample, G, K, L, nrange, alpha=1.0, beta=1.0, gamma=5.0, delta=1.0):
    """.simulateCellscelda_CG (R/simulateCells.R:302-435): theta_s ~ Dir(alpha), z_c ~ Cat(theta_s);
    phi_k ~ Dir(beta 1_L); eta ~ Dir(gamma 1_L), y_g ~ Cat(eta); psi_l ~ Dir(delta) over the genes of module l;
    N_c ~ U{nrange}; transcripts ~ two-stage multinomial.  Genes with zero total are dropped like the reference.
    Returns dict(nG, nM, p(int32), i, x(float64), s, z, y, K, L)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nM = S * cells_per_sample
    s = np.repeat(np.arange(1, S + 1, dtype=np.int32), cells_per_sample)
    theta = rng.dirichlet(np.full(K, alpha), S)
    cum = np.cumsum(theta, axis=1)
    z = (rng.random(nM)[:, None] > cum[s - 1]).sum(1).clip(0, K - 1).astype(np.int32) + 1
    z[:K] = np.arange(1, K + 1)  # every population present
    phi = rng.dirichlet(np.full(L, beta), K)
    eta = rng.dirichlet(np.full(L, gamma))
    y = rng.choice(L, size=G, p=eta).astype(np.int32) + 1
    y[:L] = np.arange(1, L + 1)  # every module present
    order = np.argsort(y, kind="stable").astype(np.int32)
    mod_off = np.zeros(L + 1, dtype=np.int32)
    np.cumsum(np.bincount(y - 1, minlength=L), out=mod_off[1:])
    psi = np.concatenate([rng.dirichlet(np.full(mod_off[l + 1] - mod_off[l], delta)) for l in range(L)])
    N = rng.integers(nrange[0], nrange[1] + 1, nM).astype(np.int32)
    lib = C.CDLL(build())
    lib.sim_counts.restype = C.c_longlong
    colptr = np.zeros(nM + 1, dtype=np.int64)
    rows, vals = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    nnz = lib.sim_counts(C.c_int(G), C.c_longlong(nM), C.c_int(K), C.c_int(L), z.ctypes.data_as(ip),
                         N.ctypes.data_as(ip), np.ascontiguousarray(phi).ctypes.data_as(dp),
                         mod_off.ctypes.data_as(ip), order.ctypes.data_as(ip), psi.ctypes.data_as(dp),
                         C.c_uint64(seed), C.c_int(_threads()), C.c_longlong(0), colptr.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(rows), C.byref(vals))
    ri = np.ctypeslib.as_array(rows, shape=(max(nnz, 1),))[:nnz].copy()
    xv = np.ctypeslib.as_array(vals, shape=(max(nnz, 1),))[:nnz].astype(np.float64)
    lib.sim_free(rows)
    lib.sim_free(vals)
    # drop all-zero genes (R/simulateCells.R:404-411), relabel
    present = np.zeros(G, dtype=bool)
    present[ri] = True
    if not present.all():
        remap = np.cumsum(present) - 1
        ri = remap[ri].astype(np.int32)
        y = y[present]
        G = int(present.sum())
    return dict(nG=G, nM=nM, p=colptr.astype(np.int32), i=ri.astype(np.int32), x=xv, s=s, z=z, y=y, K=K, L=L)


def simulate_cells_c(seed, S, cells_per_sample, G, K, nrange, alpha=1.0, beta=1.0, cell_range=None):
    """.simulateCellscelda_C (R/simulateCells.R:151-218): theta_s ~ Dir(alpha), z_c ~ Cat(theta_s), phi_k ~ Dir(beta 1_G),
    counts[, c] ~ Mult(N_c, phi_{z_c}).  cell_range = (lo, hi) generates only those columns of the same matrix (one
    RNG stream per global cell index), which is how the ranks of a cell-sharded run each build their own shard.
    Returns dict(nG, nM (columns generated), p, i, x, s, z, lo, hi, nM_total)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nM = S * cells_per_sample
    s = np.repeat(np.arange(1, S + 1, dtype=np.int32), cells_per_sample)
    theta = rng.dirichlet(np.full(K, alpha), S)
    cum = np.cumsum(theta, axis=1)
    z = (rng.random(nM)[:, None] > cum[s - 1]).sum(1).clip(0, K - 1).astype(np.int32) + 1
    z[:K] = np.arange(1, K + 1)
    phi = rng.dirichlet(np.full(G, beta), K)  # K x G gene distributions
    N = rng.integers(nrange[0], nrange[1] + 1, nM).astype(np.int32)
    lo, hi = (0, nM) if cell_range is None else cell_range
    # reuse the two-stage sampler: "module" k = population k's own gene distribution, chosen with probability 1
    onehot = np.eye(K)
    mod_off = (np.arange(K + 1) * G).astype(np.int32)
    mod_genes = np.tile(np.arange(G, dtype=np.int32), K)
    lib = C.CDLL(build())
    lib.sim_counts.restype = C.c_longlong
    n_loc = hi - lo
    colptr = np.zeros(n_loc + 1, dtype=np.int64)
    rows, vals = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    zl, Nl = np.ascontiguousarray(z[lo:hi]), np.ascontiguousarray(N[lo:hi])
    nnz = lib.sim_counts(C.c_int(G), C.c_longlong(n_loc), C.c_int(K), C.c_int(K), zl.ctypes.data_as(ip),
                         Nl.ctypes.data_as(ip), np.ascontiguousarray(onehot).ctypes.data_as(dp),
                         mod_off.ctypes.data_as(ip), mod_genes.ctypes.data_as(ip),
                         np.ascontiguousarray(phi).ravel().ctypes.data_as(dp), C.c_uint64(seed),
                         C.c_int(_threads()), C.c_longlong(lo), colptr.ctypes.data_as(C.POINTER(C.c_longlong)),
                         C.byref(rows), C.byref(vals))
    ri = np.ctypeslib.as_array(rows, shape=(max(nnz, 1),))[:nnz].copy()
    xv = np.ctypeslib.as_array(vals, shape=(max(nnz, 1),))[:nnz].astype(np.float64)
    lib.sim_free(rows)
    lib.sim_free(vals)
    return dict(nG=G, nM=n_loc, nM_total=nM, lo=lo, hi=hi, p=colptr.astype(np.int64), i=ri.astype(np.int32), x=xv,
                s=s, z=z, K=K)
