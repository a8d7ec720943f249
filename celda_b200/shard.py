"""Cell sharding for the EM z step and count aggregation across GPUs.

One process per GPU.  The count matrix is split into contiguous column (cell) ranges balanced by stored entries;
each rank owns a celda_C chain over its range; the partial G x K / K x S tables are all-reduced inside
libcelda_cuda with NCCL after every (re-)decomposition, so all ranks hold bit-identical global tables and compute
phi / theta / the log-likelihood redundantly.
"""
import ctypes as C

import numpy as np

from . import CSC, MODEL_C, Chain
from ._lib import check, lib


def partition_columns_by_nnz(p, nranks):
    """Column boundaries b[0..nranks] with b[0] = 0, b[-1] = ncol, chosen so every rank holds ~nnz / nranks stored
    entries (each range non-empty when ncol >= nranks)."""
    p = np.asarray(p, dtype=np.int64)
    ncol, nnz = p.size - 1, int(p[-1])
    b = [0]
    for r in range(1, nranks):
        c = int(np.searchsorted(p, nnz * r / nranks, side="left"))
        c = min(max(c, b[-1] + 1), ncol - (nranks - r))
        b.append(c)
    b.append(ncol)
    return b


def shard_csc(counts, lo, hi):
    """Columns [lo, hi) of a CSC matrix (row dimension unchanged)."""
    a, e = int(counts.p[lo]), int(counts.p[hi])
    return CSC(counts.nrow, hi - lo, counts.p[lo:hi + 1] - a, counts.i[a:e], counts.x[a:e])


class Comm:
    """NCCL communicator owned by libcelda_cuda (celda_comm_*)."""

    def __init__(self, unique_id, nranks, rank, device):
        self._c = C.c_void_p()
        check(lib().celda_comm_create(C.byref(self._c), unique_id, nranks, rank, device))
        self.nranks, self.rank, self.device = nranks, rank, device

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(lib().celda_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if self._c:
            lib().celda_comm_destroy(self._c)
            self._c = C.c_void_p()


class ShardedChainC:
    """celda_C chain over this rank's cell range; the tables it exposes are the global ones."""

    def __init__(self, counts, sampleLabel, K, comm, alpha=1.0, beta=1.0, nS=None):
        self.bounds = partition_columns_by_nnz(counts.p, comm.nranks)
        self.lo, self.hi = self.bounds[comm.rank], self.bounds[comm.rank + 1]
        s = np.asarray(sampleLabel, dtype=np.int32)
        self.nS = int(s.max()) if nS is None else nS
        local = shard_csc(counts, self.lo, self.hi)
        self.chain = Chain(local, s[self.lo:self.hi], K=K, alpha=alpha, beta=beta, model=MODEL_C, device=comm.device)
        # every rank must size its K x S table for ALL samples, not only those present in its shard
        if self.chain.nS != self.nS:
            self.chain.close()
            s_loc = s[self.lo:self.hi].copy()
            self.chain = _chain_with_ns(local, s_loc, K, alpha, beta, comm.device, self.nS)
        check(lib().celda_chain_attach_comm(self.chain._h, comm._c))

    def set_state(self, z):
        self.chain.set_state(z=np.asarray(z, dtype=np.int32)[self.lo:self.hi])

    def iterate_em(self):
        return self.chain.iterate_em()

    def local_z(self):
        return self.chain.get_state()[0]

    def close(self):
        self.chain.close()


def _chain_with_ns(local, s_loc, K, alpha, beta, device, nS):
    ch = Chain.__new__(Chain)
    ch._h = C.c_void_p()
    ch.counts, ch.model, ch.K, ch.L = local, MODEL_C, int(K), 1
    ch.nG, ch.nM, ch.nS, ch.s = local.nrow, local.ncol, nS, s_loc
    check(lib().celda_chain_create(C.byref(ch._h), device, MODEL_C, local.nrow, local.ncol,
                                   local.p.ctypes.data_as(C.POINTER(C.c_int)), local.i.ctypes.data_as(C.POINTER(C.c_int)),
                                   local.x.ctypes.data_as(C.POINTER(C.c_double)),
                                   s_loc.ctypes.data_as(C.POINTER(C.c_int)), nS, int(K), 1, alpha, beta, 1.0, 1.0))
    return ch
