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
    """celda_C chain over this rank's cell range; the tables it exposes are the global (all-reduced) ones."""

    def __init__(self, local_counts, s_local, K, comm, nS, alpha=1.0, beta=1.0, lo=0):
        """local_counts: this rank's columns; s_local: their sample labels; nS: number of samples over ALL ranks
        (every rank sizes its K x S table for all samples, not only those present in its shard)."""
        self.lo, self.hi = lo, lo + local_counts.ncol
        s_loc = np.ascontiguousarray(s_local, dtype=np.int32)
        self.chain = _chain_with_ns(local_counts, s_loc, K, alpha, beta, comm.device, int(nS))
        check(lib().celda_chain_attach_comm(self.chain._h, comm._c))

    @classmethod
    def from_full(cls, counts, sampleLabel, K, comm, alpha=1.0, beta=1.0):
        """Split a full matrix into column ranges balanced by stored entries and keep this rank's."""
        b = partition_columns_by_nnz(counts.p, comm.nranks)
        lo, hi = b[comm.rank], b[comm.rank + 1]
        s = np.asarray(sampleLabel, dtype=np.int32)
        return cls(shard_csc(counts, lo, hi), s[lo:hi], K, comm, int(s.max()), alpha, beta, lo=lo)

    def set_state(self, z_full=None, z_local=None):
        z = np.asarray(z_full, dtype=np.int32)[self.lo:self.hi] if z_local is None else z_local
        self.chain.set_state(z=z)

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
