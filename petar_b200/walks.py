"""Host-side container for what FDPS hands to the multiwalk dispatch functor.

A :class:`WalkBatch` owns exactly the arrays FDPS ``calcForceAllAndWriteBackMultiWalkIndex`` passes
to ``dispatch`` (reference call site ``src/petar.hpp:894-899``; argument meaning per
``src/force_gpu_cuda.hpp:120-132``):

* ``epj`` / ``spj``: the shared sorted j arrays (``epj_sorted``, ``spj_sorted``),
* per walk ``iw``: ``epi[iw][0:n_epi[iw]]``, ``id_epj[iw][0:n_epj[iw]]``, ``id_spj[iw][0:n_spj[iw]]``
  (indices into the shared arrays),

stored concatenated with offset tables, and hands out the pointer-to-pointer tables of that
signature (:meth:`pointer_tables`).
"""
from types import SimpleNamespace

import numpy as np

from .types import EPISoft, EPJSoft, SPJQuad, ForceSoft


class WalkBatch:
    def __init__(self, epj, spj, epi, i_off, id_epj, ej_off, id_spj, sj_off):
        self.epj = np.ascontiguousarray(epj, dtype=EPJSoft)
        self.spj = np.ascontiguousarray(spj, dtype=SPJQuad)
        self.epi = np.ascontiguousarray(epi, dtype=EPISoft)
        self.i_off = np.ascontiguousarray(i_off, dtype=np.int64)
        self.id_epj = np.ascontiguousarray(id_epj, dtype=np.int32)
        self.ej_off = np.ascontiguousarray(ej_off, dtype=np.int64)
        self.id_spj = np.ascontiguousarray(id_spj, dtype=np.int32)
        self.sj_off = np.ascontiguousarray(sj_off, dtype=np.int64)
        n = len(self.i_off) - 1
        assert len(self.ej_off) == n + 1 and len(self.sj_off) == n + 1
        assert self.i_off[-1] == len(self.epi)
        assert self.ej_off[-1] == len(self.id_epj) and self.sj_off[-1] == len(self.id_spj)
        if len(self.id_epj):
            assert self.id_epj.min() >= 0 and self.id_epj.max() < len(self.epj)
        if len(self.id_spj):
            assert self.id_spj.min() >= 0 and self.id_spj.max() < len(self.spj)

    # ---- sizes -----------------------------------------------------------------------
    @property
    def n_walk(self):
        return len(self.i_off) - 1

    @property
    def n_epi(self):
        return np.diff(self.i_off).astype(np.int32)

    @property
    def n_epj(self):
        return np.diff(self.ej_off).astype(np.int32)

    @property
    def n_spj(self):
        return np.diff(self.sj_off).astype(np.int32)

    @property
    def n_epi_total(self):
        return int(self.i_off[-1])

    def interactions(self, walk_slice=None):
        """(I_ep, I_sp) = (sum n_epi*n_epj, sum n_epi*n_spj): PeTar's Ep-Ep_sum / Ep-Sp_sum
        (reference src/petar.hpp:943-946)."""
        s = walk_slice or slice(None)
        ni = self.n_epi[s].astype(np.int64)
        return int((ni * self.n_epj[s]).sum()), int((ni * self.n_spj[s]).sum())

    # ---- FDPS-style pointer tables ------------------------------------------------------
    def pointer_tables(self, force, walk_slice=None):
        """Pointer-of-pointer tables in the dispatch/retrieve signature for walks in `walk_slice`.
        `force` is a ForceSoft array of length n_epi_total (whole batch); force[iw] pointers alias it."""
        assert force.dtype == ForceSoft and len(force) == self.n_epi_total and force.flags["C_CONTIGUOUS"]
        s = walk_slice or slice(None)
        io, eo, so = self.i_off[:-1][s], self.ej_off[:-1][s], self.sj_off[:-1][s]
        t = SimpleNamespace()
        t.n_walk = len(io)
        t.n_epi = np.ascontiguousarray(self.n_epi[s])
        t.n_epj = np.ascontiguousarray(self.n_epj[s])
        t.n_spj = np.ascontiguousarray(self.n_spj[s])
        t.epi_ptrs = (self.epi.ctypes.data + io * EPISoft.itemsize).astype(np.uint64)
        t.id_epj_ptrs = (self.id_epj.ctypes.data + eo * 4).astype(np.uint64)
        t.id_spj_ptrs = (self.id_spj.ctypes.data + so * 4).astype(np.uint64)
        t.force_ptrs = (force.ctypes.data + io * ForceSoft.itemsize).astype(np.uint64)
        t.i_begin = int(io[0]) if len(io) else 0
        t.i_end = int(io[-1] + t.n_epi[-1]) if len(io) else 0
        return t

    @staticmethod
    def single(epi, epj, spj):
        """One walk over flat arrays with identity index lists (reference src/simd_test.cxx:139-150)."""
        return WalkBatch(epj, spj, epi, [0, len(epi)], np.arange(len(epj), dtype=np.int32), [0, len(epj)],
                         np.arange(len(spj), dtype=np.int32), [0, len(spj)])
