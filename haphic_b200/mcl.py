"""Host side of the Markov-cluster path (hh_mcl handle).

Mirrors ``run_mcl_clustering`` / ``mcl`` / ``interpret_result``
(scripts/HapHiC_cluster.py:2026-2095, 2132-2162); the matrix work runs on the GPU.
"""

from __future__ import annotations

import ctypes as C
from decimal import Decimal

import numpy as np

from ._lib import HH_PREEXP_AUTO, HH_PREEXP_DENSE, HH_PREEXP_SPARSE, MclResult, PreexpInfo, check, load, ptr
from .links import LinkMatrix


class Mcl:
    """M0 = column-normalised link matrix and M1 = M0^expansion resident on the device, shared
    by every inflation of the sweep (HapHiC_cluster.py:2144-2158)."""

    PREEXP = {"auto": HH_PREEXP_AUTO, "sparse": HH_PREEXP_SPARSE, "dense": HH_PREEXP_DENSE}
    _close_order = 2

    def __init__(self, matrix: LinkMatrix, expansion: int = 2, col_lo: int = 0, col_hi: int | None = None,
                 preexp: str = "auto"):
        self.ctx = matrix.ctx
        self.n = matrix.n
        self.col_lo = int(col_lo)
        self.col_hi = self.n if col_hi is None else int(col_hi)
        self._own = (self.col_lo, self.col_hi)
        self._h = C.c_void_p()
        check(load().hh_mcl_create_ex(matrix._h, int(expansion), self.col_lo, self.col_hi, self.PREEXP[preexp],
                                      C.byref(self._h)))
        self.ctx.adopt(self)
        n = C.c_int32()
        nnz0 = C.c_int64()
        pre = C.c_int64()
        t0, t1 = C.c_float(), C.c_float()
        check(load().hh_mcl_info(self._h, C.byref(n), C.byref(nnz0), C.byref(pre), C.byref(t0), C.byref(t1)))
        self.nnz_m0 = int(nnz0.value)
        self.preexp_products = int(pre.value)
        self.normalize_ms, self.preexp_ms = float(t0.value), float(t1.value)
        pi = PreexpInfo()
        check(load().hh_mcl_preexp_info(self._h, C.byref(pi)))
        self.preexp = {k: getattr(pi, k) for k, _t in PreexpInfo._fields_}
        self.preexp["mode"] = {HH_PREEXP_SPARSE: "sparse", HH_PREEXP_DENSE: "dense"}.get(pi.mode, "?")
        self.last = None

    # -- inspection (parity tests) ------------------------------------------------------------
    def _fetch_csc(self, fn):
        import scipy.sparse as sp
        indptr = np.empty(self.n + 1, np.int64)
        check(fn(self._h, ptr(indptr), None, None))
        nnz = int(indptr[-1])
        indices = np.empty(nnz, np.int32)
        data = np.empty(nnz, np.float32)
        check(fn(self._h, ptr(indptr), ptr(indices), ptr(data)))
        return sp.csc_matrix((data, indices, indptr), shape=(self.n, self.n))

    def m0(self):
        return self._fetch_csc(load().hh_mcl_fetch_m0)

    def m1(self) -> np.ndarray:
        """Owned block of the pre-expanded matrix, dense [n, col_hi-col_lo] (column-major on device)."""
        ncols = self.col_hi - self.col_lo
        buf = np.empty((ncols, self.n), np.float32)
        check(load().hh_mcl_fetch_m1(self._h, ptr(buf)))
        return buf.T

    # -- one mcl() call on a single GPU ----------------------------------------------------------
    def run(self, inflation: float, max_iter: int = 200, pruning: float = 1e-4) -> dict:
        res = MclResult()
        it_nnz = np.zeros(max_iter, np.int64)
        it_prod = np.zeros(max_iter, np.int64)
        it_delta = np.zeros(max_iter, np.float32)
        it_ms = np.zeros(max_iter, np.float32)
        check(load().hh_mcl_run(self._h, float(inflation), int(max_iter), float(pruning), C.byref(res), ptr(it_nnz),
                                ptr(it_prod), ptr(it_delta), ptr(it_ms)))
        r = int(res.rounds)
        self.last = {
            "rounds": r, "converged": bool(res.converged), "nnz": int(res.nnz), "products": int(res.products),
            "bytes": int(res.bytes), "iter_nnz": it_nnz[:r].copy(), "iter_products": it_prod[:r].copy(),
            "iter_delta": it_delta[:r].copy(), "iter_ms": it_ms[:r].copy(),
        }
        return self.last

    def result(self):
        """The matrix the last run / committed step left, canonical CSC on the host."""
        return self._fetch_csc(load().hh_mcl_fetch_result)

    # -- step interface (column shards) -------------------------------------------------------
    def begin(self, inflation: float, pruning: float = 1e-4):
        check(load().hh_mcl_begin(self._h, float(inflation), float(pruning)))
        self.col_lo, self.col_hi = self._own

    def step(self, it: int):
        nnz = C.c_int64()
        prod = C.c_int64()
        delta = C.c_float()
        ms = C.c_float()
        check(load().hh_mcl_step(self._h, int(it), C.byref(nnz), C.byref(prod), C.byref(delta), C.byref(ms)))
        self.last_step_ms = float(ms.value)
        return int(nnz.value), int(prod.value), float(delta.value)

    def pack(self, nnz_owned: int):
        """Owned block of the pending iterate as CUDA tensors (len int32 [ncols], idx int32, val fp32)."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        ncols = self.col_hi - self.col_lo
        ln = torch.empty(ncols, dtype=torch.int32, device=dev)
        idx = torch.empty(max(nnz_owned, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(nnz_owned, 1), dtype=torch.float32, device=dev)
        check(load().hh_mcl_pack(self._h, ptr(ln), ptr(idx), ptr(val)))
        return ln, idx[:nnz_owned], val[:nnz_owned]

    def pack_flat(self, nnz_owned: int, capacity: int):
        """The same three arrays laid out back to back in ONE int32 CUDA tensor of `capacity` words
        ([ncols] lengths, [nnz] row indices, [nnz] fp32 bit patterns): one collective moves a whole block."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        ncols = self.col_hi - self.col_lo
        buf = torch.empty(max(int(capacity), ncols + 2 * nnz_owned, 1), dtype=torch.int32, device=dev)
        base = buf.data_ptr()
        check(load().hh_mcl_pack(self._h, C.c_void_p(base), C.c_void_p(base + 4 * ncols), C.c_void_p(base + 4 * (ncols + nnz_owned))))
        return buf

    def unpack(self, col_lo: int, col_hi: int, ln, idx, val):
        check(load().hh_mcl_unpack(self._h, int(col_lo), int(col_hi), ptr(ln), ptr(idx), ptr(val), int(idx.shape[0])))

    def commit(self):
        check(load().hh_mcl_commit(self._h))

    def set_block(self, col_lo: int, col_hi: int):
        """Columns the following sparse steps compute (reset by begin())."""
        check(load().hh_mcl_set_block(self._h, int(col_lo), int(col_hi)))
        self.col_lo, self.col_hi = int(col_lo), int(col_hi)

    def close(self):
        if self._h:
            load().hh_mcl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interpret_result(result):
    """Attractor rows -> clusters (HapHiC_cluster.py:2065-2095).  ``result`` is a scipy sparse
    matrix; returns a list of index tuples, or None when a node is in two clusters or in none."""
    import scipy.sparse as sp
    r = sp.csr_matrix(result)
    r.eliminate_zeros()
    r.sort_indices()
    n = r.shape[0]
    attractors = np.nonzero(r.diagonal())[0]
    clusters = set()
    for a in attractors.tolist():
        clusters.add(tuple(r.indices[r.indptr[a]:r.indptr[a + 1]].tolist()))
    seen = np.zeros(n, dtype=np.int64)
    for c in clusters:
        np.add.at(seen, list(c), 1)
    if n == 0 or seen.min() != 1 or seen.max() != 1:
        return None
    return list(clusters)


def inflation_values(min_inflation, max_inflation, step):
    """The Decimal sweep of run_mcl_clustering (2139-2141, 2155); ``str(v)`` names the output dirs."""
    start = Decimal(str(min_inflation))
    st = Decimal(str(step))
    end = Decimal(str(max_inflation)) + st
    import math
    n = max(0, math.ceil((end - start) / st))       # numpy.arange length rule, exact in Decimal
    # numpy.arange over Decimals: element 0 is `start` itself ('1.0', not '1.00'), element k is start + k * step
    return [start if k == 0 else start + k * st for k in range(n)]
