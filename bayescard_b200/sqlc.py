"""Batched SQL front end: many SQL texts -> descriptor rows in ONE native call (``bc_sqlc_compile``).

The per-query Python of :mod:`bayescard_b200.sql_front` + :class:`bayescard_b200.decode.PredicateCompiler` mirrors
the reference (``Evaluation/cardinality_estimation.py:22-119`` + ``Models/Bayescard_BN.py:279-325``) and costs
~45 us per query -- four orders of magnitude more than the kernel.  :class:`SqlBatchCompiler` hands the column tables
of a model to the C++ compiler once and then compiles whole batches; the few queries the native code declines
(``SQLC_PYTHON``: predicate shapes the reference raises on, operands with subtle Python parsing) go through the Python
mirror one by one, so the rows are identical either way (``tests/test_sqlc.py``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .decode import PredicateCompiler, _Column
from .loader import TreeModel
from .sql_front import parse_query_single_table


def _is_num(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, (bool, np.bool_))


class _BnView:
    """What parse_query_single_table reads from a BN: attr_type and domain."""

    def __init__(self, tm: TreeModel):
        self.attr_type = tm.attr_type
        self.domain = tm.domain


class SqlBatchCompiler:
    def __init__(self, tm: TreeModel, python_mirror: Optional[PredicateCompiler] = None):
        self.tm = tm
        self.py = python_mirror or PredicateCompiler(tm)
        self._view = _BnView(tm)
        lib = L.lib()
        card = np.ascontiguousarray(tm.card, dtype=np.int32)
        h = C.c_void_p()
        L.check(lib.bc_sqlc_create(tm.n_nodes, card.ctypes.data, C.byref(h)))
        self._h = h
        self.bits_stride = int(lib.bc_sqlc_bits_stride(h))
        self.dense_width = int(lib.bc_sqlc_dense_width(h))
        self.native_columns: List[str] = []
        for name, kind in tm.attr_type.items():
            node = tm._index.get(name, -1)
            if kind == "continuous":
                self._add_continuous(name, node)
            else:
                self._add_categorical(name, node)

    # ------------------------------------------------------------------ column tables
    def _add_categorical(self, name: str, node: int) -> None:
        tm = self.tm
        col: _Column = self.py.cols[name]
        keys = list(col.pair.keys())
        dom = tm.domain.get(name)
        dom = list(dom) if dom is not None else []
        # anything that is neither a plain number nor a str (None, bool, tuple ...) keeps the column on the Python path
        if not all(_is_num(k) or isinstance(k, str) for k in keys) or not all(_is_num(d) or isinstance(d, str) for d in dom) \
                or any(_is_num(k) and np.isnan(float(k)) for k in keys):
            return self._add_python_only(name, node)

        def tables(vals):
            is_str = np.asarray([isinstance(v, str) for v in vals], dtype=np.uint8)
            num = np.asarray([0.0 if isinstance(v, str) else float(v) for v in vals], dtype=np.float64)
            raw = [v.encode("utf-8") if isinstance(v, str) else b"" for v in vals]
            arr = (C.c_char_p * max(1, len(vals)))(*raw) if vals else (C.c_char_p * 1)()
            return is_str, num, arr

        e_str, e_num, e_arr = tables(keys)
        e_bin = np.asarray([int(col.pair[k][0]) for k in keys], dtype=np.int32)
        e_w = np.asarray([float(col.pair[k][1]) for k in keys], dtype=np.float64)
        d_str, d_num, d_arr = tables(dom)
        L.check(L.lib().bc_sqlc_add_categorical(
            self._h, name.encode("utf-8"), node, 1 if col.enc is not None else 0, len(keys),
            e_str.ctypes.data, e_num.ctypes.data, C.cast(e_arr, C.c_void_p), e_bin.ctypes.data, e_w.ctypes.data,
            len(dom), d_str.ctypes.data, d_num.ctypes.data, C.cast(d_arr, C.c_void_p)))
        if col.has_null and _is_num(col.null):   # BN.null_values[name]: skipped by (lo, hi) tuples (Bayescard_BN.py:304-318)
            L.check(L.lib().bc_sqlc_set_null(self._h, name.encode("utf-8"), float(col.null)))
        self.native_columns.append(name)

    def _add_python_only(self, name: str, node: int) -> None:
        """A column whose tables hold values the native compiler does not model (None, bool, tuples ...): registered
        with an empty domain, which makes every predicate on it come back as SQLC_PYTHON."""
        z = np.zeros(1, dtype=np.float64)
        arr = (C.c_char_p * 1)()
        L.check(L.lib().bc_sqlc_add_categorical(self._h, name.encode("utf-8"), node, 0, 0, z.ctypes.data, z.ctypes.data,
                                                C.cast(arr, C.c_void_p), z.ctypes.data, z.ctypes.data, 0, z.ctypes.data,
                                                z.ctypes.data, C.cast(arr, C.c_void_p)))

    def _add_continuous(self, name: str, node: int) -> None:
        col: _Column = self.py.cols[name]
        if col.edges is None or col.domain is None:
            return self._add_python_only(name, node)
        lo = np.asarray([float(e[0]) for e in col.edges], dtype=np.float64)
        hi = np.asarray([float(e[1]) for e in col.edges], dtype=np.float64)
        nd = col.nd_map or {}
        if not all(_is_num(k) for k in nd):
            return self._add_python_only(name, node)
        k = np.asarray([float(x) for x in nd.keys()], dtype=np.float64)
        v = np.asarray([float(x) for x in nd.values()], dtype=np.float64)
        L.check(L.lib().bc_sqlc_add_continuous(self._h, name.encode("utf-8"), node, float(col.domain[0]), float(col.domain[1]),
                                               len(lo), lo.ctypes.data, hi.ctypes.data, len(k), k.ctypes.data, v.ctypes.data))
        self.native_columns.append(name)

    def close(self):
        if getattr(self, "_h", None):
            L.lib().bc_sqlc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ compile
    def compile_native(self, sqls: Sequence[str], dense_capacity: Optional[int] = None):
        """One native pass: ``(kind uint8[n], bits uint8[n, stride], dense f32[k, width], dense_index uint32[k])``."""
        n = len(sqls)
        raw = [s.encode("utf-8") for s in sqls]
        arr = (C.c_char_p * max(1, n))(*raw) if n else (C.c_char_p * 1)()
        kind = np.zeros(n, dtype=np.uint8)
        bits = np.empty((n, self.bits_stride), dtype=np.uint8)
        cap = n if dense_capacity is None else dense_capacity
        dense = np.empty((cap, self.dense_width), dtype=np.float32)
        didx = np.zeros(max(cap, 1), dtype=np.uint32)
        nd = C.c_size_t()
        L.check(L.lib().bc_sqlc_compile(self._h, n, C.cast(arr, C.c_void_p), kind.ctypes.data, bits.ctypes.data,
                                        dense.ctypes.data, cap, didx.ctypes.data, C.byref(nd)))
        return kind, bits, dense[: nd.value], didx[: nd.value]

    def column_index(self, name: str) -> int:
        return int(L.lib().bc_sqlc_column_index(self._h, name.encode("utf-8")))

    def compile_factors(self, ids: np.ndarray, pred_off: np.ndarray, pred_col: np.ndarray, pred_kind: np.ndarray, pred_a: np.ndarray,
                        pred_b: np.ndarray, fan_mask: Optional[np.ndarray], wsparse: bool = False):
        """Factors ``ids`` of a factor table (``bc_joblight_plan``) -> ``(kind uint8[n], bits uint8[n, stride], dense, dense_index)``:
        ``query_decoding`` + row packing of dicts ``{column: scalar | (lo, hi)}`` in one native call.  ``wsparse=True``: the
        factors with fractional weights come back as WSPARSE rows ``(row_off, words)`` in place of the DENSE_F32 rows."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        n = ids.size
        kind = np.zeros(n, dtype=np.uint8)
        bits = np.empty((n, self.bits_stride), dtype=np.uint8)
        didx = np.zeros(max(n, 1), dtype=np.uint32)
        nd = C.c_size_t()
        args = (self._h, n, ids.ctypes.data, pred_off.ctypes.data, pred_col.ctypes.data, pred_kind.ctypes.data, pred_a.ctypes.data,
                pred_b.ctypes.data, fan_mask.ctypes.data if fan_mask is not None else None, kind.ctypes.data, bits.ctypes.data)
        if not wsparse:
            dense = np.empty((n, self.dense_width), dtype=np.float32)
            L.check(L.lib().bc_sqlc_compile_factors(*args, dense.ctypes.data, n, didx.ctypes.data, C.byref(nd), None, None, 0, None))
            return kind, bits, dense[: nd.value], didx[: nd.value]
        cap = max(64, 48 * n)   # words; grown on overflow
        while True:
            ro = np.zeros(n + 1, dtype=np.uint32)
            words = np.empty(cap, dtype=np.uint32)
            nw = C.c_size_t()
            L.check(L.lib().bc_sqlc_compile_factors(*args, None, n, didx.ctypes.data, C.byref(nd), ro.ctypes.data, words.ctypes.data, cap,
                                                    C.byref(nw)))
            if (kind == L.SQLC_OVERFLOW).any():
                cap *= 4
                continue
            return kind, bits, (ro[: nd.value + 1], words[: nw.value]), didx[: nd.value]

    def compile(self, sqls: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Descriptors of a batch: ``(bits_idx, bits_rows, dense_idx, dense_rows, zero_idx)``.

        Queries the native compiler declines are compiled by the Python mirror (and raise what it raises).
        """
        kind, bits, dense, didx = self.compile_native(sqls)
        if (kind == L.SQLC_OVERFLOW).any():  # cannot happen with dense_capacity = n; kept for callers that cap it
            raise L.BayesCardError("dense capacity exceeded")
        bits_idx = np.nonzero(kind == L.SQLC_BITS)[0]
        bits_rows = bits[bits_idx]
        dense_idx = didx.astype(np.int64)
        dense_rows = dense
        zero = list(np.nonzero(kind == L.SQLC_ZERO)[0])
        py_idx = np.nonzero(kind == L.SQLC_PYTHON)[0]
        if len(py_idx):
            eb, ed, ib, idn = [], [], [], []
            for i in py_idx:
                q = parse_query_single_table(sqls[i], self._view)
                b, w = self.py.decode(q)
                if b is None or not any(k in self.tm._index for k in b):
                    zero.append(int(i))
                    continue
                bi, bd, di, dd, _ = self.py.pack([(b, w)])
                if len(bi):
                    eb.append(bd[0]); ib.append(int(i))
                else:
                    ed.append(dd[0]); idn.append(int(i))
            if eb:
                bits_idx = np.concatenate([bits_idx, np.asarray(ib, dtype=np.int64)])
                bits_rows = np.concatenate([bits_rows, np.stack(eb)])
            if ed:
                dense_idx = np.concatenate([dense_idx, np.asarray(idn, dtype=np.int64)])
                dense_rows = np.concatenate([dense_rows, np.stack(ed)])
        return bits_idx, bits_rows, dense_idx, dense_rows, np.asarray(sorted(zero), dtype=np.int64)
