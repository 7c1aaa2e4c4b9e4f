"""``Bayescard_BN`` drop-in: same pickle, same ``query(...)`` / ``expectation(...)`` API, GPU underneath.

Mirrors the public surface of the reference class for the exact-jit path
(``Models/Bayescard_BN.py``): ``init_inference_method`` (``:122-178``), ``query_decoding`` (``:279-325``),
``query`` (``:496-556``), ``expectation`` (``:559-621``), ``align_cpds_in_topological`` (``:340-358``) and the
metadata attributes drivers read (``nrows``, ``attr_type``, ``domain``, ``encoding``, ``fanouts``, ...).
Only ``infer_algo='exact-jit'`` exists here; the other algorithms of the reference are out of scope.

New next to the scalar methods: ``query_batch`` / ``expectation_batch`` (one CUDA launch per batch).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .infer import VariableEliminationB200
from .loader import TreeModel, load_model


class Bayescard_BN:
    def __init__(self, tree: TreeModel, device=0, infer_algo: Optional[str] = None, specialize: bool = True,
                 kernel: int = L.KERNEL_AUTO):
        self.tree = tree
        self.device = device
        self.specialize = specialize
        self.kernel = kernel
        # reference attributes (Models/BN_single_model.py:18-41, Models/Bayescard_BN.py:43-51)
        self.table_name = tree.table_name
        self.nrows = tree.nrows
        self.node_names = tree.node_names
        self.structure = tree.structure
        self.attr_type = tree.attr_type
        self.algorithm = tree.algorithm
        self.max_parents = tree.max_parents
        self.n_mcv = tree.n_mcv
        self.n_bins = tree.n_bins
        self.root = tree.root
        self.encoding = tree.encoding
        self.n_in_bin = tree.n_in_bin
        self.mapping = tree.mapping
        self.domain = tree.domain
        self.null_values = tree.null_values
        self.n_distinct_mapping = tree.n_distinct_mapping
        self.fanouts = tree.fanouts
        self.fanout_attr = tree.fanout_attr
        self.fanout_attr_inverse = tree.fanout_attr_inverse
        self.fanout_attr_positive = tree.fanout_attr_positive
        self.infer_algo = infer_algo
        self.infer_machine = None
        self.cpds = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def load(cls, path: str, device=0, **kw) -> "Bayescard_BN":
        """Load a reference pickle (``pickle.dump(Bayescard_BN)``) or the flat ``.npz`` form."""
        return cls(load_model(path), device=device, **kw)

    def __str__(self):
        return f"bn{self.table_name}.{self.algorithm}-{self.max_parents}-{self.root}-{self.n_mcv}-{self.n_bins}"

    # ------------------------------------------------------------------ init
    def align_cpds_in_topological(self):
        t = self.tree
        order = [t.node_names.index(n) for n in t.topo_names]
        return list(t.cpts), order, list(t.topo_names)

    def init_inference_method(self, algorithm: Optional[str] = None):
        if algorithm:
            self.infer_algo = algorithm
        if self.infer_algo is None:
            self.infer_algo = "exact-jit"
        if self.infer_algo != "exact-jit":
            raise NotImplementedError(
                f"infer_algo={self.infer_algo!r}: bayescard_b200 implements the exact-jit path only")
        assert self.algorithm == "chow-liu", "Currently JIT only supports CLT"
        self.cpds, self.topological_order, self.topological_order_node = self.align_cpds_in_topological()
        self.infer_machine = VariableEliminationB200(self.tree, device=self.device, specialize=self.specialize,
                                                     kernel=self.kernel)

    def _machine(self) -> VariableEliminationB200:
        assert self.infer_algo is not None and self.infer_machine is not None, \
            "must call .init_inference_method() first"
        return self.infer_machine

    # ------------------------------------------------------------------ decode
    def query_decoding(self, query, coverage=None, epsilon=0.5):
        return self._machine().compiler.decode(query, coverage, epsilon)

    # ------------------------------------------------------------------ scalar API
    def query(self, query, num_samples=1, n_distinct=None, coverage=None, return_prob=False, sample_size=1000,
              hard_sample=False):
        m = self._machine()
        nrows = self.nrows
        if n_distinct is None:
            query, n_distinct = m.compiler.decode(query, coverage)
        if query is None:
            return (0, nrows) if return_prob else 0
        p = m.query(query, n_distinct)
        return (p, nrows) if return_prob else p * nrows

    def expectation(self, query, fanout_attrs, num_samples=1, n_distinct=None, coverage=None, return_prob=False,
                    sample_size=1000, hard_sample=False):
        if fanout_attrs is None or len(fanout_attrs) == 0:
            return self.query(query, num_samples, n_distinct, coverage, return_prob, sample_size)
        m = self._machine()
        if n_distinct is None:
            query, n_distinct = m.compiler.decode(query, coverage)
        if query is None:
            # the reference does not test for an undecodable predicate here and fails inside
            # VariableEliminationJIT.expectation on None.keys() (Models/Bayescard_BN.py:581-583)
            raise AttributeError("'NoneType' object has no attribute 'keys'")
        e = m.expectation(query, fanout_attrs, n_distinct)
        return (e, self.nrows) if return_prob else e * self.nrows

    # ------------------------------------------------------------------ batch API
    def query_batch(self, queries: Sequence[dict], return_prob: bool = False) -> np.ndarray:
        """Raw predicate dicts -> cardinalities (or probabilities); undecodable queries give 0."""
        return self.expectation_batch(queries, None, return_prob)

    def expectation_batch(self, queries: Sequence[dict], fanout_attrs: Optional[Sequence[Sequence[str]]],
                          return_prob: bool = False) -> np.ndarray:
        m = self._machine()
        decoded, keep = [], []
        for i, q in enumerate(queries):
            b, w = m.compiler.decode(q)
            if b is not None:
                decoded.append((b, w))
                keep.append(i)
        out = np.zeros(len(queries), dtype=np.float64)
        if decoded:
            if fanout_attrs is None:
                res = m.query_batch([d[0] for d in decoded], [d[1] for d in decoded])
            else:
                fans = [list(fanout_attrs[i] or []) for i in keep]
                res = m.expectation_batch([d[0] for d in decoded], fans, [d[1] for d in decoded])
                for j, i in enumerate(keep):  # expectation with no fan-out column is query (:568-569)
                    if not fans[j] and not any(k in self.tree._index for k in decoded[j][0]):
                        res[j] = 0.0
            out[np.asarray(keep)] = res
        return out if return_prob else out * self.nrows

    def query_sql_batch(self, sqls: Sequence[str], return_prob: bool = False) -> np.ndarray:
        """SQL texts (``SELECT COUNT(*) FROM t WHERE ...``, the language of ``parse_query_single_table``) ->
        cardinalities.  Parsing, decoding and descriptor packing of the whole batch happen in one native call
        (``bayescard_b200.sqlc``); the results equal ``[self.query(parse_query_single_table(s, self)) for s in sqls]``."""
        m = self._machine()
        if getattr(self, "_sqlc", None) is None:
            from .sqlc import SqlBatchCompiler

            self._sqlc = SqlBatchCompiler(self.tree, m.compiler)
        bits_idx, bits_rows, dense_idx, dense_rows, _zero = self._sqlc.compile(sqls)
        out = np.zeros(len(sqls), dtype=np.float64)
        if len(bits_idx):
            out[bits_idx] = m.dev.run_host(bits_rows, L.DESC_BITS, None, m.kernel)
        if len(dense_idx):
            out[dense_idx] = m.dev.run_host(dense_rows, L.DESC_DENSE_F32, None, m.kernel)
        return out if return_prob else out * self.nrows

    def close(self):
        if getattr(self, "_sqlc", None) is not None:
            self._sqlc.close()
            self._sqlc = None
        if self.infer_machine is not None:
            self.infer_machine.close()
            self.infer_machine = None


def load_BN_single(path: str, device=0, **kw) -> Bayescard_BN:
    """Name-compatible with reference ``Models/BN_single_model.py:219-223``."""
    return Bayescard_BN.load(path, device=device, **kw)
