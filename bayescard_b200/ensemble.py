"""``BN_ensemble``: the IMDB join-cardinality glue around per-table-join BNs.

Follows reference ``Models/BN_ensemble_model.py``: ``parse_query_all`` (``:192-225``) resolves
``bn_index``, drops a factor that repeats its neighbour and pre-decodes the predicates;
``cardinality`` (``:228-252``) multiplies ``join_size`` by ``p`` (or ``1/p`` for ``inverse`` factors),
returns 1 as soon as a factor is 0 and clamps the result to >= 1.

``cardinality_batch`` is new: it gathers the factors of MANY join queries per BN, evaluates each
BN's factors in one CUDA batch, and combines on the host with the same rules.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .model import Bayescard_BN


class BN_ensemble:
    def __init__(self, schema_graph=None, bns: Dict[int, Bayescard_BN] = None):
        self.schema_graph = schema_graph
        self.bns = dict() if bns is None else bns
        self.join_size = dict()

    def add_BN(self, bn: Bayescard_BN):
        self.bns[bn.table_name] = bn

    # ------------------------------------------------------------------ parse_query_all
    def parse_query_all(self, table_queries: Sequence[list]) -> List[list]:
        parsed_all = []
        for tq in table_queries:
            factors = tq[1:]
            for f in factors:
                if type(f["bn_index"]) != int:
                    for j, bn in self.bns.items():
                        if set(bn.table_name) == f["bn_index"]:
                            f["bn_index"] = j
                            break
                assert type(f["bn_index"]) == int, f["bn_index"]

            def same(a, b):
                return (a["bn_index"] == b["bn_index"] and a["query"] == b["query"]
                        and a["expectation"] == b["expectation"])

            parsed = [tq[0]]
            for i, f in enumerate(factors):
                if i + 1 < len(factors) and same(f, factors[i + 1]):
                    continue
                if i > 0 and same(f, factors[i - 1]):
                    continue
                bins, wts = self.bns[f["bn_index"]].query_decoding(f["query"])
                parsed.append({"bn_index": f["bn_index"], "inverse": f["inverse"], "expectation": f["expectation"],
                               "query": bins, "n_distinct": wts})
            parsed_all.append(parsed)
        return parsed_all

    # ------------------------------------------------------------------ cardinality
    _UNDECODABLE = object()  # factor the reference would crash on IF its loop reaches it

    @staticmethod
    def _combine(join_size, factors, probs):
        card = join_size
        for f, p in zip(factors, probs):
            if p is BN_ensemble._UNDECODABLE:
                # Bayescard_BN.expectation has no `query is None` guard (Models/Bayescard_BN.py:581-583);
                # the reference only fails if no earlier factor already returned 1 (:244-245)
                raise AttributeError("'NoneType' object has no attribute 'keys'")
            if p == 0:
                return 1
            card = card * (1 / p) if f["inverse"] else card * p
        return 1 if card <= 1 else card

    def cardinality(self, table_query, sample_size=1000, hard_sample=False):
        probs = []
        for f in table_query[1:]:
            bn = self.bns[f["bn_index"]]
            if len(f["expectation"]) == 0:
                p, _ = bn.query(f["query"], n_distinct=f["n_distinct"], return_prob=True)
            else:
                p, _ = bn.expectation(f["query"], f["expectation"], n_distinct=f["n_distinct"], return_prob=True)
            p = float(np.asarray(p).reshape(-1)[0]) if not isinstance(p, int) else p
            probs.append(p)
            if p == 0:
                return 1
        return self._combine(table_query[0], table_query[1:], probs)

    def cardinality_batch(self, table_queries: Sequence[list]) -> np.ndarray:
        """Cardinalities of many parsed join queries with one device batch per BN."""
        per_bn: Dict[int, list] = {}
        for qi, tq in enumerate(table_queries):
            for fi, f in enumerate(tq[1:]):
                per_bn.setdefault(f["bn_index"], []).append((qi, fi, f))
        probs = [[None] * (len(tq) - 1) for tq in table_queries]
        for j, items in per_bn.items():
            m = self.bns[j]._machine()
            ok = [(qi, fi, f) for qi, fi, f in items if f["query"] is not None]
            for qi, fi, f in items:
                if f["query"] is None:
                    probs[qi][fi] = 0 if len(f["expectation"]) == 0 else self._UNDECODABLE
            if ok:
                res = m.expectation_batch([f["query"] for _, _, f in ok], [list(f["expectation"]) for _, _, f in ok],
                                          [f["n_distinct"] for _, _, f in ok])
                for (qi, fi, _), p in zip(ok, res):
                    probs[qi][fi] = float(p)
        return np.asarray([self._combine(tq[0], tq[1:], probs[qi]) for qi, tq in enumerate(table_queries)],
                          dtype=np.float64)
