"""job-light star-join planner for the shipped IMDB ensemble.

The reference plans join queries with DeepDB's ensemble machinery (``Evaluation/parse_query_imdb.py:54-325`` over
``DeepDBUtils/ensemble_compilation``), which cannot run here (``ensemble_loader.pkl`` is missing, spflow / sqlparse are
not installed).  This module restates what that machinery produces for the ONLY shape job-light has -- a star on
``title.id`` over the five two-table models ``title x X`` (``Schemas/imdb/schema.py:59-63``; BN index = relationship
order) -- and emits factor lists in the reference's own format, ``[full_join_size, {"bn_index", "inverse", "query",
"expectation"}, ...]`` (``parse_query_imdb.py:301-323``), so that ``BN_ensemble.parse_query_all`` / ``.cardinality``
consume them unchanged.

For a query over ``title`` and the tables ``a, b, ...`` with conditions ``C_t, C_a, C_b, ...`` ``generate_factors``
(``:54-240``) yields, with two-table models:

    card = |J_a| * E_a[ 1{C_t, C_a, a not null} * prod_{b != a} F_b ]                    (first model, ``:81-88``; the
                                                                                          outgoing fan-outs F_b =
                                                                                          title.mul_b.movie_id are
                                                                                          merged into it, ``:124-130``)
           * prod_{b != a}  P_b(C_b, C_t, b not null) / P_b(C_t, b not null)             (``:187-223``: overlap = title)

``relevant_conditions`` adds the NOT NULL condition of every merged table (``spn_ensemble.py:96-101``);
``factor_refine`` (``:243-254``) drops a nominator / denominator pair that cancels (no condition on b).  Strict
comparisons get ``epsilon = 0.1`` (``prepare_single_query``, ``:19-44``).

NOT PINNED against the reference (it cannot run); validated against the 70 true cardinalities shipped in
``Benchmark/IMDB/job-light.sql`` and the paper's q-error row (Table 9: 1.30 / 3.534 / 4.836 / 19.13 at 50 / 90 / 95 / 100 %;
this planner over the shipped models gives 1.301 / 3.535 / 4.837 / 19.14).
The first model is chosen as ``_greedily_select_first_cardinality_spn`` does with ``rdc_spn_selection=True``
(``DeepDBUtils/ensemble_compilation/spn_ensemble.py:1312-1350``): the candidate vector (sum of the pairwise RDC values of
the conditioned columns the model covers, number of covered tables with conditions, ...) is maximised, the models are
visited in BN-index order and a later one must be strictly better.  The RDC values are the twelve title-x-X entries of the
shipped ``Benchmark/IMDB/pairwise_rdc.pkl`` (data, copied below).
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Sequence, Tuple

ALIASES = {"t": "title", "mc": "movie_companies", "mi": "movie_info", "mi_idx": "movie_info_idx", "mk": "movie_keyword",
           "ci": "cast_info"}
# BN index = order of the relationships in Schemas/imdb/schema.py:59-63 (== the shipped {i}_chow-liu_1.pkl files)
BN_INDEX = {"movie_info_idx": 0, "movie_info": 1, "cast_info": 2, "movie_keyword": 3, "movie_companies": 4}
EPSILON = 0.1

# Benchmark/IMDB/pairwise_rdc.pkl, the pairs inside one two-table model (no entry exists for two columns of one table)
PAIRWISE_RDC = {
    ("movie_info_idx.info_type_id", "title.kind_id"): 0.39070659303540217,
    ("movie_info_idx.info_type_id", "title.production_year"): 0.14469894415051307,
    ("movie_info.info_type_id", "title.kind_id"): 0.5492260005642482,
    ("movie_info.info_type_id", "title.production_year"): 0.2902598439357767,
    ("cast_info.role_id", "title.kind_id"): 0.1293990618864897,
    ("cast_info.role_id", "title.production_year"): 0.16279968513002951,
    ("movie_keyword.keyword_id", "title.kind_id"): 0.5536931277256608,
    ("movie_keyword.keyword_id", "title.production_year"): 0.19072312253292326,
    ("movie_companies.company_id", "title.kind_id"): 0.6260741802831398,
    ("movie_companies.company_type_id", "title.kind_id"): 0.553063881600441,
    ("movie_companies.company_id", "title.production_year"): 0.34941105786412013,
    ("movie_companies.company_type_id", "title.production_year"): 0.257376245147359,
}

_COND = re.compile(r"^\s*(\w+)\.(\w+)\s*(<=|>=|=|<|>)\s*(-?\d+(?:\.\d+)?)\s*$")
_JOIN = re.compile(r"^\s*(\w+)\.(\w+)\s*=\s*(\w+)\.(\w+)\s*$")


def parse_job_light(sql: str) -> Tuple[List[str], Dict[str, List[Tuple[str, str, float]]]]:
    """``SELECT COUNT(*) FROM a x, b y WHERE ...`` -> (joined tables in FROM order without title, {table: [(column, op, value)]})."""
    sql = sql.strip().rstrip(";")
    m = re.match(r"(?is)^select\s+count\(\*\)\s+from\s+(.*?)\s+where\s+(.*)$", sql)
    if not m:
        raise ValueError("not a job-light query: " + sql[:80])
    alias_of = {}
    order = []
    for part in m.group(1).split(","):
        toks = part.split()
        table, alias = toks[0], toks[-1]
        alias_of[alias] = table
        if table != "title":
            if table not in BN_INDEX:
                raise ValueError("table outside the job-light star: " + table)
            order.append(table)
    if "title" not in alias_of.values():
        raise ValueError("job-light queries join through title")
    conds: Dict[str, List[Tuple[str, str, float]]] = {}
    for c in re.split(r"(?i)\s+and\s+", m.group(2)):
        if _JOIN.match(c):
            continue
        mc = _COND.match(c)
        if not mc:
            raise ValueError("unsupported condition: " + c)
        alias, col, op, val = mc.groups()
        conds.setdefault(alias_of[alias], []).append((col, op, float(val)))
    return order, conds


def _table_query(table: str, conds: Sequence[Tuple[str, str, float]]) -> dict:
    """Conditions of one table -> the ``{column: scalar | (lo, hi)}`` dict of ``prepare_single_query``."""
    per_col: Dict[str, List[Tuple[str, float]]] = {}
    for col, op, val in conds:
        per_col.setdefault(col, []).append((op, val))
    q = {}
    for col, ops in per_col.items():
        lo, hi, eq = -math.inf, math.inf, None
        for op, val in ops:
            if op == "=":
                eq = val
            elif op == ">":
                lo = max(lo, val + EPSILON)
            elif op == ">=":
                lo = max(lo, val)
            elif op == "<":
                hi = min(hi, val - EPSILON)
            elif op == "<=":
                hi = min(hi, val)
        q[f"{table}.{col}"] = eq if eq is not None else (lo, hi)
    return q


def _first_table(order: Sequence[str], conds: Dict[str, list]) -> str:
    """``_greedily_select_first_cardinality_spn`` over the five two-table models (start table title or X: same vector)."""
    best, best_vec = None, None
    for table in sorted(order, key=lambda t: BN_INDEX[t]):   # self.spns order
        cols = {f"{t}.{c}" for t in ("title", table) for c, _, _ in conds.get(t, [])}
        rdc = 0.0
        for x in cols:
            for y in cols:
                if x < y:
                    rdc += PAIRWISE_RDC.get((x, y), PAIRWISE_RDC.get((y, x), 0.0))
        n_where = sum(1 for t in ("title", table) if conds.get(t))
        vec = (rdc, n_where, 2, 0)
        if best_vec is None or vec > best_vec:
            best, best_vec = table, vec
    return best


def plan_star_query(sql: str, join_sizes: Dict[int, float]) -> list:
    """One job-light SQL text -> ``[full_join_size, factor, ...]`` in the reference's format."""
    order, conds = parse_job_light(sql)
    if not order:
        raise ValueError("no joined table")
    c_t = _table_query("title", conds.get("title", []))
    first = _first_table(order, conds)
    a = BN_INDEX[first]

    def nn(table):   # the NOT NULL condition of relevant_conditions: the null marker of <table>_nn is 0
        return {f"{table}.{table}_nn": 1}

    q_first = dict(c_t)
    q_first.update(_table_query(first, conds.get(first, [])))
    q_first.update(nn(first))
    others = [t for t in order if t != first]
    factors = [{"bn_index": a, "inverse": False, "query": q_first,
                "expectation": [f"title.mul_{b}.movie_id" for b in others]}]
    for b in others:
        c_b = _table_query(b, conds.get(b, []))
        if not c_b:
            continue   # factor_refine: nominator and denominator cancel
        nom = dict(c_t)
        nom.update(c_b)
        nom.update(nn(b))
        den = dict(c_t)
        den.update(nn(b))
        factors.append({"bn_index": BN_INDEX[b], "inverse": False, "query": nom, "expectation": []})
        factors.append({"bn_index": BN_INDEX[b], "inverse": True, "query": den, "expectation": []})
    return [join_sizes[a]] + factors


def plan_workload(sqls: Sequence[str], join_sizes: Dict[int, float]) -> List[list]:
    return [plan_star_query(s, join_sizes) for s in sqls]
