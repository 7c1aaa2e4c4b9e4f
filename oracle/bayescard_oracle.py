"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

A CPU (numpy, fp64) restatement of the reference's ``--infer_algo exact-jit`` path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module, and only as the checker / the reported CPU baseline.  The product package
(``bayescard_b200``) never imports it and has no CPU fallback.

Parity is PINNED: ``tests/golden/*`` holds outputs of the unmodified reference (imported from
``/root/reference`` by ``tools/make_golden.py``) and ``tests/test_oracle_vs_golden.py`` checks every
function below against them (DMV 1965 queries, Census 468, seeded IMDB query/expectation sets,
decode quirk cases).

Every function cites the reference lines it follows (paths relative to the reference root).
The model argument ``m`` is duck-typed: any object with the attributes of
``bayescard_b200.loader.TreeModel`` (infer_names, parent, card, cpts, fanouts, encoding, n_in_bin,
mapping, domain, null_values, n_distinct_mapping, attr_type, nrows).
"""
from __future__ import annotations

import ast
import math

import numpy as np

# --------------------------------------------------------------------------------------------
# 1. Variable elimination, pruned to the Steiner tree  (Pgmpy/inference/ExactInference.py)
# --------------------------------------------------------------------------------------------


def _steiner_nodes(m, targets):
    """Nodes on the root->target paths (ExactInference.py:55-75).

    The reference walks the successors of ``self.root`` depth first and unions the path to every
    target it meets; targets outside the root's component are never met and are ignored.
    """
    wanted = set(t for t in targets if t in m._index)
    keep = set()
    for t in wanted:
        v = m._index[t]
        while v >= 0:
            keep.add(v)
            v = int(m.parent[v])
    return keep


def _plan(m, targets):
    """Elimination order and per-node child lists (ExactInference.py:78-109).

    Order: reverse topological, restricted to the Steiner nodes (:96-102).  A node's working
    factors are its own CPD followed by the CPDs of its children that are in the Steiner tree, in
    the order those children were appended, i.e. reverse topological (:104-107).
    """
    keep = _steiner_nodes(m, targets)
    order = [v for v in range(len(m.infer_names) - 1, -1, -1) if v in keep]
    kids = {v: [c for c in order if int(m.parent[c]) == v] for v in order}
    return order, kids


def _run(m, query, n_distinct, fanout_attrs, is_expectation):
    targets = list(query.keys()) + (list(fanout_attrs) if is_expectation else [])
    order, kids = _plan(m, targets)
    msg = {}  # node -> message vector over the parent's states (the reference overwrites cpd.values)
    for i, v in enumerate(order):
        name = m.infer_names[v]
        table = np.asarray(m.cpts[v], dtype=np.float64)
        last = i == len(order) - 1
        if not kids[v]:
            # ---- leaf of the Steiner tree (:128-151 / :207-236)
            if name in query:
                sel = query[name]
                if n_distinct:
                    nd = n_distinct[name]
                    if len(nd) == 1:
                        out = table[sel] * nd[0]                       # :134-135 / :214-216
                        if is_expectation:
                            out = out.reshape(-1)
                    else:
                        out = np.dot(nd, table[sel])                    # :136-137 / :217-218
                else:
                    out = np.sum(table[sel], axis=0)                    # :138-139 / :219-220
                out = out.reshape(-1)
                if last:
                    return out                                          # :141-142 / :222-223
            elif is_expectation and name in fanout_attrs:
                assert not last, "no querying variables"                 # :225
                out = np.dot(np.asarray(m.fanouts[name], dtype=np.float64), table)   # :226-228
            else:
                if last:
                    return 1                                            # :146-147 / :230-231
                out = np.ones(table.shape[-1])                          # :148 / :232
            assert out.ndim == 1
            msg[v] = out
        else:
            # ---- internal node (:152-196 / :237-286)
            if name in query:
                if type(query[name]) != list:
                    assert type(query[name]) == int, f"invalid query {query[name]}"
                    query[name] = [query[name]]                         # :154-156 (mutates, as the reference)
                sel = query[name]
                own = table[sel]                                         # :157
                if n_distinct:
                    own = (own.transpose() * n_distinct[name]).transpose()   # :158-159
                parts = [msg[c][sel] for c in kids[v]]                  # :162-168
            else:
                own = table                                              # :179 / :265
                if is_expectation and name in fanout_attrs:
                    own = (own.transpose() * np.asarray(m.fanouts[name], dtype=np.float64)).transpose()  # :266-267
                parts = [msg[c] for c in kids[v]]                       # :182-186
            for p_ in parts:
                assert p_.ndim == 1
            below = parts[0] if len(parts) == 1 else np.prod(np.stack(parts), axis=0)   # :169-173
            if last:
                return np.dot(own, below)                                # :174-176
            out = np.dot(np.transpose(own), below)                       # :177
            assert out.ndim == 1
            msg[v] = out
    return 0                                                             # :197 / :287


def ve_query(m, query, n_distinct=None):
    """``VariableEliminationJIT.query`` (ExactInference.py:112-197)."""
    return _run(m, query, n_distinct, [], False)


def ve_expectation(m, query, fanout_attrs, n_distinct=None):
    """``VariableEliminationJIT.expectation`` (ExactInference.py:199-287)."""
    return _run(m, query, n_distinct, fanout_attrs, True)


# --------------------------------------------------------------------------------------------
# 2. The dense full-tree form  (equals (1) to <= 1e-15 in fp64)
# --------------------------------------------------------------------------------------------


def dense_weights(m, query, n_distinct, fanout_attrs=()):
    """Per-node dense weight vectors w_v: predicate weights win over fan-out weights
    (ExactInference.py:209,:238 are tested before :224,:266); unconstrained columns are all-ones."""
    W = []
    for v, name in enumerate(m.infer_names):
        k = int(m.card[v])
        if name in query:
            w = np.zeros(k)
            sel = query[name] if isinstance(query[name], list) else [query[name]]
            nd = None if not n_distinct else n_distinct[name]
            for j, b in enumerate(sel):
                # fancy indexing T[sel] with repeated bins sums their contributions
                w[b] += 1.0 if nd is None else (nd[0] if len(nd) == 1 else nd[j])
        elif name in fanout_attrs:
            w = np.asarray(m.fanouts[name], dtype=np.float64).copy()
        else:
            w = np.ones(k)
        W.append(w)
    return W


def dense_tree(m, W, dtype=np.float64):
    """result = sum_x prod_v w_v[x_v] T_v[x_v, x_pa(v)], evaluated leaf -> root.

    ``W`` is a list (per node) of arrays shaped ``[card_v]`` or ``[B, card_v]``; the result is a scalar
    or ``[B]``.  Per edge: lambda_pa *= (w_v * lambda_v) @ T_v.
    """
    n = len(m.infer_names)
    lam = [np.asarray(W[v], dtype=dtype) for v in range(n)]
    lam = [x.copy() for x in lam]
    for v in range(n - 1, 0, -1):
        p = int(m.parent[v])
        t = np.asarray(m.cpts[v], dtype=dtype)
        lam[p] = lam[p] * (lam[v] @ t)
    return lam[0] @ np.asarray(m.cpts[0], dtype=dtype).reshape(-1)


def range_weights(m, lo, hi, fan_mask=None, dtype=np.float64):
    """Dense weights of a batch of range descriptors: w_v[q, c] = [lo <= c <= hi] * (fan_v[c] if bit v)."""
    lo = np.asarray(lo)
    hi = np.asarray(hi)
    W = []
    for v in range(len(m.infer_names)):
        c = np.arange(int(m.card[v]))[None, :]
        w = ((c >= lo[:, v:v + 1]) & (c <= hi[:, v:v + 1])).astype(dtype)
        if fan_mask is not None:
            name = m.infer_names[v]
            f = m.fanouts.get(name)
            if f is not None and len(f) == int(m.card[v]):
                bit = ((np.asarray(fan_mask)[:, v // 32] >> np.uint32(v % 32)) & 1).astype(bool)
                w = np.where(bit[:, None], w * np.asarray(f, dtype=dtype)[None, :], w)
        W.append(w)
    return W


# --------------------------------------------------------------------------------------------
# 3. Predicate decoding  (Models/Bayescard_BN.py, Models/BN_single_model.py)
# --------------------------------------------------------------------------------------------


def encode_values(m, value, col):
    """``apply_encoding_to_value`` (BN_single_model.py:98-117): unknown values become None."""
    enc = m.encoding.get(col) if col in m.encoding else None
    if col not in m.encoding:
        return None
    if type(value) == list:
        return [enc[x] if x in enc else None for x in value]
    return enc[value] if value in enc else None


def ndistinct_of(m, enc_value, value, col):
    """``apply_ndistinct_to_value`` (BN_single_model.py:119-139)."""
    if col not in m.n_in_bin:
        return 1
    table = m.n_in_bin[col]
    if type(enc_value) != list:
        enc_value, value = [enc_value], [value]
    else:
        assert len(enc_value) == len(value), "incorrect number of values"
    out = []
    for j, b in enumerate(enc_value):
        if b not in table:
            out.append(1)
        elif type(table[b]) == int:
            out.append(1 / table[b])
        elif value[j] not in table[b]:
            out.append(1)
        else:
            out.append(table[b][value[j]])
    return np.asarray(out)


def realign(enc_value, n_distinct):
    """``realign`` (Bayescard_BN.py:53-72): drop None, merge duplicate bins capping the weight at 1."""
    if type(enc_value) != list and type(n_distinct) != list:
        return enc_value, n_distinct
    assert len(enc_value) == len(n_distinct)
    bins, wts = [], []
    for j, b in enumerate(enc_value):
        if b is None:
            continue
        if b in bins:
            k = bins.index(b)
            wts[k] = min(wts[k] + n_distinct[j], 1)
        else:
            bins.append(b)
            wts.append(n_distinct[j])
    return bins, wts


def _coverage(l, r, tl, tr):
    """``cal_coverage`` (Bayescard_BN.py:181-195): fraction of the bin (tl, tr] covered by [l, r]."""
    if l >= tr or r <= tl:
        return 0
    if r > tr:
        return 1 if l < tl else (tr - l) / (tr - tl)
    return (r - l) / (tr - tl) if l > tl else (r - tl) / (tr - tl)


def continuous_range_map(m, col, rng):
    """``continuous_range_map`` (Bayescard_BN.py:180-239), quirks included: expansion needs BOTH
    cursors in range (:220-221) and bins come back in discovery order."""
    bins_ = m.mapping[col]
    left, right = rng
    left = -np.inf if left is None else left
    right = np.inf if right is None else right

    def search(i, j):                                                    # :197-208
        if i == j:
            return i
        mid = int(i + (j - i) / 2)
        tl, tr = bins_[mid]
        if left >= tr:
            return search(mid, j)
        if right <= tl:
            return search(i, mid)
        return mid

    start = search(0, len(bins_))
    a, b = start, start + 1
    go_a = go_b = True
    picked, cover = [], []
    while a >= 0 and b < len(bins_) and (go_a or go_b):                  # :220-221
        if go_a:
            c = _coverage(left, right, *bins_[a])
            if c != 0:
                picked.append(a); cover.append(c); a -= 1
            else:
                go_a = False
        if go_b:
            c = _coverage(left, right, *bins_[b])
            if c != 0:
                picked.append(b); cover.append(c); b += 1
            else:
                go_b = False
    return picked, np.asarray(cover)


def query_decoding(m, query, coverage=None, epsilon=0.5):
    """``query_decoding`` (Bayescard_BN.py:279-325).  Mutates ``query`` like the reference."""
    n_distinct = dict()
    for attr in query:
        if m.attr_type[attr] == 'continuous':
            if coverage is None:
                mult = None
                if type(query[attr]) == tuple:
                    l = max(m.domain[attr][0], query[attr][0])          # :289-290
                    r = min(m.domain[attr][1], query[attr][1])
                else:
                    l = query[attr] - epsilon                            # :292-293
                    r = query[attr] + epsilon
                    if attr in m.n_distinct_mapping and query[attr] in m.n_distinct_mapping[attr]:
                        mult = m.n_distinct_mapping[attr][query[attr]]   # :294-296
                if l > r:
                    return None, None                                    # :297-298
                query[attr], n_distinct[attr] = continuous_range_map(m, attr, (l, r))
                if mult is not None:
                    n_distinct[attr] = n_distinct[attr] * mult           # :300-301
            else:
                n_distinct[attr] = coverage[attr]                        # :303
        elif type(query[attr]) == tuple:
            lo, hi = query[attr][0], query[attr][1]
            nv = m.null_values
            skip_null = not (nv is None or len(nv) == 0 or attr not in nv)   # :306
            vals = []
            for val in m.encoding[attr]:
                if skip_null and val == nv[attr]:
                    continue
                if lo <= val <= hi:
                    vals.append(val)                                     # :307-313
            enc = encode_values(m, vals, attr)
            if enc is None or enc == []:
                return None, None                                        # :315-316
            nd = ndistinct_of(m, enc, vals, attr)
            query[attr], n_distinct[attr] = realign(enc, nd)
        else:
            enc = encode_values(m, query[attr], attr)
            if enc is None or enc == []:
                return None, None                                        # :321-322
            nd = ndistinct_of(m, enc, query[attr], attr)
            query[attr], n_distinct[attr] = realign(enc, nd)
    return query, n_distinct


def bn_query(m, query, n_distinct=None, coverage=None, return_prob=False):
    """``Bayescard_BN.query`` exact-jit branch (Bayescard_BN.py:496-533)."""
    if n_distinct is None:
        query, n_distinct = query_decoding(m, query, coverage)
    if query is None:
        return (0, m.nrows) if return_prob else 0                        # :517-521
    p = ve_query(m, query, n_distinct)
    return (p, m.nrows) if return_prob else p * m.nrows                  # :529-533


def bn_expectation(m, query, fanout_attrs, n_distinct=None, coverage=None, return_prob=False):
    """``Bayescard_BN.expectation`` exact-jit branch (Bayescard_BN.py:559-587)."""
    if fanout_attrs is None or len(fanout_attrs) == 0:
        return bn_query(m, query, n_distinct, coverage, return_prob)    # :568-569
    if n_distinct is None:
        query, n_distinct = query_decoding(m, query, coverage)           # :581-582
    e = ve_expectation(m, query, fanout_attrs, n_distinct)
    return (e, m.nrows) if return_prob else e * m.nrows


# --------------------------------------------------------------------------------------------
# 4. Single-table SQL front end  (Evaluation/cardinality_estimation.py:12-119)
# --------------------------------------------------------------------------------------------

_OPS = {'>': np.greater, '<': np.less, '>=': np.greater_equal, '<=': np.less_equal,
        '=': np.equal, '==': np.equal}


def split_predicate(s):
    """``str_pattern_matching`` (:22-58)."""
    if len(s.split(' IN ')) != 1:
        lhs, rhs = s.split(' IN ')[0], s.split(' IN ')[1]
        try:
            value = list(ast.literal_eval(rhs.strip()))
        except Exception:
            value = [t.strip() for t in rhs.strip()[1:][:-1].split(',')]   # :32-35 (keeps '' items)
        return lhs.strip(), 'in', value
    start = end = 0
    for i, ch in enumerate(s):
        if ch in _OPS:
            start = i
            end = i + 1 if (i + 1 < len(s) and s[i + 1] in _OPS) else i
            break
    attr, op, raw = s[:start], s[start:end + 1], s[end + 1:].strip()
    # (:48-49 tries literal_eval(s[1]) first; for any attribute name that is a NameError / not
    #  iterable, so the cascade int -> float -> str below is what runs.)
    try:
        value = list(ast.literal_eval(s[1].strip()))
    except Exception:
        try:
            value = int(raw)
        except Exception:
            try:
                value = float(raw)
            except Exception:
                value = raw
    return attr.strip(), op.strip(), value


def add_predicate(m, table_query, attr, ops, val, epsilon=1e-6):
    """``construct_table_query`` (:60-111)."""
    if m is None or attr not in m.attr_type:
        return None
    if m.attr_type[attr] == 'continuous':
        if ops == ">=":
            dom = (val, np.inf)
        elif ops == ">":
            dom = (val + epsilon, np.inf)
        elif ops == "<=":
            dom = (-np.inf, val)
        elif ops == "<":
            dom = (-np.inf, val - epsilon)
        elif ops in ("=", "=="):
            dom = val
        else:
            assert False, f"operation {ops} is invalid for continous domain"
        if attr in table_query:
            prev = table_query[attr]
            dom = (max(prev[0], dom[0]), min(prev[1], dom[1]))           # :87-91
        table_query[attr] = dom
    else:
        attr_domain = m.domain[attr]
        if type(attr_domain[0]) != str:
            attr_domain = np.asarray(attr_domain)
        if ops == "in":
            assert type(val) == list, "use list for in query"
            dom = val
        elif ops in ("=", "=="):
            dom = val if type(val) == list else [val]
        else:
            if type(val) == list:
                assert len(val) == 1
                val = val[0]
                assert type(val) == int or type(val) == float
            dom = list(attr_domain[_OPS[ops](attr_domain, val)])        # :97-103
        if attr in table_query:
            dom = [x for x in dom if x in table_query[attr]]            # :105-109
        table_query[attr] = dom
    return table_query


def parse_query_single_table(sql, m):
    """``parse_query_single_table`` (:113-119)."""
    where = sql.split(' WHERE ')[-1].strip()
    out = dict()
    for part in where.split(' AND '):
        attr, ops, value = split_predicate(part.strip())
        add_predicate(m, out, attr, ops, value)
    return out


# --------------------------------------------------------------------------------------------
# 5. IMDB ensemble glue  (Models/BN_ensemble_model.py:192-252)
# --------------------------------------------------------------------------------------------


def ensemble_parse_query_all(bns, table_queries):
    """``parse_query_all`` (:192-225): drop a factor equal to its neighbour, pre-decode the rest."""
    out_all = []
    for tq in table_queries:
        out = [tq[0]]
        for q in tq[1:]:
            if type(q["bn_index"]) != int:
                for j in bns:
                    if set(bns[j].table_name) == q["bn_index"]:
                        q["bn_index"] = j
                        break
            assert type(q["bn_index"]) == int, q["bn_index"]
        same = lambda a, b: (a["bn_index"] == b["bn_index"] and a["query"] == b["query"]
                             and a["expectation"] == b["expectation"])
        for i, q in enumerate(tq[1:]):
            ind = i + 1
            if ind + 1 < len(tq) and same(q, tq[ind + 1]):
                continue
            if i > 0 and same(q, tq[i]):
                continue
            new = {"bn_index": q["bn_index"], "inverse": q["inverse"], "expectation": q["expectation"]}
            new["query"], new["n_distinct"] = query_decoding(bns[q["bn_index"]], q["query"])
            out.append(new)
        out_all.append(out)
    return out_all


def ensemble_cardinality(bns, table_query):
    """``cardinality`` (:228-252)."""
    card = table_query[0]
    for q in table_query[1:]:
        bn = bns[q["bn_index"]]
        if len(q["expectation"]) == 0:
            p, _ = bn_query(bn, q["query"], n_distinct=q["n_distinct"], return_prob=True)
        else:
            p, _ = bn_expectation(bn, q["query"], q["expectation"], n_distinct=q["n_distinct"], return_prob=True)
        if p == 0:
            return 1
        card = card * (1 / p) if q["inverse"] else card * p
    return 1 if card <= 1 else card


# --------------------------------------------------------------------------------------------
# 6. Evaluation loop  (Testing/BN_testing.py:21-51)
# --------------------------------------------------------------------------------------------


def q_error(pred, true):
    """q-error exactly as Testing/BN_testing.py:34-44."""
    if pred == 0 and true == 0:
        return 1.0
    if np.isnan(pred) or pred == 0:
        pred = 1
    elif true == 0:
        true = 1
    return max(pred / true, true / pred)


# --------------------------------------------------------------------------------------------
# 6. CPT fitting for a fixed tree  (Pgmpy/estimators/MLE.py:61-104, Pgmpy/estimators/base.py:63-135)
# --------------------------------------------------------------------------------------------


def fit_counts(parent, card, table):
    """``state_counts`` for every node: counts[v][c, p] = #rows with column v == c and column parent(v) == p
    (base.py:101-135; the reference reaches it through a pandas group-by).  Rows holding an id >= card are skipped."""
    table = np.asarray(table).astype(np.int64)
    card = np.asarray(card, dtype=np.int64)
    ok = (table < card[None, :]).all(axis=1) if table.size else np.zeros(0, dtype=bool)
    t = table[ok]
    out = []
    for v in range(len(card)):
        if parent[v] < 0:
            out.append(np.bincount(t[:, v], minlength=int(card[v])).astype(np.uint64))
        else:
            cp = int(card[parent[v]])
            flat = np.bincount(t[:, v] * cp + t[:, parent[v]], minlength=int(card[v]) * cp)
            out.append(flat.reshape(int(card[v]), cp).astype(np.uint64))
    return out, int((~ok).sum())


def fit_cpts(parent, card, table):
    """``MaximumLikelihoodEstimator.estimate_cpd`` (MLE.py:61-104): counts, all-zero columns -> ones (:77-79),
    ``cpd.normalize()`` = values / values.sum(axis=0)."""
    counts, _ = fit_counts(parent, card, table)
    cpts = []
    for v, c in enumerate(counts):
        t = c.astype(np.float64).reshape(int(card[v]), -1)
        t[:, (t == 0).all(axis=0)] = 1.0
        t = t / t.sum(axis=0)
        cpts.append(t.reshape(-1) if parent[v] < 0 else t)
    return cpts
