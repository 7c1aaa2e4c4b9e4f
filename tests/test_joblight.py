"""job-light star-join planner (bayescard_b200/joblight.py).

The reference's planner cannot run here (missing ensemble blob, spflow, sqlparse), so these tests are NOT pinned to its
outputs; they pin the restated planner to what the reference ships and publishes: the 70 true cardinalities of
Benchmark/IMDB/job-light.sql and the paper's q-error row for BayesCard on JOB-light (Table 9: 1.30 / 3.53 / 4.84 / 19.1
at 50 / 90 / 95 / 100 %)."""
import copy

import numpy as np
import pytest

import golden_util as G
from bayescard_b200.joblight import BN_INDEX, parse_job_light, plan_star_query, plan_workload
from oracle import bayescard_oracle as O


def _workload():
    w = G.load("job_light.json.gz")
    return [r["sql"] for r in w["queries"]], np.asarray([float(r["true"]) for r in w["queries"]]), w["paper_table9_qerror_50_90_95_100"]


def _oracle_estimates(tqs):
    bns = {i: G.model(f"imdb{i}") for i in range(5)}
    parsed = O.ensemble_parse_query_all(bns, copy.deepcopy(tqs))
    return np.asarray([float(np.asarray(O.ensemble_cardinality(bns, tq)).reshape(-1)[0]) for tq in parsed])


def test_parse_and_factor_shapes():
    order, conds = parse_job_light("SELECT COUNT(*) FROM movie_companies mc,title t,movie_info_idx mi_idx WHERE t.id=mc.movie_id AND "
                                   "t.id=mi_idx.movie_id AND mi_idx.info_type_id=113 AND mc.company_type_id=2 AND "
                                   "t.production_year>2005 AND t.production_year<2010")
    assert order == ["movie_companies", "movie_info_idx"]
    assert conds["title"] == [("production_year", ">", 2005.0), ("production_year", "<", 2010.0)]
    js = {i: float(G.model(f"imdb{i}").nrows) for i in range(5)}
    tq = plan_star_query("SELECT COUNT(*) FROM movie_companies mc,title t,movie_info_idx mi_idx WHERE t.id=mc.movie_id AND "
                         "t.id=mi_idx.movie_id AND mi_idx.info_type_id=113 AND mc.company_type_id=2 AND t.production_year>2005 "
                         "AND t.production_year<2010", js)
    assert tq[0] == js[BN_INDEX["movie_companies"]]
    first, nom, den = tq[1:]
    assert first["bn_index"] == 4 and first["expectation"] == ["title.mul_movie_info_idx.movie_id"] and not first["inverse"]
    assert first["query"] == {"title.production_year": (2005.1, 2009.9), "movie_companies.company_type_id": 2.0,
                              "movie_companies.movie_companies_nn": 1}
    assert nom["bn_index"] == den["bn_index"] == 0 and den["inverse"] and not nom["inverse"]
    assert den["query"] == {"title.production_year": (2005.1, 2009.9), "movie_info_idx.movie_info_idx_nn": 1}
    # a joined table without a condition contributes only its fan-out (the cancelling pair is dropped, factor_refine)
    tq = plan_star_query("SELECT COUNT(*) FROM title t,cast_info ci,movie_keyword mk WHERE t.id=ci.movie_id AND t.id=mk.movie_id "
                         "AND mk.keyword_id=117", js)
    assert len(tq) == 2 and tq[1]["bn_index"] == 3 and tq[1]["expectation"] == ["title.mul_cast_info.movie_id"]
    # first model: the pairwise-RDC vector of _greedily_select_first_cardinality_spn (kind_id x company_type_id 0.553 beats
    # kind_id x info_type_id 0.391), not the FROM order
    tq = plan_star_query("SELECT COUNT(*) FROM movie_info_idx mi_idx,title t,movie_companies mc WHERE t.id=mi_idx.movie_id AND "
                         "t.id=mc.movie_id AND t.kind_id=1 AND mi_idx.info_type_id=101 AND mc.company_type_id=2", js)
    assert tq[1]["bn_index"] == 4 and tq[2]["bn_index"] == tq[3]["bn_index"] == 0
    with pytest.raises(ValueError):
        parse_job_light("SELECT COUNT(*) FROM title t,name n WHERE t.id=n.id")


def test_operators_and_single_join():
    js = {i: float(G.model(f"imdb{i}").nrows) for i in range(5)}
    tq = plan_star_query("SELECT COUNT(*) FROM title t,cast_info ci WHERE t.id=ci.movie_id AND t.kind_id=1 AND ci.role_id>=2 AND "
                         "ci.role_id<=4 AND t.production_year<=2000;", js)
    assert len(tq) == 2 and tq[0] == js[2] and tq[1]["expectation"] == [] and not tq[1]["inverse"]
    assert tq[1]["query"] == {"title.kind_id": 1.0, "title.production_year": (-np.inf, 2000.0), "cast_info.role_id": (2.0, 4.0),
                              "cast_info.cast_info_nn": 1}
    # the planner's factor lists go through the reference-format glue unchanged (CPU port of parse_query_all / cardinality)
    est = _oracle_estimates([tq])
    assert np.isfinite(est[0]) and est[0] >= 1.0
    # conditions only on title: every joined table contributes its fan-out, no nominator / denominator pair survives
    tq = plan_star_query("SELECT COUNT(*) FROM title t,movie_keyword mk,movie_info mi WHERE t.id=mk.movie_id AND t.id=mi.movie_id "
                         "AND t.production_year>1990", js)
    assert len(tq) == 2 and tq[1]["bn_index"] in (1, 3) and len(tq[1]["expectation"]) == 1


def test_q_errors_match_the_published_row():
    sqls, true, paper = _workload()
    js = {i: float(G.model(f"imdb{i}").nrows) for i in range(5)}
    est = _oracle_estimates(plan_workload(sqls, js))
    qe = np.asarray([O.q_error(e, t) for e, t in zip(est, true)])
    got = [float(np.percentile(qe, p)) for p in (50, 90, 95, 100)]
    # measured here: 1.301 / 3.535 / 4.837 / 19.14 against 1.30 / 3.534 / 4.836 / 19.13 published: the row is reproduced to
    # the digits the paper prints (first model chosen by the shipped pairwise RDC values, as the reference does)
    for g, p in zip(got, paper):
        assert abs(g - p) / p < 0.005, (got, paper)


@pytest.mark.gpu
def test_job_light_on_the_gpu_equals_the_oracle():
    from bayescard_b200.ensemble import BN_ensemble
    from bayescard_b200.model import Bayescard_BN
    import os

    sqls, true, _ = _workload()
    bns = {}
    for i in range(5):
        bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", f"imdb{i}.npz"), device=0)
        bn.infer_algo = "exact-jit"
        bn.init_inference_method()
        bns[i] = bn
    ens = BN_ensemble(bns=bns)
    tqs = plan_workload(sqls, {i: float(bns[i].nrows) for i in range(5)})
    ref = _oracle_estimates(tqs)
    parsed = ens.parse_query_all(copy.deepcopy(tqs))
    batch = ens.cardinality_batch(parsed)
    one = np.asarray([float(np.asarray(ens.cardinality(tq)).reshape(-1)[0]) for tq in parsed])
    assert np.max(np.abs(batch - ref) / np.maximum(ref, 1e-300)) < 2e-5   # up to 7 fp32 factors multiplied per query
    assert np.max(np.abs(one - batch) / np.maximum(batch, 1e-300)) < 2e-5
    qe = np.asarray([O.q_error(e, t) for e, t in zip(batch, true)])
    assert abs(float(np.percentile(qe, 50)) - 1.301) < 0.01 and abs(float(qe.max()) - 19.14) < 0.05
    # the native path (C++ planner + factor compiler + combine): SQL texts in, cardinalities out; the 70 q-errors are unchanged
    from bayescard_b200.joblight import NativeJobLight

    nat = NativeJobLight(ens)
    native = nat.cardinality_sql_batch(sqls)
    assert np.max(np.abs(native - batch) / np.maximum(batch, 1e-300)) < 1e-6
    qe_n = np.asarray([O.q_error(e, t) for e, t in zip(native, true)])
    assert np.allclose(np.percentile(qe_n, [50, 90, 95, 100]), np.percentile(qe, [50, 90, 95, 100]), rtol=1e-6)
    fuzz = _fuzz_star_queries(3000, 11) + ["SELECT COUNT(*) FROM title t,cast_info ci WHERE t.id=ci.movie_id AND t.kind_id=1;"]
    good, want, bad = [], [], []
    for s_, tq in zip(fuzz, ens.parse_query_all(plan_workload(fuzz, nat.join_sizes))):
        try:   # an undecodable predicate on the expectation factor: the reference (and the mirror) fail with AttributeError
            want.append(float(np.asarray(ens.cardinality(tq)).reshape(-1)[0]))
            good.append(s_)
        except AttributeError:
            bad.append(s_)
    want = np.asarray(want)
    assert len(good) > 1000 and np.max(np.abs(nat.cardinality_sql_batch(good) - want) / np.maximum(want, 1e-300)) < 1e-5
    if bad:
        with pytest.raises(AttributeError):
            nat.cardinality_sql_batch(bad[:3])
    nat.close()
    for bn in bns.values():
        bn.close()


# ------------------------------------------------------------------------------------ native planner + factor compiler
def _fuzz_star_queries(n, seed):
    """Synthetic job-light-shaped queries: random subsets of the five tables, random conditions on the columns job-light uses."""
    rng = np.random.default_rng(seed)
    cols = {"title": [("kind_id", 1, 7), ("production_year", 1880, 2019)], "movie_companies": [("company_id", 1, 230000), ("company_type_id", 1, 2)],
            "movie_info": [("info_type_id", 1, 110)], "movie_info_idx": [("info_type_id", 99, 113)], "movie_keyword": [("keyword_id", 1, 134000)],
            "cast_info": [("role_id", 1, 11)]}
    alias = {"title": "t", "movie_companies": "mc", "movie_info": "mi", "movie_info_idx": "mi_idx", "movie_keyword": "mk", "cast_info": "ci"}
    out = []
    for _ in range(n):
        tabs = list(rng.choice(list(BN_INDEX), size=int(rng.integers(1, 5)), replace=False))
        frm = ["title t"] + [f"{t} {alias[t]}" for t in tabs]
        rng.shuffle(frm)
        conds = [f"t.id={alias[t]}.movie_id" for t in tabs]
        for t in ["title"] + tabs:
            for c, lo, hi in cols[t]:
                for _ in range(int(rng.integers(0, 3))):
                    op = str(rng.choice(["=", "<", ">", "<=", ">="]))
                    conds.append(f"{alias[t]}.{c}{op}{int(rng.integers(lo, hi + 1))}")
        out.append("SELECT COUNT(*) FROM " + ",".join(frm) + " WHERE " + " AND ".join(conds))
    return out


def _host_ensemble():
    """The five IMDB BNs as host-only models (no GPU): enough for the planner and the factor compiler."""
    from bayescard_b200.ensemble import BN_ensemble
    from bayescard_b200.joblight import NativeJobLight
    from bayescard_b200.model import Bayescard_BN

    ens = BN_ensemble()
    for i in range(5):
        bn = Bayescard_BN(G.model(f"imdb{i}"), device=-1, infer_algo="exact-jit")
        bn.init_inference_method()
        ens.bns[i] = bn
    return ens, NativeJobLight(ens)


def test_native_planner_and_factor_rows_equal_the_python_mirror():
    """bc_joblight_plan + bc_sqlc_compile_factors (C++) against plan_star_query + query_decoding + PredicateCompiler.pack (the
    Python mirror): same factors (BN, inverse, fan-out columns, predicates in the same order with the same values) and the same
    descriptor rows, on the 70 shipped job-light queries and 1 500 fuzzed star queries."""
    from bayescard_b200 import _lib as L

    sqls, _, _ = _workload()
    sqls = sqls + _fuzz_star_queries(1500, 7)
    ens, nat = _host_ensemble()
    js = nat.join_sizes
    plan = nat.plan(sqls)
    assert not plan["status"].any()
    rows = nat.factor_rows(plan)
    where = {}
    for b, (ids, kind, bits, dense, didx) in rows.items():
        dpos = {int(i): k for k, i in enumerate(didx)}
        for j, f in enumerate(ids):
            where[int(f)] = (b, int(kind[j]), bits[j], dense[dpos[j]] if j in dpos else None)
    n_dense = n_zero = 0
    for q, sql in enumerate(sqls):
        tq = plan_star_query(sql, js)
        f0, f1 = int(plan["first_factor"][q]), int(plan["first_factor"][q + 1])
        assert plan["join_size"][q] == tq[0] and f1 - f0 == len(tq) - 1, sql
        for f, fac in zip(range(f0, f1), tq[1:]):
            b = int(plan["factor_bn"][f])
            assert b == fac["bn_index"] and bool(plan["factor_inverse"][f]) == fac["inverse"], sql
            tm = ens.bns[b].tree
            want_mask = 0
            for name in fac["expectation"]:
                want_mask |= 1 << tm._index[name]
            assert int(plan["factor_fan_mask"][f]) == want_mask, sql
            p0, p1 = int(plan["pred_off"][f]), int(plan["pred_off"][f + 1])
            got = []
            for p in range(p0, p1):
                name = nat.sqlc[b].py.cols and [n for n in tm.attr_type if nat.sqlc[b].column_index(n) == int(plan["pred_col"][p])][0]
                val = float(plan["pred_a"][p]) if plan["pred_kind"][p] == 0 else (float(plan["pred_a"][p]), float(plan["pred_b"][p]))
                got.append((name, val))
            want = [(k, float(v) if not isinstance(v, tuple) else (float(v[0]), float(v[1]))) for k, v in fac["query"].items()]
            assert got == want, (sql, got, want)
            # rows: the Python mirror decodes and packs the same dict
            bn_, kind, bits_row, dense_row = where[f]
            m = ens.bns[b]._machine()
            dq, dw = m.compiler.decode(dict(fac["query"]))
            if dq is None:
                assert kind == L.SQLC_ZERO, sql
                n_zero += 1
                continue
            bi, bd, di, dd, mask = m.compiler.pack([(dq, dw)], [list(fac["expectation"])])
            if len(bi):
                assert kind == L.SQLC_BITS and np.array_equal(bits_row, np.asarray(bd[0]).view(np.uint8).reshape(-1)), sql
            else:
                assert kind == L.SQLC_DENSE and np.array_equal(dense_row, dd[0]), sql
                n_dense += 1
    assert n_dense > 20 and n_zero > 0   # fractional weights (continuous columns) and undecodable predicates both occur
    # the same factors as WSPARSE rows straight from the native compiler == the Python packer on the DENSE rows
    from bayescard_b200.decode import dense_to_wsparse

    rows_ws = nat.factor_rows(plan, wsparse=True)
    for b, (ids, kind, bits, dense, didx) in rows.items():
        ids2, kind2, bits2, (ro, words), didx2 = rows_ws[b]
        assert np.array_equal(kind, kind2) and np.array_equal(didx, didx2) and np.array_equal(bits[kind == L.SQLC_BITS], bits2[kind2 == L.SQLC_BITS])
        ro_py, words_py = dense_to_wsparse(ens.bns[b].tree, dense)
        assert np.array_equal(ro, ro_py) and np.array_equal(words, words_py)
    # what the native planner declines goes to the mirror
    bad = nat.plan(["SELECT COUNT(*) FROM title t,name n WHERE t.id=n.id", "SELECT 1", sqls[0]])
    assert bad["status"].tolist() == [1, 1, 0]
    # the three ways in -- a list of strings, one text buffer + offsets, the char* array of bc_joblight_plan -- give one factor table
    import ctypes as C
    blob, off = nat.join_texts(sqls)
    plan_b = nat.plan(blob, off)
    raw = [x.encode() for x in sqls]
    arr = (C.c_char_p * len(raw))(*raw)
    n = len(sqls)
    st, jn, ff = np.zeros(n, np.uint8), np.zeros(n), np.zeros(n + 1, np.uint32)
    nf, npred = plan["factor_bn"].size, plan["pred_col"].size
    fb, fi, fm, po = np.zeros(nf, np.int32), np.zeros(nf, np.uint8), np.zeros(nf, np.uint32), np.zeros(nf + 1, np.uint32)
    pc, pk, pa, pb = np.zeros(npred, np.int32), np.zeros(npred, np.uint8), np.zeros(npred), np.zeros(npred)
    gf, gp = C.c_size_t(), C.c_size_t()
    L.check(L.lib().bc_joblight_plan(nat._h, n, C.cast(arr, C.c_void_p), st.ctypes.data, jn.ctypes.data, ff.ctypes.data, nf, fb.ctypes.data,
                                     fi.ctypes.data, fm.ctypes.data, po.ctypes.data, npred, pc.ctypes.data, pk.ctypes.data, pa.ctypes.data,
                                     pb.ctypes.data, C.byref(gf), C.byref(gp)))
    plan_c = {"status": st, "join_size": jn, "first_factor": ff, "factor_bn": fb, "factor_inverse": fi, "factor_fan_mask": fm, "pred_off": po,
              "pred_col": pc, "pred_kind": pk, "pred_a": pa, "pred_b": pb}
    for key, want in plan_c.items():
        assert np.array_equal(plan[key], want) and np.array_equal(plan_b[key], want), key
    assert (gf.value, gp.value) == (nf, npred)
    # non-ASCII text: byte offsets, not character counts; such a text is simply not a job-light query
    odd = ["SELECT COUNT(*) FROM title t WHERE t.kind_id=1 AND t.note='caf\u00e9 \u4e2d\u6587'", sqls[0], "\u00e9", sqls[1]]
    blob2, off2 = nat.join_texts(odd)
    assert [blob2[int(off2[i]):int(off2[i + 1])].decode("utf-8").rstrip("\n") for i in range(4)] == odd
    p_odd = nat.plan(odd)
    assert p_odd["status"].tolist() == [1, 0, 1, 0]
    for key in ("factor_bn", "factor_inverse", "factor_fan_mask"):
        assert np.array_equal(p_odd[key], np.concatenate([plan[key][int(plan["first_factor"][q]):int(plan["first_factor"][q + 1])] for q in (0, 1)]))
    assert nat.plan([])["n_queries"] == 0 and nat.plan(b"", np.zeros(1, dtype=np.uint64))["factor_bn"].size == 0
    # combine: BN_ensemble.cardinality's rules
    first = np.asarray([0, 2, 4, 5], dtype=np.uint32)
    inv = np.asarray([0, 1, 0, 0, 0], dtype=np.uint8)
    prob = np.asarray([0.5, 0.25, 0.0, 0.5, 1e-9], dtype=np.float64)
    out = np.zeros(3)
    L.check(L.lib().bc_joblight_combine(3, None, np.asarray([100.0, 100.0, 100.0]).ctypes.data, first.ctypes.data, inv.ctypes.data,
                                        prob.ctypes.data, out.ctypes.data))
    assert out.tolist() == [200.0, 1.0, 1.0]
    nat.close()
