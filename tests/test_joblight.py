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
    for bn in bns.values():
        bn.close()
