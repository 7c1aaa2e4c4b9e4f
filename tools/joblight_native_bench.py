#!/usr/bin/env python
"""Config 3 end to end through the NATIVE job-light path: SQL texts of star joins -> C++ planner
(bc_joblight_plan) -> C++ factor decoder / packer (bc_sqlc_compile_factors) -> one device batch per BN and descriptor kind ->
C++ combine, against the Python path (plan_workload -> parse_query_all -> cardinality_batch) on the same queries.

    python tools/joblight_native_bench.py [--queries 330000] [--out profiles/r2_joblight_native.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=330000)
    ap.add_argument("--python-queries", type=int, default=20000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import golden_util as G
    from bayescard_b200.ensemble import BN_ensemble
    from bayescard_b200.joblight import NativeJobLight, plan_workload
    from bayescard_b200.model import Bayescard_BN
    from test_joblight import _fuzz_star_queries, _workload

    bns = {}
    for i in range(5):
        bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", f"imdb{i}.npz"), device=0)
        bn.infer_algo = "exact-jit"
        bn.init_inference_method()
        bns[i] = bn
    ens = BN_ensemble(bns=bns)
    nat = NativeJobLight(ens)
    sqls70, true, _ = _workload()
    sqls = _fuzz_star_queries(args.queries, 5)
    # drop the queries the reference itself fails on (undecodable predicate on the expectation factor: AttributeError)
    from bayescard_b200 import _lib as L
    plan0 = nat.plan(sqls)
    bad = set()
    for b, (ids, kind, bits, dense, didx) in nat.factor_rows(plan0).items():
        z = ids[(kind == L.SQLC_ZERO) & (plan0["factor_fan_mask"][ids] != 0)]
        bad.update((np.searchsorted(plan0["first_factor"], z, side="right") - 1).tolist())
    sqls = [s for i, s in enumerate(sqls) if i not in bad]
    nat.cardinality_sql_batch(sqls[:4096])   # warm-up (K3 plans, pipes)
    nat.cardinality_sql_batch(sqls)
    t = time.perf_counter()
    plan = nat.plan(sqls)
    t_plan = time.perf_counter() - t
    nf = int(plan["factor_bn"].size)
    t = time.perf_counter()
    rows = nat.factor_rows(plan)
    t_rows = time.perf_counter() - t
    phases = {}
    t = time.perf_counter()
    est = nat.cardinality_sql_batch(sqls, timing=phases)
    t_all = time.perf_counter() - t
    # the same from ONE text buffer (what a caller that already holds the workload as text hands over)
    blob, off = nat.join_texts(sqls)
    phases_text = {}
    t = time.perf_counter()
    est_text = nat.cardinality_sql_batch(blob, off, timing=phases_text)
    t_text = time.perf_counter() - t
    assert np.array_equal(est, est_text)
    # the Python path on a subset (it is ~100x slower)
    sub = sqls[: args.python_queries]
    t = time.perf_counter()
    tqs = ens.parse_query_all(plan_workload(sub, nat.join_sizes))
    ref = ens.cardinality_batch(tqs)
    t_py = time.perf_counter() - t
    nf_sub = sum(len(tq) - 1 for tq in tqs)
    rel = float(np.max(np.abs(est[: len(sub)] - ref) / np.maximum(ref, 1e-300)))
    est70 = nat.cardinality_sql_batch(sqls70)
    qe = np.maximum(est70 / true, true / est70)
    rec = {"queries": len(sqls), "factors": nf, "dense_factors": int(sum(len(r[4]) for r in rows.values())),
           "native": {"seconds": t_all, "queries_per_s": len(sqls) / t_all, "factors_per_s": nf / t_all,
                      "plan_seconds": t_plan, "decode_pack_seconds": t_rows, "phases": phases},
           "native_text_buffer": {"seconds": t_text, "queries_per_s": len(sqls) / t_text, "factors_per_s": nf / t_text, "phases": phases_text},
           "python": {"queries": len(sub), "factors": nf_sub, "seconds": t_py, "queries_per_s": len(sub) / t_py, "factors_per_s": nf_sub / t_py},
           "max_rel_diff_native_vs_python": rel,
           "job_light_q_error_50_90_95_100": [float(np.percentile(qe, p)) for p in (50, 90, 95, 100)]}
    print(json.dumps(rec))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rec, f, indent=1)
    nat.close()


if __name__ == "__main__":
    main()
