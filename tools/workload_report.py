#!/usr/bin/env python
"""BASELINE.json configs 1 and 3 on one B200, next to the CPU restatement of the reference.

config 1  shipped DMV / Census models x the shipped SQL workloads (tests/golden/*_workload.json.gz: SQL text,
          true cardinality, the reference's own estimate).  Through the drop-in API exactly as
          ``Testing/BN_testing.py:21-46`` drives the reference: parse, then time ``BN.query`` only.  Reports the
          q-error percentiles (must equal the reference's: SURVEY.md section 8c), the worst relative difference
          to the reference's estimates, p50/p99 latency of the scalar call, the batch call, and the oracle port
          (numpy fp64, one process) on the same parsed queries.
config 3  the five IMDB BNs as a join ensemble: seeded factor lists as SURVEY.md section 8d describes them -- 1-3
          predicated non-fan-out columns (random contiguous bin range, n_distinct weights ~ U(0.2, 1)), 1-3
          fan-out columns, 15 % of the factors with a predicate ON a fan-out column (the predicate wins), random
          ``inverse`` flags -- evaluated with fan-out-weighted expectations (DENSE_F32 rows + fan-out bitmask) and
          combined as ``BN_ensemble.cardinality`` (Models/BN_ensemble_model.py:228-252).  Device-resident and
          host-buffer throughput per BN, parity against the fp64 oracle on a sub-sample.

    python tools/workload_report.py [--out profiles/r1_workload_report.json] [--factors 262144]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def config1(name, reps=3):
    import golden_util as G
    from bayescard_b200.model import Bayescard_BN
    from bayescard_b200.sql_front import parse_query_single_table
    from oracle import bayescard_oracle as O  # checker + CPU baseline only

    bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", name + ".npz"), device=0)
    bn.infer_algo = "exact-jit"
    bn.init_inference_method()
    rows = G.load(f"{name}_workload.json.gz")["queries"]
    t0 = time.perf_counter()
    parsed = [parse_query_single_table(r["sql"], bn) for r in rows]
    parse_s = time.perf_counter() - t0
    for q in parsed[:50]:
        bn.query(q)  # warm-up
    lat, est = [], []
    for q in parsed:
        t = time.perf_counter()
        c = bn.query(q)
        lat.append(time.perf_counter() - t)
        est.append(float(np.asarray(c).reshape(-1)[0]))
    est = np.asarray(est)
    ref = np.asarray([float(np.asarray(r["card"]["value"]).reshape(-1)[0]) for r in rows])
    true = np.asarray([float(r["true"]) for r in rows])
    qerr = np.asarray([O.q_error(e, t) for e, t in zip(est, true)])
    qerr_ref = np.asarray([O.q_error(e, t) for e, t in zip(ref, true)])
    rel = np.abs(est - ref) / np.maximum(np.abs(ref), 1e-300)
    tb = []
    for _ in range(reps):
        t = time.perf_counter()
        batch = bn.query_batch(parsed)
        tb.append(time.perf_counter() - t)
    rel_b = np.abs(batch - ref) / np.maximum(np.abs(ref), 1e-300)
    sqls = [r["sql"] for r in rows]
    bn.query_sql_batch(sqls)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        sql_batch = bn.query_sql_batch(sqls)
        ts.append(time.perf_counter() - t)
    rel_s = np.abs(sql_batch - ref) / np.maximum(np.abs(ref), 1e-300)
    big = sqls * (200000 // len(sqls) + 1)
    t = time.perf_counter()
    bn.query_sql_batch(big)
    big_s = time.perf_counter() - t
    # CPU: the oracle port of Bayescard_BN.query (decode + VariableEliminationJIT.query), one process
    tm = bn.tree
    t = time.perf_counter()
    cpu = [float(np.asarray(O.bn_query(tm, dict(q))).reshape(-1)[0]) for q in parsed]
    cpu_s = time.perf_counter() - t
    bn.close()
    pct = (50, 90, 95, 99, 100)
    return {
        "config": f"{name} shipped Chow-Liu BN x shipped query.sql ({len(rows)} queries)",
        "q_error_percentiles_50_90_95_99_max": [float(np.percentile(qerr, p)) for p in pct],
        "q_error_percentiles_reference": [float(np.percentile(qerr_ref, p)) for p in pct],
        "max_rel_diff_vs_reference_estimates": float(rel.max()),
        "max_rel_diff_batch_vs_reference": float(rel_b.max()),
        "scalar_query_latency_us": {"p50": float(np.percentile(lat, 50) * 1e6), "p99": float(np.percentile(lat, 99) * 1e6),
                                    "mean": float(np.mean(lat) * 1e6), "note": "Bayescard_BN.query: decode + H2D + kernel + D2H, B=1"},
        "scalar_queries_per_s": len(rows) / float(np.sum(lat)),
        "batch_api_s": float(np.median(tb)), "batch_api_queries_per_s": len(rows) / float(np.median(tb)),
        "sql_text_batch_api": {"s": float(np.median(ts)), "queries_per_s": len(rows) / float(np.median(ts)),
                               "max_rel_diff_vs_reference_estimates": float(rel_s.max()),
                               "queries_per_s_at_%d" % len(big): len(big) / big_s, "host_threads": os.cpu_count(),
                               "note": "SQL text -> native parse/decode/pack (bc_sqlc_compile) -> H2D -> kernel -> D2H"},
        "sql_parse_us_per_query": parse_s / len(rows) * 1e6,
        "cpu_oracle_port": {"queries_per_s": len(rows) / cpu_s, "ms_per_query": cpu_s / len(rows) * 1e3, "cores": 1,
                            "max_rel_diff_vs_reference_estimates": float(np.max(np.abs(np.asarray(cpu) - ref) / np.maximum(np.abs(ref), 1e-300)))},
    }


def make_factors(tm, n, seed):
    """Seeded factor batch for one IMDB BN: dense weight rows, fan-out masks, and what the oracle needs."""
    rng = np.random.default_rng(seed)
    nn = tm.n_nodes
    card = tm.card.astype(np.int64)
    fan_nodes = np.asarray([v for v in range(nn) if tm.fan_vector(v) is not None])
    plain = np.asarray([v for v in range(nn) if tm.fan_vector(v) is None])
    pad = -(-card // 4) * 4
    off = np.concatenate([[0], np.cumsum(pad)[:-1]])
    W = np.zeros((n, int(pad.sum())), dtype=np.float32)
    for v in range(nn):
        W[:, off[v]: off[v] + card[v]] = 1.0
    mask = np.zeros((n, (nn + 31) // 32), dtype=np.uint32)

    def constrain(rows_idx, v):
        m = len(rows_idx)
        lo = (rng.random(m) * card[v]).astype(np.int64)
        hi = lo + (rng.random(m) * (card[v] - lo)).astype(np.int64)
        c = np.arange(card[v])[None, :]
        sel = (c >= lo[:, None]) & (c <= hi[:, None])
        w = rng.uniform(0.2, 1.0, size=(m, int(card[v]))).astype(np.float32)
        W[rows_idx, off[v]: off[v] + card[v]] = np.where(sel, w, 0.0)

    # 1-3 predicated non-fan-out columns
    k = rng.integers(1, 4, size=n)
    order = np.argsort(rng.random((n, len(plain))), axis=1)
    for j in range(3):
        rows_idx = np.nonzero(k > j)[0]
        for v in np.unique(plain[order[rows_idx, j]]):
            constrain(rows_idx[plain[order[rows_idx, j]] == v], int(v))
    # 1-3 fan-out columns
    kf = rng.integers(1, 4, size=n)
    forder = np.argsort(rng.random((n, len(fan_nodes))), axis=1)
    fan_pred = rng.random(n) < 0.15
    for j in range(3):
        rows_idx = np.nonzero(kf > j)[0]
        vs = fan_nodes[forder[rows_idx, j]]
        for v in np.unique(vs):
            r = rows_idx[vs == v]
            if j == 0:  # 15 %: a predicate on the (first) fan-out column -- it wins, the mask bit stays clear
                rp = r[fan_pred[r]]
                if len(rp):
                    constrain(rp, int(v))
                r = r[~fan_pred[r]]
            mask[r, v >> 5] |= np.uint32(1 << (v & 31))
    return W, mask, off, card


def oracle_dense(tm, W, mask, off, card):
    from oracle import bayescard_oracle as O

    Wl = []
    for v in range(tm.n_nodes):
        w = W[:, off[v]: off[v] + card[v]].astype(np.float64)
        f = tm.fan_vector(v)
        if f is not None:
            bit = ((mask[:, v >> 5] >> np.uint32(v & 31)) & 1).astype(bool)
            w = np.where(bit[:, None], w * np.asarray(f, dtype=np.float64)[None, :], w)
        Wl.append(w)
    return O.dense_tree(tm, Wl)


def config3(n_factors, reps=5):
    import torch

    import golden_util as G
    from bayescard_b200 import _lib as L
    from bayescard_b200.engine import DeviceModel

    st = torch.cuda.current_stream().cuda_stream
    per_bn, probs, checks = [], [], []
    for i in range(5):
        tm = G.model(f"imdb{i}")
        dm = DeviceModel(tm, device=0, specialize=True)
        W, mask, off, card = make_factors(tm, n_factors, seed=3000 + i)
        assert W.shape[1] * 4 == dm.desc_stride(L.DESC_DENSE_F32)
        d_w = torch.from_numpy(W).cuda()
        d_m = torch.from_numpy(mask.view(np.int32)).cuda()
        out = torch.empty(n_factors, dtype=torch.float32, device="cuda")
        per_kernel = {}
        for kname, kernel in (("specialised", L.KERNEL_SPEC), ("fused_tensor_core", L.KERNEL_FUSED)):
            ts = []
            for r in range(reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dm.run_device(d_w.data_ptr(), n_factors, L.DESC_DENSE_F32, out.data_ptr(), mask_ptr=d_m.data_ptr(), kernel=kernel, stream=st)
                e1.record()
                torch.cuda.synchronize()
                if r >= 2:
                    ts.append(e0.elapsed_time(e1))
            kms = float(np.median(ts))
            per_kernel[kname] = {"device_ms": round(kms, 3), "device_factors_per_s": n_factors / (kms * 1e-3),
                                 "dense_tflops": dm.flops_dense * n_factors / (kms * 1e-3) / 1e12}
        # the library's own choice (AUTO) is what the API below uses
        ts = []
        for r in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dm.run_device(d_w.data_ptr(), n_factors, L.DESC_DENSE_F32, out.data_ptr(), mask_ptr=d_m.data_ptr(), stream=st)
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        got = out.cpu().numpy().astype(np.float64)
        hp = torch.from_numpy(W).pin_memory().numpy()
        hm = torch.from_numpy(mask.view(np.int32)).pin_memory().numpy().view(np.uint32)
        dm.run_host(hp, L.DESC_DENSE_F32, hm)
        t = time.perf_counter()
        host = dm.run_host(hp, L.DESC_DENSE_F32, hm)
        e2e_s = time.perf_counter() - t
        assert np.array_equal(host.astype(np.float64), got)
        # the same factors as weighted runs (WSPARSE): what the public batch API sends over PCIe
        from bayescard_b200.decode import dense_to_wsparse
        row_off, words = dense_to_wsparse(tm, W)
        p_off, p_words = torch.from_numpy(row_off.view(np.int32)).pin_memory().numpy().view(np.uint32), torch.from_numpy(words.view(np.int32)).pin_memory().numpy().view(np.uint32)
        dm.run_wsparse_host(p_off, p_words, hm)
        t = time.perf_counter()
        hostw = dm.run_wsparse_host(p_off, p_words, hm)
        e2e_w_s = time.perf_counter() - t
        assert np.array_equal(hostw.astype(np.float64), got)
        s = min(4096, n_factors)
        ref = oracle_dense(tm, W[:s], mask[:s], off, card)
        rel = float(np.max(np.abs(got[:s] - ref) / np.maximum(np.abs(ref), 1e-300)))
        ms = float(np.median(ts))
        per_bn.append({"bn": f"imdb{i}", "n_nodes": tm.n_nodes, "bytes_per_factor": int(W.shape[1] * 4 + mask.shape[1] * 4 + 4),
                       "kernel": "auto (fused tensor-core kernel K3 on these models)", "per_kernel": per_kernel, "device_ms": round(ms, 3),
                       "flops_dense_per_factor": dm.flops_dense, "dense_tflops": dm.flops_dense * n_factors / (ms * 1e-3) / 1e12,
                       "device_factors_per_s": n_factors / (ms * 1e-3),
                       "hbm_GBps": (W.shape[1] * 4 + mask.shape[1] * 4 + 4) * n_factors / (ms * 1e-3) / 1e9,
                       "e2e_host_dense_rows_factors_per_s": n_factors / e2e_s,
                       "e2e_host_factors_per_s": n_factors / e2e_w_s, "e2e_format": "WSPARSE (weighted runs) + fan-out mask",
                       "e2e_bytes_per_factor": (row_off.nbytes + words.nbytes + hm.nbytes + 4 * n_factors) / n_factors,
                       "max_rel_err_vs_fp64_oracle": rel})
        probs.append(got)
        checks.append(rel)
        dm.close()
        del d_w, d_m, out
    # join queries: one factor of each of 2-3 BNs, random inverse flags, combined as BN_ensemble.cardinality
    rng = np.random.default_rng(77)
    nq = n_factors
    t = time.perf_counter()
    card_est = np.full(nq, 1.0e6)
    dead = np.zeros(nq, dtype=bool)
    for i in range(5):
        use = rng.random(nq) < 0.5
        inv = rng.random(nq) < 0.3
        p = probs[i]
        dead |= use & (p == 0)
        f = np.where(inv, 1.0 / np.maximum(p, 1e-300), p)
        card_est = np.where(use, card_est * f, card_est)
    card_est = np.where(dead | (card_est <= 1), 1.0, card_est)
    combine_s = time.perf_counter() - t
    return {"config": f"IMDB ensemble, 5 shipped BNs, {n_factors} seeded expectation factors per BN (DENSE_F32 + fan-out mask)",
            "per_bn": per_bn, "join_queries_combined": int(nq), "host_combine_s": combine_s,
            "max_rel_err_vs_fp64_oracle": max(checks)}


def config3_job_light(reps=5, replicate=2000):
    """The real workload of BASELINE config 3: the 70 job-light join queries (tests/golden/job_light.json.gz, copied from
    Benchmark/IMDB/job-light.sql) planned by bayescard_b200/joblight.py into the reference's factor-list format and evaluated
    by BN_ensemble: q-errors against the shipped true cardinalities (paper Table 9 next to them), per-query latency of the
    scalar call (the paper reports 5.4 ms per query), batch throughput, and the CPU port on the same factor lists."""
    import copy

    import golden_util as G
    from bayescard_b200.ensemble import BN_ensemble
    from bayescard_b200.joblight import plan_workload
    from bayescard_b200.model import Bayescard_BN
    from oracle import bayescard_oracle as O

    w = G.load("job_light.json.gz")
    sqls = [r["sql"] for r in w["queries"]]
    true = np.asarray([float(r["true"]) for r in w["queries"]])
    bns = {}
    for i in range(5):
        bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", f"imdb{i}.npz"), device=0)
        bn.infer_algo = "exact-jit"
        bn.init_inference_method()
        bns[i] = bn
    ens = BN_ensemble(bns=bns)
    t0 = time.perf_counter()
    tqs = plan_workload(sqls, {i: float(bns[i].nrows) for i in range(5)})
    plan_s = time.perf_counter() - t0
    parsed = ens.parse_query_all(copy.deepcopy(tqs))
    for tq in parsed[:10]:
        ens.cardinality(tq)
    lat, est1 = [], []
    for tq in parsed:
        t = time.perf_counter()
        c = ens.cardinality(tq)
        lat.append(time.perf_counter() - t)
        est1.append(float(np.asarray(c).reshape(-1)[0]))
    ens.cardinality_batch(parsed)
    big = parsed * replicate
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        est = ens.cardinality_batch(big)
        ts.append(time.perf_counter() - t)
    est = est[:len(parsed)]
    # the CPU port (numpy fp64) on the same factor lists
    tms = {i: G.model(f"imdb{i}") for i in range(5)}
    oparsed = O.ensemble_parse_query_all(tms, copy.deepcopy(tqs))
    t = time.perf_counter()
    ref = np.asarray([float(np.asarray(O.ensemble_cardinality(tms, tq)).reshape(-1)[0]) for tq in oparsed])
    cpu_s = time.perf_counter() - t
    qe = np.asarray([O.q_error(e, t_) for e, t_ in zip(est, true)])
    for bn in bns.values():
        bn.close()
    return {"config": "job-light: 70 star joins over the 5 shipped IMDB BNs, planner = bayescard_b200/joblight.py (restates "
                      "Evaluation/parse_query_imdb.py:54-325 for the star; NOT pinned to the reference planner, which cannot run)",
            "q_error_50_90_95_99_100": [float(np.percentile(qe, p)) for p in (50, 90, 95, 99, 100)],
            "paper_table9_q_error_50_90_95_100": w["paper_table9_qerror_50_90_95_100"],
            "max_rel_diff_gpu_vs_cpu_port": float(np.max(np.abs(est - ref) / np.maximum(ref, 1e-300))),
            "factors_per_query_mean": float(np.mean([len(tq) - 1 for tq in parsed])),
            "plan_us_per_query": plan_s / len(sqls) * 1e6,
            "scalar_call_latency_ms_p50_p99": [float(np.percentile(lat, 50) * 1e3), float(np.percentile(lat, 99) * 1e3)],
            "paper_latency_ms": 5.4,
            "batch_queries_per_s": len(big) / float(np.median(ts)), "batch_size": len(big),
            "cpu_port_queries_per_s_1core": len(oparsed) / cpu_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--factors", type=int, default=262144)
    ap.add_argument("--only", default="", choices=["", "config1", "config3"])
    args = ap.parse_args()
    rep = {}
    if args.only != "config3":
        rep["config1"] = [config1("dmv"), config1("census")]
    if args.only != "config1":
        rep["config3_job_light"] = config3_job_light()
        rep["config3"] = config3(args.factors)
    print(json.dumps(rep, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
