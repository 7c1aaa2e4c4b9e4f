#!/usr/bin/env python
"""CPT fitting (bc_fit_counts) on DMV- and Census-shaped tables resident in HBM, next to the numpy restatement of the
reference's counting on the host.

    python tools/fit_bench.py [--out profiles/r1_fit_bench.json]

The table is an ancestral sample of the shipped model (its real skew and structural zeros) at the dataset's real row
count.  roofline: bound = HBM, algorithmic bytes = n_rows x n_cols x 1 B read once (the counters stay on chip).
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


def device_sample(tm, n, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    cols = []
    for v in range(tm.n_nodes):
        t = torch.tensor(np.asarray(tm.cpts[v], dtype=np.float64).reshape(int(tm.card[v]), -1), device="cuda")
        cdf = torch.cumsum(t / t.sum(dim=0, keepdim=True), dim=0).T.contiguous()
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        for a in range(0, n, 1 << 21):  # chunked: the gathered CDF rows are n x card doubles
            b = min(n, a + (1 << 21))
            u = torch.rand(b - a, generator=g, device="cuda", dtype=torch.float64)
            idx = cols[tm.parent[v]][a:b].long() if tm.parent[v] >= 0 else torch.zeros(b - a, dtype=torch.long, device="cuda")
            out[a:b] = torch.clamp((cdf[idx] < u[:, None]).sum(dim=1), max=int(tm.card[v]) - 1).to(torch.uint8)
        cols.append(out)
    return torch.stack(cols, dim=1).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    import torch

    import golden_util as G
    from bayescard_b200 import fit as F
    from oracle import bayescard_oracle as O  # CPU baseline + checker

    peaks = {"hbm_gbs": 6650.0}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    res = []
    for name, n_rows in (("dmv", 11_591_877), ("census", 2_458_285)):
        tm = G.model(name)
        table = device_sample(tm, n_rows, seed=1)
        off, total = F.count_layout(tm.parent, tm.card)
        counts = torch.empty(total + 1, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ts, hs = [], []
        for r in range(args.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            h0 = time.perf_counter()
            F.fit_counts_device(tm.parent, tm.card, table.data_ptr(), n_rows, 1, tm.n_nodes, counts.data_ptr(), 0, st)
            h1 = time.perf_counter()
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
                hs.append((h1 - h0) * 1e3)
        ms = float(np.median(ts))
        # end to end from a pinned host table: H2D + count + D2H of the counters + fp64 normalisation
        h = table.cpu().pin_memory()
        t0 = time.perf_counter()
        d = h.to("cuda", non_blocking=True)
        F.fit_counts_device(tm.parent, tm.card, d.data_ptr(), n_rows, 1, tm.n_nodes, counts.data_ptr(), 0, st)
        c = counts[:total].cpu().numpy()
        cpts = F.counts_to_cpts(tm.parent, tm.card, c)
        e2e_s = time.perf_counter() - t0
        # CPU: numpy restatement of state_counts on a bounded sample of the same table
        sample = h.numpy()[: min(n_rows, 2_000_000)]
        t0 = time.perf_counter()
        want, _ = O.fit_counts(tm.parent, tm.card, sample)
        cpu_s = time.perf_counter() - t0
        _, chk, _ = F.fit_cpts(tm.parent, tm.card, sample, device=0)
        exact = bool(np.array_equal(chk, np.concatenate([w.reshape(-1) for w in want])))
        bytes_alg = n_rows * tm.n_nodes
        rec = {"table": f"{name}-shaped ancestral sample, {n_rows} rows x {tm.n_nodes} uint8 columns", "counters": int(total),
               "kernel_ms": round(ms, 4), "host_enqueue_ms": round(float(np.median(hs)), 4), "rows_per_s": n_rows / (ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": bytes_alg},
               "e2e_from_pinned_host_s": e2e_s, "e2e_rows_per_s": n_rows / e2e_s,
               "cpu_baseline": {"rows_per_s": len(sample) / cpu_s, "cores": 1, "kind": "port",
                                "sample": f"first {len(sample)} rows, oracle.fit_counts (numpy bincount)"},
               "counts_bit_exact_vs_oracle_on_sample": exact,
               "reference_log_seconds": {"dmv": 140, "census": 717}[name]}
        print(json.dumps(rec), flush=True)
        res.append(rec)
        del table, counts
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
