#!/usr/bin/env python
"""BASELINE.json config 4: synthetic tree-BN sweep (10-100 columns, domains 10-10k bins) on ONE B200.

For every (n_cols, card) point: build a random recursive tree with Dirichlet CPT columns
(bayescard_b200.synth), upload it, run the batched large-domain path (K2) with the FP32 SIMT per-edge GEMM
(BC_KERNEL_GEMM_SIMT) and with the tcgen05 3xTF32 GEMM (BC_KERNEL_GEMM) on the same resident RANGE_U16 rows,
time both with CUDA events, compare the two result vectors, and (card <= 1000) check a sub-sample against the
fp64 oracle.  Where the model is small enough the warp-per-query kernel K1 and the specialised kernel are timed too.

    python tools/k2_sweep.py [--points 10x100,20x1000,...] [--nq 65536] [--reps 5] [--out profiles/r1_k2_sweep.txt]

The per-edge GEMM flops counted are the ALGORITHMIC ones: 2 * rows * card(v) * card(pa(v)) per internal non-root
edge (leaf edges are prefix-sum differences, O(rows * card_pa)).  The tensor kernel executes 3x that in TF32.
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

DEFAULT_POINTS = "10x10,100x10,10x100,20x100,50x100,100x100,10x1000,20x1000,50x1000,100x1000,10x10000"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default=DEFAULT_POINTS)
    ap.add_argument("--nq", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--oracle-sample", type=int, default=64)
    args = ap.parse_args()

    import torch

    from bayescard_b200 import _lib as L
    from bayescard_b200.engine import DeviceModel
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16, random_range_queries
    from oracle import bayescard_oracle as O  # checker only

    assert torch.cuda.is_available()
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for pt in args.points.split(","):
        n_cols, card = (int(x) for x in pt.split("x"))
        t0 = time.time()
        m = make_tree_model(n_cols, card, seed=n_cols, dtype=np.float32)
        t_build = time.time() - t0
        dm = DeviceModel(m, device=0, specialize=False)
        nq = args.nq
        lo, hi = random_range_queries(m, nq, seed=1, kmax=10)
        desc_h = pack_ranges_u16(lo, hi)
        desc = torch.from_numpy(desc_h.view(np.int16)).cuda()
        out = torch.empty(nq, dtype=torch.float32, device="cuda")
        is_internal = np.zeros(n_cols, dtype=bool)
        is_internal[m.parent[1:]] = True
        gemm_flop_q = sum(2.0 * int(m.card[v]) * int(m.card[m.parent[v]]) for v in range(1, n_cols) if is_internal[v])
        n_int = int(sum(1 for v in range(1, n_cols) if is_internal[v]))
        rec = {"n_cols": n_cols, "card": card, "nq": nq, "internal_edges": n_int, "gemm_flop_per_query": gemm_flop_q,
               "arena_MB": float(sum(c.size for c in m.cpts) * 4 / 1e6), "host_build_s": round(t_build, 1)}
        results = {}
        kernels = [("simt", L.KERNEL_GEMM_SIMT), ("umma", L.KERNEL_GEMM)]
        small = card <= 256 and sum(-(-int(c) // 4) * 4 for c in m.card) * 4 < 40000 and float(rec["arena_MB"]) < 0.2
        if small:
            kernels.append(("k1", L.KERNEL_GENERIC))
        if n_cols <= 128 and card <= 256:   # the fused tensor-core tree kernel (K3): Lambda stays in tensor memory
            kernels.append(("fused", L.KERNEL_FUSED))
        rec["flops_dense_per_query"] = dm.flops_dense
        for name, k in kernels:
            try:
                for _ in range(2):  # warm-up: builds the prefix sums / transposed split CPTs on first use
                    dm.run_device(desc.data_ptr(), nq, L.DESC_RANGE_U16, out.data_ptr(), kernel=k, stream=st)
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dm.run_device(desc.data_ptr(), nq, L.DESC_RANGE_U16, out.data_ptr(), kernel=k, stream=st)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                results[name] = out.cpu().numpy().astype(np.float64)
                rec[name + "_ms"] = round(ms, 3)
                rec[name + "_qps"] = nq / (ms * 1e-3)
                if name == "fused":
                    rec["fused_tflops_dense"] = round(dm.flops_dense * nq / (ms * 1e-3) / 1e12, 2)
                elif name != "k1" and gemm_flop_q:
                    rec[name + "_tflops_alg"] = round(gemm_flop_q * nq / (ms * 1e-3) / 1e12, 2)
            except Exception as e:  # noqa: BLE001
                rec[name + "_error"] = str(e)[:200]
        if "simt" in results and "umma" in results:
            a, b = results["simt"], results["umma"]
            rec["umma_vs_simt_max_rel"] = float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)))
            rec["speedup_umma"] = round(rec["simt_ms"] / rec["umma_ms"], 2)
        if card <= 1000 and args.oracle_sample and results:
            s = min(args.oracle_sample, nq)
            ref = O.dense_tree(m, O.range_weights(m, lo[:s], hi[:s]))
            for name, r in results.items():
                rec[name + "_max_rel_vs_fp64"] = float(np.max(np.abs(r[:s] - ref) / np.maximum(np.abs(ref), 1e-300)))
        dm.close()
        del desc, out
        torch.cuda.empty_cache()
        print(json.dumps(rec), flush=True)
        rows.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
