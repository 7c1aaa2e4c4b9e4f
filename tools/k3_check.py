#!/usr/bin/env python
"""K3 (fused tensor-core tree kernel): parity on the golden infer cases and device-resident throughput next to the
specialised kernel, BITS and DENSE_F32 rows (+ fan-out mask on the models that have fan-out columns).

    python tools/k3_check.py [--models dmv,imdb0,...] [--nq 1048576] [--skip-parity]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="dmv,imdb0,imdb1,imdb2,imdb3,imdb4,census")
    ap.add_argument("--nq", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-bench", action="store_true")
    args = ap.parse_args()
    import torch

    import golden_util as G
    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import PredicateCompiler, unpack_ranges
    from bayescard_b200.engine import DeviceModel
    from oracle import bayescard_oracle as O

    FUSED = L.KERNEL_FUSED
    st = torch.cuda.current_stream().cuda_stream
    for name in args.models.split(","):
        tm = G.model(name)
        dm = DeviceModel(tm, device=0, specialize=True)
        rec = {"model": name}
        if not args.skip_parity:
            pc = PredicateCompiler(tm)
            cases = [r for r in G.load("infer_cases.json.gz")[name] if "error" not in r]
            decoded = [({k: list(v) for k, v in r["bins"].items()}, {k: np.asarray(v) for k, v in r["weights"].items()}) for r in cases]
            ref = np.asarray([np.asarray(r["p"]["value"]).reshape(-1)[0] for r in cases])
            for force_dense in (False, True):
                r_idx, r_desc, d_idx, d_desc, mask = pc.pack(decoded, [r["fanout"] for r in cases], force_dense=force_dense)
                got = np.zeros(len(cases))
                try:
                    if len(r_idx):
                        got[r_idx] = dm.run_host(r_desc, L.DESC_BITS, mask[r_idx], FUSED)
                    if len(d_idx):
                        got[d_idx] = dm.run_host(d_desc, L.DESC_DENSE_F32, mask[d_idx], FUSED)
                except L.BayesCardError as e:
                    rec["error"] = str(e)
                    break
                err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
                err[(ref == 0) & (got == 0)] = 0
                rec["parity_dense" if force_dense else "parity_mixed"] = {"n": len(cases), "n_bits": int(len(r_idx)), "max_rel": float(err.max()),
                                                                         "mean_signed_rel": float(np.mean((got - ref)[ref > 0] / ref[ref > 0]))}
        if "error" in rec or args.skip_bench:
            print(json.dumps(rec), flush=True)
            dm.close()
            continue
        nq = args.nq
        kmax = min(14, tm.n_nodes)
        ranges = torch.empty((nq, dm.desc_stride(L.DESC_RANGE_U8)), dtype=torch.uint8, device="cuda")
        dm.gen_range_queries_device(0, 0, nq, 1, kmax, ranges.data_ptr(), st)
        bits = torch.empty((nq, dm.desc_stride(L.DESC_BITS)), dtype=torch.uint8, device="cuda")
        dm.convert_device(ranges.data_ptr(), L.DESC_RANGE_U8, bits.data_ptr(), L.DESC_BITS, nq, st)
        lo_hi = ranges[:, : 2 * tm.n_nodes].reshape(nq, tm.n_nodes, 2).to(torch.int32)
        width = dm.dense_width
        dense = torch.zeros((nq, width), dtype=torch.float32, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(1)
        for v in range(tm.n_nodes):
            c = torch.arange(int(tm.card[v]), device="cuda", dtype=torch.int32)[None, :]
            sel = (c >= lo_hi[:, v, 0:1]) & (c <= lo_hi[:, v, 1:2])
            o = int(dm.dense_offset[v])
            frac = 0.2 + 0.8 * torch.rand((nq, int(tm.card[v])), device="cuda", generator=g)   # fractional n_distinct weights
            dense[:, o: o + int(tm.card[v])] = sel.to(torch.float32) * frac
        fan_nodes = [v for v in range(tm.n_nodes) if tm.infer_names[v] in tm.fanouts]
        mask = None
        if fan_nodes:
            mv = torch.zeros(nq, dtype=torch.int64, device="cuda")
            for v in fan_nodes:
                mv |= (torch.rand(nq, device="cuda", generator=g) < 0.4).to(torch.int64) << v
            mask = mv.to(torch.int32)
        outs = {}
        for label, ptr, fmt, mptr in (("bits", bits.data_ptr(), L.DESC_BITS, 0), ("dense", dense.data_ptr(), L.DESC_DENSE_F32, 0),
                                      ("dense_fan", dense.data_ptr(), L.DESC_DENSE_F32, mask.data_ptr() if mask is not None else -1)):
            if mptr == -1:
                continue
            for kname, kernel in (("spec", L.KERNEL_SPEC), ("k3", FUSED)):
                out = torch.empty(nq, dtype=torch.float32, device="cuda")
                ts = []
                for r in range(args.reps + 2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dm.run_device(ptr, nq, fmt, out.data_ptr(), mask_ptr=mptr, kernel=kernel, stream=st)
                    e1.record()
                    torch.cuda.synchronize()
                    if r >= 2:
                        ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                outs[(label, kname)] = out
                rec[f"{label}_{kname}_ms"] = round(ms, 3)
                rec[f"{label}_{kname}_qps"] = round(nq / (ms * 1e-3) / 1e9, 4)
                rec[f"{label}_{kname}_tflops_dense"] = round(dm.flops_dense * nq / (ms * 1e-3) / 1e12, 2)
            a, b = outs[(label, "spec")], outs[(label, "k3")]
            rec[f"{label}_max_rel_k3_vs_spec"] = float(((a - b).abs() / a.abs().clamp_min(1e-30)).max())
        # oracle spot check of the BITS run
        sub = np.arange(0, nq, max(1, nq // 2000))[:2000]
        lo, hi = unpack_ranges(tm, ranges[torch.as_tensor(sub, device="cuda")].cpu().numpy())
        ref = O.dense_tree(tm, O.range_weights(tm, lo, hi))
        got = outs[("bits", "k3")][torch.as_tensor(sub, device="cuda")].cpu().numpy().astype(np.float64)
        rel = (got - ref) / np.maximum(np.abs(ref), 1e-300)
        rec["bits_k3_vs_fp64_oracle"] = {"max_abs_rel": float(np.abs(rel).max()), "mean_signed_rel": float(rel.mean())}
        rec["flops_dense"] = dm.flops_dense
        print(json.dumps(rec), flush=True)
        dm.close()
        del ranges, bits, dense, lo_hi


if __name__ == "__main__":
    main()
