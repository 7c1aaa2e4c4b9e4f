#!/usr/bin/env python
"""The same range queries through the two specialised entry points: BITS rows (predicated FMAs) and DENSE_F32 rows
(u = w * lambda, plain FMAs), device resident, CUDA events.  Separates "the dense format is slower" from "the IMDB
workload is different" (profiles/r1_dense_vs_bits.txt).

    python tools/dense_vs_bits.py [--models census,dmv,imdb0,imdb1] [--nq 1048576]
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
    ap.add_argument("--models", default="census,dmv,imdb0,imdb1")
    ap.add_argument("--nq", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    import golden_util as G
    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import unpack_ranges
    from bayescard_b200.engine import DeviceModel

    st = torch.cuda.current_stream().cuda_stream
    for name in args.models.split(","):
        tm = G.model(name)
        dm = DeviceModel(tm, device=0, specialize=True)
        nq = args.nq
        kmax = min(14, tm.n_nodes)
        rstride = dm.desc_stride(L.DESC_RANGE_U8)
        ranges = torch.empty((nq, rstride), dtype=torch.uint8, device="cuda")
        dm.gen_range_queries_device(0, 0, nq, 1, kmax, ranges.data_ptr(), st)
        bits = torch.empty((nq, dm.desc_stride(L.DESC_BITS)), dtype=torch.uint8, device="cuda")
        dm.convert_device(ranges.data_ptr(), L.DESC_RANGE_U8, bits.data_ptr(), L.DESC_BITS, nq, st)
        # DENSE rows with the same 0/1 weights, built on the device with torch (plumbing, untimed)
        lo_hi = ranges[:, : 2 * tm.n_nodes].reshape(nq, tm.n_nodes, 2).to(torch.int32)
        width = dm.dense_width
        dense = torch.zeros((nq, width), dtype=torch.float32, device="cuda")
        for v in range(tm.n_nodes):
            c = torch.arange(int(tm.card[v]), device="cuda", dtype=torch.int32)[None, :]
            sel = (c >= lo_hi[:, v, 0:1]) & (c <= lo_hi[:, v, 1:2])
            o = int(dm.dense_offset[v])
            dense[:, o: o + int(tm.card[v])] = sel.to(torch.float32)
        out_b = torch.empty(nq, dtype=torch.float32, device="cuda")
        out_d = torch.empty(nq, dtype=torch.float32, device="cuda")
        rec = {"model": name, "nq": nq, "bits_bytes_per_query": bits.shape[1], "dense_bytes_per_query": width * 4}
        for label, ptr, fmt, out in (("bits", bits.data_ptr(), L.DESC_BITS, out_b), ("dense", dense.data_ptr(), L.DESC_DENSE_F32, out_d)):
            ts = []
            for r in range(args.reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dm.run_device(ptr, nq, fmt, out.data_ptr(), kernel=L.KERNEL_SPEC, stream=st)
                e1.record()
                torch.cuda.synchronize()
                if r >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            rec[label + "_ms"] = round(ms, 3)
            rec[label + "_qps"] = nq / (ms * 1e-3)
            rec[label + "_frac_fp32_peak_72T"] = round(2 * dm.spec_ffma() * nq / (ms * 1e-3) / 72.3e12, 4)
        rec["max_rel_diff_dense_vs_bits"] = float(((out_b - out_d).abs() / out_b.abs().clamp_min(1e-30)).max())
        print(json.dumps(rec), flush=True)
        dm.close()
        del ranges, bits, dense, lo_hi


if __name__ == "__main__":
    main()
