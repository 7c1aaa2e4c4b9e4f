#!/usr/bin/env python
"""BASELINE.json config 5: large-batch sweep, 1e6-1e9 random range queries on the DMV tree at 1/2/4/8 B200.

Queries are generated ON THE DEVICE from the counter-based RNG keyed by (seed, query index) -- no PCIe on the
path (PCIe cannot feed the kernel's roof) -- in chunks of --chunk queries:
gen (RANGE_U8 rows) -> convert to BITS -> specialised kernel -> fp32 result kept in HBM.  Any query can be
regenerated on the host from its index, so parity is spot-checked on --check sampled indices per batch size
against the fp64 oracle.

Two device times per batch size (CUDA events, max over ranks):
  total_ms   generator + descriptor conversion + inference kernel
  infer_ms   the inference kernel launches only (measured in a second pass over chunks whose BITS rows are
             resident; the rows of 8 chunks rotate so the working set exceeds L2)

    python tools/large_batch_sweep.py [--sizes 1e6,1e7,1e8,1e9] [--model dmv]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/large_batch_sweep.py

With N ranks every batch is split into N contiguous index ranges (bayescard_b200.sharding.rank_range): strong
scaling of a fixed batch, no collective on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e6,1e7,1e8,1e9")
    ap.add_argument("--model", default="dmv")
    ap.add_argument("--chunk", type=int, default=1 << 22)
    ap.add_argument("--kmin", type=int, default=1)
    ap.add_argument("--kmax", type=int, default=5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--check", type=int, default=10000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import unpack_ranges
    from bayescard_b200.engine import DeviceModel, gen_range_queries_host
    from bayescard_b200.loader import TreeModel
    from bayescard_b200.sharding import max_over_ranks, rank_range
    from oracle import bayescard_oracle as O  # checker only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    tm = TreeModel.load(os.path.join(ROOT, "tests", "golden", "models", args.model + ".npz"))
    dm = DeviceModel(tm, device=local, specialize=True)
    if not dm.has_spec:
        raise SystemExit("specialised kernel unavailable: " + str(dm.spec_error))
    kmax = min(args.kmax, tm.n_nodes)
    rstride, bstride = dm.desc_stride(L.DESC_RANGE_U8), dm.desc_stride(L.DESC_BITS)
    chunk = args.chunk
    NROT = 8
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    ranges = torch.empty((chunk, rstride), dtype=torch.uint8, device=dev)
    bits = [torch.empty((chunk, bstride), dtype=torch.uint8, device=dev) for _ in range(NROT)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lines = []
    for size in (int(float(s)) for s in args.sizes.split(",")):
        a, b = rank_range(size, rank, world)
        mine = b - a
        out = torch.empty(max(mine, 1), dtype=torch.float32, device=dev)

        def full_pass():
            for k, q0 in enumerate(range(a, b, chunk)):
                n = min(chunk, b - q0)
                dm.gen_range_queries_device(args.seed, q0, n, args.kmin, kmax, ranges.data_ptr(), st)
                dm.convert_device(ranges.data_ptr(), L.DESC_RANGE_U8, bits[k % NROT].data_ptr(), L.DESC_BITS, n, st)
                dm.run_device(bits[k % NROT].data_ptr(), n, L.DESC_BITS, out.data_ptr() + 4 * (q0 - a), kernel=L.KERNEL_SPEC,
                              stream=st)

        def infer_pass():  # BITS rows of the last <= NROT chunks are resident; same launch sizes as full_pass
            for k, q0 in enumerate(range(a, b, chunk)):
                n = min(chunk, b - q0)
                dm.run_device(bits[k % NROT].data_ptr(), n, L.DESC_BITS, out.data_ptr() + 4 * (q0 - a), kernel=L.KERNEL_SPEC,
                              stream=st)

        full_pass()  # warm-up (also the pass whose results are checked)
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        full_pass()
        e1.record(stream)
        barrier()
        checked = out.clone()
        infer_pass()  # overwrites `out` with rotated rows: timing only
        barrier()
        e1b = torch.cuda.Event(enable_timing=True)
        e1b.record(stream)
        infer_pass()
        e2.record(stream)
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
        infer_ms = max_over_ranks(e1b.elapsed_time(e2), device=dev)
        # ---- oracle spot check on sampled indices of this rank's range
        rel = 0.0
        if args.check and mine:
            rng = np.random.default_rng(size + rank)
            idx = np.unique(rng.integers(a, b, size=min(args.check // world + 1, mine)))
            rows = np.concatenate([gen_range_queries_host(tm, args.seed, int(i), 1, args.kmin, kmax) for i in idx])
            lo, hi = unpack_ranges(tm, rows)
            ref = O.dense_tree(tm, O.range_weights(tm, lo, hi))
            got = checked[torch.from_numpy(idx - a).to(dev)].cpu().numpy().astype(np.float64)
            rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
        rel = max_over_ranks(rel, device=dev)
        if rank == 0:
            line = {"workload": f"{args.model} tree, {size} device-generated range queries, k~U{{{args.kmin}..{kmax}}}, seed {args.seed}",
                    "n_gpus": world, "batch": size, "chunk": chunk,
                    "total_ms": round(total_ms, 3), "qps_total": size / (total_ms * 1e-3),
                    "infer_ms": round(infer_ms, 3), "qps_infer": size / (infer_ms * 1e-3),
                    "checked": int(args.check), "max_rel_err_vs_fp64_oracle": rel}
            print(json.dumps(line), flush=True)
            lines.append(line)
        del out, checked
    if rank == 0 and args.out:
        with open(args.out, "a") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
