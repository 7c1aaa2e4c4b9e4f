#!/usr/bin/env python
"""Run specialised-kernel images (cubin / PTX) on the GPU box: parity vs the fp64 oracle + CUDA-event timing.

  python tools/spec_experiment.py census /path/a.cubin /path/b.ptx ... [--batch 1000000]
Prints one line per image: rel-err, ms/launch, q/s.  Development tool (not part of the product).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("images", nargs="+")
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--kmax", type=int, default=14)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch

    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import unpack_ranges
    from bayescard_b200.engine import DeviceModel
    from bayescard_b200.loader import TreeModel
    from oracle import bayescard_oracle as O

    tm = TreeModel.load(os.path.join(ROOT, "tests", "golden", "models", a.model + ".npz"))
    B = a.batch
    st = torch.cuda.current_stream().cuda_stream
    kmax = min(a.kmax, tm.n_nodes)
    base = DeviceModel(tm, device=0, specialize=False)
    NB = 4
    rs, bs = base.desc_stride(L.DESC_RANGE_U8), base.desc_stride(L.DESC_BITS)
    ranges = [torch.empty((B, rs), dtype=torch.uint8, device="cuda") for _ in range(NB)]
    bits = [torch.empty((B, bs), dtype=torch.uint8, device="cuda") for _ in range(NB)]
    for i in range(NB):
        base.gen_range_queries_device(0, i * B, B, 1, kmax, ranges[i].data_ptr(), st)
        base.convert_device(ranges[i].data_ptr(), L.DESC_RANGE_U8, bits[i].data_ptr(), L.DESC_BITS, B, st)
    torch.cuda.synchronize()
    out = torch.empty(B, dtype=torch.float32, device="cuda")
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(B, size=min(B, 4000), replace=False))
    lo, hi = unpack_ranges(tm, ranges[0].cpu().numpy()[idx])
    ref = O.dense_tree(tm, O.range_weights(tm, lo, hi))
    # generic kernel on BITS as a cross-check of the conversion path
    base.run_device(bits[0].data_ptr(), B, L.DESC_BITS, out.data_ptr(), kernel=L.KERNEL_GENERIC, stream=st)
    torch.cuda.synchronize()
    got = out.cpu().numpy()[idx].astype(np.float64)
    print(f"generic/BITS rel-err {np.max(np.abs(got - ref) / np.maximum(ref, 1e-300)):.3e}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for path in a.images:
        dm = DeviceModel(tm, device=0, specialize=False)
        with open(path, "rb") as f:
            dm.load_image(f.read())
        fmt, bufs = L.DESC_BITS, bits
        out.zero_()
        dm.run_device(bufs[0].data_ptr(), B, fmt, out.data_ptr(), kernel=L.KERNEL_SPEC, stream=st)
        torch.cuda.synchronize()
        got = out.cpu().numpy()[idx].astype(np.float64)
        err = np.max(np.abs(got - ref) / np.maximum(ref, 1e-300))
        for k in range(3):
            dm.run_device(bufs[k % NB].data_ptr(), B, fmt, out.data_ptr(), kernel=L.KERNEL_SPEC, stream=st)
        torch.cuda.synchronize()
        e0.record()
        for k in range(a.reps):
            dm.run_device(bufs[k % NB].data_ptr(), B, fmt, out.data_ptr(), kernel=L.KERNEL_SPEC, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"{os.path.basename(path):32s} rel-err {err:.3e}  {ms * 1e3:8.1f} us/launch  {B / ms / 1e6:8.3f} Gq/s")
        dm.close()


if __name__ == "__main__":
    main()
