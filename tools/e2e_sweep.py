#!/usr/bin/env python
"""Sweep the chunk size of the SPARSE host pipeline (BC_SPARSE_CHUNK) on the bench workload. Dev tool."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bayescard_b200 import _lib as L
from bayescard_b200.engine import DeviceModel
from bayescard_b200.loader import TreeModel

name = sys.argv[1] if len(sys.argv) > 1 else "census"
tm = TreeModel.load(os.path.join(ROOT, "tests", "golden", "models", name + ".npz"))
dm = DeviceModel(tm, device=0, specialize=True)
B = 1_000_000
ro, en = dm.gen_sparse_queries_host(0, 0, B, 1, 14)
h_off = torch.from_numpy(ro.view(np.int32)).pin_memory().numpy().view(np.uint32)
h_ent = torch.from_numpy(en.view(np.int32)).pin_memory().numpy().view(np.uint32)
h_out = torch.empty(B, dtype=torch.float32).pin_memory().numpy()
print(f"{name}: {h_off.nbytes + h_ent.nbytes} B in, {h_out.nbytes} B out per step")
for chunk in (16384, 32768, 65536, 131072, 262144, 524288, 1048576):
    os.environ["BC_SPARSE_CHUNK"] = str(chunk)
    for _ in range(3):
        dm.run_sparse_host(h_off, h_ent, None, L.KERNEL_AUTO, out=h_out)
    t = time.perf_counter()
    for _ in range(10):
        dm.run_sparse_host(h_off, h_ent, None, L.KERNEL_AUTO, out=h_out)
    dt = (time.perf_counter() - t) / 10
    print(f"chunk {chunk:8d}: {dt * 1e3:7.3f} ms/step  {B / dt / 1e9:6.3f} Gq/s  H2D {(h_off.nbytes + h_ent.nbytes) / dt / 1e9:5.1f} GB/s")
# raw H2D copy speed of the same bytes for reference
d = torch.empty(h_ent.nbytes, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(h_ent.view(np.uint8))
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    d.copy_(src, non_blocking=True)
torch.cuda.synchronize()
print(f"raw pinned H2D: {h_ent.nbytes * 10 / (time.perf_counter() - t) / 1e9:.1f} GB/s")
