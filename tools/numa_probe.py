#!/usr/bin/env python
"""Host-side probe for the end-to-end leg: NUMA placement of pinned buffers against H2D bandwidth.

Prints the GPU's local CPU list, the process affinity, and the H2D rate of a 16 MB pinned copy with the buffer allocated (and
first touched) under each NUMA node's CPU set.  The e2e number of bench.py varies 2x between boxes of the pool; this shows why.
"""
import glob
import os
import subprocess
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def main():
    dev = torch.cuda.current_device()
    bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    print("affinity now:", sorted(os.sched_getaffinity(0))[:8], "... n =", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:1500])
    except Exception as e:
        print("topo failed", e)
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes:", [os.path.basename(n) for n in nodes])
    node_cpus = {os.path.basename(n): cpulist(open(n + "/cpulist").read()) for n in nodes}
    q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev)], capture_output=True, text=True).stdout.strip()
    sysfs = "/sys/bus/pci/devices/" + q.lower().replace("00000000:", "0000:")
    for f in ("local_cpulist", "numa_node"):
        try:
            print(f, open(sysfs + "/" + f).read().strip())
        except Exception as e:
            print(f, "unreadable", e)
    allowed = os.sched_getaffinity(0)
    nbytes = 16 << 20
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for name, cpus in list(node_cpus.items()) + [("all", sorted(allowed))]:
        use = set(cpus) & allowed
        if not use:
            print(name, "no allowed cpus")
            continue
        os.sched_setaffinity(0, use)
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        torch.cuda.synchronize()
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(20):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 20
        # latency of a small launch + sync from this cpu set
        t = time.perf_counter()
        for _ in range(200):
            d[:4].zero_()
            torch.cuda.synchronize()
        lat = (time.perf_counter() - t) / 200
        print(f"{name}: {len(use)} cpus, H2D {nbytes / dt / 1e9:.1f} GB/s, launch+sync {lat * 1e6:.1f} us")
        os.sched_setaffinity(0, allowed)


if __name__ == "__main__":
    main()
