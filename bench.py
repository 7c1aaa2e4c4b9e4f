#!/usr/bin/env python
"""bench.py -- queries/sec of exact tree inference (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic range-predicate queries:
BASELINE.json configs[1], "Census Chow-Liu BN, batch of 1M synthetic multi-column range queries"
(seeded counter-based generator, k ~ U{1..14} constrained columns).  With N > 1
every rank runs the same per-GPU batch on its own replica of the CPT arena (weak scaling; the path
has no exchange step, so no collective runs inside the timed region -- NCCL is used for the barrier
and the max-over-ranks reduction only).

  value      whole-job q/s, BITS descriptors (one bit per column state) already resident in HBM,
             CUDA events around K steps
  e2e        same metric through the host-buffer C-ABI call (bc_query_batch_sparse_host): pinned host
             SPARSE (CSR) queries -> H2D -> expand to BITS -> kernel -> D2H of the fp32 results, all
             inside the timed region
  roofline   dominant kernel vs the FP32 FFMA peak measured in this same run (the kernel keeps the
             CPTs in the instruction stream, so HBM is not its bound; the HBM view is reported too)
  cpu_baseline  the oracle port of the reference algorithm (numpy fp64, per query, as the reference
             loops) timed on this host on a bounded sample -- a reported baseline, not the target
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec exact tree inference"
UNIT = "queries/s"
KMIN, KMAX = 1, 14
SEED = 0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="census")
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "spec"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


def load_tree(name):
    from bayescard_b200.loader import TreeModel

    return TreeModel.load(os.path.join(ROOT, "tests", "golden", "models", name + ".npz"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------ oracle legs
def queries_as_dicts(tm, desc):
    """RANGE_U8 rows -> the (bins, n_distinct) dicts VariableEliminationJIT.query takes."""
    from bayescard_b200.decode import unpack_ranges

    lo, hi = unpack_ranges(tm, desc)
    out = []
    for i in range(lo.shape[0]):
        q = {}
        for v in range(tm.n_nodes):
            if lo[i, v] > 0 or hi[i, v] < tm.card[v] - 1:
                q[tm.infer_names[v]] = list(range(int(lo[i, v]), int(hi[i, v]) + 1))
        out.append(q)
    return out


def _oracle_chunk(args):
    name, first, n = args
    from bayescard_b200.engine import gen_range_queries_host
    from oracle import bayescard_oracle as O

    tm = load_tree(name)
    qs = queries_as_dicts(tm, gen_range_queries_host(tm, SEED, first, n, KMIN, KMAX))
    t = time.perf_counter()
    for q in qs:
        O.ve_query(tm, q, {k: np.ones(len(b)) for k, b in q.items()})
    return time.perf_counter() - t


def cpu_baseline_single(tm, name, seconds):
    """Oracle port, one core, bounded sample of the bench workload."""
    from bayescard_b200.engine import gen_range_queries_host
    from oracle import bayescard_oracle as O

    probe = 500
    qs = queries_as_dicts(tm, gen_range_queries_host(tm, SEED, 0, probe, KMIN, KMAX))
    t = time.perf_counter()
    for q in qs:
        O.ve_query(tm, q, {k: np.ones(len(b)) for k, b in q.items()})
    rate = probe / (time.perf_counter() - t)
    n = int(max(probe, min(rate * seconds, 200_000)))
    qs = queries_as_dicts(tm, gen_range_queries_host(tm, SEED, 0, n, KMIN, KMAX))
    t = time.perf_counter()
    for q in qs:
        O.ve_query(tm, q, {k: np.ones(len(b)) for k, b in q.items()})
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n} queries of the {name} bench batch, oracle/bayescard_oracle.py ve_query "
                      f"(numpy fp64 restatement of VariableEliminationJIT.query, one Python process)"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    per_core = 1500
    chunks_per_step = cores
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(k):
            jobs = [(args.model, (k * chunks_per_step + c) * per_core, per_core) for c in range(chunks_per_step)]
            t = time.perf_counter()
            pool.map(_oracle_chunk, jobs)
            return time.perf_counter() - t

        for k in range(args.warmup):
            step(k)
        t_total = sum(step(args.warmup + k) for k in range(args.steps))
    nq = args.steps * chunks_per_step * per_core
    value = nq / t_total
    sample = (f"{chunks_per_step * per_core} queries of the {args.model} bench workload per step, "
              f"oracle port (numpy fp64) in {cores} worker processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.model} Chow-Liu BN, synthetic range queries k~U{{{KMIN}..{KMAX}}}",
                       "queries_per_step": chunks_per_step * per_core},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                smax = float(f[2])
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[1]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         f[4:8]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:  # timed region shorter than the sampling period: use every sample we have
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ our arm
def imdb_expectation_leg(device, fp32_peak, n=262144, reps=5):
    """BASELINE.json configs[2] in short (not the headline): fan-out weighted expectation factors on the shipped IMDB-1 BN
    (fractional n_distinct weights on random ranges + fan-out mask) through the fused tensor-core kernel (K3), next to the
    per-model straight-line kernel, and end to end from pinned host memory as WSPARSE runs."""
    import torch

    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import dense_to_wsparse, unpack_ranges
    from bayescard_b200.engine import DeviceModel
    from oracle import bayescard_oracle as O

    tm = load_tree("imdb1")
    dm = DeviceModel(tm, device=device, specialize=True)
    st = torch.cuda.current_stream().cuda_stream
    ranges = dm.gen_range_queries_host(SEED + 1, 0, n, 1, 4)
    lo, hi = unpack_ranges(tm, ranges)
    rng = np.random.default_rng(SEED + 2)
    W = np.zeros((n, dm.dense_width), dtype=np.float32)
    for v in range(tm.n_nodes):
        c = np.arange(int(tm.card[v]))[None, :]
        sel = (c >= lo[:, v:v + 1]) & (c <= hi[:, v:v + 1])
        con = (lo[:, v] > 0) | (hi[:, v] < int(tm.card[v]) - 1)
        w = np.where(con[:, None], rng.uniform(0.2, 1.0, sel.shape), 1.0)
        o = int(dm.dense_offset[v])
        W[:, o:o + int(tm.card[v])] = sel * w
    fan_nodes = [v for v in range(tm.n_nodes) if tm.infer_names[v] in tm.fanouts and tm.fan_vector(v) is not None]
    mask = np.zeros((n, 1), dtype=np.uint32)
    for v in fan_nodes:
        mask[:, 0] |= (rng.random(n) < 0.4).astype(np.uint32) << np.uint32(v)
    d_w, d_m = torch.from_numpy(W).cuda(device), torch.from_numpy(mask.view(np.int32)).cuda(device)
    out = torch.empty(n, dtype=torch.float32, device=f"cuda:{device}")
    res = {}
    for kname, kernel in (("k_spec", L.KERNEL_SPEC), ("k3_fused_tcgen05", L.KERNEL_FUSED)):
        ts = []
        for r in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dm.run_device(d_w.data_ptr(), n, L.DESC_DENSE_F32, out.data_ptr(), mask_ptr=d_m.data_ptr(), kernel=kernel, stream=st)
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        res[kname] = n / (float(np.median(ts)) * 1e-3)
    got = out.cpu().numpy().astype(np.float64)
    sub = np.arange(0, n, n // 2048)[:2048]
    Wl = []
    for v in range(tm.n_nodes):
        o = int(dm.dense_offset[v])
        w = W[sub, o:o + int(tm.card[v])].astype(np.float64)
        f = tm.fan_vector(v)
        if f is not None:
            on = ((mask[sub, 0] >> np.uint32(v)) & 1).astype(bool)
            w = np.where(on[:, None], w * f[None, :], w)
        Wl.append(w)
    ref = O.dense_tree(tm, Wl)
    rel = float(np.max(np.abs(got[sub] - ref) / np.maximum(np.abs(ref), 1e-300)))
    row_off, words = dense_to_wsparse(tm, W)
    p_off = torch.from_numpy(row_off.view(np.int32)).pin_memory().numpy().view(np.uint32)
    p_words = torch.from_numpy(words.view(np.int32)).pin_memory().numpy().view(np.uint32)
    p_mask = torch.from_numpy(mask.view(np.int32)).pin_memory().numpy().view(np.uint32)
    p_out = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    dm.run_wsparse_host(p_off, p_words, p_mask, out=p_out)
    t = time.perf_counter()
    for _ in range(3):
        dm.run_wsparse_host(p_off, p_words, p_mask, out=p_out)
    e2e = 3 * n / (time.perf_counter() - t)
    flops = dm.flops_dense
    dm.close()
    k3 = res["k3_fused_tcgen05"]
    return {"workload": f"shipped IMDB BN #1 (Benchmark/IMDB/1_chow-liu_1.pkl), {n} fan-out weighted expectation factors "
                        "(DENSE_F32 rows + fan-out mask), device resident",
            "factors_per_s": k3, "factors_per_s_k_spec": res["k_spec"], "kernel": "k3_kernel (tcgen05 3xTF32, messages in TMEM)",
            "flop_per_factor_dense": flops, "achieved_tflops": k3 * flops / 1e12,
            "frac_of_fp32_ffma_peak": k3 * flops / 1e12 / fp32_peak if fp32_peak else None,
            "e2e_factors_per_s_wsparse_host": e2e, "e2e_bytes_per_factor": (row_off.nbytes + words.nbytes + mask.nbytes) / n + 4,
            "rel_err_max_vs_fp64_oracle": rel}


def run_ours(args):
    import torch

    from bayescard_b200 import _lib as L
    from bayescard_b200.decode import unpack_ranges
    from bayescard_b200.engine import DeviceModel, launch_count, measure_fp32_peak
    from bayescard_b200.model import Bayescard_BN
    from bayescard_b200.sharding import max_over_ranks
    from oracle import bayescard_oracle as O  # checker + cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bayescard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    kernel = {"auto": L.KERNEL_AUTO, "generic": L.KERNEL_GENERIC, "spec": L.KERNEL_SPEC}[args.kernel]

    tm = load_tree(args.model)
    dm = DeviceModel(tm, device=local, specialize=True)
    if kernel != L.KERNEL_GENERIC and not dm.has_spec:
        raise SystemExit("specialised kernel unavailable: " + str(dm.spec_error))
    B = args.batch
    rstride = dm.desc_stride(L.DESC_RANGE_U8)
    spec = kernel != L.KERNEL_GENERIC
    fmt = L.DESC_BITS
    stride = dm.desc_stride(fmt)
    NBUF = 4  # rotate over 4 resident batches: the working set (4 x B x stride) is larger than L2
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    # seeded queries are generated on the device as RANGE_U8 rows (the form the host twin of the generator
    # and the oracle read) and converted once, outside the timed region, to the resident BITS rows
    ranges0 = torch.empty((B, rstride), dtype=torch.uint8, device=dev)
    tmp = torch.empty((B, rstride), dtype=torch.uint8, device=dev)
    descs = [torch.empty((B, stride), dtype=torch.uint8, device=dev) for _ in range(NBUF)]
    out = torch.empty(B, dtype=torch.float32, device=dev)
    for b, d in enumerate(descs):
        r = ranges0 if b == 0 else tmp
        dm.gen_range_queries_device(SEED, (rank * NBUF + b) * B, B, KMIN, KMAX, r.data_ptr(), st)
        dm.convert_device(r.data_ptr(), L.DESC_RANGE_U8, d.data_ptr(), fmt, B, st)
    torch.cuda.synchronize()
    del tmp

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        dm.run_device(descs[k % NBUF].data_ptr(), B, fmt, out.data_ptr(), kernel=kernel, stream=st)

    # ---- FP32 peak of this device, measured before the timed region ---------------------------
    fp32_peak, _ = measure_fp32_peak(local)

    # ---- device-resident throughput --------------------------------------------------------
    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25 if sampler else 0)
    launches0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(args.steps):
        step(k)
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    launches = launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------
    # the same B queries as SPARSE (CSR) entries in pinned host memory -- what a caller holding the
    # reference's sparse {column: bins} dicts would hand over
    row_off_np, entries_np = dm.gen_sparse_queries_host(SEED, rank * NBUF * B, B, KMIN, KMAX)
    h_off = torch.from_numpy(row_off_np.view(np.int32)).pin_memory()
    h_ent = torch.from_numpy(entries_np.view(np.int32)).pin_memory()
    h_out = torch.empty(B, dtype=torch.float32).pin_memory()
    ho_np, he_np, ho = h_off.numpy().view(np.uint32), h_ent.numpy().view(np.uint32), h_out.numpy()
    h2d_bytes = int(ho_np.nbytes + he_np.nbytes)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        dm.run_sparse_host(ho_np, he_np, None, kernel, out=ho)
    e2e_first = ho.copy()
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        dm.run_sparse_host(ho_np, he_np, None, kernel, out=ho)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te
    e2e_s = max_over_ranks(e2e_s, device=dev)
    e2e_value = world * B * e2e_steps / e2e_s
    clocks = sampler.stop(t0, time.perf_counter()) if sampler else None

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- parity of this very batch against the fp64 oracle (sub-sample) ------------------------
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(B, size=min(B, 10000), replace=False))
    dm.run_device(descs[0].data_ptr(), B, fmt, out.data_ptr(), kernel=kernel, stream=st)
    torch.cuda.synchronize()
    full = out.cpu().numpy()
    if not np.array_equal(full, e2e_first):
        raise SystemExit("end-to-end (SPARSE host) results differ from the device-resident (BITS) results")
    got = full[idx].astype(np.float64)
    lo, hi = unpack_ranges(tm, ranges0.cpu().numpy()[idx])
    ref = O.dense_tree(tm, O.range_weights(tm, lo, hi))
    rel_err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))

    # ---- p50 latency of the scalar drop-in call (B = 1, decode included) ----------------------
    bn = Bayescard_BN(tm, device=local, infer_algo="exact-jit")
    bn.init_inference_method()
    qd = queries_as_dicts(tm, dm.gen_range_queries_host(SEED, 0, 300, KMIN, KMAX))
    raw = []
    inv = {n: {b: v for v, b in tm.encoding[n].items()} for n in tm.infer_names if tm.encoding.get(n)}
    for q in qd:
        raw.append({k: [inv[k][b] for b in bins] for k, bins in q.items()})
    lat = []
    for q in raw:
        if not q:
            continue
        tq = time.perf_counter()
        bn.query(q)
        lat.append(time.perf_counter() - tq)
    p50_us = float(np.median(lat[20:]) * 1e6)
    bn.close()

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks, peak_src = measured_peaks()
    flop_q = 2 * dm.spec_ffma() if spec else dm.flops_dense
    kernel_s = ms * 1e-3 / args.steps
    achieved_tf = flop_q * B / kernel_s / 1e12
    bytes_q = stride + 4
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    kname = "bc_spec_bits" if spec else "k1_kernel"
    if os.path.exists(tpath):  # dram__bytes_read+write per query of this kernel, from the committed ncu capture
        with open(tpath) as f:
            per_q = json.load(f).get(f"{args.model}:{kname}")
        traffic = per_q * B if per_q is not None else None
    roof = {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp32_peak if fp32_peak else None, "traffic": traffic,
            "kernel": kname,
            "flop_per_query": flop_q, "flop_per_query_dense": dm.flops_dense,
            "peak_source": "FFMA micro-benchmark (bc_measure_fp32_peak) in this run",
            "note": "exact zeros of the CPTs emit no FFMA; achieved counts executed FFMAs only"}
    roof_hbm = {"bound": "hbm", "achieved": bytes_q * B / kernel_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": bytes_q * B / kernel_s / 1e9 / peaks["hbm_gbs"], "bytes_per_query": bytes_q,
                "peak_source": f"MEASURED_PEAKS.json ({peak_src})"}

    cpu = cpu_baseline_single(tm, args.model, args.cpu_seconds)
    secondary = imdb_expectation_leg(local, fp32_peak) if world == 1 else None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} Chow-Liu BN ({tm.n_nodes} columns), {B} synthetic range queries per "
                                   f"GPU per step, k~U{{{KMIN}..{KMAX}}} constrained columns, seed {SEED}",
                       "descriptor": "BITS (resident) / SPARSE CSR (e2e)", "bytes_per_query": bytes_q,
                       "l2": f"inputs rotate over {NBUF} resident batches = {NBUF * B * stride / 1e6:.0f} MB > 126 MB L2",
                       "kernel": roof["kernel"], "parallelism": f"replica x{world}, batch sharded, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * 4,
                    "steps": e2e_steps, "api": "bc_query_batch_sparse_host (pinned host CSR -> H2D -> expand -> "
                                               "infer -> D2H)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
            "cpu_baseline": cpu, "rel_err_max_vs_fp64_oracle": rel_err, "p50_latency_us_scalar_query": p50_us,
            "fp32_peak_tflops_measured": fp32_peak}
    if secondary is not None:
        line["secondary"] = secondary
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        # launched by hand: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
