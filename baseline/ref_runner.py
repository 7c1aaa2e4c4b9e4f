"""Drive the UNMODIFIED reference exact-jit path for ``bench.py --impl reference`` and the ``cpu_baseline`` leg.

It mirrors the loop of ``Testing/BN_testing.py:9-46`` (that file itself is not importable: its line 2 pulls in the
spflow-based IMDB planner): ``pickle.load`` -> ``BN.infer_algo = 'exact-jit'`` -> ``BN.init_inference_method()`` ->
``BN.query(...)`` timed per call with ``perf_counter``.  The bench workload is defined over the discretised bins, so
the calls use the ``n_distinct=`` form of the public API, ``BN.query(query_bins, n_distinct=weights, return_prob=True)``
-- exactly how ``Models/BN_ensemble_model.py:235`` calls it.  Nothing of bayescard_b200's engine, kernels or oracle is
on this path; the query stream is regenerated here in pure Python from the same counter-based generator
(``bc_gen_row`` in ``bayescard_b200/csrc/bc_api.cu``; equality with the C generator is a CPU test).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, "_ref")
MODEL_REL = {"dmv": "Benchmark/DMV/chow-liu_1.pkl", "census": "Benchmark/Census/chow-liu_1.pkl",
             **{f"imdb{i}": f"Benchmark/IMDB/{i}_chow-liu_1.pkl" for i in range(5)}}
_M64 = (1 << 64) - 1


def available() -> bool:
    return os.path.isdir(os.path.join(STAGED, "Pgmpy")) and os.path.exists(os.path.join(STAGED, MODEL_REL["census"]))


def _mix(s):
    s = (s + 0x9E3779B97F4A7C15) & _M64
    z = s
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return s, z ^ (z >> 31)


def gen_ranges(card, seed, first, n, kmin, kmax):
    """Pure-Python twin of ``bc_gen_range_queries_host``: ``(lo, hi)`` int arrays ``[n, len(card)]``."""
    nn = len(card)
    lo = np.zeros((n, nn), dtype=np.int64)
    hi = np.tile(np.asarray(card, dtype=np.int64) - 1, (n, 1))
    for i in range(n):
        s = (seed * 0xD1342543DE82EF95 + (first + i) * 0x2545F4914F6CDD1D + 0x1234567) & _M64
        s, _ = _mix(s)
        s, r = _mix(s)
        k = min(kmin + r % (kmax - kmin + 1), nn)
        used = set()
        for _j in range(k):
            while True:
                s, r = _mix(s)
                v = r % nn
                if v not in used:
                    break
            used.add(v)
            c = int(card[v])
            s, r = _mix(s)
            l = r % c
            s, r = _mix(s)
            h = l + r % (c - l)
            lo[i, v], hi[i, v] = l, h
    return lo, hi


def queries_as_dicts(names, card, lo, hi):
    """(bins, n_distinct) dict pairs: the sparse form ``query_decoding`` produces for unit-weight range predicates."""
    out = []
    for i in range(lo.shape[0]):
        q, nd = {}, {}
        for v, name in enumerate(names):
            if lo[i, v] > 0 or hi[i, v] < card[v] - 1:
                q[name] = list(range(int(lo[i, v]), int(hi[i, v]) + 1))
                nd[name] = np.ones(len(q[name]))
        out.append((q, nd))
    return out


_BN = {}


def load_bn(model: str):
    """``Testing/BN_testing.py:11-15`` on the staged copy."""
    if model not in _BN:
        os.environ["BAYESCARD_REFERENCE"] = STAGED
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ref_harness as R

        R.REFERENCE_ROOT = STAGED
        _BN[model] = R.load_bn(MODEL_REL[model], "exact-jit")
    return _BN[model]


def run_chunk(args):
    """Worker: ``n`` queries of the seeded stream starting at ``first``; returns (seconds inside BN.query, results)."""
    model, names, card, seed, first, n, kmin, kmax = args
    bn = load_bn(model)
    lo, hi = gen_ranges(card, seed, first, n, kmin, kmax)
    qs = queries_as_dicts(names, card, lo, hi)
    res = np.zeros(n)
    t_in = 0.0
    for i, (q, nd) in enumerate(qs):
        t = time.perf_counter()
        r = bn.query(q, n_distinct=nd, return_prob=True)
        t_in += time.perf_counter() - t
        p = r[0] if isinstance(r, tuple) else r
        res[i] = float(np.asarray(p).reshape(-1)[0])
    return t_in, res
