"""N > 1 host logic on CPU: two gloo ranks shard a batch, evaluate their slices and gather on the host.

The per-slice evaluator here is the oracle (no GPU in this container); on the GPU box the same
``evaluate_sharded`` wraps ``DeviceModel.run_sparse_host`` (tests/test_gpu_parity.py::test_sharded_ranks_on_gpu)."""
import os
import socket

import numpy as np
import pytest

import golden_util as G
from bayescard_b200.decode import PredicateCompiler, unpack_ranges
from bayescard_b200.engine import gen_range_queries_host
from bayescard_b200.sharding import csr_slice, evaluate_sharded, max_over_ranks, rank_range
from oracle import bayescard_oracle as O

N_QUERIES = 2001  # odd: the last rank takes the remainder


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = G.model("dmv")
        pc = PredicateCompiler(m)
        desc = gen_range_queries_host(m, 21, 0, N_QUERIES, 1, 6)
        lo, hi = unpack_ranges(m, desc)
        row_off, entries = pc.pack_sparse(lo, hi)
        seen = []

        def evaluate(a, b):
            # what a GPU rank does: take its CSR slice; here the oracle evaluates it
            ro, en = csr_slice(row_off, entries, a, b)
            assert ro[0] == 0 and ro[-1] == en.size
            seen.append((a, b))
            return O.dense_tree(m, O.range_weights(m, lo[a:b], hi[a:b])).astype(np.float32)

        full = evaluate_sharded(evaluate, N_QUERIES)
        only0 = evaluate_sharded(evaluate, N_QUERIES, gather_to=0)
        slowest = max_over_ranks(1.0 + rank)
        q.put((rank, seen[0], full, None if only0 is None else only0.copy(), slowest))
    finally:
        dist.destroy_process_group()


def test_rank_range_matches_in_process_split():
    from bayescard_b200.engine import ShardedModel

    for n, w in [(10, 4), (3, 8), (2001, 2), (0, 3), (1_000_000, 8)]:
        assert [rank_range(n, r, w) for r in range(w)] == ShardedModel.split(n, w)
    with pytest.raises(ValueError):
        rank_range(5, 2, 2)


def test_two_gloo_ranks_shard_and_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=180)
        res[r[0]] = r
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    m = G.model("dmv")
    lo, hi = unpack_ranges(m, gen_range_queries_host(m, 21, 0, N_QUERIES, 1, 6))
    ref = O.dense_tree(m, O.range_weights(m, lo, hi)).astype(np.float32)
    assert res[0][1] == (0, 1000) and res[1][1] == (1000, 2001)
    for r in (0, 1):
        assert np.array_equal(res[r][2], ref)          # all_gather: every rank holds the full result
        assert res[r][4] == 2.0                        # max over ranks of (1 + rank)
    assert np.array_equal(res[0][3], ref) and res[1][3] is None   # gather_to=0


def test_single_process_identity():
    out = evaluate_sharded(lambda a, b: np.arange(a, b, dtype=np.float32), 17)
    assert np.array_equal(out, np.arange(17, dtype=np.float32))
    assert max_over_ranks(3.5) == 3.5
