"""CPT fitting for a fixed tree (bc_fit_counts + bayescard_b200/fit.py) against the reference's pgmpy MLE.

tests/golden/fit_dmv_shaped.npz holds a seeded DMV-shaped table and the CPDs the UNMODIFIED reference fitted on it
(tools/make_golden_fit.py).  Counting is integer work: the bar is bit-exact counts and bit-exact fp64 CPTs.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import golden_util as G  # noqa: E402
from bayescard_b200 import _lib as L  # noqa: E402
from bayescard_b200 import fit as F  # noqa: E402
from oracle import bayescard_oracle as O  # noqa: E402


def golden():
    d = np.load(os.path.join(G.GOLD, "fit_dmv_shaped.npz"))
    n = len(d["card"])
    return d["parent"], d["card"], d["table"], [d[f"cpd_{v}"] for v in range(n)]


def flat(counts):
    return np.concatenate([np.asarray(c, dtype=np.uint64).reshape(-1) for c in counts])


def test_oracle_fit_equals_reference_bit_exact():
    parent, card, table, ref = golden()
    for v, t in enumerate(O.fit_cpts(parent, card, table)):
        assert np.array_equal(t.reshape(ref[v].shape), ref[v]), v
    # the unobserved parent state (column 1, state 3) gives its children a uniform column (MLE.py:77-79)
    for v in range(1, len(card)):
        if parent[v] == 1:
            assert np.all(ref[v][:, 3] == 1.0 / card[v])


def test_host_normalisation_equals_reference():
    parent, card, table, ref = golden()
    counts, bad = O.fit_counts(parent, card, table)
    assert bad == 0
    off, total = F.count_layout(parent, card)
    assert total == sum(c.size for c in counts) and off[0] == 0
    for v, t in enumerate(F.counts_to_cpts(parent, card, flat(counts))):
        assert np.array_equal(t.reshape(ref[v].shape), ref[v]), v
    with pytest.raises(ValueError):
        F.counts_to_cpts(parent, card, np.zeros(3))


def test_bad_arguments_fail_before_any_device_work():
    parent = np.asarray([-1, 0], dtype=np.int32)
    card = np.asarray([3, 300], dtype=np.int32)
    lib = L.lib()
    # card 300 does not fit uint8 elements
    rc = lib.bc_fit_counts(0, 2, parent.ctypes.data, card.ctypes.data, None, 1, 0, 2, 1, 903, None, None)
    assert rc == -1 and b"card" in lib.bc_last_error()
    rc = lib.bc_fit_counts(0, 2, parent.ctypes.data, card.ctypes.data, None, 4, 0, 2, 1, 903, None, None)
    assert rc == -1


@pytest.mark.gpu
def test_gpu_counts_and_cpts_bit_exact():
    parent, card, table, ref = golden()
    cpts, counts, bad = F.fit_cpts(parent, card, table, device=0)
    assert bad == 0
    want, _ = O.fit_counts(parent, card, table)
    assert np.array_equal(counts, flat(want))
    for v, t in enumerate(cpts):
        assert np.array_equal(t.reshape(ref[v].shape), ref[v]), v


@pytest.mark.gpu
def test_gpu_edge_cases():
    parent, card, table, _ = golden()
    # empty table: every column of every CPT is uniform
    cpts, counts, bad = F.fit_cpts(parent, card, table[:0], device=0)
    assert counts.sum() == 0 and bad == 0
    assert all(np.allclose(t, 1.0 / card[v]) for v, t in enumerate(cpts))
    # ragged sizes around the CTA / warp granularity
    for n in (1, 31, 33, 255, 257, 4097):
        _, counts, _ = F.fit_cpts(parent, card, table[:n], device=0)
        assert np.array_equal(counts, flat(O.fit_counts(parent, card, table[:n])[0])), n
    # rows with a bin id outside the domain are skipped and reported
    dirty = table[:5000].copy()
    dirty[::7, 4] = 255
    _, counts, bad = F.fit_cpts(parent, card, dirty, device=0)
    want, want_bad = O.fit_counts(parent, card, dirty)
    assert bad == want_bad == len(dirty[::7]) and np.array_equal(counts, flat(want))


@pytest.mark.gpu
def test_gpu_uint16_large_domains_use_global_counters():
    rng = np.random.default_rng(3)
    parent = np.asarray([-1, 0, 1, 1], dtype=np.int32)
    card = np.asarray([300, 280, 310, 2], dtype=np.int32)  # 300 + 84 000 + 86 800 + 560 counters: beyond shared memory
    n = 200_003
    table = np.stack([rng.integers(0, c, n) for c in card], axis=1).astype(np.uint16)
    table[:, 3] = (table[:, 1] % 2)  # a deterministic child: structural zeros
    cpts, counts, bad = F.fit_cpts(parent, card, table, device=0)
    assert bad == 0 and np.array_equal(counts, flat(O.fit_counts(parent, card, table)[0]))
    ref = O.fit_cpts(parent, card, table)
    assert all(np.array_equal(a, b) for a, b in zip(cpts, ref))


@pytest.mark.gpu
def test_gpu_full_size_checksums_and_refit_round_trip():
    """Size-independent properties at DMV's real row count (11.6 M): every node's counts add up to the number of rows,
    the marginal of a child's table over its own states equals its parent's table marginal, and a model re-fitted
    from a sample of itself answers queries like the original."""
    import torch

    from bayescard_b200.engine import DeviceModel
    from bayescard_b200.engine import gen_range_queries_host

    tm = G.model("dmv")
    parent, card = tm.parent, tm.card
    n = 11_591_877
    g = torch.Generator(device="cuda").manual_seed(5)
    cols = []
    for v in range(tm.n_nodes):  # ancestral sample on the device (torch is plumbing here)
        t = torch.tensor(np.asarray(tm.cpts[v], dtype=np.float64).reshape(int(card[v]), -1), device="cuda")
        cdf = torch.cumsum(t / t.sum(dim=0, keepdim=True), dim=0).T.contiguous()  # [card_pa, card]
        u = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
        rows = cdf[cols[parent[v]].long()] if parent[v] >= 0 else cdf[torch.zeros(n, dtype=torch.long, device="cuda")]
        cols.append(torch.clamp((rows < u[:, None]).sum(dim=1), max=int(card[v]) - 1).to(torch.uint8))
        del rows
    table = torch.stack(cols, dim=1).contiguous()
    off, total = F.count_layout(parent, card)
    counts = torch.empty(total + 1, dtype=torch.int64, device="cuda")
    bad = F.fit_counts_device(parent, card, table.data_ptr(), n, 1, tm.n_nodes, counts.data_ptr(), 0,
                              torch.cuda.current_stream().cuda_stream, True)
    c = counts[:total].cpu().numpy()
    assert bad == 0
    per_node = [c[off[v]: off[v] + int(card[v]) * (int(card[parent[v]]) if parent[v] >= 0 else 1)] for v in range(tm.n_nodes)]
    assert all(int(x.sum()) == n for x in per_node)
    for v in range(1, tm.n_nodes):  # column sums of a child's table = the parent's own marginal counts
        pa = int(parent[v])
        pa_marg = per_node[pa].reshape(int(card[pa]), -1).sum(axis=1)
        assert np.array_equal(per_node[v].reshape(int(card[v]), int(card[pa])).sum(axis=0), pa_marg)
    assert np.array_equal(per_node[3].reshape(int(card[3]), -1),
                          torch.bincount(cols[3].long() * int(card[parent[3]]) + cols[parent[3]].long(),
                                         minlength=int(card[3]) * int(card[parent[3]])).cpu().numpy().reshape(int(card[3]), -1))
    # round trip: the re-fitted model reproduces the original's estimates (sampling noise only)
    cpts = F.counts_to_cpts(parent, card, c)
    import copy

    tm2 = copy.copy(tm)
    tm2.cpts = cpts
    a, b = DeviceModel(tm, device=0, specialize=False), DeviceModel(tm2, device=0, specialize=False)
    desc = gen_range_queries_host(tm, 1, 0, 2000, 1, 3)
    pa_, pb_ = a.run_host(desc, L.DESC_RANGE_U8), b.run_host(desc, L.DESC_RANGE_U8)
    big = pa_ > 1e-3
    assert np.max(np.abs(pa_[big] - pb_[big]) / pa_[big]) < 0.05
    a.close()
    b.close()
