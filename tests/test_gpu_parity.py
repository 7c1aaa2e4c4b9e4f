"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden outputs and
against the fp64 oracle.  Tolerance: 1e-5 relative (BASELINE.json north_star), fp32 kernels vs fp64."""
import copy

import numpy as np
import pytest

import golden_util as G
from bayescard_b200 import _lib as L
from bayescard_b200.decode import PredicateCompiler, dense_to_wsparse, unpack_ranges
from bayescard_b200.engine import DeviceModel, ShardedModel, launch_count
from bayescard_b200.ensemble import BN_ensemble
from bayescard_b200.model import Bayescard_BN
from bayescard_b200.sql_front import parse_query_single_table
from oracle import bayescard_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5
KERNELS = [L.KERNEL_GENERIC, L.KERNEL_SPEC]
_dm = {}


def dev_model(name) -> DeviceModel:
    if name not in _dm:
        _dm[name] = DeviceModel(G.model(name), device=0, specialize=True)
        assert _dm[name].has_spec, _dm[name].spec_error
    return _dm[name]


def rel_err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)


def assert_close(got, ref, what=""):
    ref = np.asarray(ref, dtype=np.float64)
    err = rel_err(got, ref)
    err[(ref == 0) & (np.asarray(got) == 0)] = 0
    assert err.max() <= RTOL, (what, float(err.max()), int(err.argmax()))


# ------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", G.MODEL_NAMES)
def test_infer_cases_both_kernels(name, kernel):
    """Decoded (bins, fractional weights, fan-out) cases: range + dense descriptors + fan-out mask."""
    m, dm = G.model(name), dev_model(name)
    pc = PredicateCompiler(m)
    cases = [r for r in G.load("infer_cases.json.gz")[name] if "error" not in r]
    decoded = [({k: list(v) for k, v in r["bins"].items()}, {k: np.asarray(v) for k, v in r["weights"].items()})
               for r in cases]
    ref = np.asarray([np.asarray(r["p"]["value"]).reshape(-1)[0] for r in cases])
    for force_dense in (False, True):
        r_idx, r_desc, d_idx, d_desc, mask = pc.pack(decoded, [r["fanout"] for r in cases], force_dense=force_dense)
        got = np.zeros(len(cases))
        before = launch_count()
        if len(r_idx):
            got[r_idx] = dm.run_host(r_desc, L.DESC_BITS, mask[r_idx], kernel)
        if len(d_idx):
            got[d_idx] = dm.run_host(d_desc, L.DESC_DENSE_F32, mask[d_idx], kernel)
        assert launch_count() > before, "no kernel was launched"
        assert_close(got, ref, (name, kernel, force_dense))


FUSED_MODELS = [n for n in G.MODEL_NAMES if n != "census"]  # K3 serves models of <= 32 columns


@pytest.mark.parametrize("name", FUSED_MODELS)
def test_infer_cases_fused_kernel(name):
    """K3 (whole tree per 128-query tile on the tensor cores, 3xTF32): the golden query / expectation cases of the
    unmodified reference through BITS rows, DENSE_F32 rows and the fan-out mask."""
    m, dm = G.model(name), dev_model(name)
    pc = PredicateCompiler(m)
    cases = [r for r in G.load("infer_cases.json.gz")[name] if "error" not in r]
    decoded = [({k: list(v) for k, v in r["bins"].items()}, {k: np.asarray(v) for k, v in r["weights"].items()})
               for r in cases]
    ref = np.asarray([np.asarray(r["p"]["value"]).reshape(-1)[0] for r in cases])
    for force_dense in (False, True):
        r_idx, r_desc, d_idx, d_desc, mask = pc.pack(decoded, [r["fanout"] for r in cases], force_dense=force_dense)
        got = np.zeros(len(cases))
        before = launch_count()
        if len(r_idx):
            got[r_idx] = dm.run_host(r_desc, L.DESC_BITS, mask[r_idx], L.KERNEL_FUSED)
        if len(d_idx):
            got[d_idx] = dm.run_host(d_desc, L.DESC_DENSE_F32, mask[d_idx], L.KERNEL_FUSED)
        assert launch_count() > before, "no kernel was launched"
        assert_close(got, ref, (name, "fused", force_dense))


@pytest.mark.parametrize("name", ["dmv", "imdb1", "imdb3"])
def test_fused_kernel_batch_vs_oracle_and_formats(name):
    """A ragged batch (not a multiple of the 128-query tile, more tiles than CTAs) through K3: fp64 oracle on every
    query; range rows, BITS rows and SPARSE rows give the same bits; AUTO picks K3 for the big models."""
    import torch

    m, dm = G.model(name), dev_model(name)
    n = 148 * 2 * 128 * 2 + 77
    kmax = min(m.n_nodes, 14)
    host = dm.gen_range_queries_host(5, 9, n, 1, kmax)
    lo, hi = unpack_ranges(m, host)
    sub = np.random.default_rng(1).choice(n, 6000, replace=False)
    sub[:300] = np.arange(n - 300, n)  # the ragged last tiles
    ref = O.dense_tree(m, O.range_weights(m, lo[sub], hi[sub]))
    got = dm.run_host(host, L.DESC_RANGE_U8, None, L.KERNEL_FUSED).astype(np.float64)
    assert_close(got[sub], ref, (name, "fused ranges"))
    pc = PredicateCompiler(m)
    via_bits = dm.run_host(pc.pack_bits(lo, hi), L.DESC_BITS, None, L.KERNEL_FUSED).astype(np.float64)
    assert np.array_equal(via_bits, got)
    row_off, entries = pc.pack_sparse(lo, hi)
    assert np.array_equal(dm.run_sparse_host(row_off, entries, None, L.KERNEL_FUSED).astype(np.float64), got)
    # tiny batches: fewer queries than one tile, one query, none
    for k in (1, 5, 127, 129):
        assert np.array_equal(dm.run_host(host[:k], L.DESC_RANGE_U8, None, L.KERNEL_FUSED).astype(np.float64), got[:k])
    assert dm.run_host(host[:0], L.DESC_RANGE_U8, None, L.KERNEL_FUSED).shape == (0,)
    # unconstrained query: total mass 1; an empty range selects nothing
    full = np.zeros((2, host.shape[1]), dtype=np.uint8)
    full[:, 1:2 * m.n_nodes:2] = np.minimum(m.card - 1, 255)
    full[1, 2 * 1], full[1, 2 * 1 + 1] = 3, 1
    r = dm.run_host(full, L.DESC_RANGE_U8, None, L.KERNEL_FUSED)
    assert abs(r[0] - 1.0) < 1e-5 and r[1] == 0.0
    auto = dm.run_host(host[:4096], L.DESC_RANGE_U8, None, L.KERNEL_AUTO).astype(np.float64)
    if dm.flops_dense >= 30000:
        assert np.array_equal(auto, got[:4096])  # AUTO = K3 on the IMDB-sized models
    else:
        assert_close(auto, got[:4096], (name, "auto"))


def test_fused_kernel_properties_full_batch():
    """1 M fan-out weighted expectation factors on an IMDB BN through K3 (BASELINE config 3 at bench size): properties
    that need no oracle -- permutation equivariance bit for bit, additivity of a split range, scaling linearity of the
    weights, agreement with the straight-line kernel -- plus the fp64 oracle on a seeded sub-sample."""
    import torch

    name = "imdb2"
    m, dm = G.model(name), dev_model(name)
    n = 1_000_000
    st = torch.cuda.current_stream().cuda_stream
    host = dm.gen_range_queries_host(21, 0, n, 1, 4)
    lo, hi = unpack_ranges(m, host)
    rng = np.random.default_rng(9)
    gen = torch.Generator(device="cuda:0").manual_seed(4)
    d_lo, d_hi = torch.from_numpy(lo.astype(np.int32)).cuda(), torch.from_numpy(hi.astype(np.int32)).cuda()
    dense = torch.zeros((n, dm.dense_width), dtype=torch.float32, device="cuda:0")
    for v in range(m.n_nodes):
        c = torch.arange(int(m.card[v]), device="cuda:0", dtype=torch.int32)[None, :]
        sel = (c >= d_lo[:, v:v + 1]) & (c <= d_hi[:, v:v + 1])
        w = 0.25 + 0.75 * torch.rand((n, int(m.card[v])), device="cuda:0", generator=gen)
        o = int(dm.dense_offset[v])
        dense[:, o:o + int(m.card[v])] = sel.to(torch.float32) * w
    fan_nodes = [v for v in range(m.n_nodes) if m.fan_vector(v) is not None]
    mask_h = np.zeros(n, dtype=np.uint32)
    for v in fan_nodes:
        mask_h |= (rng.random(n) < 0.4).astype(np.uint32) << np.uint32(v)
    mask = torch.from_numpy(mask_h.view(np.int32)).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda:0")

    def run(rows, msk, kernel=L.KERNEL_FUSED):
        o = torch.empty(rows.shape[0], dtype=torch.float32, device="cuda:0")
        dm.run_device(rows.data_ptr(), rows.shape[0], L.DESC_DENSE_F32, o.data_ptr(), mask_ptr=msk.data_ptr(), kernel=kernel, stream=st)
        torch.cuda.synchronize()
        return o

    out = run(dense, mask)
    assert bool(torch.all(out >= 0))
    # (a) permutation: results follow their factors bit for bit (tiles are formed from different neighbours)
    perm = torch.randperm(n, device="cuda:0", generator=gen)
    assert torch.equal(run(dense[perm].contiguous(), mask[perm].contiguous()), out[perm])
    # (b) additivity: the weights of one column split into two disjoint halves
    v = max(range(1, m.n_nodes), key=lambda u: int(m.card[u]))
    o, card = int(dm.dense_offset[v]), int(m.card[v])
    sub = torch.arange(0, 200000, device="cuda:0")
    a, b = dense[sub].clone(), dense[sub].clone()
    a[:, o + card // 2:o + card] = 0
    b[:, o:o + card // 2] = 0
    tot = (run(a, mask[sub].contiguous()).double() + run(b, mask[sub].contiguous()).double())
    ref = out[sub].double()
    assert bool(torch.all((tot - ref).abs() <= 1e-5 * ref.abs() + 1e-30))
    # (c) linearity: scaling one column's weights by 0.5 halves the result (exact in binary floating point)
    h = dense[sub].clone()
    h[:, o:o + card] *= 0.5
    assert torch.equal(run(h, mask[sub].contiguous()), out[sub] * 0.5)
    # (d) the straight-line kernel agrees within the budget; (e) fp64 oracle on a sub-sample
    spec = run(dense[sub].contiguous(), mask[sub].contiguous(), L.KERNEL_SPEC).double()
    assert bool(torch.all((spec - ref).abs() <= 1e-5 * ref.abs() + 1e-30))
    idx = np.sort(rng.choice(n, 4000, replace=False))
    Wh = dense[torch.from_numpy(idx).cuda()].cpu().numpy().astype(np.float64)
    W = []
    for u in range(m.n_nodes):
        w = Wh[:, int(dm.dense_offset[u]):int(dm.dense_offset[u]) + int(m.card[u])]
        f = m.fan_vector(u)
        if f is not None:
            on = ((mask_h[idx] >> np.uint32(u)) & 1).astype(bool)
            w = np.where(on[:, None], w * f[None, :], w)
        W.append(w)
    assert_close(out.cpu().numpy()[idx], O.dense_tree(m, W), "imdb2 1M factors, fused kernel")


def test_model_from_flat_file_equals_model_from_arrays(tmp_path):
    """bc_model_create_from_file (mmap of the .bcm file, no pickle) serves the same bits as bc_model_create."""
    for name in ("dmv", "imdb1"):
        m, dm = G.model(name), dev_model(name)
        path = str(tmp_path / (name + ".bcm"))
        m.save_flat(path)
        df = DeviceModel.from_flat_file(path, device=0, specialize=True)
        desc = dm.gen_range_queries_host(3, 0, 5000, 1, min(m.n_nodes, 14))
        for kernel in (L.KERNEL_GENERIC, L.KERNEL_SPEC, L.KERNEL_FUSED):
            assert np.array_equal(df.run_host(desc, L.DESC_RANGE_U8, None, kernel), dm.run_host(desc, L.DESC_RANGE_U8, None, kernel))
        df.close()


@pytest.mark.parametrize("name", ["dmv", "imdb0", "imdb3"])
def test_wsparse_rows_expand_to_the_dense_rows(name):
    """WSPARSE (weighted runs over PCIe) == the DENSE_F32 rows they stand for: device expansion bit for bit, results bit for
    bit through every kernel, empty rows = unconstrained query, golden expectation cases through the public batch API."""
    import torch

    m, dm = G.model(name), dev_model(name)
    pc = PredicateCompiler(m)
    cases = [r for r in G.load("infer_cases.json.gz")[name] if "error" not in r]
    decoded = [({k: list(v) for k, v in r["bins"].items()}, {k: np.asarray(v) for k, v in r["weights"].items()}) for r in cases]
    _, _, d_idx, dense, mask = pc.pack(decoded, [r["fanout"] for r in cases], force_dense=True)
    # a larger seeded batch with fractional weights on random ranges, all-zero columns and untouched columns
    rng = np.random.default_rng(5)
    big = np.repeat(dense, 40, axis=0)
    big *= (rng.random(big.shape) < 0.9) * rng.uniform(0.1, 1.0, big.shape).astype(np.float32)
    off = dm.dense_offset
    for v in range(m.n_nodes):  # restore some columns to "unconstrained"
        rows = rng.random(big.shape[0]) < 0.5
        big[np.ix_(rows, np.arange(off[v], off[v] + int(m.card[v])))] = 1.0
    for rows_dense, msk in ((dense, mask), (big.astype(np.float32), None)):
        row_off, words = dense_to_wsparse(m, rows_dense)
        assert words.nbytes < rows_dense.nbytes
        d_off, d_words = torch.from_numpy(row_off.astype(np.int64)).to(torch.int32).cuda(), torch.from_numpy(words.astype(np.int64)).to(torch.int32).cuda()
        d_dense = torch.empty(rows_dense.shape, dtype=torch.float32, device="cuda:0")
        L.check(L.lib().bc_expand_wsparse(dm._h, d_off.data_ptr(), d_words.data_ptr(), rows_dense.shape[0], d_dense.data_ptr(), None))
        torch.cuda.synchronize()
        back = d_dense.cpu().numpy()
        for v in range(m.n_nodes):  # row padding is never read by a kernel
            assert np.array_equal(back[:, off[v]: off[v] + int(m.card[v])], rows_dense[:, off[v]: off[v] + int(m.card[v])])
        for kernel in (L.KERNEL_GENERIC, L.KERNEL_SPEC, L.KERNEL_FUSED):
            a = dm.run_wsparse_host(row_off, words, msk, kernel)
            b = dm.run_host(rows_dense, L.DESC_DENSE_F32, msk, kernel)
            assert np.array_equal(a, b), (name, kernel)
    one = dm.run_wsparse_host(np.zeros(3, dtype=np.uint32), np.zeros(0, dtype=np.uint32))
    assert np.all(np.abs(one - 1.0) < 1e-5)


def test_wsparse_run_of_256_states_on_the_device():
    """A 256-state column with weights at states 0 and 255 (one run of 256 states: sent as 255 + a continuation run)."""
    import torch

    from bayescard_b200.synth import make_tree_model

    m = make_tree_model(3, [256, 200, 256], seed=5)
    dm = DeviceModel(m, device=0, specialize=False)
    rng = np.random.default_rng(0)
    rows = np.ones((64, dm.dense_width), dtype=np.float32)
    o = dm.dense_offset
    rows[0, o[0]:o[0] + 256] = 0.0
    rows[0, o[0]] = 0.25
    rows[0, o[0] + 255] = 0.5
    rows[1:, o[0]:o[0] + 256] = rng.uniform(0.1, 1.0, (63, 256))
    rows[2::2, o[2]:o[2] + 256] = (rng.random((31, 256)) < 0.5)
    rows[2::2, o[2]] = 1.0
    rows[2::2, o[2] + 255] = 1.0
    row_off, words = dense_to_wsparse(m, rows)
    d_off = torch.from_numpy(row_off.astype(np.int64)).to(torch.int32).cuda()
    d_words = torch.from_numpy(words.astype(np.int64)).to(torch.int32).cuda()
    d_dense = torch.empty(rows.shape, dtype=torch.float32, device="cuda:0")
    L.check(L.lib().bc_expand_wsparse(dm._h, d_off.data_ptr(), d_words.data_ptr(), rows.shape[0], d_dense.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(d_dense.cpu().numpy(), rows)
    got = dm.run_wsparse_host(row_off, words, None, L.KERNEL_GENERIC)
    ref = O.dense_tree(m, [rows[:, o[v]:o[v] + int(m.card[v])].astype(np.float64) for v in range(3)])
    assert_close(got, ref, "wsparse 256")
    dm.close()


def test_fused_kernel_wide_model_and_refusals():
    """K3 serves trees of up to 128 columns (Census: 68 columns, three fan-out-mask words would be two) and declines, with a
    message, what does not fit tensor memory; AUTO then falls through to the other kernels."""
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16, random_range_queries

    m, dm = G.model("census"), dev_model("census")
    desc = dm.gen_range_queries_host(1, 0, 3000, 1, 14)
    lo, hi = unpack_ranges(m, desc)
    ref = O.dense_tree(m, O.range_weights(m, lo, hi))
    assert_close(dm.run_host(desc, L.DESC_RANGE_U8, None, L.KERNEL_FUSED), ref, "census through the fused kernel")
    big = make_tree_model(10, 200, seed=10, dtype=np.float32)   # two 200-state messages + a 208-column accumulator
    db = DeviceModel(big, device=0, specialize=False)
    lo, hi = random_range_queries(big, 512, seed=1, kmax=5)
    rows = pack_ranges_u16(lo, hi)
    with pytest.raises(L.BayesCardError, match="K3"):
        db.run_host(rows, L.DESC_RANGE_U16, None, L.KERNEL_FUSED)
    auto = db.run_host(rows, L.DESC_RANGE_U16, None, L.KERNEL_AUTO)
    assert_close(auto, O.dense_tree(big, O.range_weights(big, lo, hi)), "declined model through AUTO")
    db.close()
    wide = make_tree_model(100, 10, seed=100, dtype=np.float32)  # 100 columns x 10 states: beyond one mask word, small domains
    dw = DeviceModel(wide, device=0, specialize=False)
    lo, hi = random_range_queries(wide, 2000, seed=2, kmax=10)
    rows = pack_ranges_u16(lo, hi)
    assert_close(dw.run_host(rows, L.DESC_RANGE_U16, None, L.KERNEL_FUSED), O.dense_tree(wide, O.range_weights(wide, lo, hi)),
                 "100-column tree through the fused kernel")
    dw.close()


K3_SHAPES = {
    # name: (parent, cards) -- each one drives a different path of the fused kernel's epilogue / step sequence
    "star_root_split": ([-1, 0, 0, 0, 0, 0, 0], [40, 50, 33, 64, 7, 50, 21]),            # no tail; the root's message in both epilogue groups (32 + 8 columns)
    "chain": ([-1, 0, 1, 2, 3], [8, 40, 60, 50, 33]),                                     # every edge is the only message into its parent: tail = all but one
    "root_two_internal": ([-1, 0, 0, 1, 1, 2, 2, 0], [30, 45, 52, 20, 61, 33, 17, 9]),    # the root's runs are interrupted: its message visits tensor memory
    "hub_112_columns": ([-1, 0, 1, 1, 1, 1], [5, 110, 40, 50, 60, 70]),                   # a parent beyond the register accumulators: message (*)= D in tensor memory
    "root_100_states": ([-1, 0, 0, 0, 3], [100, 20, 30, 40, 50]),                         # the same for the root + an internal child
    "deep_hub_tail": ([-1, 0, 1, 2, 3, 3, 3, 3], [3, 12, 70, 84, 80, 9, 77, 50]),         # IMDB-2-like: hub under a chain of three
    "two_hubs": ([-1, 0, 1, 1, 1, 0, 5, 5, 5], [6, 60, 30, 40, 50, 72, 20, 64, 33]),      # two internal children with leaves each
}


@pytest.mark.parametrize("shape", sorted(K3_SHAPES))
@pytest.mark.parametrize("nq", [3000, 60000])
def test_fused_kernel_tree_shapes(shape, nq):
    """K3 on synthetic trees chosen to walk every branch of its planner and epilogue (register accumulators split over two warp
    groups, tensor-memory fallback for wide parents, the root folded in registers or read from tensor memory, tail edges
    interleaved with the next tile, one tile per CTA and several passes per CTA), BITS and DENSE_F32 rows, against the fp64 oracle."""
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16, random_range_queries

    parent, cards = K3_SHAPES[shape]
    tm = make_tree_model(len(cards), cards, seed=len(cards) * 7 + nq % 5, dtype=np.float32, parent=parent)
    dm = DeviceModel(tm, device=0, specialize=False)
    lo, hi = random_range_queries(tm, nq, seed=3, kmax=len(cards))
    w = O.range_weights(tm, lo, hi)
    ref = O.dense_tree(tm, w)
    rows = pack_ranges_u16(lo, hi)
    try:
        got = dm.run_host(rows, L.DESC_RANGE_U16, None, L.KERNEL_FUSED)
    except L.BayesCardError as e:
        dm.close()
        pytest.skip(f"K3 declines this model: {e}")
    assert_close(got, ref, f"{shape}: BITS rows through the fused kernel")
    # DENSE_F32 rows with fractional weights: every selected state's weight scaled by a per-(query, column) factor
    rng = np.random.default_rng(5)
    width = dm.dense_width
    dense = np.full((nq, width), np.nan, dtype=np.float32)   # the pad floats between columns are the caller's: whatever they hold must not matter
    wf = []
    for v in range(tm.n_nodes):
        f = (0.25 + 0.75 * rng.random((nq, 1))).astype(np.float32)
        wv = (np.asarray(w[v], dtype=np.float32) * f).astype(np.float32)
        wf.append(wv.astype(np.float64))
        o = int(dm.dense_offset[v])
        dense[:, o:o + int(tm.card[v])] = wv
    ref_d = O.dense_tree(tm, wf)
    got_d = dm.run_host(dense, L.DESC_DENSE_F32, None, L.KERNEL_FUSED)
    assert_close(got_d, ref_d, f"{shape}: DENSE rows through the fused kernel")
    dm.close()


@pytest.mark.parametrize("name", ["dmv", "census", "imdb1"])
def test_synthetic_batch_generic_vs_spec_vs_oracle(name):
    """Config 2/5 generator: device == host twin bit for bit; both kernels vs the fp64 dense oracle."""
    import torch

    m, dm = G.model(name), dev_model(name)
    n = 20000 + 37  # ragged: not a multiple of the warp / CTA size
    kmax = min(m.n_nodes, 14)
    stride = dm.desc_stride(L.DESC_RANGE_U8)
    desc = torch.zeros((n, stride), dtype=torch.uint8, device="cuda:0")
    out = torch.empty(n, dtype=torch.float32, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    dm.gen_range_queries_device(11, 5, n, 1, kmax, desc.data_ptr(), st)
    torch.cuda.synchronize()
    host = dm.gen_range_queries_host(11, 5, n, 1, kmax)
    assert np.array_equal(desc.cpu().numpy(), host)
    lo, hi = unpack_ranges(m, host)
    ref = O.dense_tree(m, O.range_weights(m, lo, hi))
    res = {}
    for kernel in KERNELS:
        out.zero_()
        dm.run_device(desc.data_ptr(), n, L.DESC_RANGE_U8, out.data_ptr(), kernel=kernel, stream=st)
        torch.cuda.synchronize()
        res[kernel] = out.cpu().numpy().astype(np.float64)
        assert_close(res[kernel], ref, (name, kernel))
    # host-buffer pipeline == device-buffer launch, bit for bit
    via_host = dm.run_host(host, L.DESC_RANGE_U8, None, L.KERNEL_SPEC)
    assert np.array_equal(via_host.astype(np.float64), res[L.KERNEL_SPEC])
    # the three descriptor forms of the same queries give the same bits out: device conversion == host packing
    pc = PredicateCompiler(m)
    bits_host = pc.pack_bits(lo, hi)
    bits_dev = torch.empty((n, dm.desc_stride(L.DESC_BITS)), dtype=torch.uint8, device="cuda:0")
    dm.convert_device(desc.data_ptr(), L.DESC_RANGE_U8, bits_dev.data_ptr(), L.DESC_BITS, n, st)
    torch.cuda.synchronize()
    assert np.array_equal(bits_dev.cpu().numpy(), bits_host)
    for kernel in KERNELS:
        via_bits = dm.run_host(bits_host, L.DESC_BITS, None, kernel)
        assert np.array_equal(via_bits.astype(np.float64), res[kernel])
    row_off, entries = dm.gen_sparse_queries_host(11, 5, n, 1, kmax)
    ro2, en2 = pc.pack_sparse(lo, hi)
    assert np.array_equal(row_off, ro2) and np.array_equal(entries, en2)
    for kernel in KERNELS:
        via_sparse = dm.run_sparse_host(row_off, entries, None, kernel)
        assert np.array_equal(via_sparse.astype(np.float64), res[kernel])


def test_sparse_in_lists_and_chunking():
    """SPARSE entries with the continuation bit build IN lists; > 256K queries exercise the chunked pipeline."""
    m, dm = G.model("dmv"), dev_model("dmv")
    pc = PredicateCompiler(m)
    rng = np.random.default_rng(3)
    n = 600_000 + 11
    v = 9  # a 63-state leaf
    card = int(m.card[v])
    picks = rng.integers(0, card, size=(n, 3))
    bits = np.zeros((n, card), dtype=bool)
    np.put_along_axis(bits, picks, True, axis=1)
    row_off = (np.arange(n + 1, dtype=np.uint32) * 3).astype(np.uint32)
    ent = (np.uint32(v) | (picks.astype(np.uint32) << 16) | (picks.astype(np.uint32) << 24))
    ent[:, 1:] |= np.uint32(1 << 15)  # OR into the mask started by the first entry
    got = dm.run_sparse_host(row_off, ent.reshape(-1)).astype(np.float64)
    W = [np.ones((n, int(c))) for c in m.card]
    W[v] = bits.astype(np.float64)
    sub = rng.choice(n, 5000, replace=False)
    ref = O.dense_tree(m, [w[sub] for w in W])
    assert_close(got[sub], ref, "sparse IN lists")
    # the same masks as BITS rows, built on the host
    decoded = [({m.infer_names[v]: sorted(set(map(int, picks[i])))}, {m.infer_names[v]: np.ones(len(set(picks[i])))})
               for i in sub[:300]]
    b_idx, b_desc, d_idx, _, _ = pc.pack(decoded)
    assert len(d_idx) == 0
    assert np.array_equal(dm.run_host(b_desc, L.DESC_BITS).astype(np.float64), got[sub[:300]])
    # an empty SPARSE row is the unconstrained query
    one = dm.run_sparse_host(np.zeros(2, dtype=np.uint32), np.zeros(0, dtype=np.uint32))
    assert abs(one[0] - 1.0) < 1e-5


def test_size_independent_properties_full_batch():
    """1M Census queries (BASELINE.json config 2): properties that need no oracle."""
    import torch

    m, dm = G.model("census"), dev_model("census")
    n = 1_000_000
    stride = dm.desc_stride(L.DESC_RANGE_U8)
    desc = torch.zeros((n, stride), dtype=torch.uint8, device="cuda:0")
    out = torch.empty(n, dtype=torch.float32, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    dm.gen_range_queries_device(0, 0, n, 1, 14, desc.data_ptr(), st)
    dm.run_device(desc.data_ptr(), n, L.DESC_RANGE_U8, out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    p = out.cpu().numpy().astype(np.float64)
    assert np.all(p >= 0) and np.all(p <= 1 + 1e-5)
    # (a) permutation: results follow their queries bit for bit
    perm = torch.randperm(n, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(1))
    desc2 = desc[perm].contiguous()
    out2 = torch.empty_like(out)
    dm.run_device(desc2.data_ptr(), n, L.DESC_RANGE_U8, out2.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert torch.equal(out2, out[perm])
    # (b) additivity: splitting one column's range [lo,hi] into [lo,mid] + [mid+1,hi] splits the mass
    d = desc.cpu().numpy()
    lo, hi = unpack_ranges(m, d)
    v = 6  # a column with 18 states
    wide = np.nonzero(hi[:, v] > lo[:, v])[0][:200000]
    mid = (lo[wide, v] + hi[wide, v]) // 2
    left, right = d[wide].copy(), d[wide].copy()
    left[:, 2 * v + 1] = mid
    right[:, 2 * v] = mid + 1
    pl = dm.run_host(left, L.DESC_RANGE_U8).astype(np.float64)
    pr = dm.run_host(right, L.DESC_RANGE_U8).astype(np.float64)
    assert np.allclose(pl + pr, p[wide], rtol=2e-6, atol=1e-12)
    # (c) oracle on a seeded subsample of the full batch
    idx = np.random.default_rng(0).choice(n, 10000, replace=False)
    ref = O.dense_tree(m, O.range_weights(m, lo[idx], hi[idx]))
    assert_close(p[idx], ref, "census 1M subsample")


def test_edge_cases():
    m, dm = G.model("dmv"), dev_model("dmv")
    stride = dm.desc_stride(L.DESC_RANGE_U8)
    full = np.zeros((1, stride), dtype=np.uint8)
    full[0, 1:2 * m.n_nodes:2] = m.card - 1
    for kernel in KERNELS:
        # empty batch
        assert dm.run_host(np.zeros((0, stride), dtype=np.uint8), L.DESC_RANGE_U8, None, kernel).shape == (0,)
        # unconstrained query: total mass 1
        assert abs(dm.run_host(full, L.DESC_RANGE_U8, None, kernel)[0] - 1.0) < 1e-5
        # lo > hi selects nothing
        none = full.copy()
        none[0, 2 * 4], none[0, 2 * 4 + 1] = 5, 2
        assert dm.run_host(none, L.DESC_RANGE_U8, None, kernel)[0] == 0.0
        # hi beyond the domain is clamped
        over = full.copy()
        over[0, 1:2 * m.n_nodes:2] = 255
        assert abs(dm.run_host(over, L.DESC_RANGE_U8, None, kernel)[0] - 1.0) < 1e-5
        # single state of every column = product along the tree
        one = full.copy()
        one[0, 0:2 * m.n_nodes:2] = 0
        one[0, 1:2 * m.n_nodes:2] = 0
        ref = O.dense_tree(m, O.range_weights(m, np.zeros((1, m.n_nodes), int), np.zeros((1, m.n_nodes), int)))[0]
        assert_close(dm.run_host(one, L.DESC_RANGE_U8, None, kernel), [ref], "all-first-state")
    with pytest.raises(L.BayesCardError):
        dm.run_host(full, 99)


def test_replicas_shard_a_batch():
    """ShardedModel splits contiguously across replicas (two replicas on cuda:0 when only one GPU)."""
    import torch

    m = G.model("dmv")
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    sm = ShardedModel(m, devs)
    desc = sm.replicas[0].gen_range_queries_host(3, 0, 50001, 1, 6)
    got = sm.run_host(desc, L.DESC_RANGE_U8)
    one = dev_model("dmv").run_host(desc, L.DESC_RANGE_U8)
    assert np.array_equal(got, one)
    sm.close()


@pytest.mark.parametrize("shape", [(8, 300, 6000), (20, 10, 5000), (12, [7, 130, 33, 257, 64, 5, 90, 200, 17, 128, 40, 3], 4001),
                                   (1, 50, 100), (10, 1000, 1024), (40, 260, 1500)])
def test_batched_large_domain_path(shape):
    """K2 (BASELINE.json config 4): synthetic random trees, RANGE_U16 rows, FP32 SIMT and tensor-core GEMM edges.
    The (10, 1000) and (40, 260) shapes are the ones that expose accumulator truncation inside the tensor core: with
    one TMEM accumulator over the whole K loop they came out 3e-5 low (see UmmaCfg in k2_umma.cu)."""
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16, random_range_queries

    n_cols, card, nq = shape
    m = make_tree_model(n_cols, card, seed=n_cols)
    dm = DeviceModel(m, device=0, specialize=False)
    lo, hi = random_range_queries(m, nq, seed=1, kmax=10)
    lo[0], hi[0] = 0, m.card - 1                       # unconstrained query
    if nq > 2:
        lo[1, 0], hi[1, 0] = 3, 2                      # empty range on the root
    desc = pack_ranges_u16(lo, hi)
    ref = O.dense_tree(m, O.range_weights(m, lo, hi))
    assert abs(ref[0] - 1) < 1e-9
    for kernel in (L.KERNEL_GEMM_SIMT, L.KERNEL_GEMM):
        got = dm.run_host(desc, L.DESC_RANGE_U16, None, kernel)
        assert_close(got, ref, (shape, kernel))
    if int(m.card.max()) <= 256 and sum(-(-int(c) // 4) * 4 for c in m.card) * 4 < 40000:
        # the same queries through the generic kernel (RANGE_U16 rows)
        assert_close(dm.run_host(desc, L.DESC_RANGE_U16, None, L.KERNEL_GENERIC), ref, (shape, "generic"))
    # small workspace: many tiles
    import os
    os.environ["BC_K2_WORKSPACE_MB"] = "16"
    try:
        assert_close(dm.run_host(desc, L.DESC_RANGE_U16, None, L.KERNEL_GEMM_SIMT), ref, (shape, "tiled"))
    finally:
        del os.environ["BC_K2_WORKSPACE_MB"]
    dm.close()


@pytest.mark.parametrize("name", ["census", "dmv", "imdb1"])
def test_packed_wire_format_equals_sparse(name):
    """PACKED (bit-packed entries, 17 B per Census query) against SPARSE (CSR): identical BITS rows on the device, identical
    probabilities through the host pipeline -- ragged batch, many sub-chunks, empty queries, AUTO and the generic kernel."""
    import os

    import torch

    m, dm = G.model(name), dev_model(name)
    n = 128 * 37 + 51
    row_off, entries = dm.gen_sparse_queries_host(9, 3, n, 0, min(14, m.n_nodes))
    klen, blk, payload = dm.pack_sparse(row_off, entries)
    assert payload.nbytes + klen.nbytes + blk.nbytes < 0.62 * (row_off.nbytes + entries.nbytes)
    words = dm.desc_stride(L.DESC_BITS) // 4
    d_a = torch.zeros((n, words), dtype=torch.int32, device="cuda:0")
    d_b = torch.ones((n, words), dtype=torch.int32, device="cuda:0")
    t = lambda a, dt: torch.from_numpy(a.view(dt) if a.dtype != dt else a).cuda()
    d_off, d_ent = t(row_off.view(np.int32), np.int32), t(entries.view(np.int32), np.int32)
    d_klen, d_blk, d_pay = torch.from_numpy(klen).cuda(), t(blk.view(np.int32), np.int32), torch.from_numpy(payload).cuda()
    dm.expand_sparse_device(d_off.data_ptr(), d_ent.data_ptr(), n, d_a.data_ptr())
    L.check(L.lib().bc_expand_packed(dm._h, d_klen.data_ptr(), d_blk.data_ptr(), d_pay.data_ptr(), n, d_b.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(d_a, d_b)
    want = dm.run_sparse_host(row_off, entries)
    assert np.array_equal(dm.run_packed_host(klen, blk, payload), want)
    os.environ["BC_PACKED_CHUNK"] = "1024"   # 5 sub-chunks through the three pipeline slots
    try:
        assert np.array_equal(dm.run_packed_host(klen, blk, payload), want)
        assert np.array_equal(dm.run_packed_host(klen, blk, payload, None, L.KERNEL_GENERIC), dm.run_sparse_host(row_off, entries, None, L.KERNEL_GENERIC))
    finally:
        del os.environ["BC_PACKED_CHUNK"]
    assert dm.run_packed_host(klen[:0], blk[:1], payload[:8]).size == 0


@pytest.mark.gpu
def test_packed_submissions_in_flight():
    """bc_query_batch_packed_host_submit / bc_pipe_wait: several batches enqueued back to back (more chunks than pipeline slots,
    different sizes, so slots are reused and buffers re-allocated while work is in flight), results equal the synchronous call;
    a synchronous call of another wire format in between first drains the pipe; pageable result buffers are refused."""
    import os

    import torch

    dm = dev_model("census")
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    batches = []
    for i, n in enumerate([128 * 40 + 5, 128 * 9, 128 * 64 + 77, 300, 128 * 40 + 5]):
        row_off, entries = dm.gen_sparse_queries_host(20 + i, 0, n, 0, 12)
        klen, blk, payload = dm.pack_sparse(row_off, entries)
        want = dm.run_sparse_host(row_off, entries)
        batches.append((pin(klen), pin(blk), pin(payload), pin(np.zeros(n, dtype=np.float32)), want, (row_off, entries)))
    os.environ["BC_PACKED_CHUNK"] = "2048"   # 1-5 chunks per batch over three slots
    try:
        tickets = [dm.submit_packed_host(k, b, p, out=o) for k, b, p, o, _, _ in batches]
        assert tickets == sorted(tickets) and tickets[0] > 0
        dm.wait(tickets[1])
        assert np.array_equal(batches[0][3], batches[0][4]) and np.array_equal(batches[1][3], batches[1][4])
        dm.wait(tickets[-1])
        for k, b, p, o, want, _ in batches:
            assert np.array_equal(o, want)
        # two in flight, then a synchronous call of another format: it waits for them first
        for _, _, _, o, _, _ in batches:
            o[:] = -1
        t0 = dm.submit_packed_host(*batches[2][:3], out=batches[2][3])
        t1 = dm.submit_packed_host(*batches[0][:3], out=batches[0][3])
        again = dm.run_sparse_host(*batches[3][5])
        assert np.array_equal(again, batches[3][4])
        assert np.array_equal(batches[2][3], batches[2][4]) and np.array_equal(batches[0][3], batches[0][4])
        dm.wait(t1)
        dm.wait(t0)          # an old ticket: nothing to wait for
        dm.wait(0)
    finally:
        del os.environ["BC_PACKED_CHUNK"]
    with pytest.raises(L.BayesCardError, match="PINNED"):
        dm.submit_packed_host(*batches[1][:3], out=np.zeros(batches[1][0].size, dtype=np.float32))
    with pytest.raises(L.BayesCardError, match="not issued"):
        dm.wait(10 ** 12)


def _rare_equality_queries(m, nq, n_pred, seed, pairs=False):
    """Equality predicates on ``n_pred`` columns whose joint probability is far below the fp32 range: either many columns,
    each on a state from the rarer half of its CPT's row mass (``pairs=False``), or parent-child PAIRS where the child's
    state is one of the least likely given the parent's (``pairs=True``: ten predicates on wide domains reach 1e-60)."""
    rng = np.random.default_rng(seed)
    n = m.n_nodes
    lo = np.zeros((nq, n), dtype=np.int32)
    hi = np.tile(m.card.astype(np.int32) - 1, (nq, 1))
    rare = []
    for v in range(n):
        t = np.asarray(m.cpts[v], dtype=np.float64)
        mass = t if t.ndim == 1 else t.sum(axis=1)
        rare.append(np.argsort(mass)[: max(1, len(mass) // 2)])
    for q in range(nq):
        if not pairs:
            for v in rng.choice(n, size=n_pred, replace=False):
                lo[q, v] = hi[q, v] = int(rng.choice(rare[v]))
            continue
        used = set()
        while len(used) < n_pred:
            v = int(rng.integers(1, n))
            pa = int(m.parent[v])
            if v in used or (pa in used and len(used) + 1 > n_pred):
                continue
            if pa not in used:
                lo[q, pa] = hi[q, pa] = int(rng.integers(0, int(m.card[pa])))
                used.add(pa)
            col = np.asarray(m.cpts[v], dtype=np.float64)[:, lo[q, pa]]
            lo[q, v] = hi[q, v] = int(rng.choice(np.argsort(col)[: max(1, len(col) // 100)]))
            used.add(v)
    return lo, hi


@pytest.mark.parametrize("shape", [(100, 16, 60, "k1"), (50, 16, 40, "k1"), (50, 1000, 10, "k2"), (100, 300, 10, "k2")])
def test_results_below_the_fp32_range(shape):
    """VERDICT r1 missing #1: the reference multiplies in fp64 (ExactInference.py:157-177).  Wide synthetic trees with many
    equality predicates on rare states: fp64 oracle results far below 1e-45 (fp32 flushes them to 0); the *_scaled entry
    points carry a per-query exponent and must match to 1e-5 relative -- through the generic kernel (small domains) and the
    batched large-domain path (FP32 SIMT and tensor-core GEMM edges)."""
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16

    n_cols, card, n_pred, path = shape
    m = make_tree_model(n_cols, card, seed=n_cols + card, dtype=np.float32)
    dm = DeviceModel(m, device=0, specialize=False)
    nq = 300
    lo, hi = _rare_equality_queries(m, nq, n_pred, seed=3, pairs=path == "k2")
    lo[0], hi[0] = 0, m.card - 1                              # unconstrained: exactly 1, exponent 0
    lo[1, 0], hi[1, 0] = 3, 2                                 # empty range: exactly 0
    ref = O.dense_tree(m, O.range_weights(m, lo, hi))
    tiny = (ref > 0) & (ref < 1e-45)
    assert tiny.sum() > nq // 2, (shape, float(np.median(ref)))   # the case the plain fp32 result cannot represent
    desc = pack_ranges_u16(lo, hi)
    # (small domains: also the fused tensor-core kernel, which is what AUTO picks for them)
    kernels = (L.KERNEL_GENERIC, L.KERNEL_FUSED, L.KERNEL_AUTO) if path == "k1" else (L.KERNEL_GEMM_SIMT, L.KERNEL_GEMM, L.KERNEL_AUTO)
    for kernel in kernels:
        got = dm.run_host_scaled(desc, L.DESC_RANGE_U16, None, kernel)
        assert got.dtype == np.float64 and got[1] == 0.0 and abs(got[0] - 1.0) < 1e-5
        err = rel_err(got, ref)
        err[ref == 0] = 0
        assert err.max() <= RTOL, (shape, kernel, float(err.max()), float(ref[err.argmax()]))
    # the plain fp32 entry point flushes these to zero / denormals: that is what the scaled one is for
    plain = dm.run_host(desc, L.DESC_RANGE_U16, None, kernels[0])
    assert np.all(plain[tiny] < 1e-37)
    dm.close()


@pytest.mark.parametrize("shape", ["deep_hub_tail", "hub_112_columns", "two_hubs", "star_root_split"])
def test_scaled_results_through_the_fused_kernel(shape):
    """The exponent K3 carries through its epilogue (register accumulators, tensor-memory fallback, tail edges that belong to the
    previous tile of the CTA): several passes per CTA, results compared in fp64 -- mantissa * 2^exponent must equal the oracle
    whether or not the plain fp32 result would have survived."""
    from bayescard_b200.synth import make_tree_model, pack_ranges_u16, random_range_queries

    parent, cards = K3_SHAPES[shape]
    tm = make_tree_model(len(cards), cards, seed=5, dtype=np.float32, parent=parent)
    dm = DeviceModel(tm, device=0, specialize=False)
    nq = 50000
    lo, hi = random_range_queries(tm, nq, seed=9, kmax=len(cards))
    narrow = np.random.default_rng(1).random(lo.shape) < 0.5       # half of the predicates collapse to one (often rare) state
    hi = np.where(narrow, lo, hi)
    ref = O.dense_tree(tm, O.range_weights(tm, lo, hi))
    got = dm.run_host_scaled(pack_ranges_u16(lo, hi), L.DESC_RANGE_U16, None, L.KERNEL_FUSED)
    err = rel_err(got, ref)
    err[ref == 0] = 0
    assert err.max() <= RTOL, (shape, float(err.max()), float(ref[err.argmax()]))
    assert np.array_equal(got == 0, ref == 0)
    dm.close()


def test_in_process_multi_gpu_sharding():
    """north_star item 4: one host thread + stream set per device inside ONE process, results gathered in one host array.
    Needs two GPUs (skipped on a one-GPU box, where `test_sharded_model_same_device` still exercises the threading with
    two replicas on device 0): every wire format through ShardedModel equals the single-device result bit for bit, and the
    drop-in API routes through it."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    devs = list(range(min(torch.cuda.device_count(), 8)))
    for name in ("census", "imdb1"):
        m, dm = G.model(name), dev_model(name)
        sm = ShardedModel(m, devs, specialize=True)
        n = 128 * 53 + 77
        row_off, entries = dm.gen_sparse_queries_host(4, 11, n, 0, min(14, m.n_nodes))
        want = dm.run_sparse_host(row_off, entries)
        assert np.array_equal(sm.run_sparse_host(row_off, entries), want)
        klen, blk, payload = dm.pack_sparse(row_off, entries)
        assert np.array_equal(sm.run_packed_host(klen, blk, payload), want)
        desc = dm.gen_range_queries_host(4, 11, n, 0, min(14, m.n_nodes))
        assert np.array_equal(sm.run_host(desc, L.DESC_RANGE_U8), dm.run_host(desc, L.DESC_RANGE_U8))
        if name == "imdb1":
            pc = PredicateCompiler(m)
            cases = [r for r in G.load("infer_cases.json.gz")[name] if "error" not in r]
            decoded = [({k: list(v) for k, v in r["bins"].items()}, {k: np.asarray(v) for k, v in r["weights"].items()}) for r in cases]
            _, _, _, dense, mask = pc.pack(decoded, [r["fanout"] for r in cases], force_dense=True)
            ro, words = dense_to_wsparse(m, dense)
            assert np.array_equal(sm.run_wsparse_host(ro, words, mask), dm.run_wsparse_host(ro, words, mask))
        sm.close()
    bn = Bayescard_BN(G.model("dmv"), device=devs, infer_algo="exact-jit")
    bn.init_inference_method()
    one = Bayescard_BN(G.model("dmv"), device=0, infer_algo="exact-jit")
    one.init_inference_method()
    rows = G.load("dmv_workload.json.gz")["queries"][:400]
    qs = [parse_query_single_table(r["sql"], bn) for r in rows]
    assert np.array_equal(bn.query_batch(copy.deepcopy(qs)), one.query_batch(copy.deepcopy(qs)))
    bn.close()
    one.close()


def test_sharded_model_same_device():
    """ShardedModel's slicing of every wire format with two replicas on device 0 (runs on a one-GPU box)."""
    m, dm = G.model("census"), dev_model("census")
    sm = ShardedModel(m, [0, 0], specialize=True)
    n = 128 * 9 + 5
    row_off, entries = dm.gen_sparse_queries_host(6, 0, n, 0, 14)
    want = dm.run_sparse_host(row_off, entries)
    assert np.array_equal(sm.run_sparse_host(row_off, entries), want)
    klen, blk, payload = dm.pack_sparse(row_off, entries)
    assert np.array_equal(sm.run_packed_host(klen, blk, payload), want)
    assert sm.run_sparse_host(np.zeros(1, dtype=np.uint32), np.zeros(0, dtype=np.uint32)).size == 0
    # asynchronous submissions: two batches in flight on every replica
    import torch

    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    pk, pb, pp = pin(klen), pin(blk), pin(payload)
    outs = [pin(np.zeros(n, dtype=np.float32)) for _ in range(2)]
    t0 = sm.submit_packed_host(pk, pb, pp, out=outs[0])
    t1 = sm.submit_packed_host(pk, pb, pp, out=outs[1])
    sm.wait(t0)
    assert np.array_equal(outs[0], want)
    sm.wait(t1)
    assert np.array_equal(outs[1], want)
    sm.close()


def test_sharded_ranks_on_gpu():
    """The per-rank slice evaluator of bayescard_b200.sharding on the real device (single rank = whole batch;
    the two-rank split itself is covered on CPU by tests/test_sharding_gloo.py)."""
    from bayescard_b200.sharding import csr_slice, evaluate_sharded, rank_range

    m, dm = G.model("census"), dev_model("census")
    n = 30011
    row_off, entries = dm.gen_sparse_queries_host(2, 0, n, 1, 14)
    whole = dm.run_sparse_host(row_off, entries)

    def evaluate(a, b):
        ro, en = csr_slice(row_off, entries, a, b)
        return dm.run_sparse_host(ro, en)

    assert np.array_equal(evaluate_sharded(evaluate, n), whole)
    parts = [evaluate(*rank_range(n, r, 4)) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), whole)


# ------------------------------------------------------------------------------------ drop-in API
@pytest.mark.parametrize("name", ["dmv", "census"])
def test_workload_end_to_end_through_dropin_api(name):
    """BASELINE.json config 1: shipped model + real SQL through Bayescard_BN.query, as Testing/BN_testing.py."""
    import os

    bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", name + ".npz"), device=0)
    bn.infer_algo = "exact-jit"
    bn.init_inference_method()
    rows = G.load(f"{name}_workload.json.gz")["queries"]
    qerrs, parsed_all = [], []
    for r in rows:
        parsed = parse_query_single_table(r["sql"], bn)
        parsed_all.append(parsed)
        card = bn.query(parsed)
        ref, kind = r["card"]["value"], r["card"]["kind"]
        if kind == "array":
            assert isinstance(card, np.ndarray) and card.shape == (1,)
        elif kind == "int":
            assert isinstance(card, int) and card == ref
            continue
        else:
            assert np.ndim(card) == 0
        assert_close(np.asarray(card).reshape(-1), np.asarray(ref).reshape(-1), r["sql"][:60])
        qerrs.append(O.q_error(float(np.asarray(card).reshape(-1)[0]), r["true"]))
    pins = {"dmv": [1.0012, 1.0243, 1.0498, 1.3361, 7.6408],
            "census": [1.0635, 1.4844, 2.0523, 15.6009, 227.5043]}[name]
    assert np.allclose([np.percentile(qerrs, p) for p in (50, 90, 95, 99, 100)], pins, rtol=1e-4)
    # the batch entry point gives the same numbers in one go
    batch = bn.query_batch(parsed_all)
    ref_all = np.asarray([np.asarray(r["card"]["value"]).reshape(-1)[0] for r in rows], dtype=np.float64)
    assert_close(batch, ref_all, name + " batch")
    # ... and so does the SQL-text batch entry point (native parse + decode + pack, one launch per descriptor kind)
    sql_batch = bn.query_sql_batch([r["sql"] for r in rows])
    assert_close(sql_batch, ref_all, name + " sql batch")
    assert np.array_equal(sql_batch, batch)
    bn.close()


@pytest.mark.parametrize("i", range(5))
def test_imdb_query_and_expectation_api(i):
    import os

    bn = Bayescard_BN.load(os.path.join(G.GOLD, "models", f"imdb{i}.npz"), device=0, infer_algo="exact-jit")
    bn.init_inference_method()
    for r in G.load("imdb_cases.json.gz")[f"imdb{i}"]:
        q = G.unjson(r["query"])
        if "error" in r:
            with pytest.raises(Exception):
                bn.expectation(copy.deepcopy(q), list(r["fanout"]), return_prob=True)
            continue
        p, nrows = bn.expectation(copy.deepcopy(q), list(r["fanout"]), return_prob=True)
        assert nrows == r["nrows"]
        ref, kind = r["p"]["value"], r["p"]["kind"]
        if kind == "int":
            assert isinstance(p, int) and p == ref
            continue
        if kind == "array":
            assert isinstance(p, np.ndarray) and p.shape == (1,)
        assert_close(np.asarray(p).reshape(-1), np.asarray(ref).reshape(-1), str(q)[:80])
    bn.close()


def test_quirks_through_api():
    import os

    bns = {}
    for r in G.load("quirk_cases.json.gz")["cases"]:
        name = r["model"]
        if name not in bns:
            bns[name] = Bayescard_BN.load(os.path.join(G.GOLD, "models", name + ".npz"), device=0, infer_algo="exact-jit")
            bns[name].init_inference_method()
        bn, q = bns[name], G.unjson(r["query"])
        if "error" in r:
            with pytest.raises(Exception):
                bn.query(q)
            continue
        if r["kind"] == "decode":
            continue
        got = bn.query(q, return_prob=r["return_prob"]) if r["kind"] == "query" else \
            bn.expectation(q, list(r["fanout"]), return_prob=r["return_prob"])
        if r["return_prob"]:
            assert got[1] == r["nrows"]
            got = got[0]
        ref, kind = r["result"]["value"], r["result"]["kind"]
        if kind == "int":
            assert isinstance(got, int) and got == ref, (q, got)
        else:
            if kind == "array":
                assert isinstance(got, np.ndarray) and got.shape == (1,), (q, got)
            assert_close(np.asarray(got).reshape(-1), np.asarray(ref).reshape(-1), str(q))
    for bn in bns.values():
        bn.close()


def test_ensemble_cardinality():
    import os

    bns = {}
    for i in range(5):
        bns[i] = Bayescard_BN.load(os.path.join(G.GOLD, "models", f"imdb{i}.npz"), device=0, infer_algo="exact-jit")
        bns[i].init_inference_method()
    ens = BN_ensemble(None, bns)
    good, parsed_good = [], []
    for r in G.load("ensemble_cases.json.gz")["cases"]:
        tq = G.unjson(r["table_query"])
        if "error" in r:
            with pytest.raises(Exception):
                ens.cardinality(ens.parse_query_all([copy.deepcopy(tq)])[0])
            continue
        parsed = ens.parse_query_all([copy.deepcopy(tq)])[0]
        assert len(parsed) - 1 == r["n_factors_kept"]
        assert_close([ens.cardinality(parsed)], [r["card"]], "ensemble")
        good.append(r["card"])
        parsed_good.append(parsed)
    assert_close(ens.cardinality_batch(parsed_good), good, "ensemble batch")
    for bn in bns.values():
        bn.close()
