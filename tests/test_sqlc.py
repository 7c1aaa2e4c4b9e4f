"""The native batched SQL -> descriptor compiler (bc_sqlc_*, host only) against its Python mirror.

The mirror (bayescard_b200.sql_front + bayescard_b200.decode) is itself pinned to the unmodified reference by
tests/test_host_logic.py::test_sql_and_decode_match_reference, so equality of descriptor ROWS here pins the native
compiler to the reference's parse_query_single_table + query_decoding.  No GPU is involved.
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
from bayescard_b200.sql_front import parse_query_single_table  # noqa: E402
from bayescard_b200.sqlc import SqlBatchCompiler  # noqa: E402


def mirror_row(sc, sql):
    """('bits'|'dense', row) / ('zero',) / ('error', type) through the Python path."""
    tm, pc = sc.tm, sc.py
    try:
        q = parse_query_single_table(sql, sc._view)
        b, w = pc.decode(q)
    except Exception as e:  # noqa: BLE001
        return ("error", type(e).__name__)
    if b is None or not any(k in tm._index for k in b):
        return ("zero",)
    bi, bd, di, dd, _ = pc.pack([(b, w)])
    return ("bits", bd[0]) if len(bi) else ("dense", dd[0])


def check_batch(sc, sqls, allow_python=True):
    kind, bits, dense, didx = sc.compile_native(sqls)
    assert len(didx) == int((kind == L.SQLC_DENSE).sum())
    dense_of = {int(q): k for k, q in enumerate(didx)}
    n_py = 0
    for i, sql in enumerate(sqls):
        ref = mirror_row(sc, sql)
        k = int(kind[i])
        if k == L.SQLC_PYTHON:
            n_py += 1
            assert allow_python, sql
            continue
        assert ref[0] != "error", (sql, ref, k)  # whatever the reference raises on must be left to Python
        if k == L.SQLC_ZERO:
            assert ref[0] == "zero", (sql, ref[0])
        elif k == L.SQLC_BITS:
            assert ref[0] == "bits", (sql, ref[0])
            assert np.array_equal(bits[i], ref[1]), sql
        else:
            assert k == L.SQLC_DENSE and ref[0] == "dense", (sql, ref[0], k)
            assert np.array_equal(dense[dense_of[i]], ref[1]), sql
    return n_py


@pytest.mark.parametrize("name", ["dmv", "census"])
def test_shipped_workloads_compile_natively(name):
    sc = SqlBatchCompiler(G.model(name))
    sqls = [r["sql"] for r in G.load(f"{name}_workload.json.gz")["queries"]]
    assert check_batch(sc, sqls, allow_python=False) == 0
    # geometry agrees with the Python packer (and therefore with the C ABI of the kernels)
    _, row_bytes, _, width = sc.py.geometry()
    assert (sc.bits_stride, sc.dense_width) == (row_bytes, width)
    sc.close()


def _fmt(v, rng):
    if isinstance(v, str):
        r = rng.random()
        return v if r < 0.7 else ("'" + v + "'" if r < 0.85 else '"' + v + '"')
    if float(v) == int(v) and rng.random() < 0.6:
        return str(int(v))
    return repr(float(v))


def fuzz_sql(tm, rng, n):
    cols = list(tm.attr_type.keys())
    out = []
    for _ in range(n):
        preds = []
        for _ in range(int(rng.integers(1, 5))):
            if rng.random() < 0.05:
                preds.append("no_such_column = 3")
                continue
            c = cols[int(rng.integers(len(cols)))]
            if tm.attr_type[c] == "continuous":
                lo, hi = tm.domain[c]
                x = float(rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo)))
                op = ["<", ">", "<=", ">=", "=", "=="][int(rng.integers(6))]
                preds.append(f"{c} {op} {int(x) if rng.random() < 0.5 else x}")
                continue
            enc = tm.encoding.get(c) or {}
            vals = list(enc.keys()) or [0]
            r = rng.random()
            pick = lambda: vals[int(rng.integers(len(vals)))] if rng.random() < 0.9 else ("NOPE" if isinstance(vals[0], str) else 987654)
            if r < 0.35:
                items = [_fmt(pick(), rng) for _ in range(int(rng.integers(0, 5)))]
                body = ", ".join(items) + ("," if rng.random() < 0.03 else "")
                preds.append(f"{c} IN [{body}]")
            elif r < 0.7 or (isinstance(vals[0], str) and rng.random() < 0.9):  # inequalities on text columns raise
                preds.append(f"{c} {'=' if rng.random() < 0.7 else '=='} {_fmt(pick(), rng)}")
            else:
                op = ["<", ">", "<=", ">="][int(rng.integers(4))]
                v = pick()
                preds.append(f"{c}{op}{_fmt(v, rng)}" if rng.random() < 0.3 else f"{c} {op} {_fmt(v, rng)}")
        weird = rng.random()
        if weird < 0.02:
            preds.append("a_b")  # no operator: NameError in the reference
        elif weird < 0.04:
            preds.append(f"{cols[0]} = 1_000")
        elif weird < 0.06:
            preds.append(f"{cols[0]} <> 3")
        elif weird < 0.08:
            preds.append(f"{cols[0]} = inf")
        out.append("SELECT COUNT(*) FROM t WHERE " + " AND ".join(preds))
    return out


@pytest.mark.parametrize("name", G.MODEL_NAMES)
def test_fuzzed_sql_rows_equal_python_mirror(name):
    tm = G.model(name)
    sc = SqlBatchCompiler(tm)
    rng = np.random.default_rng(G.MODEL_NAMES.index(name) + 5)
    sqls = fuzz_sql(tm, rng, 1500)
    n_py = check_batch(sc, sqls)
    assert n_py < 0.5 * len(sqls)  # the native path carries the bulk
    # the full compile() (native + Python for the declined ones) reproduces the mirror for every query that the
    # reference can evaluate at all
    ok = [s for s in sqls if mirror_row(sc, s)[0] != "error"]
    bits_idx, bits_rows, dense_idx, dense_rows, zero = sc.compile(ok)
    assert len(bits_idx) + len(dense_idx) + len(zero) == len(ok)
    where = {int(i): ("bits", r) for i, r in zip(bits_idx, bits_rows)}
    where.update({int(i): ("dense", r) for i, r in zip(dense_idx, dense_rows)})
    where.update({int(i): ("zero",) for i in zero})
    for i, s in enumerate(ok):
        ref = mirror_row(sc, s)
        assert where[i][0] == ref[0], s
        if ref[0] != "zero":
            assert np.array_equal(where[i][1], ref[1]), s
    sc.close()


def test_hand_written_cases():
    tm = G.model("dmv")
    sc = SqlBatchCompiler(tm)
    sqls = [
        "SELECT COUNT(*) FROM dmv WHERE Record_Type = VEH",
        "SELECT COUNT(*) FROM dmv WHERE Record_Type IN [VEH, NOPE]",          # unknown member silently dropped
        "SELECT COUNT(*) FROM dmv WHERE Record_Type IN [NOPE1, NOPE2]",       # all unknown: a row selecting nothing
        "SELECT COUNT(*) FROM dmv WHERE Record_Type IN []",                   # empty list: undecodable -> 0
        "SELECT COUNT(*) FROM dmv WHERE Record_Type = 'VEH'",                 # quotes are part of the bare word
        "SELECT COUNT(*) FROM dmv WHERE Model_Year >= 2005 AND Model_Year < 2010",
        "SELECT COUNT(*) FROM dmv WHERE Model_Year = 2005 AND Model_Year = 2006",  # intersection is empty
        "SELECT COUNT(*) FROM dmv WHERE nothing = 1",                         # no column of the BN
        "SELECT COUNT(*) FROM dmv WHERE Model_Year>=2005",
        "SELECT COUNT(*) FROM dmv WHERE Record_Type > 3",                     # string domain: the reference raises
        "SELECT COUNT(*) FROM dmv",                                           # no WHERE
    ]
    kind, *_ = sc.compile_native(sqls)
    assert kind[3] == L.SQLC_ZERO and kind[6] == L.SQLC_ZERO and kind[7] == L.SQLC_ZERO
    assert kind[9] == L.SQLC_PYTHON and kind[10] == L.SQLC_PYTHON
    assert all(k in (L.SQLC_BITS, L.SQLC_DENSE) for k in kind[[0, 1, 2, 4, 5, 8]])
    check_batch(sc, sqls)
    sc.close()
