"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference.

Run here (the reference cannot travel to the GPU box):  ``python tools/make_golden.py``.
Deterministic: fixed seeds, sorted iteration.  Outputs
  tests/golden/models/{dmv,census,imdb0..imdb4}.npz   flat TreeModel files converted from the shipped pickles
  tests/golden/dmv_workload.json.gz                   1965 queries: sql, true card, reference parse/decode/result
  tests/golden/census_workload.json.gz                468 queries, same
  tests/golden/imdb_cases.json.gz                     seeded query()/expectation() calls on the 5 IMDB models
  tests/golden/infer_cases.json.gz                    seeded infer_machine-level calls with fractional weights
  tests/golden/ensemble_cases.json.gz                 seeded BN_ensemble.parse_query_all + cardinality cases
  tests/golden/quirk_cases.json.gz                    decode / return-shape quirks of SURVEY.md section 8a
"""
from __future__ import annotations

import copy
import gzip
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import ref_harness as R  # noqa: E402
from bayescard_b200.loader import _jsonable, load_pickle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODELS = {
    "dmv": "Benchmark/DMV/chow-liu_1.pkl",
    "census": "Benchmark/Census/chow-liu_1.pkl",
    **{f"imdb{i}": f"Benchmark/IMDB/{i}_chow-liu_1.pkl" for i in range(5)},
}


def dump(name, obj):
    path = os.path.join(GOLD, name)
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode("utf-8"))
    print(f"  wrote {name}: {os.path.getsize(path)} bytes")


def result_record(x):
    """Reference results are np.float64 scalars, shape-(1,) arrays, or ints 0/1."""
    if isinstance(x, np.ndarray):
        return {"value": [float(v) for v in x.reshape(-1)], "kind": "array"}
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        return {"value": int(x), "kind": "int"}
    return {"value": float(x), "kind": "float"}


def decoded_record(q, nd):
    if q is None:
        return None
    return {"bins": {k: [int(b) for b in (v if isinstance(v, list) else [v])] for k, v in q.items()},
            "weights": {k: [float(w) for w in np.asarray(v, dtype=np.float64).reshape(-1)] for k, v in nd.items()}}


def convert_models():
    os.makedirs(os.path.join(GOLD, "models"), exist_ok=True)
    for name, rel in MODELS.items():
        tm = load_pickle(os.path.join(R.REFERENCE_ROOT, rel))
        bn = R.load_bn(rel)
        # cross-check the restricted-unpickler view against the object the reference itself builds
        assert tm.topo_names == list(bn.topological_order_node), name
        assert tm.infer_names[0] == bn.infer_machine.root, name
        for v, n in enumerate(tm.infer_names):
            ref = bn.cpds[bn.topological_order_node.index(n)].values
            assert np.array_equal(np.asarray(ref).reshape(tm.cpts[v].shape), tm.cpts[v]), (name, n)
        out = os.path.join(GOLD, "models", name + ".npz")
        tm.save(out)
        print(f"  model {name}: {tm.n_nodes} nodes, {tm.n_cpt_entries} CPT entries -> {os.path.getsize(out)} bytes")


def workload(name, model_rel, sql_rel):
    bn = R.load_bn(model_rel)
    rows = []
    for sql, true_card in R.read_workload(sql_rel):
        parsed = R.parse_query_single_table(sql, bn)
        q, nd = bn.query_decoding(copy.deepcopy(parsed))
        card = bn.query(copy.deepcopy(parsed))
        rows.append({"sql": sql, "true": true_card, "parsed": _jsonable(parsed),
                     "decoded": decoded_record(q, nd), "card": result_record(card)})
    dump(f"{name}_workload.json.gz", {"model": name, "queries": rows})


def _rand_raw_query(bn, rng, allow_fan_pred=True):
    """A raw-value predicate dict shaped like Evaluation/parse_query_imdb.py:18-50 emits (scalar or
    (lo, hi)), plus value lists and a few unknown values to exercise the decode corner cases."""
    cols_plain = [c for c in bn.node_names if c not in bn.fanout_attr]
    k = int(rng.integers(0, min(3, len(cols_plain)) + 1))
    chosen = list(rng.choice(cols_plain, size=k, replace=False)) if k else []
    if allow_fan_pred and bn.fanout_attr and rng.random() < 0.15:
        chosen.append(str(rng.choice(bn.fanout_attr)))
    q = {}
    for c in chosen:
        c = str(c)
        if bn.attr_type[c] == "continuous":
            lo, hi = float(bn.domain[c][0]), float(bn.domain[c][1])
            u = rng.random()
            if u < 0.6:
                a, b = sorted(rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo), size=2))
                q[c] = (float(a), float(b))
            elif u < 0.7:
                q[c] = (lo, hi)
            elif u < 0.8 and c in bn.n_distinct_mapping:
                q[c] = int(rng.choice(sorted(bn.n_distinct_mapping[c].keys())))
            else:
                q[c] = float(int(rng.uniform(lo, hi)))
        else:
            vals = sorted(bn.encoding[c].keys(), key=lambda x: (str(type(x)), x))
            u = rng.random()
            if u < 0.4 or len(vals) < 3:
                q[c] = R_py(vals[int(rng.integers(len(vals)))])
            elif u < 0.85:
                a, b = sorted(int(i) for i in rng.integers(len(vals), size=2))
                va, vb = R_py(vals[a]), R_py(vals[b])
                if isinstance(va, str):
                    q[c] = [R_py(x) for x in vals[a:b + 1]][:12]
                else:
                    q[c] = (va, vb)
            elif u < 0.95:
                n = int(rng.integers(1, min(8, len(vals)) + 1))
                q[c] = [R_py(vals[int(i)]) for i in rng.integers(len(vals), size=n)]
            else:
                q[c] = -987654 if not isinstance(vals[0], str) else "NO_SUCH_VALUE"
    return q


def R_py(v):
    return v.item() if isinstance(v, np.generic) else v


def imdb_cases(n_per_model=240):
    out = {}
    for i in range(5):
        bn = R.load_bn(MODELS[f"imdb{i}"])
        rng = np.random.default_rng(1000 + i)
        rows = []
        for _ in range(n_per_model):
            q = _rand_raw_query(bn, rng)
            nf = int(rng.integers(0, 4))
            fan = [str(x) for x in rng.choice(bn.fanout_attr, size=min(nf, len(bn.fanout_attr)), replace=False)] if nf else []
            rec = {"query": _jsonable(q), "fanout": fan}
            try:
                dq, dn = bn.query_decoding(copy.deepcopy(q))
                rec["decoded"] = decoded_record(dq, dn)
            except Exception as e:  # noqa: BLE001
                rec["decode_error"] = type(e).__name__
            try:
                if fan:
                    p, nrows = bn.expectation(copy.deepcopy(q), list(fan), return_prob=True)
                else:
                    p, nrows = bn.query(copy.deepcopy(q), return_prob=True)
                rec["p"] = result_record(p)
                rec["nrows"] = float(nrows)
            except Exception as e:  # noqa: BLE001
                rec["error"] = type(e).__name__
            rows.append(rec)
        out[f"imdb{i}"] = rows
        print(f"  imdb{i}: {len(rows)} cases, {sum('error' in r for r in rows)} raise in the reference")
    dump("imdb_cases.json.gz", out)


def infer_cases(n_per_model=200):
    """infer_machine.query / .expectation with already decoded (bins, fractional weights):
    BASELINE.json config 3's generator."""
    out = {}
    for name in MODELS:
        bn = R.load_bn(MODELS[name])
        rng = np.random.default_rng(hash(name) % 1000 + 77 if False else sum(map(ord, name)))
        topo = list(bn.topological_order_node)
        cards = {c.variable: int(c.values.shape[0]) for c in bn.cpds}
        fan_cols = [c for c in bn.fanout_attr if len(bn.fanouts[c]) == cards[c]]
        plain = [c for c in topo if c not in fan_cols]
        rows = []
        for _ in range(n_per_model):
            k = int(rng.integers(1, 4)) if name.startswith("imdb") else int(rng.integers(1, min(len(plain), 8) + 1))
            cols = [str(c) for c in rng.choice(plain, size=min(k, len(plain)), replace=False)]
            nf = int(rng.integers(0, 4)) if fan_cols else 0
            fan = [str(c) for c in rng.choice(fan_cols, size=min(nf, len(fan_cols)), replace=False)] if nf else []
            if fan and rng.random() < 0.15:
                cols.append(fan[0])  # a predicate on a fan-out column wins over the fan-out weights
            q, nd = {}, {}
            for c in cols:
                lo = int(rng.integers(cards[c]))
                hi = int(rng.integers(lo, cards[c]))
                bins = list(range(lo, hi + 1))
                if rng.random() < 0.3:
                    rng.shuffle(bins)
                q[c] = bins
                nd[c] = [float(w) for w in (rng.uniform(0.2, 1.0, size=len(bins)) if rng.random() < 0.5 else np.ones(len(bins)))]
            rec = {"bins": q, "weights": nd, "fanout": fan}
            try:
                qq = copy.deepcopy(q)
                ndd = {c: np.asarray(w) for c, w in nd.items()}
                r = bn.infer_machine.expectation(qq, list(fan), ndd) if fan else bn.infer_machine.query(qq, ndd)
                rec["p"] = result_record(r)
            except Exception as e:  # noqa: BLE001
                rec["error"] = type(e).__name__
            rows.append(rec)
        out[name] = rows
        print(f"  infer {name}: {len(rows)} cases, {sum('error' in r for r in rows)} raise")
    dump("infer_cases.json.gz", out)


def ensemble_cases(n=150):
    R.install()
    from Models.BN_ensemble_model import BN_ensemble

    bns = {i: R.load_bn(MODELS[f"imdb{i}"]) for i in range(5)}
    ens = BN_ensemble(None, bns=bns)
    rng = np.random.default_rng(4242)
    raw_all = []
    for _ in range(n):
        tq = [float(rng.integers(10 ** 5, 10 ** 9))]
        for _f in range(int(rng.integers(1, 5))):
            i = int(rng.integers(5))
            bn = bns[i]
            q = _rand_raw_query(bn, rng, allow_fan_pred=False)
            # the job-light planner only emits scalars and (lo, hi) tuples
            q = {k: v for k, v in q.items() if not isinstance(v, list)}
            nf = int(rng.integers(0, 3))
            fan = [str(x) for x in rng.choice(bn.fanout_attr, size=nf, replace=False)] if nf else []
            f = {"bn_index": i, "inverse": bool(rng.random() < 0.3), "query": q, "expectation": fan}
            tq.append(f)
            if rng.random() < 0.2:
                tq.append(copy.deepcopy(f))  # adjacent duplicate: dropped by parse_query_all
        raw_all.append(tq)
    rows = []
    for tq in raw_all:
        rec = {"table_query": _jsonable(tq)}
        try:
            parsed = ens.parse_query_all([copy.deepcopy(tq)])[0]
            rec["n_factors_kept"] = len(parsed) - 1
            rec["card"] = float(ens.cardinality(parsed))
        except Exception as e:  # noqa: BLE001
            rec["error"] = type(e).__name__
        rows.append(rec)
    print(f"  ensemble: {len(rows)} cases, {sum('error' in r for r in rows)} raise")
    dump("ensemble_cases.json.gz", {"cases": rows})


def quirk_cases():
    dmv = R.load_bn(MODELS["dmv"])
    imdb3 = R.load_bn(MODELS["imdb3"])
    imdb0 = R.load_bn(MODELS["imdb0"])
    cases = []

    def add(model, bn, kind, q, fan=None, return_prob=False):
        rec = {"model": model, "kind": kind, "query": _jsonable(q), "fanout": fan, "return_prob": return_prob}
        try:
            if kind == "query":
                r = bn.query(copy.deepcopy(q), return_prob=return_prob)
            elif kind == "expectation":
                r = bn.expectation(copy.deepcopy(q), list(fan), return_prob=return_prob)
            elif kind == "decode":
                dq, dn = bn.query_decoding(copy.deepcopy(q))
                rec["decoded"] = decoded_record(dq, dn)
                r = 0
            if return_prob:
                rec["result"] = result_record(r[0]); rec["nrows"] = float(r[1])
            else:
                rec["result"] = result_record(r)
        except Exception as e:  # noqa: BLE001
            rec["error"] = type(e).__name__
        cases.append(rec)

    add("dmv", dmv, "query", {})                                           # empty dict -> 0 (ExactInference.py:197)
    add("dmv", dmv, "query", {"Record_Type": "NOPE"})                      # unknown scalar -> 0
    add("dmv", dmv, "query", {"Record_Type": ["NOPE"]})                    # all-unknown list -> array([0.])
    add("dmv", dmv, "query", {"Record_Type": ["VEH", "NOPE"]})             # unknown members dropped
    add("dmv", dmv, "query", {"Record_Type": ["VEH"]})
    add("dmv", dmv, "query", {"Record_Type": "VEH"})                       # root-only -> shape (1,)
    add("dmv", dmv, "query", {"Record_Type": ["VEH", "TRL"]})
    add("dmv", dmv, "query", {"State": ["AA", "AE", "AK", "AA"]})          # duplicate bins capped at 1
    add("dmv", dmv, "query", {"Record_Type": []})                          # empty list -> (None, None) -> 0
    add("dmv", dmv, "query", {"Fuel_Type": ["GAS", ""]})
    add("dmv", dmv, "query", {"Model_Year": (2000, 2010)})                 # tuple on a categorical column
    add("dmv", dmv, "query", {"Model_Year": 2015})
    add("dmv", dmv, "query", {"Record_Type": "VEH"}, return_prob=True)
    add("dmv", dmv, "query", {"No_Such_Column": 1})
    kw = "movie_keyword.keyword_id"
    add("imdb3", imdb3, "query", {kw: (1, 134170)}, return_prob=True)      # whole domain loses bin 0
    add("imdb3", imdb3, "query", {kw: (50000, 1e9)}, return_prob=True)
    add("imdb3", imdb3, "query", {kw: (-5.0, 20.0)}, return_prob=True)
    add("imdb3", imdb3, "query", {kw: (134000.0, 1e9)}, return_prob=True)
    add("imdb3", imdb3, "query", {kw: 117}, return_prob=True)              # n_distinct_mapping multiplier
    add("imdb3", imdb3, "query", {kw: 398}, return_prob=True)
    add("imdb3", imdb3, "query", {kw: 5000}, return_prob=True)
    add("imdb3", imdb3, "query", {kw: (9000.0, 100.0)}, return_prob=True)  # l > r -> 0
    add("imdb3", imdb3, "decode", {kw: (1, 134170)})
    add("imdb3", imdb3, "decode", {kw: (50000, 1e9)})
    add("imdb3", imdb3, "decode", {kw: (300.0, 9000.0), "title.production_year": (1990, 2000)})
    add("imdb0", imdb0, "query", {"title.kind_id": [1.0], "title.production_year": (2000, 2010)}, return_prob=True)
    add("imdb0", imdb0, "expectation", {"title.kind_id": [1.0], "title.production_year": (2000, 2010)},
        ["title.mul_cast_info.movie_id"], return_prob=True)
    add("imdb0", imdb0, "expectation", {"title.kind_id": [1.0], "title.production_year": (2000, 2010)},
        ["title.mul_cast_info.movie_id", "title.mul_movie_keyword.movie_id"], return_prob=True)
    add("imdb0", imdb0, "expectation", {}, ["title.mul_cast_info.movie_id"], return_prob=True)
    add("imdb0", imdb0, "expectation", {}, [], return_prob=True)           # falls back to query({}) -> 0
    add("imdb0", imdb0, "expectation", {"title.kind_id": 1}, [], return_prob=False)
    add("imdb0", imdb0, "expectation", {"title.mul_cast_info.movie_id": (0, 3)},
        ["title.mul_cast_info.movie_id"], return_prob=True)               # predicate wins over fan-out
    add("imdb0", imdb0, "expectation", {"title.kind_id": 1}, ["title.mul_cast_info.movie_id"], return_prob=False)
    dump("quirk_cases.json.gz", {"cases": cases})
    for c in cases:
        print("   ", c["model"], c["kind"], str(c["query"])[:60], c.get("result", c.get("error")))


def main():
    os.makedirs(GOLD, exist_ok=True)
    print("models"); convert_models()
    print("workloads")
    workload("dmv", MODELS["dmv"], "Benchmark/DMV/query.sql")
    workload("census", MODELS["census"], "Benchmark/Census/query.sql")
    print("imdb"); imdb_cases()
    print("infer"); infer_cases()
    print("ensemble"); ensemble_cases()
    print("quirks"); quirk_cases()


if __name__ == "__main__":
    main()
