#!/usr/bin/env python
"""Golden fixture for the CPT-fitting path: the UNMODIFIED reference's ``BayesianModel.fit`` (pgmpy MLE) on a seeded
synthetic DMV-shaped table.  Runs in the build container only (imports /root/reference through tools/ref_harness.py);
the fixture travels: ``tests/golden/fit_dmv_shaped.npz`` = table (uint8, topological column order), parent, card and
the reference's CPDs.

The table is an ancestral sample of the shipped DMV tree (so it has DMV's skew and its structural zeros) with two
twist that exercises the estimator's corner case: the rows of one parent state are removed and the state sets are
passed explicitly (``state_names``), so that state's count column is all zero and becomes uniform (MLE.py:77-79).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def sample_table(tm, n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((n, tm.n_nodes), dtype=np.int64)
    for v in range(tm.n_nodes):
        t = np.asarray(tm.cpts[v], dtype=np.float64)
        if tm.parent[v] < 0:
            cdf = np.cumsum(t / t.sum())
            out[:, v] = np.minimum(np.searchsorted(cdf, rng.random(n)), len(cdf) - 1)
        else:
            cdf = np.cumsum(t / t.sum(axis=0), axis=0)  # [card, card_pa]
            u = rng.random(n)
            pa = out[:, tm.parent[v]]
            out[:, v] = np.minimum((cdf[:, pa] < u[None, :]).sum(axis=0), t.shape[0] - 1)
    return out


def main():
    import golden_util as G
    import pandas as pd
    import ref_harness as R

    R.install()
    from Pgmpy.models import BayesianModel

    tm = G.model("dmv")
    n = 60000
    table = sample_table(tm, n, seed=11)
    # an unobserved parent state: drop every row where column 1 (parent of several columns) is in state 3; with the
    # state sets given explicitly, its children get an all-zero count column there, which MLE.py:77-79 turns uniform
    table = table[table[:, 1] != 3]
    kids = [v for v in range(tm.n_nodes) if tm.parent[v] == 1]
    names = [f"c{v}" for v in range(tm.n_nodes)]
    df = pd.DataFrame({names[v]: table[:, v] for v in range(tm.n_nodes)})
    spec = [(names[int(tm.parent[v])], names[v]) for v in range(1, tm.n_nodes)]
    model = BayesianModel(spec)
    model.fit(df, state_names={names[v]: list(range(int(tm.card[v]))) for v in range(tm.n_nodes)})
    cpds = {}
    for cpd in model.get_cpds():
        v = names.index(cpd.variable)
        vals = np.asarray(cpd.values, dtype=np.float64)
        assert vals.shape[0] == int(tm.card[v]), (cpd.variable, vals.shape)
        cpds[v] = vals
    zero_cols = 0
    for v in range(1, tm.n_nodes):
        pa = int(tm.parent[v])
        seen = np.zeros(int(tm.card[pa]), dtype=bool)
        seen[np.unique(table[:, pa])] = True
        zero_cols += int((~seen).sum())
    out = os.path.join(ROOT, "tests", "golden", "fit_dmv_shaped.npz")
    np.savez_compressed(out, table=table.astype(np.uint8), parent=tm.parent.astype(np.int32), card=tm.card.astype(np.int32),
                        **{f"cpd_{v}": cpds[v] for v in range(tm.n_nodes)})
    print(f"{out}: {table.shape[0]} rows x {table.shape[1]} columns, {os.path.getsize(out) / 1e3:.0f} KB, kids of column 1: {kids}")


if __name__ == "__main__":
    main()
