"""Import the UNMODIFIED reference (wuziniu/BayesCard) from /root/reference in this container.

Test/fixture infrastructure only.  Nothing in the product package imports this module and nothing
on the GPU box can (``/root/reference`` does not exist there).  It is used by
``tools/make_golden.py`` to generate the golden vectors under ``tests/golden/`` and by the
``not gpu`` tests that re-validate the oracle against the live reference when it is present.

Four shims are needed, none of which touches arithmetic:
  1. a stub ``pomegranate`` module (training only, ``Models/BN_single_model.py:1``);
  2. a meta-path alias ``pgmpy.* -> Pgmpy.*`` (the vendored tree imports upstream names, e.g.
     ``Pgmpy/factors/distributions/CustomDistribution.py:4``);
  3. ``np.product / np.Inf / np.infty`` for numpy >= 2 (``Pgmpy/factors/discrete/DiscreteFactor.py:73``,
     ``Models/Bayescard_BN.py:211-212``, ``Evaluation/cardinality_estimation.py:65-71``);
  4. a stub ``sqlparse`` module (``Evaluation/utils.py:5-6``; multi-table parser only).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import os
import pickle
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _default_root() -> str:
    """``/root/reference`` in the build container, the staged byte-for-byte copy ``baseline/_ref`` on the GPU box
    (``baseline/stage_reference.py``)."""
    env = os.environ.get("BAYESCARD_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/Pgmpy"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Pgmpy"))


class _PgmpyAlias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolve ``pgmpy.x.y`` to the already vendored ``Pgmpy.x.y`` module object."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname == "pgmpy" or fullname.startswith("pgmpy."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        real = importlib.import_module("Pgmpy" + spec.name[len("pgmpy"):])
        return real

    def exec_module(self, module):
        return None


_installed = False


def install() -> None:
    """Put the reference on sys.path with the four shims.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import numpy as np

    if not hasattr(np, "product"):
        np.product = np.prod
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if not hasattr(np, "infty"):
        np.infty = np.inf
    for name in ("pomegranate", "sqlparse"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.__dict__["__stub__"] = True
            sys.modules[name] = stub
    sq = sys.modules["sqlparse"]
    if getattr(sq, "__stub__", False):
        # Evaluation/utils.py does `from sqlparse.tokens import Token` style imports
        tok = types.ModuleType("sqlparse.tokens")
        tok.Token = object()
        sys.modules.setdefault("sqlparse.tokens", tok)
        sq.tokens = tok
    sys.meta_path.insert(0, _PgmpyAlias())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_bn(rel_path: str, infer_algo: str = "exact-jit"):
    """pickle.load a shipped model and initialise it exactly as Testing/BN_testing.py:11-15 does."""
    install()
    with open(os.path.join(REFERENCE_ROOT, rel_path), "rb") as f:
        bn = pickle.load(f)
    bn.infer_algo = infer_algo
    bn.init_inference_method()
    return bn


def parse_query_single_table(sql: str, bn):
    install()
    from Evaluation.cardinality_estimation import parse_query_single_table as p

    return p(sql, bn)


def read_workload(rel_path: str):
    """Lines of ``<sql>||<true cardinality>`` (Testing/BN_testing.py:21-23)."""
    out = []
    with open(os.path.join(REFERENCE_ROOT, rel_path)) as f:
        for line in f.readlines():
            true_card = int(line.split("||")[-1])
            sql = line.split("||")[0].strip()
            out.append((sql, true_card))
    return out
