"""Stage the UNMODIFIED reference (wuziniu/BayesCard) under ``baseline/_ref/`` so that ``bench.py --impl reference`` and
the ``cpu_baseline`` leg can run the reference's own exact-jit path on the GPU box, where ``/root/reference`` does not
exist.  ``baseline/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so it travels with the snapshot like the built ``.so`` files.  Called by ``__graft_entry__.build()``.

The reference is pure Python without a ``setup.py`` / ``pyproject.toml`` (``pip install /root/reference`` has nothing to
build), so "install" is a file copy of the packages the path imports -- byte for byte, nothing patched:
  Models/  Pgmpy/  Evaluation/{cardinality_estimation,utils}.py  Schemas/  DataPrepare/  DeepDBUtils/ (import closure of
  Models/*), the shipped model pickles and workloads under Benchmark/{DMV,Census,IMDB}.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = os.environ.get("BAYESCARD_REFERENCE", "/root/reference")

TREES = ("Models", "Pgmpy", "Schemas", "DataPrepare", "DeepDBUtils", "Inference")
FILES = ("__init__.py", "Evaluation/__init__.py", "Evaluation/cardinality_estimation.py", "Evaluation/utils.py",
         "Testing/BN_testing.py",
         "Benchmark/DMV/chow-liu_1.pkl", "Benchmark/DMV/query.sql", "Benchmark/Census/chow-liu_1.pkl",
         "Benchmark/Census/query.sql", "Benchmark/IMDB/job-light.sql",
         *[f"Benchmark/IMDB/{i}_chow-liu_1.pkl" for i in range(5)])


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def stage(src: str = SRC_DEFAULT, dest: str = DEST) -> bool:
    """Copy the path's import closure.  Returns False (and leaves ``dest`` alone) when the reference is not mounted."""
    if not os.path.isdir(os.path.join(src, "Pgmpy")):
        return False
    manifest = {}
    for t in TREES:
        s = os.path.join(src, t)
        if not os.path.isdir(s):
            continue
        d = os.path.join(dest, t)
        shutil.rmtree(d, ignore_errors=True)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
    for f in FILES:
        s = os.path.join(src, f)
        if not os.path.exists(s):
            continue
        d = os.path.join(dest, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    for root, _, names in os.walk(dest):
        for n in sorted(names):
            if n == "MANIFEST.json" or n.endswith(".pyc"):
                continue
            p = os.path.join(root, n)
            manifest[os.path.relpath(p, dest)] = _sha(p)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=0, sort_keys=True)
    return True


def staged(dest: str = DEST) -> bool:
    return os.path.isdir(os.path.join(dest, "Pgmpy")) and os.path.exists(os.path.join(dest, "Benchmark", "Census", "chow-liu_1.pkl"))


if __name__ == "__main__":
    print("staged" if stage() else "reference not mounted; baseline/_ref left as it is", DEST)
