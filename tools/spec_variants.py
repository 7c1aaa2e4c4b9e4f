#!/usr/bin/env python
"""Build ahead-of-time cubins of K-spec code-generator VARIANTS (the BC_SPEC_* knobs of spec_codegen.cc) in parallel.

    python tools/spec_variants.py dmv,imdb1 "SYNC_EVERY=1024" "SYNC_EVERY=1024 THREADS=256 MIN_BLOCKS=1" ...

Every variant hashes to its own cache entry, so `BC_SPEC_SYNC_EVERY=1024 python bench.py --model dmv` on the GPU box
finds the matching image.  Prints ptxas' register / spill line per variant.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys; sys.path.insert(0, %r)
from bayescard_b200 import aot
from bayescard_b200.loader import TreeModel
import os
print(aot.build_cubin(TreeModel.load(os.path.join(%r, 'tests', 'golden', 'models', sys.argv[1] + '.npz'))))
"""


def build(model, variant):
    env = dict(os.environ)
    for kv in variant.split():
        k, v = kv.split("=")
        env["BC_SPEC_" + k] = v
    r = subprocess.run([sys.executable, "-c", CODE % (ROOT, ROOT), model], env=env, capture_output=True, text=True)
    return model, variant, (r.stdout.strip() or r.stderr.strip()[-400:])


def main():
    models = sys.argv[1].split(",")
    variants = sys.argv[2:] or [""]
    jobs = [(m, v) for m in models for v in variants]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for m, v, out in ex.map(lambda j: build(*j), jobs):
            print(f"{m:8s} [{v}] -> {out}", flush=True)


if __name__ == "__main__":
    main()
