"""Ahead-of-time build of the specialised kernels (K-spec) for known models.

``build_cubin`` runs the library's code generator on a HOST-ONLY model (no GPU needed), assembles the
PTX with ``ptxas -arch=sm_100a`` and stores ``_spec_cache/spec_<hash>.cubin`` -- exactly the file
``bc_model_specialize`` looks for before it hands the PTX to the driver JIT.  ``__graft_entry__.build()``
does this for the models under ``tests/golden/models`` so the GPU box does not pay JIT time for them.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

from . import _lib as L
from .engine import DeviceModel
from .loader import TreeModel

PTXAS = os.environ.get("PTXAS", "/usr/local/cuda/bin/ptxas")


def cubin_path(tm: TreeModel, cache_dir: str = None) -> str:
    dm = DeviceModel(tm, device=-1, specialize=False)
    try:
        return os.path.join(cache_dir or L.SPEC_CACHE_DIR, "spec_%016x.cubin" % dm.spec_hash())
    finally:
        dm.close()


def build_cubin(tm: TreeModel, cache_dir: str = None, force: bool = False, keep_source: bool = False) -> str:
    cache_dir = cache_dir or L.SPEC_CACHE_DIR
    os.makedirs(cache_dir, exist_ok=True)
    dm = DeviceModel(tm, device=-1, specialize=False)
    try:
        out = os.path.join(cache_dir, "spec_%016x.cubin" % dm.spec_hash())
        if os.path.exists(out) and not force:
            return out
        src = dm.spec_source()
    finally:
        dm.close()
    with tempfile.TemporaryDirectory() as td:
        ptx = os.path.join(td, "spec.ptx")
        with open(ptx, "w") as f:
            f.write(src)
        if keep_source:
            with open(out[:-6] + ".ptx", "w") as f:
                f.write(src)
        cmd = [PTXAS, "-arch=sm_100a", "-O3", "-o", out + ".tmp", ptx]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise L.BayesCardError("ptxas failed for the specialised kernel:\n" + res.stderr[-2000:])
        os.replace(out + ".tmp", out)
    return out
