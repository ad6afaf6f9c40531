"""TEST INFRASTRUCTURE -- executes the reference's own hot-path files, unmodified, in this container.

Only usable where ``/root/reference`` exists (the build container); it does not travel to the GPU
box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may depend on it.  It is used by
``tests/golden/make_golden*.py`` to generate the committed golden vectors (outputs of the reference's
own code), which ``tests/test_oracle.py`` then uses to pin ``oracle/xmris_oracle.py`` bit-exactly.

How (strategy 2): ``import xmris`` cannot work here (its ``__init__`` pulls
matplotlib / pyAMARES, and xarray itself is not installed).  We therefore
  1. install ``xmris_b200.xarray_lite`` as ``sys.modules["xarray"]`` (only if real xarray is absent),
  2. pre-seed ``sys.modules["xmris"]`` with an empty namespace package whose ``__path__`` points at
     ``/root/reference/src/xmris`` (this skips ``xmris/__init__.py``),
  3. import ``xmris.core.config``, ``xmris.core.utils``, ``xmris.processing.fourier``,
     ``xmris.processing.fid``, ``xmris.processing.phasing`` and ``xmris.vendor.bruker`` from the read-only tree.
Numeric results depend only on numpy/scipy; the stand-in affects metadata plumbing only.
No reference source is copied into this repository.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("XMRIS_REFERENCE_SRC", "/root/reference/src/xmris")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "processing", "phasing.py"))


_loaded = None


def load_reference():
    """Return a namespace with the reference's hot-path modules (fid, fourier, phasing, config, utils, xr)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")

    try:
        import xarray as xr  # noqa: F401  (real xarray, if it ever becomes available)
    except ModuleNotFoundError:
        repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if repo_root not in sys.path:
            sys.path.insert(0, repo_root)
        from xmris_b200 import xarray_lite

        sys.modules["xarray"] = xarray_lite
        xr = xarray_lite

    def _pkg(name, path):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        mod.__package__ = name
        sys.modules[name] = mod
        return mod

    # Namespace shells: skip the reference's eager __init__ files (they import plotting/fitting stacks).
    _pkg("xmris", REFERENCE_SRC)
    _pkg("xmris.core", os.path.join(REFERENCE_SRC, "core"))
    _pkg("xmris.processing", os.path.join(REFERENCE_SRC, "processing"))
    _pkg("xmris.vendor", os.path.join(REFERENCE_SRC, "vendor"))

    ns = types.SimpleNamespace()
    ns.xr = xr
    ns.config = importlib.import_module("xmris.core.config")
    ns.utils = importlib.import_module("xmris.core.utils")
    ns.fourier = importlib.import_module("xmris.processing.fourier")
    ns.fid = importlib.import_module("xmris.processing.fid")
    ns.phasing = importlib.import_module("xmris.processing.phasing")
    ns.bruker = importlib.import_module("xmris.vendor.bruker")      # remove_digital_filter ("next" row N4)
    ns.baseline = importlib.import_module("xmris.processing.baseline")   # baseline_als (the step after autophase)
    _loaded = ns
    return ns
