"""CPU: the C-ABI library builds, loads and exports every symbol ``include/xmris_b200.h`` declares (no compute calls)."""

import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "xmris_b200.h")
CSRC = os.path.join(ROOT, "xmris_b200", "csrc")


@pytest.fixture(scope="module")
def lib_path():
    path = os.path.join(CSRC, "libxmris_b200.so")
    if not os.path.isfile(path):
        subprocess.run(["make", "-C", CSRC, "-j", "8"], check=True, capture_output=True)
    return path


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xmr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib_path):
    from xmris_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 10
    lib = ctypes.CDLL(lib_path)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in xmris_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_library_loads_without_gpu_and_reports_version(lib_path):
    from xmris_b200 import _lib

    lib = _lib.load()
    assert lib.xmr_version() >= 100
    assert isinstance(lib.xmr_last_error(), bytes)
    assert lib.xmr_autophase_workspace_bytes() > 0


def test_argument_validation_needs_no_gpu(lib_path):
    """Argument errors are detected before any CUDA call and map to ValueError like the reference's validation."""
    from xmris_b200 import _lib

    lib = _lib.load()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(8)   # never dereferenced: validation fails first
    rc = lib.xmr_fid_to_spectrum_c64(one, one, 4, 1972, 1972, 0, 0, null, null, 1.0, 0, 0, 0, null, null, 0, 0.0, 0.0, null)
    assert rc == _lib.XMR_ERR_UNSUPPORTED_N
    with pytest.raises(ValueError, match="power of two"):
        _lib.check(rc)
    rc = lib.xmr_fid_to_spectrum_c64(one, one, 4, 4096, 2048, 0, 0, null, null, 1.0, 0, 0, 0, null, null, 0, 0.0, 0.0, null)
    assert rc == _lib.XMR_ERR_BAD_ARG
    rc = lib.xmr_fid_to_spectrum_c64(null, one, 4, 64, 64, 0, 0, null, null, 1.0, 0, 0, 0, null, null, 0, 0.0, 0.0, null)
    assert rc == _lib.XMR_ERR_BAD_ARG
    rc = lib.xmr_autophase_search_c64(one, 4096, 0.0, 1.0, 7, 0, 1, 0, one, one, null)
    assert rc == _lib.XMR_ERR_BAD_ARG
    assert lib.xmr_zero_fill_c64(one, one, 2, 10, 5, 0, null) == _lib.XMR_ERR_BAD_ARG


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from xmris_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.XmrisB200LibraryError, match="no CPU fallback"):
        _lib.load()
