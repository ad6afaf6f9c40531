"""ctypes binding of ``libxmris_b200.so`` (the C ABI declared in ``include/xmris_b200.h``).

There is no CPU fallback: if the shared library is missing or does not load, every operation raises
``XmrisB200LibraryError`` telling the user to build it (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C xmris_b200/csrc``).
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libxmris_b200.so")

XMR_OK = 0
XMR_ERR_BAD_ARG = 1
XMR_ERR_UNSUPPORTED_N = 2
XMR_ERR_CUDA = 3

WIN_NONE, WIN_TABLE, WIN_SEPARABLE = 0, 1, 2
PHASE_NONE, PHASE_UNIFORM = 0, 1
METHOD_ACME, METHOD_PEAK_MINIMA, METHOD_POSITIVITY = 0, 1, 2
METHODS = {"acme": METHOD_ACME, "peak_minima": METHOD_PEAK_MINIMA, "positivity": METHOD_POSITIVITY}


class XmrisB200LibraryError(RuntimeError):
    pass


_vp, _i, _i64, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); must list every symbol of include/xmris_b200.h (tests/test_abi.py checks this)
SIGNATURES = {
    "xmr_version": (_i, []),
    "xmr_last_error": (ctypes.c_char_p, []),
    "xmr_fid_to_spectrum_c64": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _i, _d, _d, _vp]),
    "xmr_fid_absmax_pruned_c64": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _i, _vp]),
    "xmr_zero_fill_c64": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "xmr_roll_rows_c64": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "xmr_scale_rows_c64": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "xmr_rotate_rows_c64": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "xmr_rotate_rows_shift_c64": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _i, _i, _vp]),
    "xmr_phase_each_c64": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "xmr_global_argmax": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "xmr_row_absmax_c64": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "xmr_chain_single_workspace_bytes": (_i64, [_i64, _i]),
    "xmr_chain_single_graph_launches": (_i64, []),
    "xmr_chain_single_last_timing": (_i, [_vp]),
    "xmr_chain_single_slot_bytes": (_i64, [_i]),
    "xmr_chain_single_front_c64": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i64, _vp, _vp]),
    "xmr_chain_single_back_c64": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "xmr_chain_single_dev_c64": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "xmr_baseline_als_workspace_bytes": (_i64, [_i64, _i]),
    "xmr_baseline_als": (_i, [_vp, _i, _vp, _i64, _i, _d, _d, _i, _vp, _i64, _vp]),
    "xmr_autophase_workspace_bytes": (_i64, []),
    "xmr_autophase_search_tuning": (_i, [_d, _d, _i, _i, _i, _i, _d]),
    "xmr_autophase_search_polish": (_i, [_i, _i, _i]),
    "xmr_autophase_search_c64": (_i, [_vp, _i, _d, _d, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmr_autophase_score_c64": (_i, [_vp, _i, _d, _d, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "xmr_chain_each_c64": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _f, _i, _d, _i, _d, _i, _i, _i,
                                _vp, _vp, _vp, _vp, _vp]),

    "xmr_chain_host_c64": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "xmr_host_workspace_release": (_i, []),
    "xmr_host_chain_resident_limit": (_i, [_i64]),
}


class HostChainDesc(ctypes.Structure):
    """``xmr_host_chain_desc`` of include/xmris_b200.h."""

    _fields_ = [("n_in", _i), ("n_out", _i), ("pad_left", _i), ("window_host", _vp), ("scale", _f),
                ("autophase_mode", _i), ("method", _i), ("index_width", _i), ("p0_only", _i), ("fixed_pivot", _i),
                ("u0_fixed", _d), ("fixed_target", _i), ("du", _d), ("chunk", _i)]


_lib = None


def load():
    """Load the shared library once; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise XmrisB200LibraryError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `make -C xmris_b200/csrc -j8` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). xmris_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise XmrisB200LibraryError(f"could not load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError -> a symbol of the header is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    """Map a C status to the exception type the reference would raise (ValueError for argument problems)."""
    if rc == XMR_OK:
        return
    msg = load().xmr_last_error().decode("utf-8", "replace")
    if rc in (XMR_ERR_BAD_ARG, XMR_ERR_UNSUPPORTED_N):
        raise ValueError(f"xmris_b200: {msg}")
    raise RuntimeError(f"xmris_b200: {msg} (status {rc}) -- a working CUDA device is required; there is no CPU fallback")
