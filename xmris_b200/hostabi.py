"""numpy in, numpy out: the whole chain through ONE C-ABI call on host buffers (``xmr_chain_host_c64``).

No torch anywhere on this path: ctypes + numpy only, exactly what a maintainer-side binding inside the reference would
use.  The C side pipelines H2D / kernels / D2H over voxel chunks and page-locks pageable buffers for the
duration of the call.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, chain


def chain_host(fid, time_coord, target_points=None, position="end", lb=None, autophase=None, out=None, chunk=0, gb=None):
    """``zero_fill -> apodize_exp -> to_spectrum [-> autophase]`` on a host array ``fid[..., n_in]`` (complex64).

    ``autophase``: ``None`` (stop after ``to_spectrum``) or a dict with the reference's keyword names
    (``method, mode, peak_width, target_coord, p0_only``).  Returns ``(spectrum ndarray, freqs, info or None)``.
    """
    lib = _lib.load()
    fid = np.ascontiguousarray(fid, dtype=np.complex64)
    n_in = fid.shape[-1]
    bshape = fid.shape[:-1]
    batch = int(np.prod(bshape)) if bshape else 1
    geo = chain.chain_geometry(n_in, time_coord, target_points, position, lb, gb)
    n_out, freqs = geo["n_out"], geo["freqs"]
    if n_out not in chain.D.SUPPORTED_N:
        raise ValueError(f"xmris_b200: the host chain needs a power-of-two length in [16, 8192], got {n_out}")
    if out is None:
        out = np.empty(bshape + (n_out,), dtype=np.complex64)
    elif out.shape != bshape + (n_out,) or out.dtype != np.complex64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous complex64 array of the output shape")
    desc = _lib.HostChainDesc()
    desc.n_in, desc.n_out, desc.pad_left = n_in, n_out, int(geo["pad_left"])
    window = None
    if geo["window"] is not None:
        window = np.ascontiguousarray(geo["window"], dtype=np.float64)
        desc.window_host = window.ctypes.data
    desc.scale = 0.0
    desc.chunk = int(chunk)
    mode = 0
    info = None
    result = np.zeros(4, dtype=np.float64)
    p0 = p1 = piv = fun = None
    if autophase is not None:
        kw = dict(autophase)
        m = kw.get("mode", "single")
        if m not in ("single", "all"):
            raise ValueError("Mode must be 'single' or 'all'.")
        method = kw.get("method", "acme")
        if method not in _lib.METHODS:
            raise ValueError("Method must be 'acme', 'peak_minima', or 'positivity'")
        mode = 1 if m == "single" else 2
        x_range = float(freqs.max()) - float(freqs.min())
        desc.du = (freqs[-1] - freqs[0]) / (n_out - 1) / x_range
        step = abs(freqs[1] - freqs[0])
        desc.index_width = max(1, int(round((kw.get("peak_width", 0.5) / 2.0) / step)))
        desc.method = _lib.METHODS[method]
        desc.p0_only = int(bool(kw.get("p0_only", False)))
        tc = kw.get("target_coord")
        if tc is not None:
            desc.fixed_pivot = 1
            desc.u0_fixed = (freqs[0] - float(tc)) / x_range
            desc.fixed_target = int(np.argmin(np.abs(freqs - tc)))
        if mode == 2:
            p0, p1 = np.empty(batch, np.float64), np.empty(batch, np.float64)
            piv, fun = np.empty(batch, np.int32), np.empty(batch, np.float32)
    desc.autophase_mode = mode

    def ptr(a):
        return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)

    rc = lib.xmr_chain_host_c64(ctypes.byref(desc), ptr(fid), ptr(out), batch, ptr(result), ptr(p0), ptr(p1), ptr(piv),
                                ptr(fun))
    _lib.check(rc)
    if mode == 1:
        tc = autophase.get("target_coord")
        pivot = float(tc) if tc is not None else float(freqs[int(result[2])])
        info = dict(p0=float(result[0]), p1=float(result[1]), pivot=pivot, fun=float(result[3]))
    elif mode == 2:
        tc = autophase.get("target_coord")
        pivot = np.full(bshape, float(tc)) if tc is not None else freqs[piv.reshape(bshape)]
        info = dict(p0=p0.reshape(bshape), p1=p1.reshape(bshape), pivot=pivot, fun=fun.reshape(bshape).astype(np.float64))
    return out, freqs, info


def release_workspace():
    """Free the calling thread's device workspace of the host chain."""
    _lib.check(_lib.load().xmr_host_workspace_release())


def set_resident_limit(nbytes=0):
    """``mode="single"`` keeps the FIDs on the device between its two passes only up to ``nbytes`` (0: whatever is free);
    larger batches are uploaded twice, chunk by chunk -- data sets beyond HBM, or a bound on what a long-lived process holds."""
    _lib.check(_lib.load().xmr_host_chain_resident_limit(int(nbytes)))
