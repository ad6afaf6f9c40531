"""Per-spectrum autophase (``mode="all"``): every voxel is searched and phased on its own, in one fused pass.

The reference declares this mode but raises ``NotImplementedError`` (``processing/phasing.py:219-222``); its meaning
here is "the reference's ``autophase`` applied to every 1-D spectrum separately" (the oracle's ``autophase_each``):
per-spectrum pivot = that spectrum's own ``|S|`` maximum (or ``target_coord``), the same objective and box bounds.
The device kernel is ``k2_kernel`` behind ``xmr_chain_each_c64``.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, chain
from . import device as D
from .vocab import ATTRS


def _launch(in_t, out_t, n_in, n_out, pad_left, spec_in, window, method, du, fixed_pivot, u0_fixed, fixed_target,
            index_width, p0_only, stream=None):
    torch = D._torch()
    lib = _lib.load()
    batch = in_t.numel() // n_in
    dev = in_t.device
    p0 = torch.empty(batch, dtype=torch.float64, device=dev)
    p1 = torch.empty(batch, dtype=torch.float64, device=dev)
    piv = torch.empty(batch, dtype=torch.int32, device=dev)
    fun = torch.empty(batch, dtype=torch.float32, device=dev)
    win_mode, win_dev, rows = _lib.WIN_NONE, None, None
    if isinstance(window, D.PreparedWindow):
        win_mode, win_dev, rows = window.mode, window.dev, window.rows
    elif window is not None:
        win_mode, table, rows = D.split_window(window, n_out)
        win_dev = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    rows_arr = (ctypes.c_float * 32)(*([1.0] * 32))
    if rows is not None:
        for i, r in enumerate(rows):
            rows_arr[i] = float(r)
    with torch.cuda.device(dev):
        rc = lib.xmr_chain_each_c64(
            D._ptr(in_t), D._ptr(out_t), batch, n_in, n_out, int(pad_left), int(bool(spec_in)), win_mode,
            D._ptr(win_dev), ctypes.cast(rows_arr, ctypes.c_void_p), ctypes.c_float(1.0 / np.sqrt(n_out)),
            _lib.METHODS[method], float(du), int(bool(fixed_pivot)), float(u0_fixed), int(fixed_target),
            int(index_width), int(bool(p0_only)), D._ptr(p0), D._ptr(p1), D._ptr(piv), D._ptr(fun),
            D._stream_ptr(stream))
    _lib.check(rc)
    return p0, p1, piv, fun


def _search_geometry(coords, peak_width, target_coord):
    """du, fixed-pivot data and ROI half-width from the coordinate along ``dim`` (uniform axis required)."""
    from .processing import _affine_ramp, _index_width

    u0_first, du = _affine_ramp(coords, coords[0])     # u with the pivot at sample 0: u0_first == 0
    index_width = _index_width(coords, peak_width)
    if target_coord is not None:
        x_range = float(coords.max()) - float(coords.min())
        u0_fixed = (coords[0] - float(target_coord)) / x_range
        fixed_target = int(np.argmin(np.abs(coords - target_coord)))
        return du, True, u0_fixed, fixed_target, index_width
    return du, False, 0.0, 0, index_width


def chain_all_device(fid_t, time_coord, target_points=None, position="end", lb=None, out=None, geo=None,
                     method="acme", peak_width=0.5, target_coord=None, p0_only=False, stream=None, gb=None):
    """FID batch ``[batch, n_in]`` (device) -> phased spectra + per-voxel angles, all on the device, one kernel."""
    torch = D._torch()
    D._require_cuda(fid_t, "fid")
    n_in = fid_t.shape[-1]
    if geo is None:
        geo = chain.chain_geometry(n_in, time_coord, target_points, position, lb, gb)
    n_out = geo["n_out"]
    flat = fid_t.reshape(-1, n_in)
    if out is None:
        out = torch.empty((flat.shape[0], n_out), dtype=torch.complex64, device=fid_t.device)
    sg = geo.get("_search")
    key = (method, peak_width, target_coord)
    if sg is None or sg[0] != key:
        sg = (key, _search_geometry(geo["freqs"], peak_width, target_coord))
        geo["_search"] = sg
    du, fixed, u0_fixed, fixed_target, index_width = sg[1]
    p0, p1, piv, fun = _launch(flat, out, n_in, n_out, geo["pad_left"], False, chain._win(geo, fid_t.device), method, du,
                               fixed, u0_fixed, fixed_target, index_width, p0_only, stream)
    return dict(out=out, p0=p0, p1=p1, pivot_index=piv, fun=fun, freqs=geo["freqs"], launches=1, n_out=n_out)


def chain_all(fid_t, time_coord, target_points=None, position="end", lb=None, method="acme", peak_width=0.5,
              target_coord=None, p0_only=False, gb=None):
    """Like :func:`chain_all_device` with host copies of the per-voxel results.  Returns ``(spec, freqs, info)``."""
    r = chain_all_device(fid_t, time_coord, target_points, position, lb, None, None, method, peak_width, target_coord,
                         p0_only, gb=gb)
    bshape = tuple(fid_t.shape[:-1])
    freqs = r["freqs"]
    piv_idx = r["pivot_index"].cpu().numpy().reshape(bshape)
    pivot = np.full(bshape, float(target_coord)) if target_coord is not None else freqs[piv_idx]
    info = dict(p0=r["p0"].cpu().numpy().reshape(bshape), p1=r["p1"].cpu().numpy().reshape(bshape), pivot=pivot,
                fun=r["fun"].cpu().numpy().reshape(bshape).astype(np.float64))
    return r["out"].reshape(bshape + (r["n_out"],)), freqs, info


def autophase_spectra_device(spec_t, coords, method="acme", peak_width=0.5, target_coord=None, p0_only=False, lb=0.0):
    """Per-spectrum autophase of device-resident spectra ``[batch, n]`` with coordinate ``coords`` along the last axis."""
    torch = D._torch()
    n = spec_t.shape[-1]
    flat = spec_t.reshape(-1, n)
    du, fixed, u0_fixed, fixed_target, index_width = _search_geometry(coords, peak_width, target_coord)
    out = torch.empty_like(flat)
    if lb > 0:
        # phasing.py:250-255: the SEARCH sees a smoothed copy (to_fid -> apodize_exp -> to_spectrum); the phase is
        # applied to the un-smoothed data
        fid, _, _ = D.fid_to_spectrum(flat, inverse=True, in_shift=n // 2, out_shift=0)
        df = abs(coords[1] - coords[0])
        t = np.arange(n) * (1.0 / (n * df))
        smooth, _, _ = D.fid_to_spectrum(fid, window=np.exp(-np.pi * lb * t) / np.sqrt(n))
        p0, p1, piv, fun = _launch(smooth, out, n, n, 0, True, None, method, du, fixed, u0_fixed, fixed_target, index_width,
                                   p0_only)
        # the reference pivots on the UN-smoothed spectrum's maximum (phasing.py:229-238 run before the smoothing)
        _, argmax = D.row_absmax(flat)
        if not fixed:
            # re-anchor: phase ramp about the un-smoothed maximum; p1 found about the smoothed maximum is kept, p0 moves
            shift = (argmax.to(torch.float64) - piv.to(torch.float64)) * du
            p0 = p0 + p1 * shift
            p0 = torch.remainder(p0 + 180.0, 360.0) - 180.0
            piv = argmax
        u0 = torch.full_like(p0, u0_fixed) if fixed else -du * piv.to(torch.float64)
        out = D.phase_each(flat, p0 / 360.0 + (p1 / 360.0) * u0, (p1 / 360.0) * du)
    else:
        p0, p1, piv, fun = _launch(flat, out, n, n, 0, True, None, method, du, fixed, u0_fixed, fixed_target, index_width,
                                   p0_only)
    return out.reshape(spec_t.shape), p0, p1, piv, fun


def attach_per_spectrum_coords(res, dim, info):
    """Per-voxel lineage: ``phase_p0`` / ``phase_p1`` / ``phase_pivot`` as non-index coordinates over the batch dims."""
    bdims = tuple(d for d in res.dims if d != dim)
    coords = {ATTRS.phase_p0: (bdims, np.asarray(info["p0"], dtype=np.float64)),
              ATTRS.phase_p1: (bdims, np.asarray(info["p1"], dtype=np.float64)),
              ATTRS.phase_pivot: (bdims, np.asarray(info["pivot"], dtype=np.float64))}
    res = res.assign_coords(coords)
    res.attrs[ATTRS.phase_pivot_coord] = dim
    return res


def autophase_all(da, dim, axis, flat2d, coords, method, peak_width, target_coord, p0_only, lb):
    """DataArray front end of ``autophase(mode="all")`` (called by :func:`xmris_b200.processing.autophase`)."""
    from . import processing as P

    out, p0, p1, piv, fun = autophase_spectra_device(flat2d, coords, method, peak_width, target_coord, p0_only, lb)
    bshape = tuple(s for i, s in enumerate(da.shape) if i != axis)
    moved_shape = bshape + (flat2d.shape[-1],)
    res = da.copy(data=P._from_device(out.reshape(moved_shape), axis))
    if da.name != dim:
        res.name = None
    res.attrs = dict(da.attrs)
    piv_idx = piv.cpu().numpy().reshape(bshape)
    pivot = np.full(bshape, float(target_coord)) if target_coord is not None else coords[piv_idx]
    info = dict(p0=p0.cpu().numpy().reshape(bshape), p1=p1.cpu().numpy().reshape(bshape), pivot=pivot)
    return attach_per_spectrum_coords(res, dim, info)
