"""The fused FID -> phased-spectrum chain on device-resident data (reference semantics, ``mode="single"``).

Two fused passes over the FID batch, no intermediate array in HBM:

  pass 1   K1(statistics only): zero-fill + window + FFT, per-spectrum max |S| / argmax        reads  8*n_in B
           global first-occurrence argmax (phasing.py:229-231) -> winning spectrum + pivot
           one-spectrum (p0, p1) search (phasing.py:270-287)
  pass 2   K1(store + fused phase): same transform, rotated by exp(i*phi_m) on the way out       reads  8*n_in B
                                                                                                 writes 8*n_out B
The reference's global-argmax dependency makes the second read of the FID compulsory; the un-phased spectrum is
never written.  Multi-GPU: each rank runs pass 1 on its shard, the (max, index) pairs are exchanged
(``xmris_b200.sharding``), the owning rank searches, and (p0, p1, pivot) is broadcast before pass 2.
"""

from __future__ import annotations

import threading

import numpy as np

from . import device as D
from .vocab import ATTRS, COORDS, DIMS


_GEO_CACHE: dict = {}
_GEO_LOCK = threading.RLock()     # the cache and the per-geometry "_prepared" / "_search" entries are shared by threads


def chain_geometry(n_in, time_coord, target_points, position, lb, gb=None):
    """Host metadata of the chain: padded time axis, window (incl. 1/sqrt(N)), frequency axis, pad_left.

    ``gb`` given: the Lorentz-to-Gauss window of ``apodize_lg`` (``exp(+pi*lb*t) * exp(-t^2/t_g^2)``, fid.py:176-193) instead
    of ``apodize_exp``'s ``exp(-pi*lb*t)`` (fid.py:136).  Memoised on the exact time coordinate (a repeated call on the same axis reuses the float64 tables and the window
    already uploaded to the device: for small batches these host steps cost more than the kernels)."""
    t = np.ascontiguousarray(time_coord, dtype=np.float64)
    key = (int(n_in), None if target_points is None else int(target_points), position, None if lb is None else float(lb),
           None if gb is None else float(gb), t.tobytes())
    with _GEO_LOCK:
        hit = _GEO_CACHE.get(key)
        if hit is not None:
            return hit
        geo = _chain_geometry(n_in, t, target_points, position, lb, gb)
        if len(_GEO_CACHE) >= 32:
            _GEO_CACHE.pop(next(iter(_GEO_CACHE)))
        _GEO_CACHE[key] = geo
        return geo


def _chain_geometry(n_in, t, target_points, position, lb, gb=None):
    n_out, pad_left = n_in, 0
    t_pad = t
    if target_points is not None and target_points > n_in:
        n_out = int(target_points)
        pad = n_out - n_in
        if position == "end":
            pad_left = 0
        elif position == "symmetric":
            pad_left = pad // 2
        else:
            raise ValueError("`position` must be either 'end' or 'symmetric'.")
        delta = t[1] - t[0]
        t_pad = (t[0] - pad_left * delta) + np.arange(n_out) * delta   # fid.py:254-263
    D.check_length(n_out)
    window = None
    if gb is not None:
        w = np.exp(np.pi * (1.0 if lb is None else lb) * t_pad)          # fid.py:176-193 (apodize_lg's default lb is 1.0)
        if gb != 0:
            t_g = (2 * np.sqrt(np.log(2))) / (np.pi * gb)
            w = w * np.exp(-(t_pad**2) / (t_g**2))
        window = w / np.sqrt(n_out)
    elif lb is not None:
        window = np.exp(-np.pi * lb * t_pad) / np.sqrt(n_out)           # fid.py:136 with the ortho norm folded in
    delta = (t_pad[1] - t_pad[0]) if n_out > 1 else 1.0
    freqs = np.roll(np.fft.fftfreq(n_out, d=delta), n_out // 2)         # fourier.py:95, 31-32
    for arr in (t_pad, window, freqs):
        if arr is not None:
            arr.setflags(write=False)        # shared by every caller of the memoised geometry
    return dict(n_out=n_out, pad_left=pad_left, t_pad=t_pad, window=window, freqs=freqs)


def phase_turns(freqs, p0, p1, pivot):
    """Affine phase in turns of the stored bin index m: ``turns(m) = a + b*m`` (phasing.py:56-69, uniform axis)."""
    x_range = float(freqs.max()) - float(freqs.min())
    du = (freqs[-1] - freqs[0]) / (len(freqs) - 1) / x_range
    u0 = (freqs[0] - pivot) / x_range
    return p0 / 360.0 + (p1 / 360.0) * u0, (p1 / 360.0) * du, u0, du


def _win(geo, device):
    """The chain's window, uploaded once per (geometry, device)."""
    if geo["window"] is None:
        return None
    if geo["n_out"] not in D.SUPPORTED_N:
        return geo["window"]          # chirp-z path takes the float64 window as is
    with _GEO_LOCK:
        cache = geo.setdefault("_prepared", {})
        key = (device.type, device.index)
        if key not in cache:
            cache[key] = D.PreparedWindow(geo["window"], geo["n_out"], device)
        return cache[key]


def local_stats(fid_t, geo):
    """Pass 1 on this rank's shard.  Returns ``(max |S|, row * n_out)`` of the shard (host scalars).

    Only the per-spectrum maxima are recorded (the cheapest statistics pass); the position of the maximum inside the
    winning row is recovered when that one row is transformed again for the search."""
    if geo["n_out"] in D.SUPPORTED_N:
        # branch and bound: spectra that provably cannot hold the global maximum skip their last FFT stage (entry 0)
        absmax, _ = D.fid_absmax_pruned(fid_t, n_out=geo["n_out"], pad_left=geo["pad_left"], window=_win(geo, fid_t.device))
    else:
        _, absmax, _ = D.fid_to_spectrum(fid_t, n_out=geo["n_out"], pad_left=geo["pad_left"],
                                         window=_win(geo, fid_t.device), store=False, want_stats=True, want_index=False)
    return D.global_argmax(absmax.reshape(-1), None, geo["n_out"])


def search_on_row(fid_row_t, geo, flat_index, method="acme", peak_width=0.5, target_coord=None, p0_only=False,
                  lb=0.0):
    """Transform the winning FID alone and run the (p0, p1) search on it.  Returns ``(p0, p1, pivot, fun)``."""
    from .processing import _index_width, _smooth_slice

    n_out, freqs = geo["n_out"], geo["freqs"]
    spec, _, argmax = D.fid_to_spectrum(fid_row_t.reshape(1, -1), n_out=n_out, pad_left=geo["pad_left"],
                                        window=_win(geo, fid_row_t.device), want_stats=True)
    work = spec.reshape(n_out)
    argmax_idx = int(argmax.reshape(-1)[0].item())
    if target_coord is not None:
        target_idx = int(np.argmin(np.abs(freqs - target_coord)))
        pivot = float(target_coord)
    else:
        target_idx = int(argmax_idx)
        pivot = float(freqs[target_idx])
    if lb > 0:
        work = _smooth_slice(work, freqs, lb)
    _, _, u0, du = phase_turns(freqs, 0.0, 0.0, pivot)
    res = D.autophase_search(work.contiguous(), u0, du, method, target_idx, _index_width(freqs, peak_width),
                             p0_only).cpu().numpy()
    p0 = float(res[0])
    p1 = 0.0 if p0_only else float(res[1])
    return p0, p1, pivot, float(res[2])


def apply_pass(fid_t, geo, p0, p1, pivot, out=None):
    """Pass 2: transform again and rotate by the winning phase on the way out."""
    a, b, _, _ = phase_turns(geo["freqs"], p0, p1, pivot)
    spec, _, _ = D.fid_to_spectrum(fid_t, n_out=geo["n_out"], pad_left=geo["pad_left"],
                                   window=_win(geo, fid_t.device), phase_turns=(a, b), out=out)
    return spec


def chain_single(fid_t, time_coord, target_points=None, position="end", lb=None, method="acme", peak_width=0.5,
                 target_coord=None, p0_only=False, autophase_lb=0.0, out=None, exchange=None, gb=None, all_gather=None,
                 row_offset=0):
    """Full chain with the reference's ``mode="single"`` autophase on a ``[batch, n_in]`` device tensor.

    Multi-GPU (voxels sharded over ranks): pass ``all_gather=sharding.SlotAllGather(dist)`` and ``row_offset`` (first global
    row of this rank's shard): ONE device-side all-gather of the ranks' candidate rows, no host round trip, every rank
    searches the global winner redundantly.  ``exchange`` (legacy, host-mediated) is a callable
    ``(local_max, local_flat_index, search_fn) -> (p0, p1, pivot, fun)``: two small collectives through the host.
    Returns ``(phased spectrum tensor, freqs float64, info dict)``.
    """
    n_in = fid_t.shape[-1]
    flat = fid_t.reshape(-1, n_in)
    geo = chain_geometry(n_in, time_coord, target_points, position, lb, gb)
    if autophase_lb == 0 and geo["n_out"] in D.SUPPORTED_N and (flat.shape[0] > 0 or all_gather is not None) and (
            exchange is None):
        return _chain_single_one_call(fid_t, flat, geo, method, peak_width, target_coord, p0_only, out, all_gather, row_offset)
    vmax, findex = local_stats(flat, geo)

    def search():
        return search_on_row(flat[findex // geo["n_out"]], geo, findex, method, peak_width, target_coord, p0_only,
                             autophase_lb)

    p0, p1, pivot, fun = search() if exchange is None else exchange(vmax, findex, search)
    spec = apply_pass(flat, geo, p0, p1, pivot, out=out)
    spec = spec.reshape(tuple(fid_t.shape[:-1]) + (geo["n_out"],))
    return spec, geo["freqs"], dict(p0=p0, p1=p1, pivot=pivot, fun=fun, n_out=geo["n_out"], pad_left=geo["pad_left"])


def _chain_single_one_call(fid_t, flat, geo, method, peak_width, target_coord, p0_only, out, all_gather=None, row_offset=0):
    """Single-GPU ``mode="single"`` chain through ``xmr_chain_single_dev_c64`` (one C call, three small read-backs)."""
    from .processing import _index_width

    n_out, freqs = geo["n_out"], geo["freqs"]
    x_range = float(freqs.max()) - float(freqs.min())
    du = (freqs[-1] - freqs[0]) / (n_out - 1) / x_range
    fixed = None
    if target_coord is not None:
        fixed = ((freqs[0] - float(target_coord)) / x_range, int(np.argmin(np.abs(freqs - target_coord))))
    spec, res = D.chain_single_dev(flat, n_out, geo["pad_left"], _win(geo, flat.device), du, method,
                                   _index_width(freqs, peak_width), p0_only, fixed,
                                   out=None if out is None else out.reshape(-1, n_out), all_gather=all_gather,
                                   row_offset=row_offset)
    pivot = float(target_coord) if target_coord is not None else float(freqs[int(res[2])])
    spec = spec.reshape(tuple(fid_t.shape[:-1]) + (n_out,))
    return spec, freqs, dict(p0=res[0], p1=0.0 if p0_only else res[1], pivot=pivot, fun=res[3], n_out=n_out,
                             pad_left=geo["pad_left"], max_abs=res[4], winning_row=int(res[5]))


def chain_to_spectrum(fid_t, time_coord, target_points=None, position="end", lb=None, out=None, gb=None):
    """``zero_fill -> apodize_exp (or apodize_lg when ``gb`` is given) -> to_spectrum`` only (one fused pass)."""
    n_in = fid_t.shape[-1]
    geo = chain_geometry(n_in, time_coord, target_points, position, lb, gb)
    spec, _, _ = D.fid_to_spectrum(fid_t, n_out=geo["n_out"], pad_left=geo["pad_left"],
                                   window=_win(geo, fid_t.device), out=out)
    return spec, geo["freqs"], geo


def run_chain_dataarray(da, dim, out_dim, target_points, position, lb, autophase_kwargs, baseline_kwargs=None, gb=None):
    """DataArray front end of the fused chain: same coords / attrs / lineage as the chained accessor calls.

    ``baseline_kwargs`` (``lam, p, n_iter``): additionally run ``baseline_als`` on the device-resident spectrum (the step
    after autophase in the reference's pipeline); the result is then real-valued."""
    from . import processing as P
    from ._xr import xr

    P._check_dims(da, dim, "process_fid")
    axis = da.get_axis_num(dim)
    n_in = da.sizes[dim]
    t = np.asarray(da.coords[dim].values, dtype=np.float64)
    fid_t = None
    attrs = dict(da.attrs)
    name = da.name
    padded = target_points is not None and target_points > n_in
    if padded:
        attrs[ATTRS.zero_fill_target] = target_points
        attrs[ATTRS.zero_fill_position] = position
    if gb is not None and lb is None:
        lb = 1.0                                   # apodize_lg's default (fid.py:148)
    if lb is not None:
        attrs[ATTRS.apodization_lb] = lb
        if gb is not None:
            attrs[ATTRS.apodization_gb] = gb
        if name != dim:
            name = None
    target = out_dim if out_dim is not None else dim
    n_out_chk = int(target_points) if padded else n_in
    use_host_abi = baseline_kwargs is None and n_out_chk in D.SUPPORTED_N and (autophase_kwargs is None or (
        autophase_kwargs.get("lb", 0.0) == 0.0 and (autophase_kwargs.get("mode", "single") != "all" or n_out_chk >= 512)))
    if use_host_abi:
        # numpy in -> ONE C-ABI call on host buffers -> numpy out (no torch on this path)
        from . import hostabi

        moved = np.ascontiguousarray(np.moveaxis(np.asarray(da.values), axis, -1), dtype=np.complex64)
        out_np, freqs, info = hostabi.chain_host(moved, t, target_points if padded else None, position, lb,
                                                 autophase=autophase_kwargs, gb=gb)
        spec = None
    elif autophase_kwargs is None:
        fid_t = P._to_device(da.values, axis)
        spec, freqs, _ = chain_to_spectrum(fid_t, t, target_points if padded else None, position, lb, gb=gb)
        info = None
    else:
        fid_t = P._to_device(da.values, axis)
        kw = dict(autophase_kwargs)
        mode = kw.pop("mode", "single")
        if mode == "all":
            from .pervoxel import chain_all

            spec, freqs, info = chain_all(fid_t, t, target_points if padded else None, position, lb,
                                          method=kw.get("method", "acme"), peak_width=kw.get("peak_width", 0.5),
                                          target_coord=kw.get("target_coord"), p0_only=kw.get("p0_only", False), gb=gb)
        elif mode == "single":
            spec, freqs, info = chain_single(fid_t, t, target_points if padded else None, position, lb,
                                             method=kw.get("method", "acme"), peak_width=kw.get("peak_width", 0.5),
                                             target_coord=kw.get("target_coord"), p0_only=kw.get("p0_only", False),
                                             autophase_lb=kw.get("lb", 0.0), gb=gb)
        else:
            raise ValueError("Mode must be 'single' or 'all'.")
    dims = tuple(target if d == dim else d for d in da.dims)
    coords = {k: da.coords[k] for k in da.coords if dim not in da.coords[k].dims}
    _, var = P._spectrum_coord(dim, out_dim, freqs)
    coords[target] = var
    if spec is None:
        values = np.ascontiguousarray(np.moveaxis(out_np, -1, axis)).astype(P.OUTPUT_DTYPE, copy=False)
    elif baseline_kwargs is not None:
        bk = dict(lam=1e5, p=0.001, n_iter=10)
        bk.update(baseline_kwargs)
        real = D.baseline_als(spec, lam=bk["lam"], p=bk["p"], n_iter=bk["n_iter"])       # the spectrum never left the device
        values = np.ascontiguousarray(np.moveaxis(real.cpu().numpy(), -1, axis))
    else:
        values = P._from_device(spec, axis)
    res = xr.DataArray(values, dims=dims, coords=coords, attrs=attrs, name=name)
    if info is not None and np.ndim(info["p0"]) == 0:
        if name != target:
            res.name = None
        p1 = info["p1"] if not (autophase_kwargs or {}).get("p0_only", False) else 0.0
        res.attrs[ATTRS.phase_p0] = np.float64(info["p0"])
        res.attrs[ATTRS.phase_p1] = np.float64(p1) if p1 != 0.0 or not (autophase_kwargs or {}).get("p0_only") else 0.0
        res.attrs[ATTRS.phase_pivot] = info["pivot"]
        res.attrs[ATTRS.phase_pivot_coord] = target
    elif info is not None:
        from .pervoxel import attach_per_spectrum_coords

        res = attach_per_spectrum_coords(res, target, info)
    if baseline_kwargs is not None:
        res.attrs[ATTRS.baseline_method] = "als"
        res.attrs[ATTRS.baseline_lam] = bk["lam"]
        res.attrs[ATTRS.baseline_p] = bk["p"]
        res.attrs[ATTRS.baseline_iter] = bk["n_iter"]
    return res
