"""TEST INFRASTRUCTURE -- CPU oracle for the FID->spectrum hot path.  NOT a product code path.

A numpy/scipy float64 restatement of the reference's algorithm for
``zero_fill -> apodize_exp -> to_spectrum -> autophase`` (andrewendlinger/xmris v0.6.1), array level:
every function takes ``(values, axis, coord vector, ...)`` instead of an xarray object.  Each function
cites the reference ``file:line`` it follows (paths relative to ``/root/reference/src/xmris``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module, and only as the checker / the reported CPU baseline.  ``xmris_b200`` never
imports it; the product path fails loudly when the CUDA library is missing.

Parity status: PINNED for zero_fill / apodize_exp / to_spectrum / to_fid / phase / the three score
functions / autophase(mode="single") -- against outputs of the reference's own files executed in the
build container (``oracle/ref_loader.py``; vectors committed under ``tests/golden/`` by
``tests/golden/make_golden.py``) and against the reference's notebook known-answer tests re-expressed in
``tests/test_oracle.py``.  The optimiser is the same third-party call the reference makes
(``scipy.optimize.differential_evolution`` -- SciPy is not vendored in the reference; lock pins
1.15.3/1.17.0, this image has 1.18.1; call site ``processing/phasing.py:276-284``).
The reference's own tests do NOT pin autophase angles numerically (the assertion is commented out,
``docs/notebooks/pipeline/autophasing.md:158-162``); our pin is "the reference code run here".
``autophase_each`` (per-spectrum mode) has no reference implementation (``mode="all"`` raises
``NotImplementedError``, ``phasing.py:219-222``); its oracle is the reference's 1-D autophase applied to
every spectrum separately.
"""

from __future__ import annotations

import numpy as np
import scipy.optimize

# --------------------------------------------------------------------------------------------
# A1  zero_fill                                                     processing/fid.py:201-285
# --------------------------------------------------------------------------------------------


def zero_fill(values, axis, coord, target_points=1024, position="end"):
    """Pad ``axis`` with zeros up to ``target_points``.  Returns ``(values, coord, applied)``.

    fid.py:232-236  no-op (plain copy, no lineage) when ``target_points <= n``;
    fid.py:241-248  pad widths: end ``(0, pad)``; symmetric ``(pad//2, pad - pad//2)``; else ValueError;
    fid.py:251      constant-zero pad;
    fid.py:254-263  coordinate rebuilt as ``c0 (- pad_left*delta) + arange(N)*delta``, delta = c[1]-c[0]
                    (only when a coordinate with more than one sample exists).
    """
    values = np.asarray(values)
    n = values.shape[axis]
    if target_points <= n:
        return values.copy(), (None if coord is None else np.array(coord, copy=True)), False
    pad = target_points - n
    if position == "end":
        pw = (0, pad)
    elif position == "symmetric":
        pl = pad // 2
        pw = (pl, pad - pl)
    else:
        raise ValueError("`position` must be either 'end' or 'symmetric'.")
    widths = [(0, 0)] * values.ndim
    widths[axis] = pw
    out = np.pad(values, widths, mode="constant", constant_values=0)
    new_coord = None
    if coord is not None:
        coord = np.asarray(coord)
        if len(coord) > 1:
            delta = coord[1] - coord[0]
            if position == "end":
                new_coord = coord[0] + np.arange(target_points) * delta
            else:
                start = coord[0] - (pw[0] * delta)
                new_coord = start + np.arange(target_points) * delta
        else:
            # fid.py:256 -- a single-sample coordinate is left as xarray's pad made it (NaN filled)
            new_coord = np.pad(coord.astype(float), pw, mode="constant", constant_values=np.nan)
    return out, new_coord, True


# --------------------------------------------------------------------------------------------
# A2  apodize_exp                                                   processing/fid.py:105-144
# --------------------------------------------------------------------------------------------


def _bcast(vec, ndim, axis):
    shape = [1] * ndim
    shape[axis] = -1
    return np.asarray(vec).reshape(shape)


def apodize_exp(values, axis, coord, lb=1.0):
    """``values * exp(-pi*lb*t)`` along ``axis`` using the coordinate VALUES (fid.py:132-139)."""
    values = np.asarray(values)
    weight = np.exp(-np.pi * lb * np.asarray(coord))
    return values * _bcast(weight, values.ndim, axis)


def apodize_lg(values, axis, coord, lb=1.0, gb=1.0):
    """Lorentz-to-Gauss window (fid.py:176-193): ``exp(+pi*lb*t) * exp(-t^2/t_g^2)``, t_g = 2 sqrt(ln2)/(pi gb)."""
    values = np.asarray(values)
    t = np.asarray(coord)
    w = np.exp(np.pi * lb * t)
    if gb != 0:
        t_g = (2 * np.sqrt(np.log(2))) / (np.pi * gb)
        w = w * np.exp(-(t**2) / (t_g**2))
    return values * _bcast(w, values.ndim, axis)


# --------------------------------------------------------------------------------------------
# A3  to_spectrum = fft + fftshift                processing/fid.py:9-42, fourier.py:117-173, 10-32, 64-111
# --------------------------------------------------------------------------------------------


def fft_coords(n_points, coord):
    """Unshifted reciprocal coordinates (fourier.py:92-98): ``fftfreq(n, d=c[1]-c[0])`` (d=1.0 if one sample)."""
    coord = np.asarray(coord)
    delta = (coord[1] - coord[0]) if len(coord) > 1 else 1.0
    return np.fft.fftfreq(n_points, d=delta)


def to_spectrum(values, axis, coord):
    """Ortho FFT along ``axis`` then roll by ``N//2`` (data and coords).  Returns ``(spectrum, freq_coord)``.

    fourier.py:152-153  ``np.fft.fftn(values, axes=(axis,), norm="ortho")``
    fourier.py:31-32    ``roll({dim: N//2}, roll_coords=True)``
    """
    values = np.asarray(values)
    n = values.shape[axis]
    arr = np.fft.fftn(values, axes=(axis,), norm="ortho")
    freqs = fft_coords(n, coord)
    return np.roll(arr, n // 2, axis=axis), np.roll(freqs, n // 2)


def to_fid(values, axis, coord):
    """Inverse of :func:`to_spectrum` (fid.py:45-102): ifftshift (roll by (N+1)//2), ortho ifft, t = arange(N)/(N*df)."""
    values = np.asarray(values)
    n = values.shape[axis]
    arr = np.roll(values, (n + 1) // 2, axis=axis)
    arr = np.fft.ifftn(arr, axes=(axis,), norm="ortho")
    coord = np.asarray(coord)
    if n > 1:
        df = abs(coord[1] - coord[0])
        dt = 1.0 / (n * df)
        t = np.arange(n) * dt
    else:
        # fid.py:81 skips the rebuild; _convert_fft_coords (fourier.py:95) then used delta=1.0
        t = np.fft.fftfreq(n, d=1.0)
    return arr, t


# --------------------------------------------------------------------------------------------
# A4  phase                                                       processing/phasing.py:10-96
# --------------------------------------------------------------------------------------------


def default_pivot(values, axis, coord):
    """phasing.py:49-53 -- coordinate at the global (flat, first-occurrence) argmax of ``abs(values)``."""
    values = np.asarray(values)
    flat_idx = int(np.argmax(np.abs(values)))
    target_idx = np.unravel_index(flat_idx, values.shape)[axis]
    return float(np.asarray(coord)[target_idx])


def phase_array(coord, p0, p1, pivot):
    """phasing.py:56-69 -- ``rad(p0) + rad(p1) * ((x - pivot) / (x_max - x_min))`` (scalar rad(p0) if range is 0)."""
    coord = np.asarray(coord, dtype=float)
    x_range = float(coord.max()) - float(coord.min())
    p0_rad = np.radians(p0)
    p1_rad = np.radians(p1)
    if x_range == 0:
        return p0_rad
    return p0_rad + p1_rad * ((coord - pivot) / x_range)


def phase(values, axis, coord, p0=0.0, p1=0.0, pivot=None):
    """``values * exp(+1j*phi)`` along ``axis`` (phasing.py:73).  Returns ``(values, pivot_used)``."""
    values = np.asarray(values)
    if pivot is None:
        pivot = default_pivot(values, axis, coord)
    ph = phase_array(coord, p0, p1, pivot)
    rot = np.exp(1.0j * ph)
    if np.ndim(rot) == 0:
        return values * rot, pivot
    return values * _bcast(rot, values.ndim, axis), pivot


# --------------------------------------------------------------------------------------------
# A5/A6  score functions                                         processing/phasing.py:100-157
# --------------------------------------------------------------------------------------------


def acme_score(ph, spec1d, coord, pivot):
    """phasing.py:100-122 (a port of nmrglue's ACME score), on a 1-D complex spectrum."""
    p0 = ph[0]
    p1 = ph[1] if len(ph) > 1 else 0.0
    data = np.real(phase(spec1d, 0, coord, p0, p1, pivot)[0])
    stepsize = 1
    ds1 = np.abs((data[1:] - data[:-1]) / (stepsize * 2))
    p1_prob = ds1 / np.sum(ds1)
    p1_prob[p1_prob == 0] = 1
    h1 = -p1_prob * np.log(p1_prob)
    h1s = np.sum(h1)
    as_ = data - np.abs(data)
    sumas = np.sum(as_)
    pfun = 0.0
    if sumas < 0:
        pfun = np.sum((as_ / 2) ** 2)
    return (h1s + 1000 * pfun) / data.shape[-1] / np.max(data)


def peak_minima_score(ph, spec1d, coord, pivot, target_idx, index_width):
    """phasing.py:125-139."""
    p0 = ph[0]
    p1 = ph[1] if len(ph) > 1 else 0.0
    data = np.real(phase(spec1d, 0, coord, p0, p1, pivot)[0])
    start = max(0, target_idx - index_width)
    end = min(len(data), target_idx + index_width)
    mina = np.min(data[start:target_idx]) if start < target_idx else data[target_idx]
    minb = np.min(data[target_idx:end]) if end > target_idx else data[target_idx]
    return np.abs(mina - minb)


def roi_positivity_score(ph, spec1d, coord, pivot, target_idx, index_width):
    """phasing.py:142-157."""
    p0 = ph[0]
    p1 = ph[1] if len(ph) > 1 else 0.0
    data = np.real(phase(spec1d, 0, coord, p0, p1, pivot)[0])
    start = max(0, target_idx - index_width)
    end = min(len(data), target_idx + index_width)
    roi = data[start:end]
    pos_reward = np.sum(roi[roi > 0])
    neg_penalty = np.sum(np.abs(roi[roi < 0])) * 5.0
    return neg_penalty - pos_reward


# --------------------------------------------------------------------------------------------
# A7  autophase                                                  processing/phasing.py:161-290
# --------------------------------------------------------------------------------------------


def autophase_setup(values, axis, coord, peak_width=0.5, target_coord=None):
    """phasing.py:226-247 -- global argmax, pivot / target index, index width, and the optimisation slice index."""
    values = np.asarray(values)
    coord = np.asarray(coord)
    flat_idx = int(np.argmax(np.abs(values)))
    unr = np.unravel_index(flat_idx, values.shape)
    if target_coord is not None:
        target_idx = int(np.argmin(np.abs(coord - target_coord)))
        pivot = float(target_coord)
    else:
        target_idx = int(unr[axis])
        pivot = float(coord[target_idx])
    sl = tuple(slice(None) if i == axis else int(unr[i]) for i in range(values.ndim))
    step_size = np.abs(coord[1] - coord[0])
    index_width = int(round((peak_width / 2.0) / step_size))
    index_width = max(1, index_width)
    return sl, target_idx, pivot, index_width


def smooth_slice(spec1d, coord, lb):
    """phasing.py:250-253 -- to_fid -> apodize_exp(lb) -> to_spectrum on the 1-D optimisation slice."""
    fid, t = to_fid(spec1d, 0, coord)
    fid = apodize_exp(fid, 0, t, lb)
    spec, _ = to_spectrum(fid, 0, t)
    return spec


def autophase_search(work1d, coord, pivot, method="acme", target_idx=0, index_width=1, p0_only=False, disp=False):
    """The reference's optimiser call, verbatim arguments (phasing.py:258-287).  Returns ``(p0, p1, OptimizeResult)``."""
    if method == "acme":
        score_fn, args = acme_score, (work1d, coord, pivot)
    elif method == "peak_minima":
        score_fn, args = peak_minima_score, (work1d, coord, pivot, target_idx, index_width)
    elif method == "positivity":
        score_fn, args = roi_positivity_score, (work1d, coord, pivot, target_idx, index_width)
    else:
        raise ValueError("Method must be 'acme', 'peak_minima', or 'positivity'")
    bounds = [(-180.0, 180.0)] if p0_only else [(-180.0, 180.0), (-4000.0, 4000.0)]
    opt = scipy.optimize.differential_evolution(
        score_fn, bounds=bounds, args=args, strategy="best1bin", tol=0.01, seed=42, disp=disp
    )
    p0 = opt.x[0]
    p1 = opt.x[1] if not p0_only else 0.0
    return p0, p1, opt


def autophase(values, axis, coord, method="acme", mode="single", peak_width=0.5, target_coord=None,
              p0_only=False, lb=0.0):
    """Reference ``autophase`` (phasing.py:161-290).  Returns ``(phased values, info dict)``."""
    if mode == "all":
        raise NotImplementedError(
            "Applying autophase to each spectrum individually ('all') is not yet implemented."
        )
    elif mode != "single":
        raise ValueError("Mode must be 'single' or 'all'.")
    values = np.asarray(values)
    coord = np.asarray(coord)
    sl, target_idx, pivot, index_width = autophase_setup(values, axis, coord, peak_width, target_coord)
    opt1d = values[sl]
    work = smooth_slice(opt1d, coord, lb) if lb > 0 else opt1d
    p0, p1, opt = autophase_search(work, coord, pivot, method, target_idx, index_width, p0_only)
    out, _ = phase(values, axis, coord, p0, p1, pivot)
    info = dict(p0=float(p0), p1=float(p1), pivot=pivot, target_idx=target_idx, index_width=index_width,
                slice=sl, fun=float(opt.fun), nfev=int(opt.nfev), success=bool(opt.success))
    return out, info


def autophase_each(values, axis, coord, method="acme", peak_width=0.5, target_coord=None, p0_only=False, lb=0.0):
    """Per-spectrum autophase: the reference's ``autophase`` applied to every 1-D spectrum separately.

    This is the oracle for the north-star's per-voxel kernel (the reference's ``mode="all"`` is unimplemented).
    Returns ``(phased values, p0[batch], p1[batch], pivot[batch], fun[batch])`` with batch = all dims but ``axis``.
    """
    values = np.asarray(values)
    moved = np.moveaxis(values, axis, -1)
    bshape = moved.shape[:-1]
    flat = moved.reshape(-1, moved.shape[-1])
    out = np.empty_like(flat, dtype=np.complex128)
    p0s = np.empty(flat.shape[0])
    p1s = np.empty(flat.shape[0])
    pivs = np.empty(flat.shape[0])
    funs = np.empty(flat.shape[0])
    for i in range(flat.shape[0]):
        o, info = autophase(flat[i], 0, coord, method=method, peak_width=peak_width, target_coord=target_coord,
                            p0_only=p0_only, lb=lb)
        out[i] = o
        p0s[i], p1s[i], pivs[i], funs[i] = info["p0"], info["p1"], info["pivot"], info["fun"]
    out = np.moveaxis(out.reshape(bshape + (moved.shape[-1],)), -1, axis)
    return out, p0s.reshape(bshape), p1s.reshape(bshape), pivs.reshape(bshape), funs.reshape(bshape)


# --------------------------------------------------------------------------------------------
# "later" row  baseline_als                               reference: processing/baseline.py:10-119
# --------------------------------------------------------------------------------------------


def als_core(y, lam, p, n_iter):
    """1-D asymmetric least squares baseline (``baseline.py:10-39``): the same scipy.sparse calls as the reference."""
    from scipy import sparse
    from scipy.sparse.linalg import spsolve

    L = len(y)
    D = sparse.diags([1, -2, 1], [0, 1, 2], shape=(L - 2, L), dtype=float)
    D_T_D = (lam * D.T.dot(D)).tocsc()
    w = np.ones(L)
    z = None
    for _ in range(n_iter):
        W = sparse.diags(w, 0, format="csc", dtype=float)
        z = spsolve(W + D_T_D, w * y)
        w = p * (y > z) + (1 - p) * (y < z)
    return z


def baseline_als(values, axis, lam=1e5, p=0.001, n_iter=10):
    """Real part minus its AsLS baseline along ``axis`` (``baseline.py:81-105``).  Returns ``(corrected, baseline)``."""
    values = np.asarray(values)
    work = np.real(values) if np.iscomplexobj(values) else values
    moved = np.moveaxis(work, axis, -1)
    flat = moved.reshape(-1, moved.shape[-1])
    base = np.stack([als_core(row, lam, p, n_iter) for row in flat]).reshape(moved.shape)
    base = np.moveaxis(base, -1, axis)
    return work - base, base


# --------------------------------------------------------------------------------------------
# N4  remove_digital_filter                              reference: vendor/bruker.py:7-118
# --------------------------------------------------------------------------------------------


def remove_digital_filter(values, axis, coord, group_delay, keep_length=True):
    """Bruker group-delay removal: drop ``floor(gd)`` points, shift the rest by the fractional part with an
    FFT phase ramp, optionally pad zeros back to the original length (``vendor/bruker.py:55-118``).

    Returns ``(values, coord)``; ``group_delay <= 0`` returns copies (``bruker.py:58-59``)."""
    values = np.asarray(values)
    coord = np.asarray(coord, dtype=np.float64)
    if group_delay <= 0:
        return values.copy(), coord.copy()
    int_delay = int(np.floor(group_delay))                                   # bruker.py:62-63
    frac_delay = group_delay - int_delay
    sl = [slice(None)] * values.ndim
    sl[axis] = slice(int_delay, None)
    cut = values[tuple(sl)] if int_delay > 0 else values                      # bruker.py:67-70
    cut_coord = coord[int_delay:] if int_delay > 0 else coord
    if not np.isclose(frac_delay, 0.0):                                       # bruker.py:73-87
        n_points = cut.shape[axis]
        freqs = _bcast(np.fft.fftfreq(n_points), cut.ndim, axis)
        spectrum = np.fft.fft(cut, axis=axis)
        corrected = np.fft.ifft(spectrum * np.exp(1j * 2 * np.pi * freqs * frac_delay), axis=axis)
    else:
        corrected = cut
    if int_delay > 0 and keep_length:                                         # bruker.py:92-99
        pad_shape = list(corrected.shape)
        pad_shape[axis] = int_delay
        final = np.concatenate((corrected, np.zeros(pad_shape, dtype=corrected.dtype)), axis=axis)
        new_coord = coord
    else:
        final = corrected
        new_coord = cut_coord
    return final, new_coord - new_coord[0]                                    # bruker.py:107-108


# --------------------------------------------------------------------------------------------
# The chain                                                            README.md:66-73
# --------------------------------------------------------------------------------------------


def chain_to_spectrum(fid, axis, time_coord, target_points=None, position="end", lb=None):
    """zero_fill -> apodize_exp -> to_spectrum.  Returns ``(spectrum, freq_coord)``."""
    vals, coord = np.asarray(fid), np.asarray(time_coord)
    if target_points is not None:
        vals, coord, _ = zero_fill(vals, axis, coord, target_points, position)
    if lb is not None:
        vals = apodize_exp(vals, axis, coord, lb)
    return to_spectrum(vals, axis, coord)


def chain(fid, axis, time_coord, target_points=None, position="end", lb=None, mode="single", **autophase_kw):
    """Full chain.  ``mode="single"`` is the reference's semantics; ``mode="each"`` the per-spectrum loop."""
    spec, freqs = chain_to_spectrum(fid, axis, time_coord, target_points, position, lb)
    if mode == "single":
        out, info = autophase(spec, axis, freqs, **autophase_kw)
        return out, freqs, info
    out, p0, p1, piv, fun = autophase_each(spec, axis, freqs, **autophase_kw)
    return out, freqs, dict(p0=p0, p1=p1, pivot=piv, fun=fun)
