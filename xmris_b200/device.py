"""Array-level API on device-resident torch tensors -> C ABI (``libxmris_b200.so``).

Everything here takes / returns ``torch`` CUDA tensors (complex64, transform axis last, contiguous) and plain
numpy/float metadata.  torch is plumbing only (device memory, streams); all arithmetic happens in the
hand-written sm_100a kernels behind the C ABI.  There is no CPU fallback: a missing library or a CPU tensor
raises.
"""

from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib

SUPPORTED_N = tuple(2**k for k in range(4, 14))  # 16 .. 8192


def _torch():
    import torch

    return torch


def _require_cuda(x, name):
    torch = _torch()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError(f"{name} must be a CUDA torch tensor (xmris_b200 has no CPU path)")
    if x.dtype != torch.complex64:
        raise TypeError(f"{name} must be complex64, got {x.dtype}")
    if not x.is_contiguous():
        raise ValueError(f"{name} must be contiguous with the transform axis last")


def _stream_ptr(stream=None):
    torch = _torch()
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


BLUESTEIN_MAX_N = 4096  # arbitrary lengths up to here run as a chirp-z transform over a power-of-two length <= 8192


def check_length(n_out: int):
    if n_out not in SUPPORTED_N and not (2 <= n_out <= BLUESTEIN_MAX_N):
        raise ValueError(
            f"xmris_b200: transform length {n_out} is not supported (powers of two in [16, 8192], or any length up to "
            f"{BLUESTEIN_MAX_N} through the chirp-z path); there is no CPU fallback"
        )


_bluestein_tables = {}


def _bluestein_plan(n: int, inverse: bool, device):
    """Host float64 tables of the chirp-z (Bluestein) transform of length n over M = 2^ceil(log2(2n-1)) points.

    X_k = c_k * sum_j (x_j c_j) conj(c)_(k-j),  c_j = exp(-/+ i pi j^2 / n)   (j^2 reduced mod 2n in integers: exact phase)
    """
    key = (n, bool(inverse), device.type, device.index)
    plan = _bluestein_tables.get(key)
    if plan is None:
        torch = _torch()
        m = 1
        while m < 2 * n - 1:
            m *= 2
        m = max(m, 16)
        j = np.arange(n, dtype=np.int64)
        ang = np.pi * ((j * j) % (2 * n)).astype(np.float64) / n
        chirp = np.exp((1j if inverse else -1j) * ang)               # c_j
        b = np.zeros(m, dtype=np.complex128)
        b[:n] = np.conj(chirp)
        b[m - n + 1:] = np.conj(chirp[1:][::-1])
        # the filter's spectrum (applied between the two power-of-two FFTs) comes from K1 as well: no host FFT on the path
        b_dev = torch.from_numpy(b.astype(np.complex64)).to(device).reshape(1, m)
        filt_dev, _, _ = fid_to_spectrum(b_dev, n_out=m, scale=1.0, out_shift=0)
        plan = dict(m=m, chirp=chirp, filt_dev=filt_dev.reshape(m).contiguous(),
                    chirp_dev=torch.from_numpy(chirp.astype(np.complex64)).to(device))
        _bluestein_tables[key] = plan
    return plan


def _fft_any_length(fid, n_out, pad_left, window, scale, inverse, in_shift, out_shift, stream=None):
    """DFT of arbitrary length n_out <= BLUESTEIN_MAX_N (e.g. the 1972-point Bruker FIDs): chirp-z over K1's
    power-of-two transforms.  Five launches of the library's own kernels (pre-chirp with the input rotation folded in, FFT_M,
    filter, IFFT_M, post-chirp with the fftshift folded in, reading the first n_out points of the M-point rows in place);
    not a fused fast path."""
    torch = _torch()
    lib = _lib.load()
    n_in = fid.shape[-1]
    batch_shape = tuple(fid.shape[:-1])
    flat = fid.reshape(-1, n_in)
    if not flat.is_contiguous():
        flat = flat.contiguous()
    batch = flat.shape[0]
    plan = _bluestein_plan(n_out, inverse, fid.device)
    m, chirp = plan["m"], plan["chirp"]
    w = np.full(n_out, 1.0 if scale is None else float(scale)) if window is None else np.asarray(window, dtype=np.float64)
    pre = torch.from_numpy(np.ascontiguousarray((w * chirp)[pad_left:pad_left + n_in].astype(np.complex64))).to(fid.device)
    y = torch.empty_like(flat)
    out = torch.empty((batch, n_out), dtype=flat.dtype, device=flat.device)
    with torch.cuda.device(fid.device):
        # y[k] = fid[(k + in_shift) mod n_in] * (w c)[pad_left + k]
        _lib.check(lib.xmr_rotate_rows_shift_c64(_ptr(flat), n_in, _ptr(y), batch, n_in, _ptr(pre), int(in_shift) % max(n_in, 1),
                                                 0, _stream_ptr(stream)))
    big, _, _ = fid_to_spectrum(y, n_out=m, pad_left=int(pad_left), scale=1.0, out_shift=0, stream=stream)
    big = _rotate_rows_dev(big, plan["filt_dev"], stream)
    conv, _, _ = fid_to_spectrum(big, inverse=True, scale=1.0 / m, in_shift=0, out_shift=0, stream=stream)
    with torch.cuda.device(fid.device):
        # out[(j + out_shift) mod n_out] = conv[b, j] * c_j,  j < n_out  (rows of conv are m points long)
        _lib.check(lib.xmr_rotate_rows_shift_c64(_ptr(conv), m, _ptr(out), batch, n_out, _ptr(plan["chirp_dev"]), 0,
                                                 int(out_shift) % max(n_out, 1), _stream_ptr(stream)))
    return out.reshape(batch_shape + (n_out,))


def _rotate_rows_dev(x, rot_dev, stream=None):
    torch = _torch()
    lib = _lib.load()
    n = x.shape[-1]
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_rotate_rows_c64(_ptr(x), _ptr(out), x.numel() // max(n, 1), n, _ptr(rot_dev), _stream_ptr(stream)))
    return out


# ---------------------------------------------------------------------------------------------------------
# windows
# ---------------------------------------------------------------------------------------------------------


def split_window(w: np.ndarray, n_out: int):
    """Factor a float64 window ``w[n_out]`` as ``cols[n2] * rows[n1]`` (n = M*n1 + n2, M = min(n_out, 256)).

    Exponential windows on a uniform time axis factor exactly (``apodize_exp``, ``fid.py:132-136``); anything else
    falls back to the full table.  Returns ``(mode, table_float32, rows_float32_or_None)``.
    """
    w = np.asarray(w, dtype=np.float64)
    if w.shape != (n_out,):
        raise ValueError(f"window must have shape ({n_out},), got {w.shape}")
    m = min(n_out, 256)
    r0 = n_out // m
    if r0 == 1:
        return _lib.WIN_SEPARABLE, w.astype(np.float32), np.ones(1, np.float32)
    if w[0] != 0.0 and np.all(np.isfinite(w)):
        cols = w[:m]
        rows = w[::m] / w[0]
        recon = (rows[:, None] * cols[None, :]).ravel()
        if np.allclose(recon, w, rtol=1e-9, atol=1e-300):
            return _lib.WIN_SEPARABLE, cols.astype(np.float32), rows.astype(np.float32)
    return _lib.WIN_TABLE, w.astype(np.float32), None


class PreparedWindow:
    """A window already factored and uploaded to one device (avoids a small H2D copy per launch)."""

    def __init__(self, window, n_out, device):
        torch = _torch()
        self.n_out = n_out
        self.mode, table, rows = split_window(window, n_out)
        self.dev = torch.from_numpy(np.ascontiguousarray(table)).to(device)
        self.rows = rows
        self.device = device


# ---------------------------------------------------------------------------------------------------------
# K1: fused FID -> spectrum
# ---------------------------------------------------------------------------------------------------------


def fid_to_spectrum(fid, n_out=None, pad_left=0, window=None, scale=None, inverse=False, in_shift=0, out_shift=None,
                    want_stats=False, phase_turns=None, out=None, store=True, stream=None, want_index=True):
    """Fused zero-fill -> window -> FFT -> shift [-> phase] on the device.

    fid          complex64 CUDA tensor ``[..., n_in]``
    n_out        transform length (>= n_in); zero filling is implicit
    window       float64 numpy ``[n_out]`` (or a :class:`PreparedWindow`) multiplied before the transform (should include ``1/sqrt(n_out)`` for the
                 reference's ortho norm); ``None`` -> multiply by ``scale`` (default ``1/sqrt(n_out)``)
    out_shift    index rotation of the stored bins; default ``n_out//2`` (= fftshift, ``fourier.py:31-32``)
    phase_turns  ``(a, b)``: multiply stored bin m by ``exp(2 pi i (a + b m))``
    Returns ``(spectrum or None, absmax or None, argmax or None)``.
    """
    torch = _torch()
    lib = _lib.load()
    _require_cuda(fid, "fid")
    n_in = fid.shape[-1]
    n_out = n_in if n_out is None else int(n_out)
    check_length(n_out)
    batch_shape = tuple(fid.shape[:-1])
    batch = int(np.prod(batch_shape)) if batch_shape else 1
    if out_shift is None:
        out_shift = n_out // 2
    if scale is None:
        scale = 1.0 / math.sqrt(n_out)
    if n_out not in SUPPORTED_N:
        # arbitrary length: chirp-z composition of the power-of-two kernels (statistics / phase applied afterwards)
        if isinstance(window, PreparedWindow):
            raise ValueError("PreparedWindow is only defined for power-of-two lengths")
        spec = _fft_any_length(fid, n_out, int(pad_left), window, scale, bool(inverse), int(in_shift), int(out_shift),
                               stream)
        if phase_turns is not None:
            m = np.arange(n_out)
            spec = rotate_rows(spec, np.exp(2j * np.pi * (float(phase_turns[0]) + float(phase_turns[1]) * m)), stream)
        absmax = argmax = None
        if want_stats:
            absmax, argmax = row_absmax(spec, stream)
            if not want_index:
                argmax = None
        if out is not None and store:
            out.copy_(spec)
            spec = out
        return (spec if store else None), absmax, argmax
    dev = fid.device
    spec = None
    if store:
        if out is None:
            spec = torch.empty(batch_shape + (n_out,), dtype=torch.complex64, device=dev)
        else:
            _require_cuda(out, "out")
            if tuple(out.shape) != batch_shape + (n_out,):
                raise ValueError("out has the wrong shape")
            spec = out
    absmax = argmax = None
    if want_stats:
        absmax = torch.empty(batch_shape, dtype=torch.float32, device=dev)
        if want_index:
            argmax = torch.empty(batch_shape, dtype=torch.int32, device=dev)
    win_mode, win_dev, rows = _lib.WIN_NONE, None, None
    if isinstance(window, PreparedWindow):
        if window.n_out != n_out or window.device != dev:
            raise ValueError("PreparedWindow was built for another length / device")
        win_mode, win_dev, rows = window.mode, window.dev, window.rows
    elif window is not None:
        win_mode, table, rows = split_window(window, n_out)
        win_dev = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    rows_arr = (ctypes.c_float * 32)(*([1.0] * 32))
    if rows is not None:
        for i, r in enumerate(rows):
            rows_arr[i] = float(r)
    pa, pb = (0.0, 0.0) if phase_turns is None else (float(phase_turns[0]), float(phase_turns[1]))
    with torch.cuda.device(dev):
        rc = lib.xmr_fid_to_spectrum_c64(
            _ptr(fid), _ptr(spec), batch, n_in, n_out, int(pad_left), win_mode, _ptr(win_dev),
            ctypes.cast(rows_arr, ctypes.c_void_p), ctypes.c_float(scale), int(bool(inverse)), int(in_shift),
            int(out_shift), _ptr(absmax), _ptr(argmax),
            _lib.PHASE_NONE if phase_turns is None else _lib.PHASE_UNIFORM, pa, pb, _stream_ptr(stream))
    _lib.check(rc)
    return spec, absmax, argmax


def fid_absmax_pruned(fid, n_out=None, pad_left=0, window=None, absmax=None, running=None, reset=True, stream=None):
    """Per-spectrum ``max |S|`` for the global-argmax pass with branch-and-bound pruning (entries of spectra that provably
    cannot hold the global maximum are 0).  ``running``: 1-element float32 CUDA tensor shared by the chunks of one data set.
    Returns ``(absmax, running)``."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(fid, "fid")
    n_in = fid.shape[-1]
    n_out = n_in if n_out is None else int(n_out)
    if n_out not in SUPPORTED_N:
        raise ValueError("fid_absmax_pruned needs a power-of-two transform length")
    batch = fid.numel() // n_in
    dev = fid.device
    if absmax is None:
        absmax = torch.empty(tuple(fid.shape[:-1]), dtype=torch.float32, device=dev)
    if running is None:
        running = torch.zeros(1, dtype=torch.float32, device=dev)
    win_mode, win_dev, rows = _lib.WIN_NONE, None, None
    if isinstance(window, PreparedWindow):
        win_mode, win_dev, rows = window.mode, window.dev, window.rows
    elif window is not None:
        win_mode, table, rows = split_window(window, n_out)
        win_dev = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    rows_arr = (ctypes.c_float * 32)(*([1.0] * 32))
    if rows is not None:
        for i, r in enumerate(rows):
            rows_arr[i] = float(r)
    with torch.cuda.device(dev):
        _lib.check(lib.xmr_fid_absmax_pruned_c64(_ptr(fid), batch, n_in, n_out, int(pad_left), win_mode, _ptr(win_dev),
                                                 ctypes.cast(rows_arr, ctypes.c_void_p), ctypes.c_float(1.0 / math.sqrt(n_out)),
                                                 _ptr(absmax), _ptr(running), int(bool(reset)), _stream_ptr(stream)))
    return absmax, running


# ---------------------------------------------------------------------------------------------------------
# un-fused elementwise pieces (each accessor call on its own)
# ---------------------------------------------------------------------------------------------------------


def zero_fill(x, n_out, pad_left=0, stream=None):
    torch = _torch()
    lib = _lib.load()
    _require_cuda(x, "x")
    n_in = x.shape[-1]
    batch = x.numel() // max(n_in, 1)
    out = torch.empty(tuple(x.shape[:-1]) + (n_out,), dtype=torch.complex64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_zero_fill_c64(_ptr(x), _ptr(out), batch, n_in, n_out, int(pad_left), _stream_ptr(stream)))
    return out


def roll_rows(x, shift, stream=None):
    """``np.roll(x, shift, axis=-1)`` on the device (fftshift / ifftshift)."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(x, "x")
    n = x.shape[-1]
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_roll_rows_c64(_ptr(x), _ptr(out), x.numel() // max(n, 1), n, int(shift), _stream_ptr(stream)))
    return out


def scale_rows(x, weights, stream=None):
    """``x * weights`` along the last axis; ``weights`` float64 numpy (rounded to float32 on upload)."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(x, "x")
    n = x.shape[-1]
    w = torch.from_numpy(np.ascontiguousarray(np.asarray(weights, dtype=np.float64).astype(np.float32))).to(x.device)
    if w.shape != (n,):
        raise ValueError("weights must match the last axis")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_scale_rows_c64(_ptr(x), _ptr(out), x.numel() // max(n, 1), n, _ptr(w), _stream_ptr(stream)))
    return out


def rotate_rows(x, rot, stream=None):
    """``x * rot`` along the last axis; ``rot`` complex128 numpy of unit phasors (rounded to complex64)."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(x, "x")
    n = x.shape[-1]
    r = torch.from_numpy(np.ascontiguousarray(np.asarray(rot, dtype=np.complex128).astype(np.complex64))).to(x.device)
    if r.shape != (n,):
        raise ValueError("rot must match the last axis")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_rotate_rows_c64(_ptr(x), _ptr(out), x.numel() // max(n, 1), n, _ptr(r), _stream_ptr(stream)))
    return out


def phase_each(x, a_turns, b_turns, stream=None):
    """Per-spectrum phase: ``x[b, m] * exp(2 pi i (a_turns[b] + b_turns[b] m))`` (float64 CUDA tensors [batch])."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(x, "x")
    n = x.shape[-1]
    batch = x.numel() // max(n, 1)
    a = a_turns.to(device=x.device, dtype=torch.float64).contiguous().reshape(-1)
    b = b_turns.to(device=x.device, dtype=torch.float64).contiguous().reshape(-1)
    if a.numel() != batch or b.numel() != batch:
        raise ValueError("a_turns / b_turns must have one entry per spectrum")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_phase_each_c64(_ptr(x), _ptr(out), batch, n, _ptr(a), _ptr(b), _stream_ptr(stream)))
    return out


def global_argmax(absmax, argmax, n, stream=None):
    """First-occurrence global argmax over per-spectrum maxima.  Returns ``(max_value, flat_index)`` (host sync).

    ``argmax=None``: only the winning row is wanted; the flat index is ``row * n``."""
    torch = _torch()
    lib = _lib.load()
    out = torch.zeros(16, dtype=torch.uint8, device=absmax.device)
    with torch.cuda.device(absmax.device):
        _lib.check(lib.xmr_global_argmax(_ptr(absmax), _ptr(argmax), absmax.numel(), int(n), _ptr(out),
                                         _stream_ptr(stream)))
    host = out.cpu().numpy()
    return float(host[:4].view(np.float32)[0]), int(host[8:16].view(np.int64)[0])


def row_absmax(spec, stream=None):
    """Per-spectrum ``max |S|`` (float32) and its first index (int32) along the last axis."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(spec, "spec")
    n = spec.shape[-1]
    bshape = tuple(spec.shape[:-1])
    absmax = torch.empty(bshape, dtype=torch.float32, device=spec.device)
    argmax = torch.empty(bshape, dtype=torch.int32, device=spec.device)
    with torch.cuda.device(spec.device):
        _lib.check(lib.xmr_row_absmax_c64(_ptr(spec), spec.numel() // max(n, 1), n, _ptr(absmax), _ptr(argmax),
                                          _stream_ptr(stream)))
    return absmax, argmax


_workspaces = {}


def _workspace(device):
    torch = _torch()
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.empty(int(_lib.load().xmr_autophase_workspace_bytes()), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_als_workspaces: dict = {}
_ALS_KEEP_BYTES = 256 << 20


def baseline_als(x, lam=1e5, p=0.001, n_iter=10, out=None, stream=None):
    """Real part of ``x[..., n]`` (complex64 or float32 CUDA tensor) minus its asymmetric-least-squares baseline along the
    last axis (``xmr_baseline_als``).  Returns a float32 tensor of the same shape."""
    torch = _torch()
    lib = _lib.load()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("x must be a CUDA torch tensor (xmris_b200 has no CPU path)")
    if x.dtype not in (torch.complex64, torch.float32):
        raise TypeError(f"baseline_als works on complex64 or float32 device tensors, got {x.dtype}")
    x = x.contiguous()
    n = x.shape[-1]
    batch = x.numel() // max(n, 1)
    if out is None:
        out = torch.empty(tuple(x.shape), dtype=torch.float32, device=x.device)
    need = int(lib.xmr_baseline_als_workspace_bytes(batch, int(n)))
    key = (x.device.type, x.device.index)
    ws = _als_workspaces.get(key)
    if ws is None or ws.numel() < need:
        # the factor scratch can reach many GB (9 GB at n = 4096 with every SM busy): only small ones are kept between
        # calls, large ones go back to torch's allocator when this call returns
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=x.device)
        if need <= _ALS_KEEP_BYTES:
            _als_workspaces[key] = ws
        elif stream is not None and hasattr(stream, "cuda_stream"):
            ws.record_stream(stream)          # freed on return: keep the allocator from reusing it before the kernel is done
    with torch.cuda.device(x.device):
        _lib.check(lib.xmr_baseline_als(_ptr(x), int(x.dtype == torch.complex64), _ptr(out), batch, int(n), float(lam),
                                        float(p), int(n_iter), _ptr(ws), int(ws.numel()), _stream_ptr(stream)))
    return out


_chain_workspaces: dict = {}


_exchange_buffers: dict = {}


def chain_single_dev(fid, n_out, pad_left, window, du, method="acme", index_width=1, p0_only=False, fixed=None, out=None,
                     stream=None, all_gather=None, row_offset=0):
    """``mode="single"`` chain on a device-resident ``[batch, n_in]`` tensor through ONE C-ABI call
    (``xmr_chain_single_dev_c64``: pass 1, argmax, winning spectrum, search, pass 2 -- no Python between the launches,
    no host read-back between the passes, the front part replayed as a CUDA graph).

    ``window``: a :class:`PreparedWindow` or None; ``fixed``: ``None`` or ``(u0, target_idx)`` when ``target_coord`` is
    given.  ``all_gather`` (multi-GPU, voxels sharded over ranks): a callable ``(recv [world, slot] uint8, send [slot]
    uint8) -> None`` that all-gathers the ranks' candidate slots on the current stream (``torch.distributed``); the chain
    then runs as ``xmr_chain_single_front_c64`` -> that ONE collective -> ``xmr_chain_single_back_c64`` and every rank
    searches the global winner redundantly; ``row_offset`` = first global row of this rank's shard.
    Returns ``(spectrum, result)`` with ``result = [p0, p1, pivot index, fun, max |S|, winning (global) row]`` (host)."""
    torch = _torch()
    lib = _lib.load()
    _require_cuda(fid, "fid")
    if method not in _lib.METHODS:
        raise ValueError("Method must be 'acme', 'peak_minima', or 'positivity'")
    n_in = fid.shape[-1]
    batch = fid.numel() // n_in
    dev = fid.device
    if out is None:
        out = torch.empty((batch, n_out), dtype=torch.complex64, device=dev)
    desc = _lib.HostChainDesc()
    desc.n_in, desc.n_out, desc.pad_left = int(n_in), int(n_out), int(pad_left)
    desc.scale = 0.0
    desc.autophase_mode = 1
    desc.method = _lib.METHODS[method]
    desc.index_width = int(index_width)
    desc.p0_only = int(bool(p0_only))
    desc.du = float(du)
    if fixed is not None:
        desc.fixed_pivot = 1
        desc.u0_fixed = float(fixed[0])
        desc.fixed_target = int(fixed[1])
    win_mode, win_dev, rows = _lib.WIN_NONE, None, None
    if window is not None:
        if not isinstance(window, PreparedWindow):
            window = PreparedWindow(window, n_out, dev)
        win_mode, win_dev, rows = window.mode, window.dev, window.rows
    rows_arr = (ctypes.c_float * 32)(*([1.0] * 32))
    if rows is not None:
        for i, r in enumerate(rows):
            rows_arr[i] = float(r)
    need = int(lib.xmr_chain_single_workspace_bytes(batch, int(n_out)))
    key = (dev.type, dev.index)
    ws = _chain_workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _chain_workspaces[key] = ws
    result = (ctypes.c_double * 6)()
    with torch.cuda.device(dev):
        if all_gather is None:
            _lib.check(lib.xmr_chain_single_dev_c64(ctypes.byref(desc), _ptr(fid), _ptr(out), batch, win_mode, _ptr(win_dev),
                                                    ctypes.cast(rows_arr, ctypes.c_void_p), _ptr(ws),
                                                    ctypes.cast(result, ctypes.c_void_p), _stream_ptr(stream)))
        else:
            world = int(all_gather.world_size)
            slot_bytes = int(lib.xmr_chain_single_slot_bytes(int(n_in)))
            bkey = key + (slot_bytes, world)
            bufs = _exchange_buffers.get(bkey)
            if bufs is None:
                bufs = (torch.zeros(slot_bytes, dtype=torch.uint8, device=dev),
                        torch.zeros((world, slot_bytes), dtype=torch.uint8, device=dev))
                _exchange_buffers[bkey] = bufs
            send, recv = bufs
            _lib.check(lib.xmr_chain_single_front_c64(ctypes.byref(desc), _ptr(fid), batch, win_mode, _ptr(win_dev),
                                                      ctypes.cast(rows_arr, ctypes.c_void_p), _ptr(ws), int(row_offset),
                                                      _ptr(send), _stream_ptr(stream)))
            all_gather(recv, send)                       # the ONE collective of the chain
            _lib.check(lib.xmr_chain_single_back_c64(ctypes.byref(desc), _ptr(fid), _ptr(out), batch, win_mode, _ptr(win_dev),
                                                     ctypes.cast(rows_arr, ctypes.c_void_p), _ptr(ws), _ptr(recv), world,
                                                     ctypes.cast(result, ctypes.c_void_p), _stream_ptr(stream)))
    return out, [float(v) for v in result]


def chain_single_last_timing():
    """``(front ms, pass 2 ms)`` of the calling thread's last device chain (CUDA events on its stream; waits for pass 2)."""
    ms = (ctypes.c_double * 2)()
    _lib.check(_lib.load().xmr_chain_single_last_timing(ctypes.cast(ms, ctypes.c_void_p)))
    return float(ms[0]), float(ms[1])


def autophase_search(spec1d, u0, du, method="acme", target_idx=0, index_width=1, p0_only=False, stream=None):
    """Global (p0, p1) search on one device-resident spectrum.  Returns a CUDA float64 tensor ``[p0, p1, fun, 0]``.

    ``u_m = u0 + du*m`` is the reference's normalised phase ramp ``(x_m - pivot)/(x_max - x_min)``.
    """
    torch = _torch()
    lib = _lib.load()
    _require_cuda(spec1d, "spec1d")
    if spec1d.dim() != 1:
        raise ValueError("autophase_search works on one 1-D spectrum")
    if method not in _lib.METHODS:
        raise ValueError("Method must be 'acme', 'peak_minima', or 'positivity'")
    n = spec1d.shape[0]
    result = torch.empty(4, dtype=torch.float64, device=spec1d.device)
    ws = _workspace(spec1d.device)
    with torch.cuda.device(spec1d.device):
        _lib.check(lib.xmr_autophase_search_c64(_ptr(spec1d), n, float(u0), float(du), _lib.METHODS[method],
                                                int(target_idx), int(index_width), int(bool(p0_only)), _ptr(result),
                                                _ptr(ws), _stream_ptr(stream)))
    return result


def autophase_score(spec1d, u0, du, p0, p1, method="acme", target_idx=0, index_width=1, float64=True, stream=None):
    """The autophase objective at the candidates ``(p0[k], p1[k])`` (degrees) on one device-resident spectrum.

    What the searches minimise, evaluated by the same device code (``xmr_autophase_score_c64``); the reference's
    ``_acme_score`` / ``_peak_minima_score`` / ``_roi_positivity_score`` (``phasing.py:100-157``).  Returns a CUDA float64
    tensor ``[k]``.
    """
    torch = _torch()
    lib = _lib.load()
    _require_cuda(spec1d, "spec1d")
    if spec1d.dim() != 1:
        raise ValueError("autophase_score works on one 1-D spectrum")
    if method not in _lib.METHODS:
        raise ValueError("Method must be 'acme', 'peak_minima', or 'positivity'")
    dev = spec1d.device
    p0_t = torch.as_tensor(np.atleast_1d(np.asarray(p0, dtype=np.float64))).to(dev)
    p1_t = torch.as_tensor(np.atleast_1d(np.asarray(p1, dtype=np.float64))).to(dev)
    if p0_t.shape != p1_t.shape:
        raise ValueError("p0 and p1 must have the same length")
    out = torch.empty(p0_t.shape[0], dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xmr_autophase_score_c64(_ptr(spec1d.contiguous()), spec1d.shape[0], float(u0), float(du),
                                               _lib.METHODS[method], int(target_idx), int(index_width), _ptr(p0_t), _ptr(p1_t),
                                               int(p0_t.shape[0]), int(bool(float64)), _ptr(out), _stream_ptr(stream)))
    return out
