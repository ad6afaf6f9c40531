"""CPU: the kernel's thread choreography (csrc/fft_stages.cuh, shared with the CUDA kernel) emulated thread by thread
in float32 and compared with the oracle -- index algebra, shared-memory layouts, twiddles, zero-fill, shifts."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_l2
from oracle import xmris_oracle as orc

CSRC = os.path.join(ROOT, "xmris_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    path = os.path.join(CSRC, "libxmris_emul.so")
    src = [os.path.join(CSRC, f) for f in ("host_emul.cpp", "fft_stages.cuh", "fft_regs.cuh")]
    if not os.path.isfile(path) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in src):
        subprocess.run(["make", "-C", CSRC, "emul"], check=True, capture_output=True)
    lib = ctypes.CDLL(path)
    f = lib.xmr_emul_fft_c64
    vp, i = ctypes.c_void_p, ctypes.c_int
    f.argtypes = [vp, vp, ctypes.c_longlong, i, i, i, vp, ctypes.c_float, vp, i, i, i, i]
    f.restype = i
    return f


def run(emul, x, n_out, pad_left=0, w=None, inverse=0, in_shift=0, out_shift=None, persist=1):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    batch, n_in = x.shape
    out = np.zeros((batch, n_out), np.complex64)
    tw = np.exp(-2j * np.pi * np.arange(n_out) / n_out).astype(np.complex64)
    wt = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
    rc = emul(x.ctypes.data, out.ctypes.data, batch, n_in, n_out, pad_left, None if wt is None else wt.ctypes.data,
              1.0 / np.sqrt(n_out), tw.ctypes.data, inverse, in_shift, n_out // 2 if out_shift is None else out_shift,
              persist)
    assert rc == 0
    return out


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("persist", [0, 1])
def test_to_spectrum_all_lengths(emul, n, persist):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    got = run(emul, x, n, persist=persist)
    ref, _ = orc.to_spectrum(x.astype(np.complex64).astype(np.complex128), 1, np.arange(n) * 1e-3)
    assert rel_l2(got, ref) < 5e-7


@pytest.mark.parametrize("n_in,n_out,position", [(1024, 2048, "end"), (4096, 8192, "end"), (32, 128, "symmetric"),
                                                   (1000, 4096, "symmetric"), (37, 64, "symmetric"), (5, 16, "end")])
def test_chain_with_zero_fill_and_window(emul, n_in, n_out, position):
    rng = np.random.default_rng(n_in + n_out)
    x = (rng.standard_normal((3, n_in)) + 1j * rng.standard_normal((3, n_in))).astype(np.complex64)
    t = np.arange(n_in) * 2e-4
    ref, _ = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out, position, 5.0)
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, position)
    pad_left = 0 if position == "end" else (n_out - n_in) // 2
    w = np.exp(-np.pi * 5.0 * t_pad) / np.sqrt(n_out)
    got = run(emul, x, n_out, pad_left=pad_left, w=w)
    assert rel_l2(got, ref) < 5e-7


@pytest.mark.parametrize("n", [16, 128, 256, 1024, 4096])
def test_to_fid_inverse_with_input_unshift(emul, n):
    rng = np.random.default_rng(n + 3)
    s = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    freqs = np.roll(np.fft.fftfreq(n, d=1e-3), n // 2)
    ref, _ = orc.to_fid(s.astype(np.complex128), 1, freqs)
    got = run(emul, s, n, inverse=1, in_shift=n // 2, out_shift=0)
    assert rel_l2(got, ref) < 5e-7


@pytest.mark.parametrize("n_out,zf", [(512, 2), (1024, 2), (2048, 2), (4096, 2), (8192, 2), (1024, 4), (2048, 4), (4096, 4),
                                      (8192, 4)])
@pytest.mark.parametrize("persist", [0, 1])
def test_zero_filled_fast_variant_choreography(emul, n_out, zf, persist):
    """K1_FAST_ZF2 / ZF4: the zero rows of a stage-0 column are never read and the first butterfly layers degenerate."""
    lib = ctypes.CDLL(os.path.join(CSRC, "libxmris_emul.so"))
    f = lib.xmr_emul_fft_zf_c64
    vp, i = ctypes.c_void_p, ctypes.c_int
    f.argtypes = [vp, vp, ctypes.c_longlong, i, i, vp, ctypes.c_float, vp, i]
    f.restype = i
    n_in = n_out // zf
    rng = np.random.default_rng(n_out + zf)
    x = (rng.standard_normal((3, n_in)) + 1j * rng.standard_normal((3, n_in))).astype(np.complex64)
    t = np.arange(n_in) * 2e-4
    ref, _ = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out, "end", 5.0)
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, "end")
    w = np.ascontiguousarray(np.exp(-np.pi * 5.0 * t_pad) / np.sqrt(n_out), dtype=np.float32)
    # poison the part of the table that belongs to the zero rows: the fast variant must never touch it
    w[n_in:] = np.nan
    out = np.zeros((3, n_out), np.complex64)
    tw = np.exp(-2j * np.pi * np.arange(n_out) / n_out).astype(np.complex64)
    assert f(x.ctypes.data, out.ctypes.data, 3, n_out, zf, w.ctypes.data, 1.0 / np.sqrt(n_out), tw.ctypes.data, persist) == 0
    assert rel_l2(out, ref) < 5e-7
