"""GPU parity tests of K1 (fused zero-fill -> window -> FFT -> fftshift [-> stats] [-> phase]) against the oracle.

Tolerance: north_star states spectra within 1e-5 relative L2 (complex64 device vs the reference's complex128).
"""

import numpy as np
import pytest

from conftest import rel_l2
from oracle import xmris_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    import torch

    assert torch.cuda.is_available()
    from xmris_b200 import _lib

    _lib.load()  # fail loudly if the CUDA library is not built
    return torch.device("cuda:0")


def _rand(rng, shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


@pytest.mark.parametrize("n_out", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_plain_fft_all_lengths(dev, n_out):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(n_out)
    for batch in (1, 3, 70, 301):
        x = _rand(rng, (batch, n_out))
        t = np.arange(n_out) * 1e-3
        spec, _, _ = D.fid_to_spectrum(torch.from_numpy(x).to(dev))
        ref, _ = orc.to_spectrum(x.astype(np.complex128), 1, t)
        got = spec.cpu().numpy()
        errs = [rel_l2(got[i], ref[i]) for i in range(batch)]
        assert max(errs) < TOL, (n_out, batch, max(errs))


@pytest.mark.parametrize("n_in", [8192, 4096, 2048, 1000])
def test_grouped_ctas_at_8192_every_tail(dev, n_in):
    """N = 8192 runs ONE CTA per SM made of two independent thread groups that share a three-buffer TMA ring: batches
    below / at / just above one, two and three tiles per CTA (148 SMs), idle second groups, ragged tails; every row against
    the oracle, with the statistics and the phase epilogues."""
    import torch
    from xmris_b200 import device as D

    n_out = 8192
    rng = np.random.default_rng(n_in)
    t = np.arange(n_in) / 5000.0
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, "end")
    w = np.exp(-np.pi * 3.0 * t_pad) / np.sqrt(n_out)
    for batch in (1, 2, 147, 148, 149, 295, 296, 297, 445, 593):
        x = _rand(rng, (batch, n_in))
        ref, freqs = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out, "end", 3.0)
        xd = torch.from_numpy(x).to(dev)
        spec, amax, _ = D.fid_to_spectrum(xd, n_out=n_out, window=w, want_stats=True)
        got = spec.cpu().numpy()
        assert max(rel_l2(got[i], ref[i]) for i in range(batch)) < TOL, batch
        np.testing.assert_allclose(amax.cpu().numpy(), np.abs(ref).max(axis=1), rtol=3e-5)
        p0, p1, pivot = -33.0, 1234.5, float(freqs[n_out // 3])
        refp, _ = orc.phase(ref, 1, freqs, p0, p1, pivot)
        span = freqs.max() - freqs.min()
        b = (p1 / 360.0) * (freqs[1] - freqs[0]) / span
        a = p0 / 360.0 + (p1 / 360.0) * (freqs[0] - pivot) / span
        specp, _, _ = D.fid_to_spectrum(xd, n_out=n_out, window=w, phase_turns=(a, b))
        gotp = specp.cpu().numpy()
        assert max(rel_l2(gotp[i], refp[i]) for i in range(batch)) < TOL, batch


@pytest.mark.parametrize("n_in,n_out,position", [(1024, 2048, "end"), (4096, 8192, "end"), (64, 512, "end"),
                                                   (32, 128, "symmetric"), (1000, 4096, "symmetric"),
                                                   (1023, 2048, "end"), (37, 64, "symmetric"), (2048, 2048, "end")])
def test_chain_zero_fill_window(dev, n_in, n_out, position):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(n_in * 7 + n_out)
    batch = 37
    x = _rand(rng, (batch, n_in))
    t = 2e-4 * np.arange(n_in)
    lb = 5.0
    ref, freqs = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out if n_out > n_in else None, position, lb)
    # host-side metadata exactly as the product computes it
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, position)
    pad_left = 0 if position == "end" else (n_out - n_in) // 2
    w = np.exp(-np.pi * lb * t_pad) / np.sqrt(n_out)
    spec, amax, imax = D.fid_to_spectrum(torch.from_numpy(x).to(dev), n_out=n_out, pad_left=pad_left, window=w,
                                         want_stats=True)
    got = spec.cpu().numpy()
    errs = [rel_l2(got[i], ref[i]) for i in range(batch)]
    assert max(errs) < TOL, max(errs)
    # statistics: max |S| and its first index per spectrum
    ref_abs = np.abs(ref)
    np.testing.assert_allclose(amax.cpu().numpy(), ref_abs.max(axis=1), rtol=2e-5)
    gi = imax.cpu().numpy()
    for i in range(batch):
        assert ref_abs[i, gi[i]] >= ref_abs[i].max() * (1 - 2e-5)


def test_non_separable_window_table(dev):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(5)
    n = 4096
    x = _rand(rng, (19, n))
    w = rng.uniform(0.2, 1.0, n)   # arbitrary (non-uniform time axis) window -> full table path
    spec, _, _ = D.fid_to_spectrum(torch.from_numpy(x).to(dev), window=w / np.sqrt(n))
    ref = np.roll(np.fft.fft(x.astype(np.complex128) * w, axis=1, norm="ortho"), n // 2, axis=1)
    assert rel_l2(spec.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("n", [256, 1024, 4096, 8192])
def test_fused_uniform_phase(dev, n):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(n + 1)
    x = _rand(rng, (23, n))
    t = np.arange(n) / 5000.0
    ref, freqs = orc.to_spectrum(x.astype(np.complex128), 1, t)
    p0, p1, pivot = -137.25, 2891.5, float(freqs[n // 3])
    refp, _ = orc.phase(ref, 1, freqs, p0, p1, pivot)
    df = freqs[1] - freqs[0]
    rng_ = freqs.max() - freqs.min()
    b = (p1 / 360.0) * df / rng_
    a = p0 / 360.0 + (p1 / 360.0) * (freqs[0] - pivot) / rng_
    spec, _, _ = D.fid_to_spectrum(torch.from_numpy(x).to(dev), phase_turns=(a, b))
    assert rel_l2(spec.cpu().numpy(), refp) < TOL


@pytest.mark.parametrize("n", [64, 256, 2048, 4096])
def test_inverse_round_trip(dev, n):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(n + 2)
    x = _rand(rng, (11, n))
    t = np.arange(n) / 4000.0
    xd = torch.from_numpy(x).to(dev)
    spec, _, _ = D.fid_to_spectrum(xd)
    back, _, _ = D.fid_to_spectrum(spec, inverse=True, in_shift=n // 2, out_shift=0)
    ref_spec, freqs = orc.to_spectrum(x.astype(np.complex128), 1, t)
    ref_back, _ = orc.to_fid(ref_spec, 1, freqs)
    assert rel_l2(back.cpu().numpy(), ref_back) < TOL
    assert rel_l2(back.cpu().numpy(), x) < TOL


def test_elementwise_ops(dev):
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(9)
    x = _rand(rng, (13, 300))
    xd = torch.from_numpy(x).to(dev)
    z = D.zero_fill(xd, 777, pad_left=100).cpu().numpy()
    ref = np.zeros((13, 777), np.complex64)
    ref[:, 100:400] = x
    np.testing.assert_array_equal(z, ref)
    w = rng.uniform(0, 1, 300)
    np.testing.assert_allclose(D.scale_rows(xd, w).cpu().numpy(), x * w.astype(np.float32), rtol=1e-6)
    rot = np.exp(1j * rng.uniform(-3, 3, 300))
    assert rel_l2(D.rotate_rows(xd, rot).cpu().numpy(), x * rot) < 1e-6
    # the shifted / strided form (pre- and post-chirp of the chirp-z path): out[(j + so) % n] = in[b, (j + si) % n] * rot[j]
    from xmris_b200 import _lib
    lib = _lib.load()
    n = 211                                                     # rows of 300 points, the first 211 of each are used
    rd = torch.from_numpy(rot[:n].astype(np.complex64)).to(dev)
    for si, so in [(0, 0), (5, 0), (0, 105), (17, 200), (-3, -4)]:
        od = torch.empty((13, n), dtype=torch.complex64, device=dev)
        _lib.check(lib.xmr_rotate_rows_shift_c64(xd.data_ptr(), 300, od.data_ptr(), 13, n, rd.data_ptr(), si, so, None))
        want = np.roll(np.roll(x[:, :n], -si, axis=1) * rot[:n].astype(np.complex64), so, axis=1)
        assert rel_l2(od.cpu().numpy(), want) < 1e-6, (si, so)
    assert lib.xmr_rotate_rows_shift_c64(xd.data_ptr(), 100, xd.data_ptr(), 13, n, rd.data_ptr(), 0, 0, None) != 0   # stride < n
    a = torch.from_numpy(rng.uniform(-1, 1, 13)).to(dev)
    b = torch.from_numpy(rng.uniform(-0.01, 0.01, 13)).to(dev)
    got = D.phase_each(xd, a, b).cpu().numpy()
    refp = x * np.exp(2j * np.pi * (a.cpu().numpy()[:, None] + b.cpu().numpy()[:, None] * np.arange(300)[None, :]))
    assert rel_l2(got, refp) < 2e-6


def test_global_argmax_first_occurrence(dev):
    import torch
    from xmris_b200 import device as D

    v = np.array([1.0, 5.0, 3.0, 5.0, 2.0], np.float32)
    i = np.array([7, 11, 2, 1, 0], np.int32)
    val, flat = D.global_argmax(torch.from_numpy(v).to(dev), torch.from_numpy(i).to(dev), 100)
    assert val == 5.0 and flat == 1 * 100 + 11


def test_errors(dev):
    import torch
    from xmris_b200 import device as D

    x = torch.zeros((2, 5000), dtype=torch.complex64, device=dev)
    with pytest.raises(ValueError, match="not supported"):
        D.fid_to_spectrum(x)
    with pytest.raises(TypeError):
        D.fid_to_spectrum(torch.zeros((2, 64), dtype=torch.complex64))


@pytest.mark.parametrize("n_in,n_out", [(1972, 1972), (1000, 1000), (37, 37), (1972, 3000), (100, 1234), (3, 3)])
def test_arbitrary_length_chirp_z(dev, n_in, n_out):
    """Lengths that are not powers of two (the 1972-point Bruker FIDs of docs/notebooks/vendor/bruker_fid_loader.md)."""
    import torch
    from xmris_b200 import device as D

    rng = np.random.default_rng(n_in + n_out)
    x = _rand(rng, (7, n_in))
    t = 2e-4 * np.arange(n_in)
    ref, freqs = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out if n_out > n_in else None, "end", 3.0)
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, "end")
    w = np.exp(-np.pi * 3.0 * (t_pad if n_out > n_in else t)) / np.sqrt(n_out)
    spec, amax, imax = D.fid_to_spectrum(torch.from_numpy(x).to(dev), n_out=n_out, window=w, want_stats=True)
    got = spec.cpu().numpy()
    assert max(rel_l2(got[i], ref[i]) for i in range(7)) < TOL
    np.testing.assert_allclose(amax.cpu().numpy(), np.abs(ref).max(axis=1), rtol=3e-5)
    back, _, _ = D.fid_to_spectrum(spec, inverse=True, in_shift=n_out // 2, out_shift=0)
    ref_back, _ = orc.to_fid(ref, 1, freqs)
    assert rel_l2(back.cpu().numpy(), ref_back) < TOL


# ---- pass 1 of mode="single": branch-and-bound statistics must pick exactly the plain pass's winner --------------------
def _pruned_vs_plain(dev, fid_np, window, n_out=None, pad_left=0):
    import torch

    from xmris_b200 import device as D

    fid = torch.from_numpy(np.ascontiguousarray(fid_np.astype(np.complex64))).to(dev)
    n = fid.shape[-1] if n_out is None else n_out
    pruned, running = D.fid_absmax_pruned(fid, n_out=n, pad_left=pad_left, window=window)
    _, plain, _ = D.fid_to_spectrum(fid, n_out=n, pad_left=pad_left, window=window, store=False, want_stats=True,
                                    want_index=False)
    vp, ip = D.global_argmax(pruned.reshape(-1), None, n)
    vq, iq = D.global_argmax(plain.reshape(-1), None, n)
    pruned, plain = pruned.cpu().numpy(), plain.cpu().numpy()
    kept = pruned > 0
    assert ip == iq, "branch and bound changed the winning row"
    # (N = 8192: both kernels form the stage-0 twiddles by a float32 power chain, contracted differently by the compiler --
    #  the same maximum to an ulp; up to 4096 points the twiddles are table values and the maxima identical bit for bit)
    assert vp == vq or (n >= 8192 and abs(vp - vq) <= 3e-7 * vq)
    assert np.allclose(pruned[kept], plain[kept], rtol=1e-6, atol=0)
    assert np.all(plain[~kept] <= vq)                       # pruned rows could not have won
    assert abs(float(running.item()) ** 0.5 - vq) <= 1e-6 * max(vq, 1e-30)
    return kept.mean()


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
def test_pruned_statistics_pick_the_exact_winner(dev, n):
    from xmris_b200.synth import make_fids_numpy

    rng = np.random.default_rng(n)
    t = np.arange(n) / 5000.0
    window = np.exp(-np.pi * 5.0 * t) / np.sqrt(n)
    # (a) MRSI-like decaying multi-line FIDs (odd batch: a ragged last tile)
    fid, _, _ = make_fids_numpy("1H", 3001, n, seed=n)
    kept = _pruned_vs_plain(dev, fid, window)
    assert kept < 0.9                                        # the bounds do prune on this kind of data
    # (b) undamped single tones of ascending amplitude: every bound is tight, almost nothing can be pruned
    k = rng.integers(0, n, size=257)
    tones = (np.arange(1, 258)[:, None] * np.exp(2j * np.pi * k[:, None] * np.arange(n)[None, :] / n))
    _pruned_vs_plain(dev, tones, None)
    # (c) identical rows: ties resolve to the first row (numpy argmax order, phasing.py:229-231)
    same = np.repeat(fid[:1], 300, axis=0)
    _pruned_vs_plain(dev, same, window)
    # (d) zeros, one impulse row, noise rows
    mixed = np.zeros((65, n), dtype=np.complex128)
    mixed[17, 3] = 5.0
    mixed[40:] = rng.standard_normal((25, n)) + 1j * rng.standard_normal((25, n))
    _pruned_vs_plain(dev, mixed, None)
    _pruned_vs_plain(dev, np.zeros((9, n), dtype=np.complex128), window)
    # (e) separable windows with negative factors (the level-0 bound uses |w|)
    _pruned_vs_plain(dev, fid[:500], -window)
    _pruned_vs_plain(dev, fid[:500], window * np.where((np.arange(n) // min(n, 256)) % 2 == 1, -1.0, 1.0))


@pytest.mark.parametrize("n_in,n_out,pad_left,table", [(2048, 4096, 0, False), (1024, 4096, 0, False), (256, 512, 0, False),
                                                       (256, 1024, 0, False), (512, 2048, 0, False), (128, 512, 0, False),
                                                       (4096, 8192, 0, False), (8192, 8192, 0, False), (2048, 8192, 0, False), (1024, 2048, 0, False),
                                                       (1024, 4096, 1536, False), (2048, 2048, 0, True), (100, 256, 0, False),
                                                       (64, 64, 0, False), (1000, 4096, 7, True)])
def test_pruned_statistics_any_geometry(dev, n_in, n_out, pad_left, table):
    """Zero-filled / long / short transforms and non-separable windows prune on the level-0 bound inside the generic
    statistics kernel: same winner as the plain pass."""
    from xmris_b200.synth import make_fids_numpy

    fid, _, _ = make_fids_numpy("1H", 1501, n_in, seed=n_in + n_out)
    t = (np.arange(n_out) - pad_left) / 5000.0
    window = np.exp(-np.pi * 5.0 * np.abs(t)) / np.sqrt(n_out)
    if table:
        window = window * (1.0 + 0.3 * np.cos(np.arange(n_out) * 0.37))         # does not factor into rows x columns
    kept = _pruned_vs_plain(dev, fid, window, n_out=n_out, pad_left=pad_left)
    if n_in >= 1024:
        assert kept < 0.9
    same = np.repeat(fid[:1], 70, axis=0)
    _pruned_vs_plain(dev, same, window, n_out=n_out, pad_left=pad_left)
    _pruned_vs_plain(dev, np.zeros((5, n_in), dtype=np.complex128), window, n_out=n_out, pad_left=pad_left)


@pytest.mark.parametrize("n_out,factor", [(512, 2), (1024, 2), (2048, 2), (4096, 2), (8192, 2), (1024, 4), (2048, 4),
                                          (4096, 4), (8192, 4)])
def test_zero_filled_fast_variants(dev, n_out, factor):
    """Input zero-filled at the end to 2x / 4x its length (zero_fill's default geometry) takes compile-time specialised
    store / store+phase kernels that never load the zero part: same results as the oracle."""
    import torch
    from xmris_b200 import device as D

    n_in = n_out // factor
    rng = np.random.default_rng(n_out + 3)
    x = _rand(rng, (301, n_in))
    t = np.arange(n_in) / 5000.0
    lb = 5.0
    ref, freqs = orc.chain_to_spectrum(x.astype(np.complex128), 1, t, n_out, "end", lb)
    _, t_pad, _ = orc.zero_fill(np.zeros(n_in), 0, t, n_out, "end")
    w = np.exp(-np.pi * lb * t_pad) / np.sqrt(n_out)
    xd = torch.from_numpy(x).to(dev)
    spec, _, _ = D.fid_to_spectrum(xd, n_out=n_out, window=w)
    assert max(rel_l2(g, r) for g, r in zip(spec.cpu().numpy(), ref)) < TOL
    p0, p1, pivot = 77.5, -1903.25, float(freqs[n_out // 5])
    refp, _ = orc.phase(ref, 1, freqs, p0, p1, pivot)
    rng_ = freqs.max() - freqs.min()
    b = (p1 / 360.0) * (freqs[1] - freqs[0]) / rng_
    a = p0 / 360.0 + (p1 / 360.0) * (freqs[0] - pivot) / rng_
    specp, _, _ = D.fid_to_spectrum(xd, n_out=n_out, window=w, phase_turns=(a, b))
    assert max(rel_l2(g, r) for g, r in zip(specp.cpu().numpy(), refp)) < TOL
