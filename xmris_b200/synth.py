"""Seeded synthetic Lorentzian-metabolite FIDs for the parity tests and the benchmark.

The signal model is the one the reference's simulator uses (AMARES eq. 6, a sum of damped complex
exponentials; ``fitting/simulation.py:87-96``) with its time-domain noise convention
(``sigma_channel = mean(abs(x[0:10])) / SNR / sqrt(2)``, ``fitting/simulation.py:180-196``):

    x[b, k] = sum_p a[b,p] * exp(i*phi0[b]) * exp((-d[p] + 2*pi*i*f[b,p]) * t[b,k]) + noise,
    t[b, k] = k / sw + t_dead[b]

``phi0`` is a per-voxel zero-order phase error and ``t_dead`` a per-voxel acquisition delay, i.e. a
first-order phase error of ``p1 = 360 * sw * t_dead`` degrees across the band -- the distortions
``autophase`` exists to remove.  Noise keeps ``max(Re) > 0`` for every phase pair, which keeps the ACME
objective away from its pole (SURVEY.md finding 5).

Two implementations of the same formula: numpy/float64 on the host (tests, CPU baseline) and torch on
the device (benchmark inputs are created in HBM, never copied from the host).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Family:
    name: str
    sw: float
    amplitudes: tuple
    frequencies: tuple
    dampings: tuple
    amp_scale: tuple = (0.5, 1.5)      # per-voxel uniform amplitude scale
    freq_jitter: float = 3.0           # per-voxel, per-peak N(0, sigma) Hz
    p0_range: tuple = (-180.0, 180.0)  # per-voxel zero-order error, degrees
    p1_range: tuple = (-1000.0, 1000.0)  # per-voxel first-order error over the band, degrees
    snr_range: tuple = (5.0, 30.0)


# 1H-like MRSI (configs C2, C3, C5) and 13C-like dynamic series (C4; docs/notebooks/pipeline/autophasing.md:179-183)
PROTON = Family("1H", 5000.0, (100.0, 60.0, 40.0, 20.0), (-700.0, -300.0, 250.0, 900.0), (30.0, 25.0, 25.0, 40.0))
CARBON = Family("13C", 5000.0, (100.0, 20.0), (-128.4, 256.8), (15.0, 15.0),
                amp_scale=(0.1, 1.0), freq_jitter=1.0, p1_range=(-600.0, 600.0), snr_range=(4.0, 15.0))

FAMILIES = {"1H": PROTON, "13C": CARBON}


def time_coord(n_points: int, sw: float) -> np.ndarray:
    """The time coordinate the accessor sees: ``arange(n)/sw`` (dead time is a hidden acquisition error)."""
    return np.arange(n_points, dtype=np.float64) / sw


def _draw_params_numpy(fam: Family, batch: int, rng: np.random.Generator):
    P = len(fam.amplitudes)
    scale = rng.uniform(*fam.amp_scale, size=(batch, 1))
    amps = scale * np.asarray(fam.amplitudes)[None, :]
    freqs = np.asarray(fam.frequencies)[None, :] + fam.freq_jitter * rng.standard_normal((batch, P))
    p0 = rng.uniform(*fam.p0_range, size=batch)
    p1 = rng.uniform(*fam.p1_range, size=batch)
    snr = rng.uniform(*fam.snr_range, size=batch)
    return amps, freqs, p0, p1, snr


def make_fids_numpy(family: str | Family, batch: int, n_points: int, seed: int = 1234, dtype=np.complex128):
    """Host generator.  Returns ``(fid[batch, n_points], time_coord[n_points], params dict)``."""
    fam = FAMILIES[family] if isinstance(family, str) else family
    rng = np.random.default_rng(seed)
    amps, freqs, p0, p1, snr = _draw_params_numpy(fam, batch, rng)
    t_dead = p1 / 360.0 / fam.sw
    t = np.arange(n_points)[None, :] / fam.sw + t_dead[:, None]                     # (B, N)
    damp = np.asarray(fam.dampings)
    x = np.zeros((batch, n_points), dtype=np.complex128)
    for p in range(len(fam.amplitudes)):
        x += amps[:, p, None] * np.exp((-damp[p] + 2j * np.pi * freqs[:, p, None]) * t)
    x *= np.exp(1j * np.radians(p0))[:, None]
    sig = np.mean(np.abs(x[:, : min(10, n_points)]), axis=1) / snr / np.sqrt(2.0)
    noise = rng.standard_normal((batch, n_points)) + 1j * rng.standard_normal((batch, n_points))
    x += sig[:, None] * noise
    params = dict(amps=amps, freqs=freqs, p0=p0, p1=p1, snr=snr, family=fam.name)
    return x.astype(dtype), time_coord(n_points, fam.sw), params


def make_fids_torch(family: str | Family, batch: int, n_points: int, device, seed: int = 1234, chunk: int = 16384,
                    out=None):
    """Device generator (complex64, batch-major).  Same formula as :func:`make_fids_numpy`, torch RNG stream.

    Generated chunk-wise so that the float32 temporaries stay small next to a 32 GiB batch.
    Returns ``(fid tensor [batch, n_points] complex64 on device, time_coord float64 numpy)``.
    """
    import torch

    fam = FAMILIES[family] if isinstance(family, str) else family
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if out is None:
        out = torch.empty((batch, n_points), dtype=torch.complex64, device=device)
    P = len(fam.amplitudes)
    base_a = torch.tensor(fam.amplitudes, dtype=torch.float32, device=device)
    base_f = torch.tensor(fam.frequencies, dtype=torch.float32, device=device)
    damp = torch.tensor(fam.dampings, dtype=torch.float32, device=device)
    k = torch.arange(n_points, dtype=torch.float32, device=device) / fam.sw

    def uni(lo, hi, shape):
        return lo + (hi - lo) * torch.rand(shape, generator=gen, device=device, dtype=torch.float32)

    for s in range(0, batch, chunk):
        b = min(chunk, batch - s)
        amps = uni(*fam.amp_scale, (b, 1)) * base_a[None, :]
        freqs = base_f[None, :] + fam.freq_jitter * torch.randn((b, P), generator=gen, device=device)
        p0 = torch.deg2rad(uni(*fam.p0_range, (b,)))
        p1 = uni(*fam.p1_range, (b,))
        snr = uni(*fam.snr_range, (b,))
        t = k[None, :] + (p1 / 360.0 / fam.sw)[:, None]
        x = torch.zeros((b, n_points), dtype=torch.complex64, device=device)
        for p in range(P):
            # phase in turns reduced before the sincos keeps float32 accurate at f*t ~ 1e3 cycles
            turns = freqs[:, p, None].double() * t.double()
            turns = (turns - torch.floor(turns)).float()
            mag = amps[:, p, None] * torch.exp(-damp[p] * t)
            x += torch.polar(mag, 2.0 * torch.pi * turns)
        x *= torch.polar(torch.ones_like(p0), p0)[:, None]
        sig = x[:, : min(10, n_points)].abs().mean(dim=1) / snr / (2.0 ** 0.5)
        noise = torch.randn((b, n_points, 2), generator=gen, device=device, dtype=torch.float32)
        x += sig[:, None] * torch.view_as_complex(noise)
        out[s : s + b] = x
    return out, time_coord(n_points, fam.sw)
