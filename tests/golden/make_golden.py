"""Generate the committed golden vectors by running the REFERENCE's own code in the build container.

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.npz

The reference's five hot-path files (``core/config.py``, ``core/utils.py``, ``processing/fourier.py``,
``processing/fid.py``, ``processing/phasing.py``) are imported unmodified from ``/root/reference`` by
``oracle/ref_loader.py`` on top of the ``xarray_lite`` stand-in; all arithmetic is numpy / scipy
(numpy 2.3.5, scipy 1.18.1 here).  Every case is seeded.  The fixtures are small (inputs are kept
only where they cannot be regenerated bit-exactly from a seed by the tests).
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from xmris_b200.synth import make_fids_numpy  # noqa: E402

ref = load_reference()
xr = ref.xr


def lorentz_fid(n, sw, amps, freqs, damps, phases_deg, snr, seed, dead=0.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sw + dead
    x = np.zeros(n, complex)
    for a, f, d, ph in zip(amps, freqs, damps, phases_deg):
        x += a * np.exp(1j * np.radians(ph)) * np.exp((-d + 2j * np.pi * f) * t)
    sig = np.mean(np.abs(x[:10])) / snr / np.sqrt(2)
    x = x + sig * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x, np.arange(n) / sw


def da_fid(values, t, extra_dims=(), attrs=None):
    dims = list(extra_dims) + ["time"]
    coords = {"time": t}
    for i, d in enumerate(extra_dims):
        coords[d] = np.arange(values.shape[i])
    return xr.DataArray(values, dims=dims, coords=coords, attrs=attrs or {})


def main():
    out = {}

    # ---- C1: README quick-start (README.md:51-73), seeded instead of np.random.randn ---------------
    rng = np.random.default_rng(0)
    t = np.linspace(0, 1, 1024)
    data = rng.standard_normal((5, 1024)) + 1j * rng.standard_normal((5, 1024))
    da = xr.DataArray(data, dims=["voxel", "time"], coords={"voxel": np.arange(5), "time": t},
                      attrs={"MHz": 120.0, "sw": 10000.0})
    zf = ref.fid.zero_fill(da, target_points=2048)
    ap = ref.fid.apodize_exp(zf, lb=5.0)
    sp = ref.fid.to_spectrum(ap)
    ph = ref.phasing.autophase(sp, peak_width=100)  # accessor default peak_width (accessor.py:634)
    out["c1"] = dict(
        fid=data, time=t, zf_time=zf.coords["time"].values, spectrum=sp.values, freq=sp.coords["frequency"].values,
        phased=ph.values, p0=ph.attrs["phase_p0"], p1=ph.attrs["phase_p1"], pivot=ph.attrs["phase_pivot"],
    )

    # ---- zero_fill known answers (docs/notebooks/pipeline/zero_fill.md:176-204, 260-295) -------------
    rng = np.random.default_rng(7)
    v = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    tt = np.arange(64) * 2e-3
    z = ref.fid.zero_fill(da_fid(v, tt, attrs={"a": 1}), target_points=512)
    k = rng.standard_normal((6, 32)) + 0j
    kv = np.linspace(-16, 15, 32)
    dk = xr.DataArray(k, dims=["ky", "kx"], coords={"ky": np.arange(6.0), "kx": kv}, attrs={"domain": "k-space"})
    zs = ref.fid.zero_fill(dk, dim="kx", target_points=129, position="symmetric")
    out["zero_fill"] = dict(end_in=v, end_t=tt, end_out=z.values, end_coord=z.coords["time"].values,
                            sym_in=k, sym_kx=kv, sym_out=zs.values, sym_coord=zs.coords["kx"].values)

    # ---- apodize_exp / to_spectrum / to_fid / phase on a seeded 3-D block, transform axis in the middle --
    rng = np.random.default_rng(42)
    blk = rng.standard_normal((3, 256, 4)) + 1j * rng.standard_normal((3, 256, 4))
    tb = 1e-3 + np.arange(256) * 4e-4  # non-zero start
    dab = xr.DataArray(blk, dims=["x", "time", "coil"], coords={"time": tb}, attrs={"k": "v"}, name="blk")
    ab = ref.fid.apodize_exp(dab, lb=3.5)
    sb = ref.fid.to_spectrum(ab)
    fb = ref.fid.to_fid(sb)
    pb = ref.phasing.phase(sb, p0=33.0, p1=-725.0)
    pb2 = ref.phasing.phase(sb, p0=-170.0, p1=3990.0, pivot=123.4)
    out["block"] = dict(fid=blk, time=tb, apod=ab.values, spectrum=sb.values, freq=sb.coords["frequency"].values,
                        back=fb.values, back_time=fb.coords["time"].values,
                        phased=pb.values, phased_pivot=pb.attrs["phase_pivot"], phased2=pb2.values)

    # ---- score functions on a grid of (p0, p1) for one noisy 1H-like spectrum -----------------------
    x, t1 = lorentz_fid(2048, 5000.0, (100, 60, 40, 20), (-700, -300, 250, 900), (30, 25, 25, 40),
                        (40, 75, 20, 130), snr=10, seed=3)
    s1 = ref.fid.to_spectrum(ref.fid.apodize_exp(da_fid(x, t1), lb=5.0))
    coord = s1.coords["frequency"].values
    flat = int(np.argmax(np.abs(s1.values)))
    pivot = float(coord[flat])
    step = abs(coord[1] - coord[0])
    iw = max(1, int(round((100 / 2.0) / step)))
    grid = [(p0, p1) for p0 in (-180.0, -77.7, 0.0, 12.5, 180.0) for p1 in (-4000.0, -333.3, 0.0, 41.0, 4000.0)]
    acme = [ref.phasing._acme_score(np.array(g), s1, "frequency", pivot) for g in grid]
    pmin = [ref.phasing._peak_minima_score(np.array(g), s1, "frequency", pivot, flat, iw) for g in grid]
    posi = [ref.phasing._roi_positivity_score(np.array(g), s1, "frequency", pivot, flat, iw) for g in grid]
    acme_p0 = [ref.phasing._acme_score(np.array([g[0]]), s1, "frequency", pivot) for g in grid]
    out["scores"] = dict(spectrum=s1.values, freq=coord, pivot=pivot, target_idx=flat, index_width=iw,
                         grid=np.array(grid), acme=np.array(acme), peak_minima=np.array(pmin),
                         positivity=np.array(posi), acme_p0only=np.array(acme_p0))

    # ---- autophase on 1-D spectra: the optimiser's answers (DE seed=42 + polish) -----------------------
    cases = []
    specs = [
        ("1H_2048_snr10", dict(n=2048, sw=5000.0, amps=(100, 60, 40, 20), freqs=(-700, -300, 250, 900),
                               damps=(30, 25, 25, 40), phases_deg=(40, 75, 20, 130), snr=10, seed=3), 5.0, None),
        ("1H_4096_snr25", dict(n=4096, sw=5000.0, amps=(80, 70, 30, 10), freqs=(-705, -297, 255, 893),
                               damps=(30, 25, 25, 40), phases_deg=(-100, -140, -60, -190), snr=25, seed=4), 5.0, None),
        ("1H_1024_zf4096", dict(n=1024, sw=5000.0, amps=(100, 60, 40, 20), freqs=(-700, -300, 250, 900),
                                damps=(30, 25, 25, 40), phases_deg=(10, -35, 60, -100), snr=6, seed=5), 5.0, 4096),
        ("13C_1024_snr4", dict(n=1024, sw=5000.0, amps=(100, 20), freqs=(-128.4, 256.8), damps=(15, 15),
                               phases_deg=(-45, -83.4), snr=4, seed=6), 10.0, None),
        ("13C_1024_snr12", dict(n=1024, sw=5000.0, amps=(50, 20), freqs=(-128.4, 256.8), damps=(15, 15),
                                phases_deg=(110, 71.5), snr=12, seed=8), 10.0, None),
    ]
    for name, kw, lb, zfn in specs:
        x, tt = lorentz_fid(**kw)
        d = da_fid(x, tt)
        if zfn:
            d = ref.fid.zero_fill(d, target_points=zfn)
        d = ref.fid.apodize_exp(d, lb=lb)
        s = ref.fid.to_spectrum(d)
        variants = [dict(method="acme")]
        if name.startswith("13C"):
            variants += [dict(method="positivity", peak_width=100.0), dict(method="peak_minima", peak_width=100.0),
                         dict(method="positivity", peak_width=100.0, p0_only=True),
                         dict(method="acme", p0_only=True),
                         dict(method="positivity", peak_width=60.0, target_coord=-120.0)]
        if name == "1H_2048_snr10":
            variants += [dict(method="acme", lb=4.0), dict(method="acme", peak_width=100)]
        for var in variants:
            r = ref.phasing.autophase(s, **var)
            cases.append(dict(name=name, variant=repr(sorted(var.items())), fid=x, time=tt, lb=lb, zf=zfn or 0,
                              spectrum=s.values, freq=s.coords["frequency"].values, phased=r.values,
                              p0=float(r.attrs["phase_p0"]), p1=float(r.attrs["phase_p1"]),
                              pivot=float(r.attrs["phase_pivot"])))
            print(name, var, "->", cases[-1]["p0"], cases[-1]["p1"], cases[-1]["pivot"])
    out["autophase_cases"] = cases

    # ---- mode="single" on an N-D batch from the package's own generator (C2-like, small) ---------------
    fid, tt, _ = make_fids_numpy("1H", 12, 2048, seed=11)
    d = da_fid(fid.reshape(3, 4, 2048), tt, extra_dims=("y", "x"), attrs={"MHz": 300.0})
    s = ref.fid.to_spectrum(ref.fid.apodize_exp(d, lb=5.0))
    r = ref.phasing.autophase(s, peak_width=100)
    out["single_batch"] = dict(seed=11, batch_shape=(3, 4), n=2048, lb=5.0, spectrum=s.values,
                               freq=s.coords["frequency"].values, phased=r.values,
                               p0=float(r.attrs["phase_p0"]), p1=float(r.attrs["phase_p1"]),
                               pivot=float(r.attrs["phase_pivot"]))
    print("single_batch ->", out["single_batch"]["p0"], out["single_batch"]["p1"], out["single_batch"]["pivot"])

    # ---- write ----------------------------------------------------------------------------------------
    for key in ("c1", "zero_fill", "block", "scores", "single_batch"):
        np.savez_compressed(os.path.join(HERE, f"{key}.npz"), **out[key])
    ap_flat = {}
    for i, c in enumerate(cases):
        for k2, v2 in c.items():
            ap_flat[f"{i:02d}__{k2}"] = v2
    ap_flat["count"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "autophase_cases.npz"), **ap_flat)
    sizes = {f: os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz")}
    print(sizes, sum(sizes.values()))


if __name__ == "__main__":
    main()
