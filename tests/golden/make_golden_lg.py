"""Golden vectors for ``apodize_lg`` ("next" row N2) from the REFERENCE's own ``processing/fid.py:147-198``.

    python tests/golden/make_golden_lg.py        # needs /root/reference; writes tests/golden/apodize_lg.npz

Cases: the notebook's known-answer geometry (docs/notebooks/pipeline/apodization.md:227-251: lb=3, gb=4), a 3-D block with the
time axis in the middle and a non-zero time origin, gb=0 (pure Lorentzian cancellation), and the window inside the chain
``zero_fill -> apodize_lg -> to_spectrum`` (what ``process_fid(lb=..., gb=...)`` fuses).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402

ref = load_reference()
xr = ref.xr


def main():
    out = {}
    rng = np.random.default_rng(2024)
    # notebook geometry: 1-D FID, 1024 points, sw = 2000 Hz
    t = np.arange(1024) / 2000.0
    fid = (np.exp((-20 + 2j * np.pi * 150.0) * t) + 0.5 * np.exp((-35 + 2j * np.pi * -320.0) * t)
           + 0.02 * (rng.standard_normal(1024) + 1j * rng.standard_normal(1024)))
    da = xr.DataArray(fid, dims=["time"], coords={"time": t}, attrs={"sequence": "FID", "B0": 3.0})
    lg = ref.fid.apodize_lg(da, lb=3.0, gb=4.0)
    out.update(kat_fid=fid, kat_time=t, kat_out=lg.values)
    # 3-D block, axis in the middle, non-zero origin
    blk = rng.standard_normal((3, 256, 4)) + 1j * rng.standard_normal((3, 256, 4))
    tb = 1e-3 + np.arange(256) * 4e-4
    dab = xr.DataArray(blk, dims=["x", "time", "coil"], coords={"time": tb}, attrs={"k": "v"}, name="blk")
    out.update(blk=blk, blk_time=tb, blk_out=ref.fid.apodize_lg(dab, lb=2.5, gb=6.0).values,
               blk_gb0=ref.fid.apodize_lg(dab, lb=2.5, gb=0.0).values)
    # inside the chain: zero_fill(2048) -> apodize_lg -> to_spectrum on a [5, 1024] batch
    fids = rng.standard_normal((5, 1024)) + 1j * rng.standard_normal((5, 1024))
    fids *= np.exp(-25.0 * t)[None, :]
    dac = xr.DataArray(fids, dims=["voxel", "time"], coords={"voxel": np.arange(5), "time": t})
    sp = ref.fid.to_spectrum(ref.fid.apodize_lg(ref.fid.zero_fill(dac, target_points=2048), lb=4.0, gb=7.0))
    out.update(chain_fid=fids, chain_spec=sp.values, chain_freq=sp.coords["frequency"].values)
    np.savez_compressed(os.path.join(HERE, "apodize_lg.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
