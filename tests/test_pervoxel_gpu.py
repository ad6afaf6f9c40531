"""GPU parity of the per-voxel chain (autophase mode="all", kernel K2) against the oracle's per-spectrum loop
(``oracle.autophase_each`` = the reference's 1-D autophase, SciPy DE seed 42 + polish, on every spectrum).

What can and cannot match (DESIGN.md "Per-voxel parity"):
  * spectra: <= 1e-5 relative L2 against the float64 chain rotated by the angles the GPU reports;
  * pivot: identical;
  * optimiser: the GPU search is deterministic and global, the reference's is stochastic + L-BFGS-B on a non-smooth
    objective, which often stops a few tenths of a degree short in the flat p1 valley or in a worse local minimum.
    The test therefore requires, per well-posed voxel (f_ref > 0), EITHER |dp0|,|dp1| <= 0.1 deg OR an objective value
    (evaluated in float64 by the oracle at the GPU's angles) not worse than the reference's, and bounds the fraction of
    voxels where the GPU is worse.
"""

import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import xmris_oracle as orc

pytestmark = pytest.mark.gpu
TOL, ANG = 1e-5, 0.1


def _ref_one(args):
    spec, freqs, kw = args
    _, info = orc.autophase(spec, 0, freqs, **kw)
    return info["p0"], info["p1"], info["pivot"], info["fun"]


def _reference(spec, freqs, kw):
    workers = min(8, len(os.sched_getaffinity(0)))
    with mp.get_context("fork").Pool(workers) as pool:
        return np.array(pool.map(_ref_one, [(spec[i], freqs, kw) for i in range(spec.shape[0])]))


CONFIGS = [
    # name, family, voxels, n_in, target_points, lb  (C2-, C3-, C4-, C5-shaped, small batches)
    ("C2_2048", "1H", 24, 2048, None, 5.0),
    ("C3_4096_zf8192", "1H", 16, 4096, 8192, 5.0),
    ("C4_13C_1024", "13C", 32, 1024, None, 10.0),
    ("C5_4096", "1H", 32, 4096, None, 5.0),
]


@pytest.mark.parametrize("name,family,nvox,n_in,zf,lb", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_chain_all_matches_per_spectrum_reference(name, family, nvox, n_in, zf, lb):
    import torch
    from xmris_b200 import pervoxel
    from xmris_b200.synth import make_fids_numpy

    fid, t, _ = make_fids_numpy(family, nvox, n_in, seed=100 + n_in)
    fid = fid.astype(np.complex64)
    spec_t, freqs, info = pervoxel.chain_all(torch.from_numpy(fid).cuda(), t, zf, "end", lb, peak_width=100)
    got = spec_t.cpu().numpy()
    ref_spec, ref_freqs = orc.chain_to_spectrum(fid.astype(np.complex128), 1, t, zf, "end", lb)
    np.testing.assert_array_equal(freqs, ref_freqs)
    ref = _reference(ref_spec, ref_freqs, dict(peak_width=100))
    n_match = n_better = n_worse = n_ill = 0
    for i in range(nvox):
        p0, p1, piv = info["p0"][i], info["p1"][i], info["pivot"][i]
        assert piv == ref[i, 2], (i, piv, ref[i, 2])
        same, _ = orc.phase(ref_spec[i], 0, ref_freqs, p0, p1, piv)
        assert rel_l2(got[i], same) < TOL, (i, rel_l2(got[i], same))
        if ref[i, 3] < 0:
            n_ill += 1          # reference dived into the ACME pole (SURVEY finding 5): not comparable
            continue
        f_gpu = orc.acme_score([p0, p1], ref_spec[i], ref_freqs, piv)
        if abs(p0 - ref[i, 0]) <= ANG and abs(p1 - ref[i, 1]) <= ANG:
            n_match += 1
        elif f_gpu <= ref[i, 3] * (1 + 1e-5):
            n_better += 1
        else:
            n_worse += 1
            print(f"{name} voxel {i}: gpu ({p0:.3f}, {p1:.3f}) f={f_gpu:.6g}  ref ({ref[i,0]:.3f}, {ref[i,1]:.3f}) f={ref[i,3]:.6g}")
    print(f"{name}: match {n_match}  equal-or-better objective {n_better}  worse {n_worse}  ill-posed {n_ill}  of {nvox}")
    well = nvox - n_ill
    # small-batch smoke version of tests/test_angle_parity_gpu.py (1024 spectra per shape with the tight-DE adjudicator,
    # floors 94-99 %): here "worse" is against the UN-adjudicated reference answer only
    assert n_worse <= max(2, int(0.10 * well)), (n_match, n_better, n_worse)
    assert n_match >= 0.6 * well


def test_autophase_mode_all_accessor():
    import xmris_b200
    from xmris_b200.synth import make_fids_numpy

    fid, t, _ = make_fids_numpy("13C", 12, 1024, seed=5)
    da = xmris_b200.xr.DataArray(fid.reshape(3, 4, 1024), dims=["rep", "x", "time"], coords={"time": t}, attrs={"k": 1})
    sp = da.xmr.apodize_exp(lb=10.0).xmr.to_spectrum()
    out = sp.xmr.autophase(mode="all")
    assert out.dims == sp.dims and out.attrs["phase_pivot_coord"] == "frequency" and out.attrs["k"] == 1
    for key in ("phase_p0", "phase_p1", "phase_pivot"):
        assert out.coords[key].dims == ("rep", "x")
    np.testing.assert_allclose(np.abs(out.values), np.abs(sp.values), rtol=2e-5, atol=1e-5)
    # each voxel's own pivot and phase
    freqs = sp.coords["frequency"].values
    p0, p1, piv = (out.coords[k].values for k in ("phase_p0", "phase_p1", "phase_pivot"))
    for r in range(3):
        for x in range(4):
            s = sp.values[r, x].astype(np.complex128)
            assert piv[r, x] == freqs[int(np.argmax(np.abs(s)))]
            same, _ = orc.phase(s, 0, freqs, p0[r, x], p1[r, x], piv[r, x])
            assert rel_l2(out.values[r, x], same) < TOL
    # fused entry point agrees with the accessor chain
    fused = da.xmr.process_fid(lb=10.0, autophase_kwargs=dict(mode="all"))
    # (the fused kernel and the stand-alone K1 round their twiddles differently: in a flat p1 valley the argmin may move)
    np.testing.assert_allclose(fused.coords["phase_p0"].values, p0, atol=0.5)
    np.testing.assert_allclose(fused.coords["phase_p1"].values, p1, atol=1.5)
    # p0_only and a local method
    o2 = sp.xmr.autophase(mode="all", method="positivity", p0_only=True)
    assert np.all(o2.coords["phase_p1"].values == 0.0)
    for r in range(3):
        s = sp.values[r, 0].astype(np.complex128)
        _, i = orc.autophase(s, 0, freqs, method="positivity", peak_width=100, p0_only=True)
        assert abs(o2.coords["phase_p0"].values[r, 0] - i["p0"]) < ANG


def test_mode_all_at_8192_points_spectrum_and_fid_input():
    """N = 8192: one shared buffer is landing slot, FFT exchanges and searched spectrum (two CTAs per SM).  Spectrum input
    (padded layout written over the buffer the spectrum landed in) and FID input (4096 -> 8192 fused chain): every voxel's own
    pivot, the output equal to the reference's phase() at the GPU's angles, both entry points in the same valley, and the
    objective not worse than the reference optimiser's on a sample of voxels."""
    import xmris_b200
    from xmris_b200.synth import make_fids_numpy

    nvox, n_in, n_out = 301, 4096, 8192          # more voxels than resident CTAs: the end-of-voxel prefetch is exercised
    fid, t, _ = make_fids_numpy("1H", nvox, n_in, seed=81)
    da = xmris_b200.xr.DataArray(fid, dims=["voxel", "time"], coords={"time": t})
    sp = da.xmr.zero_fill(target_points=n_out).xmr.apodize_exp(lb=5.0).xmr.to_spectrum()
    out = sp.xmr.autophase(mode="all")
    freqs = sp.coords["frequency"].values
    p0, p1, piv = (out.coords[k].values for k in ("phase_p0", "phase_p1", "phase_pivot"))
    spv = sp.values
    for v in range(0, nvox, 7):
        s = spv[v].astype(np.complex128)
        assert piv[v] == freqs[int(np.argmax(np.abs(s)))]
        same, _ = orc.phase(s, 0, freqs, p0[v], p1[v], piv[v])
        assert rel_l2(out.values[v], same) < TOL
    fused = da.xmr.process_fid(target_points=n_out, lb=5.0, autophase_kwargs=dict(mode="all"))
    np.testing.assert_array_equal(fused.coords["phase_pivot"].values, piv)
    d0 = np.abs(fused.coords["phase_p0"].values - p0)
    d0 = np.minimum(d0, 360.0 - d0)
    assert np.mean((d0 < 0.5) & (np.abs(fused.coords["phase_p1"].values - p1) < 1.5)) > 0.97
    n_worse = 0
    for v in range(0, nvox, 60):
        s = spv[v].astype(np.complex128)
        _, info = orc.autophase(s, 0, freqs, peak_width=100)
        f_gpu = orc.acme_score([p0[v], p1[v]], s, freqs, piv[v])
        f_ref = orc.acme_score([info["p0"], info["p1"]], s, freqs, info["pivot"])
        n_worse += int(f_gpu > f_ref * (1.0 + 1e-4) + 1e-12)
    assert n_worse <= 1


@pytest.mark.parametrize("variant", ["p0_only", "target_coord"])
def test_mode_all_acme_variants_against_per_spectrum_reference(variant):
    """K2-ACME with ``p0_only=True`` (one-parameter search, p1 = 0) and with ``target_coord`` (fixed, off-grid pivot for every
    voxel: phasing.py:233-235): each voxel against the reference's autophase on that 1-D spectrum -- angles within 0.1 deg or an
    objective not worse; spectra equal to the reference's phase() at the GPU's angles."""
    import xmris_b200
    from xmris_b200.synth import make_fids_numpy

    nvox = 20
    fid, t, _ = make_fids_numpy("13C", nvox, 1024, seed=77)
    da = xmris_b200.xr.DataArray(fid.astype(np.complex64), dims=["vox", "time"], coords={"time": t})
    sp = da.xmr.apodize_exp(lb=10.0).xmr.to_spectrum()
    freqs = sp.coords["frequency"].values
    kw = dict(p0_only=True) if variant == "p0_only" else dict(target_coord=-120.3)
    out = sp.xmr.autophase(mode="all", **kw)
    p0, p1, piv = (out.coords[k].values for k in ("phase_p0", "phase_p1", "phase_pivot"))
    ref_spec = np.asarray(sp.values, dtype=np.complex128)
    ref = _reference(ref_spec, freqs, dict(peak_width=100, **kw))
    worse = 0
    for i in range(nvox):
        assert piv[i] == ref[i, 2], (i, piv[i], ref[i, 2])
        if variant == "p0_only":
            assert p1[i] == 0.0
        same, _ = orc.phase(ref_spec[i], 0, freqs, p0[i], p1[i], piv[i])
        assert rel_l2(out.values[i], same) < 2e-5
        ph = [p0[i]] if variant == "p0_only" else [p0[i], p1[i]]
        f_gpu = orc.acme_score(ph, ref_spec[i], freqs, piv[i])
        d0 = abs(((p0[i] - ref[i, 0] + 180.0) % 360.0) - 180.0)
        close = d0 <= ANG and abs(p1[i] - ref[i, 1]) <= ANG
        if not (close or f_gpu <= ref[i, 3] * (1 + 1e-5)):
            worse += 1
            print(f"{variant} voxel {i}: gpu ({p0[i]:.3f}, {p1[i]:.3f}) f={f_gpu:.6g}  ref ({ref[i,0]:.3f}, {ref[i,1]:.3f}) f={ref[i,3]:.6g}")
    assert worse <= 1, worse
