"""GPU parity of two rows the round-1 verdict found untested:

* N2 ``apodize_lg`` (reference ``processing/fid.py:147-198``; known-answer test ``docs/notebooks/pipeline/apodization.md:227-251``)
  against vectors produced by the reference's own function (``tests/golden/make_golden_lg.py``), stand-alone and fused into
  ``process_fid(lb=, gb=)``;
* A5/A6 the DEVICE objective evaluator (``csrc/autophase_eval.cuh`` through ``xmr_autophase_score_c64``) against the
  reference's ``_acme_score`` / ``_peak_minima_score`` / ``_roi_positivity_score`` outputs in ``tests/golden/scores.npz``
  (``phasing.py:100-157``).  float64 accumulation: 1e-9 relative; float32: 1e-4 relative.
"""

import numpy as np
import pytest

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xm():
    import torch

    assert torch.cuda.is_available()
    import xmris_b200
    from xmris_b200 import _lib

    _lib.load()
    return xmris_b200


def test_apodize_lg_known_answer_and_lineage(xm):
    g = load_golden("apodize_lg")
    t, fid = g["kat_time"], g["kat_fid"]
    da = xm.xr.DataArray(fid, dims=["time"], coords={"time": t}, attrs={"sequence": "FID", "B0": 3.0})
    lg = da.xmr.apodize_lg(lb=3.0, gb=4.0)
    # the notebook's formula (apodization.md:231-239) and the reference's own output
    t_g = (2 * np.sqrt(np.log(2))) / (np.pi * 4.0)
    expected = fid * (np.exp(np.pi * 3.0 * t) * np.exp(-(t**2) / (t_g**2)))
    assert rel_l2(lg.values, expected) < 2e-7
    assert rel_l2(lg.values, g["kat_out"]) < 2e-7
    assert lg.dims == da.dims
    np.testing.assert_array_equal(lg.coords["time"].values, t)
    for k, v in da.attrs.items():
        assert lg.attrs[k] == v
    assert lg.attrs["apodization_lb"] == 3.0 and lg.attrs["apodization_gb"] == 4.0


def test_apodize_lg_block_any_axis(xm):
    g = load_golden("apodize_lg")
    da = xm.xr.DataArray(g["blk"], dims=["x", "time", "coil"], coords={"time": g["blk_time"]}, attrs={"k": "v"}, name="blk")
    out = da.xmr.apodize_lg(lb=2.5, gb=6.0)
    assert out.dims == da.dims and out.attrs == {"k": "v", "apodization_lb": 2.5, "apodization_gb": 6.0}
    assert rel_l2(out.values, g["blk_out"]) < 2e-7
    out0 = da.xmr.apodize_lg(lb=2.5, gb=0.0)        # gb == 0: pure Lorentzian cancellation (fid.py:187-190)
    assert rel_l2(out0.values, g["blk_gb0"]) < 2e-7
    assert out0.attrs["apodization_gb"] == 0.0


def test_apodize_lg_fused_chain(xm):
    g = load_golden("apodize_lg")
    t = g["kat_time"]
    da = xm.xr.DataArray(g["chain_fid"], dims=["voxel", "time"], coords={"voxel": np.arange(5), "time": t})
    chained = da.xmr.zero_fill(target_points=2048).xmr.apodize_lg(lb=4.0, gb=7.0).xmr.to_spectrum()
    fused = da.xmr.process_fid(target_points=2048, lb=4.0, gb=7.0)
    for res in (chained, fused):
        assert res.dims == ("voxel", "frequency")
        np.testing.assert_array_equal(res.coords["frequency"].values, g["chain_freq"])
        assert max(rel_l2(res.values[i], g["chain_spec"][i]) for i in range(5)) < 1e-5
        assert res.attrs["apodization_lb"] == 4.0 and res.attrs["apodization_gb"] == 7.0
        assert res.attrs["zero_fill_target"] == 2048
    # the fused window (table path of K1) on the device-resident entry point too
    import torch

    from xmris_b200 import chain

    spec, freqs, _ = chain.chain_to_spectrum(torch.from_numpy(g["chain_fid"].astype(np.complex64)).cuda(), t, 2048, "end", 4.0, gb=7.0)
    assert max(rel_l2(spec.cpu().numpy()[i], g["chain_spec"][i]) for i in range(5)) < 1e-5


@pytest.mark.parametrize("f64,tol", [(True, 1e-9), (False, 1e-4)])
def test_device_objective_matches_reference_scores(xm, f64, tol):
    import torch

    from xmris_b200 import device as D
    from xmris_b200.processing import _affine_ramp

    g = load_golden("scores")
    # the evaluator works on complex64 spectra: compare against the reference's scores of the SAME rounded spectrum where
    # float64 accuracy is claimed (oracle, bit-exact restatement of the reference: tests/test_oracle.py), and against the
    # golden values themselves (computed from the complex128 spectrum) at the float32 tolerance
    from oracle import xmris_oracle as orc

    fr, pivot = g["freq"], float(g["pivot"])
    ti, iw = int(g["target_idx"]), int(g["index_width"])
    spec64 = g["spectrum"].astype(np.complex64)
    spec_t = torch.from_numpy(spec64).cuda()
    u0, du = _affine_ramp(fr, pivot)
    p0, p1 = g["grid"][:, 0], g["grid"][:, 1]
    sp128 = spec64.astype(np.complex128)
    want = {
        "acme": np.array([orc.acme_score([a, b], sp128, fr, pivot) for a, b in g["grid"]]),
        "peak_minima": np.array([orc.peak_minima_score([a, b], sp128, fr, pivot, ti, iw) for a, b in g["grid"]]),
        "positivity": np.array([orc.roi_positivity_score([a, b], sp128, fr, pivot, ti, iw) for a, b in g["grid"]]),
    }
    for method in ("acme", "peak_minima", "positivity"):
        got = D.autophase_score(spec_t, u0, du, p0, p1, method, ti, iw, float64=f64).cpu().numpy()
        scale = np.maximum(np.abs(want[method]), np.abs(want[method]).max() * 1e-6)
        err = np.abs(got - want[method]) / scale
        assert err.max() < tol, (method, f64, float(err.max()), int(err.argmax()))
        # and the reference's own golden values (complex128 input): the complex64 rounding of the input stays below 1e-5
        err_g = np.abs(got - g[method]) / np.maximum(np.abs(g[method]), np.abs(g[method]).max() * 1e-6)
        assert err_g.max() < max(tol, 2e-5), (method, float(err_g.max()))
    # p0-only evaluation (phasing.py:101-103: p1 = 0 when one parameter is given)
    got = D.autophase_score(spec_t, u0, du, p0, np.zeros_like(p0), "acme", ti, iw, float64=f64).cpu().numpy()
    err = np.abs(got - g["acme_p0only"]) / np.abs(g["acme_p0only"])
    assert err.max() < max(tol, 2e-5)
