"""GPU: the reference's own architectural tests for this path (tests/test_core.py:560-634, 709-714) re-expressed against
the drop-in accessor, plus the Fourier mixin (fft / ifft / fftshift / ifftshift / fftc / ifftc, fourier.py:10-298)."""

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xm():
    import xmris_b200

    return xmris_b200


def _fid(xm, shape=(2048,), dims=("time",)):
    # tests/test_core.py:63-174 fixtures (seeded here): 2048-point complex FID, dwell 0.5 ms, MRS attrs
    rng = np.random.default_rng(12)
    n = shape[-1]
    t = np.arange(n) * 0.5e-3
    data = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return xm.xr.DataArray(data, dims=list(dims), coords={"time": t},
                           attrs={"reference_frequency": 300.0, "carrier_ppm": 4.7, "b0_field": 7.0, "note": "x"})


def _assert_attrs_preserved(original, result):
    for key, value in original.attrs.items():
        assert key in result.attrs, f"Attribute {key!r} was silently dropped during processing."
        assert result.attrs[key] == value


def test_attrs_preservation(xm):
    fid = _fid(xm)
    spec = fid.xmr.to_spectrum()
    for result in (fid.xmr.apodize_exp(lb=5.0), spec, fid.xmr.zero_fill(target_points=4096), spec.xmr.phase(p0=10.0),
                   fid.xmr.fft(), spec.xmr.to_ppm(), spec.xmr.to_ppm().xmr.to_hz(),
                   fid.xmr.apodize_exp(lb=5.0).xmr.to_spectrum().xmr.to_ppm(), spec.xmr.autophase()):
        _assert_attrs_preserved(fid, result)
    assert fid.attrs == {"reference_frequency": 300.0, "carrier_ppm": 4.7, "b0_field": 7.0, "note": "x"}   # never mutated


def test_multidim_to_spectrum_to_ppm(xm):
    # tests/test_core.py:709-714
    fid = _fid(xm, (16, 2048), ("voxel", "time"))
    spec = fid.xmr.to_spectrum()
    assert spec.dims == ("voxel", "frequency")
    ppm = spec.xmr.to_ppm()
    assert ppm.dims == ("voxel", "chemical_shift") and ppm.shape == (16, 2048)


def test_fourier_mixin_matches_numpy(xm):
    rng = np.random.default_rng(3)
    data = rng.standard_normal((6, 128, 64)) + 1j * rng.standard_normal((6, 128, 64))
    t = 1e-3 * np.arange(128)
    ky = np.linspace(-32, 31, 64)
    da = xm.xr.DataArray(data, dims=["v", "time", "ky"], coords={"time": t, "ky": ky}, attrs={"a": 1}, name="d")
    f = da.xmr.fft()                                                   # fourier.py:117-173
    assert f.dims == ("v", "time", "ky") and f.attrs == {"a": 1} and f.name == "d"
    assert rel_l2(f.values, np.fft.fftn(data, axes=(1,), norm="ortho")) < 1e-5
    np.testing.assert_array_equal(f.coords["time"].values, np.fft.fftfreq(128, d=1e-3))
    assert f.coords["time"].attrs == {"long_name": "Frequency", "units": "Hz"}
    f2 = da.xmr.fft(dim=["time", "ky"], out_dim=["frequency", "y"])    # N-D, renamed
    assert f2.dims == ("v", "frequency", "y")
    assert rel_l2(f2.values, np.fft.fftn(data, axes=(1, 2), norm="ortho")) < 1e-5
    np.testing.assert_allclose(f2.coords["y"].values, np.fft.fftfreq(64, d=ky[1] - ky[0]))
    assert f2.coords["y"].attrs == {}
    back = f2.xmr.ifft(dim=["frequency", "y"], out_dim=["time", "ky"])
    assert rel_l2(back.values, data) < 1e-5
    assert back.coords["time"].attrs == {"long_name": "Time", "units": "s"}
    with pytest.raises(ValueError, match="same length"):
        da.xmr.fft(dim=["time", "ky"], out_dim=["frequency"])
    # shifts roll data AND coordinates (fourier.py:31-32, 57-58); odd length distinguishes the two
    odd = xm.xr.DataArray(data[:, :127, 0], dims=["v", "time"], coords={"time": t[:127]})
    sh = odd.xmr.fftshift("time")
    np.testing.assert_array_equal(sh.values, np.roll(odd.values.astype(np.complex64), 63, axis=1))
    np.testing.assert_array_equal(sh.coords["time"].values, np.roll(t[:127], 63))
    ish = sh.xmr.ifftshift("time")
    np.testing.assert_array_equal(ish.values, odd.values.astype(np.complex64))
    np.testing.assert_array_equal(ish.coords["time"].values, t[:127])
    # centred transforms (fourier.py:238-298)
    c = da.xmr.fftc(dim="ky", out_dim="y")
    ref = np.fft.fftshift(np.fft.fft(np.fft.ifftshift(data, axes=2), axis=2, norm="ortho"), axes=2)
    assert rel_l2(c.values, ref) < 1e-5
    ic = c.xmr.ifftc(dim="y", out_dim="ky")
    assert rel_l2(ic.values, data) < 1e-5


def test_real_valued_and_complex128_inputs(xm):
    t = np.arange(256) * 1e-3
    real = xm.xr.DataArray(np.cos(2 * np.pi * 50 * t), dims=["time"], coords={"time": t})
    s = real.xmr.to_spectrum()
    ref = np.fft.fftshift(np.fft.fft(np.cos(2 * np.pi * 50 * t), norm="ortho"))
    assert s.values.dtype == np.complex64 and rel_l2(s.values, ref) < 1e-5
