"""CPU: host-side logic of the drop-in layer -- signatures/defaults, error text, metadata helpers, the xarray stand-in,
window factoring, phase-ramp arithmetic, shard bookkeeping and the 2-rank exchange over gloo."""

import inspect
import os
import socket

import numpy as np
import pytest

import xmris_b200
from xmris_b200 import chain, device, processing, sharding
from xmris_b200.accessor import XmrisB200Accessor
from xmris_b200.vocab import ATTRS, COORDS, DIMS
from oracle import xmris_oracle as orc

xr = xmris_b200.xr


def _defaults(fn):
    return {k: v.default for k, v in inspect.signature(fn).parameters.items() if v.default is not inspect.Parameter.empty}


def test_accessor_defaults_match_reference():
    # reference tests/test_core.py:509-552 + accessor.py:452, 490, 526-531, 599-605, 630-638
    A = XmrisB200Accessor
    assert _defaults(A.apodize_exp) == {"dim": "time", "lb": 1.0}
    assert _defaults(A.to_spectrum) == {"dim": "time", "out_dim": "frequency"}
    assert _defaults(A.zero_fill) == {"dim": "time", "target_points": 1024, "position": "end"}
    assert _defaults(A.phase) == {"dim": "frequency", "p0": 0.0, "p1": 0.0, "pivot": None}
    assert _defaults(A.autophase) == {"dim": "frequency", "method": "acme", "peak_width": 100, "lb": 0.0,
                                      "temp_time_dim": "time"}
    assert _defaults(processing.autophase) == {"dim": "frequency", "method": "acme", "mode": "single", "peak_width": 0.5,
                                               "target_coord": None, "p0_only": False, "lb": 0.0, "temp_time_dim": "time"}
    assert DIMS.time == "time" and DIMS.frequency == "frequency" and COORDS.frequency.unit == "Hz"
    assert COORDS.chemical_shift.long_name == "Chemical Shift" and ATTRS.phase_pivot_coord == "phase_pivot_coord"


def test_reference_dim_defaults_table():
    # every (method, parameter, default) row of the reference's tests/test_core.py:509-530 that is on or beside the path
    rows = [("fft", "dim", "time"), ("ifft", "dim", "frequency"), ("fftc", "dim", "time"), ("ifftc", "dim", "frequency"),
            ("apodize_exp", "dim", "time"), ("apodize_lg", "dim", "time"), ("to_spectrum", "dim", "time"),
            ("to_spectrum", "out_dim", "frequency"), ("to_fid", "dim", "frequency"), ("to_fid", "out_dim", "time"),
            ("zero_fill", "dim", "time"), ("autophase", "dim", "frequency"), ("remove_digital_filter", "dim", "time"),
            ("to_ppm", "dim", "frequency"), ("to_hz", "dim", "chemical_shift"), ("to_real_imag", "dim", "component"),
            ("to_complex", "dim", "component"), ("baseline_als", "dim", "frequency")]
    for method, param, want in rows:
        assert inspect.signature(getattr(XmrisB200Accessor, method)).parameters[param].default == want, (method, param)
    assert _defaults(XmrisB200Accessor.remove_digital_filter) == {"dim": "time", "keep_length": True}
    assert _defaults(XmrisB200Accessor.baseline_als) == {"dim": "frequency", "lam": 1e5, "p": 0.001, "n_iter": 10}   # accessor.py:552-558


def test_real_imag_round_trip_known_answer():
    # docs/notebooks/basics/complex_numbers.md:79-151 (the notebook's strict CI cell), host data re-labelling only
    t = np.linspace(0, 1, 512)
    fid = np.exp(-t * 3.0) * np.exp(1j * 2 * np.pi * 15.0 * t)
    da = xr.DataArray(fid, dims=["time"], coords={"time": t}, attrs={"B0": 3.0}, name="Signal")
    split = da.xmr.to_real_imag()
    assert split.ndim == da.ndim + 1 and split.sizes["component"] == 2 and not np.iscomplexobj(split.values)
    assert list(split.coords["component"].values) == ["real", "imag"] and split.dims == ("time", "component")
    recon = split.xmr.to_complex()
    assert recon.ndim == da.ndim and np.iscomplexobj(recon.values) and "component" not in recon.dims
    np.testing.assert_array_equal(recon.values, da.values)
    assert recon.attrs["B0"] == 3.0 and recon.name == "Signal" and split.attrs["B0"] == 3.0
    np.testing.assert_array_equal(recon.coords["time"].values, t)
    custom = da.xmr.to_real_imag(dim="channel", coords=("ch0", "ch1"))
    assert custom.dims == ("time", "channel")
    np.testing.assert_array_equal(custom.xmr.to_complex(dim="channel", coords=("ch0", "ch1")).values, fid)
    # component axis in the middle (the layout of the reference's Bruker netCDF fixtures is (raw, component))
    mid = xr.DataArray(np.arange(12.0).reshape(2, 2, 3), dims=["v", "component", "time"],
                       coords={"component": ["real", "imag"], "time": [0.0, 0.1, 0.2]})
    c = mid.xmr.to_complex()
    assert c.dims == ("v", "time")
    np.testing.assert_array_equal(c.values, mid.values[:, 0, :] + 1j * mid.values[:, 1, :])
    with pytest.raises(ValueError, match="missing dimension"):
        da.xmr.to_complex()


def test_remove_digital_filter_host_branches():
    # branches that never reach the device: missing dim, non-positive delay (copy), whole-sample delay (slice + zeros)
    t = np.arange(32) / 1000.0
    v = np.arange(64, dtype=float).reshape(2, 32) + 1j
    da = xr.DataArray(v, dims=["rep", "time"], coords={"time": t}, attrs={"a": 1})
    with pytest.raises(ValueError, match="Dimension 'nope' missing in DataArray."):
        da.xmr.remove_digital_filter(3.0, dim="nope")
    same = da.xmr.remove_digital_filter(0.0)
    np.testing.assert_array_equal(same.values, v)
    assert same.attrs == {"a": 1}
    from oracle import xmris_oracle as orc

    for keep in (True, False):
        r = da.xmr.remove_digital_filter(5.0, keep_length=keep)
        want, coord = orc.remove_digital_filter(v, 1, t, 5.0, keep)
        np.testing.assert_array_equal(r.values, want)
        np.testing.assert_array_equal(r.coords["time"].values, coord)
        assert r.attrs == {"a": 1, "digital_filter_removed": True, "group_delay_removed": 5.0,
                           "length_retained_with_zeros": keep}
    np.testing.assert_array_equal(da.values, v)


def test_check_dims_error_text():
    # reference tests/test_core.py:411-440
    da = xr.DataArray(np.zeros((2, 4)), dims=["voxel", "t"])
    with pytest.raises(ValueError) as e:
        processing._check_dims(da, "time", "apodize_exp")
    msg = str(e.value)
    assert "missing dimension" in msg and "['voxel', 't']" in msg and "apodize_exp" in msg
    assert "obj.rename({'time': 'correct_name'})" in msg
    processing._check_dims(da, ["voxel", "t"], "fft")
    for fn in (xmris_b200.zero_fill, xmris_b200.apodize_exp, xmris_b200.to_spectrum, xmris_b200.phase, xmris_b200.autophase):
        with pytest.raises(ValueError, match="missing dimension"):
            fn(da, dim="nope")


def test_accessor_registered_and_no_cpu_fallback():
    import torch

    da = xr.DataArray(np.zeros((2, 16), complex), dims=["v", "time"], coords={"time": np.arange(16.0)})
    assert isinstance(da.xmr, XmrisB200Accessor) and da.xmr is da.xmr
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            da.xmr.to_spectrum()
        with pytest.raises(TypeError, match="CUDA"):
            device.fid_to_spectrum(torch.zeros((2, 16), dtype=torch.complex64))
    same = da.xmr.zero_fill(target_points=8)       # no-op path never touches the device (fid.py:234-236)
    assert same.attrs == {} and same.shape == (2, 16)
    with pytest.raises(ValueError, match="Mode"):
        da.xmr.autophase(dim="time", mode="nope")
    with pytest.raises(ValueError, match="Method"):
        da.xmr.autophase(dim="time", method="nope")


def test_xarray_lite_semantics():
    t = np.arange(4.0)
    da = xr.DataArray(np.arange(8.0).reshape(2, 4), dims=["v", "time"], coords={"time": ("time", t, {"units": "s"})},
                      attrs={"a": 1}, name="x")
    w = np.exp(-da.coords["time"])
    prod = (da * w).transpose(*da.dims)
    assert prod.attrs == {} and prod.name is None and prod.dims == ("v", "time")     # attrs dropped, names differ
    np.testing.assert_allclose(prod.values, da.values * np.exp(-t))
    p = da.pad({"time": (1, 2)}, mode="constant", constant_values=0)
    assert p.shape == (2, 7) and np.isnan(p.coords["time"].values[[0, 5, 6]]).all() and p.attrs == {"a": 1}
    r = da.roll({"time": 2}, roll_coords=True)
    np.testing.assert_array_equal(r.coords["time"].values, np.roll(t, 2))
    s = da.isel({"v": 1})
    assert s.dims == ("time",) and s.shape == (4,)
    rn = da.rename({"time": "frequency"})
    assert rn.dims == ("v", "frequency") and "frequency" in rn.coords and "time" not in rn.coords
    c = da.copy(data=np.zeros((2, 4)))
    assert c.attrs == {"a": 1} and c.name == "x" and c.coords["time"].attrs == {"units": "s"}
    assert float(da.coords["time"].max()) == 3.0 and da.get_axis_num("time") == 1
    with pytest.raises(KeyError):
        xr.DataArray(np.zeros(3), dims=["q"]).coords["q"]


def test_split_window_and_geometry():
    t = 1e-3 + np.arange(4096) / 5000.0
    w = np.exp(-np.pi * 5.0 * t) / 64.0
    mode, cols, rows = device.split_window(w, 4096)
    assert mode == device._lib.WIN_SEPARABLE and cols.shape == (256,) and rows.shape == (16,)
    recon = (rows[:, None].astype(np.float64) * cols[None, :]).ravel()
    np.testing.assert_allclose(recon, w, rtol=3e-7)
    mode, table, rows = device.split_window(np.linspace(1, 2, 4096), 4096)
    assert mode == device._lib.WIN_TABLE and rows is None and table.shape == (4096,)
    mode, table, rows = device.split_window(np.linspace(1, 2, 128), 128)
    assert mode == device._lib.WIN_SEPARABLE and table.shape == (128,)
    with pytest.raises(ValueError, match="not supported"):
        device.check_length(5000)
    device.check_length(1972)          # arbitrary lengths <= 4096 run through the chirp-z path
    geo = chain.chain_geometry(1024, np.linspace(0, 1, 1024), 2048, "end", 5.0)
    _, t_ref, _ = orc.zero_fill(np.zeros(1024), 0, np.linspace(0, 1, 1024), 2048, "end")
    np.testing.assert_array_equal(geo["t_pad"], t_ref)
    ref_spec, ref_freq = orc.to_spectrum(np.zeros(2048, complex), 0, t_ref)
    np.testing.assert_array_equal(geo["freqs"], ref_freq)
    geo = chain.chain_geometry(32, np.arange(32.0) - 16, 129 - 1, "symmetric", None)
    assert geo["pad_left"] == 48 and geo["window"] is None
    with pytest.raises(ValueError, match="position"):
        chain.chain_geometry(32, np.arange(32.0), 64, "middle", None)


def test_phase_turns_equal_reference_ramp():
    freqs = np.roll(np.fft.fftfreq(2048, d=1 / 5000.0), 1024)
    for p0, p1, pivot in [(33.0, -725.0, freqs[700]), (-180.0, 4000.0, 123.4), (12.0, 0.0, freqs[0])]:
        a, b, u0, du = chain.phase_turns(freqs, p0, p1, pivot)
        ref = orc.phase_array(freqs, p0, p1, pivot) / (2 * np.pi)
        np.testing.assert_allclose(a + b * np.arange(2048), ref, rtol=0, atol=1e-12)
        np.testing.assert_allclose(u0 + du * np.arange(2048), (freqs - pivot) / (freqs.max() - freqs.min()), atol=1e-14)
    desc = freqs[::-1].copy()      # descending axis (ppm-like): the ramp keeps the reference's sign convention
    u0, du = processing._affine_ramp(desc, desc[10])
    np.testing.assert_allclose(u0 + du * np.arange(2048), (desc - desc[10]) / (desc.max() - desc.min()), atol=1e-14)
    with pytest.raises(ValueError, match="uniform"):
        processing._affine_ramp(np.cumsum(np.linspace(1, 2, 64)), 3.0)


def test_shard_bounds_and_winner():
    for n, w in [(10, 3), (1 << 20, 8), (5, 8), (0, 2)]:
        spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert sharding.pick_winner([1.0, 7.0, 7.0], [5, 6, 7], [0, 100, 200]) == (1, 106)   # tie -> lowest rank
    assert sharding.pick_winner([-np.inf, 2.0], [0, 3], [0, 40]) == (1, 43)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    n_out, rows = 64, 10
    exch = sharding.make_exchange(dist, dev, n_out, rank * rows * n_out)
    called = []

    def search():
        called.append(rank)
        return 1.5 + rank, 2.5, 3.5, 4.5

    local_max, local_flat = [(3.0, 17), (9.0, 130)][rank]
    res = exch(local_max, local_flat, search)
    q.put((rank, res, called))
    dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    # rank 1 holds the global maximum: only it searches, both ranks receive its answer
    assert got[0][1] == got[1][1] == (2.5, 2.5, 3.5, 4.5)
    assert got[0][2] == [] and got[1][2] == [1]


def test_to_ppm_to_hz_coordinates_only():
    # reference tests/test_core.py:642-714 and docs/notebooks/basics/hz_and_ppm.md:169-203
    hz = np.linspace(-500, 500, 11)
    da = xr.DataArray(np.arange(22.0).reshape(2, 11), dims=["voxel", "frequency"], coords={"frequency": hz},
                      attrs={"reference_frequency": 123.2, "carrier_ppm": 4.7, "keep": 1})
    ppm = da.xmr.to_ppm()
    assert ppm.dims == ("voxel", "chemical_shift") and ppm.attrs == da.attrs
    np.testing.assert_allclose(ppm.coords["chemical_shift"].values, 4.7 + hz / 123.2)
    assert ppm.coords["chemical_shift"].attrs == {"long_name": "Chemical Shift"}
    np.testing.assert_array_equal(ppm.coords["frequency"].values, hz)       # old axis kept as a non-index coordinate
    np.testing.assert_array_equal(ppm.values, da.values)
    back = ppm.xmr.to_hz()
    assert back.dims == ("voxel", "frequency")
    np.testing.assert_allclose(back.coords["frequency"].values, hz, atol=1e-9)
    assert back.coords["frequency"].attrs == {"long_name": "Frequency", "units": "Hz"}
    with pytest.raises(ValueError, match="requires the following missing attributes"):
        xr.DataArray(np.zeros(3), dims=["frequency"], coords={"frequency": np.arange(3.0)}).xmr.to_ppm()
    with pytest.raises(ValueError, match="missing dimension"):
        da.xmr.to_ppm(dim="nope")


def test_chain_geometry_is_memoised_and_read_only():
    from xmris_b200 import chain

    t = np.arange(1024) / 5000.0
    a = chain.chain_geometry(1024, t, 2048, "end", 5.0)
    b = chain.chain_geometry(1024, t.copy(), 2048, "end", 5.0)
    assert a is b and a["n_out"] == 2048 and a["pad_left"] == 0
    for key in ("t_pad", "window", "freqs"):
        assert not a[key].flags.writeable                      # shared by every caller: must not be modified in place
    c = chain.chain_geometry(1024, t, 2048, "symmetric", 5.0)
    d = chain.chain_geometry(1024, t + 1e-6, 2048, "end", 5.0)
    assert c is not a and c["pad_left"] == 512 and d is not a
    # float64 tables as the reference builds them (fid.py:136 with the ortho norm, fourier.py:95, 31-32)
    t_pad = t[0] + np.arange(2048) * (t[1] - t[0])                # fid.py:254-263: the padded axis is rebuilt from c0 and delta
    np.testing.assert_array_equal(a["t_pad"], t_pad)
    np.testing.assert_array_equal(a["window"], np.exp(-np.pi * 5.0 * t_pad) / np.sqrt(2048))
    np.testing.assert_array_equal(a["freqs"], np.roll(np.fft.fftfreq(2048, d=t[1] - t[0]), 1024))
    with pytest.raises(ValueError, match="position"):
        chain.chain_geometry(1024, t, 2048, "middle", 5.0)


def _slot_gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from xmris_b200 import sharding

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n_in = 31                                             # odd length: the slot is padded to an even row
        rng = np.random.default_rng(5 + rank)
        row = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
        vmax = 7.5                                            # a tie: the lowest global row must win on every rank
        global_row = 1000 - 10 * rank
        slot = sharding.pack_slot(row, vmax, global_row)
        assert len(slot) == sharding.slot_elems(n_in) == 34
        send = torch.from_numpy(slot.view(np.uint8).copy())
        recv = torch.zeros((world, send.numel()), dtype=torch.uint8)
        gather = sharding.SlotAllGather(dist)
        assert gather.world_size == world
        gather(recv, send)                                    # the ONE collective of the sharded chain
        gathered = recv.numpy().view(np.complex64).reshape(world, -1)
        win, best, brow = sharding.select_winner(gathered)
        q.put((rank, win, best, brow, bool(np.array_equal(gathered[rank, :n_in], row))))
    finally:
        dist.destroy_process_group()


def test_slot_all_gather_two_ranks_gloo():
    """The sharded mode="single" exchange (sharding.SlotAllGather + the slot wire layout) on two gloo ranks: one collective,
    every rank selects the same winner, ties resolve to the lowest global row."""
    import multiprocessing as mp
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slot_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    for r in res:
        assert r[1] == 1 and r[2] == 7.5 and r[3] == 990 and r[4]      # rank 1 holds the lower global row
