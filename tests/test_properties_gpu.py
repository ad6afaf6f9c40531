"""GPU: size-independent properties of the hot path at the BASELINE configs' shapes (C2-C5), where the float64 oracle
would take too long to run in full.  Every check goes through the public device API -> C ABI; torch reductions are
used only to evaluate the property."""

import numpy as np
import pytest

from oracle import xmris_oracle as orc

pytestmark = pytest.mark.gpu


def _fids(family, batch, n, seed=77):
    import torch
    from xmris_b200.synth import make_fids_torch

    return make_fids_torch(family, batch, n, torch.device("cuda:0"), seed=seed)


SHAPES = [("C2", "1H", 64 * 64, 2048, None), ("C3", "1H", 8192, 4096, 8192), ("C4", "13C", 256 * 16 * 16, 1024, None),
          ("C5", "1H", 1 << 16, 4096, None)]


@pytest.mark.parametrize("name,family,batch,n_in,zf", SHAPES, ids=[s[0] for s in SHAPES])
def test_parseval_shard_invariance_and_round_trip(name, family, batch, n_in, zf):
    import torch
    from xmris_b200 import chain, device as D

    fid, t = _fids(family, batch, n_in)
    geo = chain.chain_geometry(n_in, t, zf, "end", 5.0)
    n_out = geo["n_out"]
    spec, _, _ = chain.chain_to_spectrum(fid, t, zf, "end", 5.0)
    # Parseval for the ortho transform (fft.md:125-133): sum |S|^2 == sum |x * w * sqrt(N)|^2 / N per spectrum
    w = torch.from_numpy(np.exp(-np.pi * 5.0 * geo["t_pad"][:n_in])).to(fid.device, torch.float32)
    e_in = (fid.abs().double() * w.double()).pow(2).sum(dim=1)
    e_out = spec.abs().double().pow(2).sum(dim=1)
    assert float(((e_out - e_in).abs() / e_in).max()) < 2e-6
    # the transform of a shard does not depend on what else is in the batch (bitwise)
    lo, hi = batch // 3, batch // 3 + 257
    part, _, _ = chain.chain_to_spectrum(fid[lo:hi].contiguous(), t, zf, "end", 5.0)
    assert torch.equal(part, spec[lo:hi])
    # a few rows against the float64 oracle
    rows = [0, batch // 2, batch - 1]
    ref, freqs = orc.chain_to_spectrum(fid[rows].cpu().numpy().astype(np.complex128), 1, t, zf, "end", 5.0)
    got = spec[rows].cpu().numpy()
    assert np.array_equal(freqs, geo["freqs"])
    assert max(np.linalg.norm(got[i] - ref[i]) / np.linalg.norm(ref[i]) for i in range(3)) < 1e-5
    # to_fid(to_spectrum(x)) == x (fid_transformations.md:144-157) on the un-windowed transform
    plain, _, _ = D.fid_to_spectrum(fid, n_out=n_in)
    back, _, _ = D.fid_to_spectrum(plain, inverse=True, in_shift=n_in // 2, out_shift=0)
    err = (back - fid).abs().pow(2).sum(dim=1).sqrt() / fid.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) < 1e-5


def test_chain_single_structure_at_scale():
    """mode="single": one (p0, p1, pivot) for every voxel, |out| == |spectrum|, the winner is the global maximum."""
    import torch
    from xmris_b200 import chain

    batch, n = 1 << 16, 4096
    fid, t = _fids("1H", batch, n, seed=5)
    spec, freqs, geo = chain.chain_to_spectrum(fid, t, None, "end", 5.0)
    out, freqs2, info = chain.chain_single(fid, t, None, "end", 5.0, peak_width=100)
    assert np.array_equal(freqs, freqs2)
    mag = spec.abs()
    flat = int(torch.argmax(mag.reshape(-1)))          # first occurrence, like numpy
    assert info["pivot"] == float(freqs[flat % n])
    assert float(((out.abs() - mag).abs().max() / mag.max())) < 2e-6
    # the applied rotation is the same unit phasor exp(i*phi_m) in every voxel
    ph = orc.phase_array(freqs, info["p0"], info["p1"], info["pivot"])
    rot = torch.from_numpy(np.exp(1j * ph)).to(out.device, torch.complex64)
    resid = (out - spec * rot[None, :]).abs().pow(2).sum(dim=1).sqrt() / mag.pow(2).sum(dim=1).sqrt()
    assert float(resid.max()) < 1e-5
    # the optimiser's answer on the winning row matches the reference's DE (one CPU search)
    row = flat // n
    ref_out, ref_info = orc.autophase(spec[row].cpu().numpy().astype(np.complex128), 0, freqs, peak_width=100)
    f_gpu = orc.acme_score([info["p0"], info["p1"]], spec[row].cpu().numpy().astype(np.complex128), freqs, info["pivot"])
    assert f_gpu <= ref_info["fun"] * (1 + 1e-6)
    if ref_info["fun"] > 0:
        ok = abs(info["p0"] - ref_info["p0"]) < 0.1 and abs(info["p1"] - ref_info["p1"]) < 0.1
        assert ok or f_gpu < ref_info["fun"], (info, ref_info)


def test_two_shard_exchange_equals_one_call():
    """Emulate the 2-rank path on one GPU: per-shard pass 1, winner by (max, lowest global index), shared angles."""
    import torch
    from xmris_b200 import chain, sharding

    batch, n = 6000, 2048
    fid, t = _fids("1H", batch, n, seed=9)
    whole, _, info = chain.chain_single(fid, t, None, "end", 5.0, peak_width=100)
    geo = chain.chain_geometry(n, t, None, "end", 5.0)
    bounds = [sharding.shard_bounds(batch, 2, r) for r in range(2)]
    stats = [chain.local_stats(fid[lo:hi].contiguous(), geo) for lo, hi in bounds]
    winner, _ = sharding.pick_winner([s[0] for s in stats], [s[1] for s in stats], [lo * n for lo, _ in bounds])
    lo, hi = bounds[winner]
    shard = fid[lo:hi].contiguous()
    p0, p1, pivot, fun = chain.search_on_row(shard[stats[winner][1] // n], geo, stats[winner][1], "acme", 100, None, False, 0.0)
    assert (p0, p1, pivot) == (info["p0"], info["p1"], info["pivot"])
    parts = [chain.apply_pass(fid[a:b].contiguous(), geo, p0, p1, pivot) for a, b in bounds]
    assert torch.equal(torch.cat(parts), whole)


def test_per_voxel_structure_at_scale():
    """mode="all" on the C4 shape: own pivot per voxel, |out| == |spectrum|, angles inside the reference's box."""
    import torch
    from xmris_b200 import chain, pervoxel

    batch, n = 256 * 16, 1024
    fid, t = _fids("13C", batch, n, seed=3)
    spec, freqs, _ = chain.chain_to_spectrum(fid, t, None, "end", 10.0)
    r = pervoxel.chain_all_device(fid, t, None, "end", 10.0, peak_width=100)
    out = r["out"]
    mag = spec.abs()
    assert torch.equal(r["pivot_index"].long(), torch.argmax(mag, dim=1))
    assert float((out.abs() - mag).abs().max() / mag.max()) < 2e-6
    p0, p1 = r["p0"], r["p1"]
    assert float(p0.abs().max()) <= 180.0 and float(p1.abs().max()) <= 4000.0
    assert bool(torch.isfinite(r["fun"]).all())
    # re-applying the reported angles to the un-phased spectrum reproduces the output
    du = (freqs[-1] - freqs[0]) / (n - 1) / (freqs.max() - freqs.min())
    u0 = -du * r["pivot_index"].double()
    m = torch.arange(n, device=out.device, dtype=torch.float64)[None, :]
    turns = (p0 / 360.0)[:, None] + (p1 / 360.0)[:, None] * (u0[:, None] + du * m)
    rot = torch.polar(torch.ones_like(turns), 2 * np.pi * turns).to(torch.complex64)
    resid = (out - spec * rot).abs().pow(2).sum(dim=1).sqrt() / mag.pow(2).sum(dim=1).sqrt()
    assert float(resid.max()) < 2e-5


@pytest.mark.parametrize("mode", ["single", "all"])
def test_host_pipeline_matches_device_chain(mode):
    """hostpipe.HostChain (pinned host buffers, chunked + pipelined) == the device-resident chain."""
    import torch
    from xmris_b200 import chain, hostpipe, pervoxel

    batch, n = (3000, 2048) if mode == "single" else (700, 1024)
    fid, t = _fids("1H", batch, n, seed=21)
    h_in = torch.empty((batch, n), dtype=torch.complex64, pin_memory=True)
    h_in.copy_(fid)
    outs = [torch.empty((batch, n), dtype=torch.complex64, pin_memory=True) for _ in range(3)]
    pipe = hostpipe.HostChain(fid.device, batch, n, t, None, "end", 5.0, mode=mode, peak_width=100, chunk=1024)
    pipe.run(h_in, outs[0])
    infos = pipe.run_many([(h_in, outs[1]), (h_in, outs[2])])
    if mode == "single":
        ref, _, info = chain.chain_single(fid, t, None, "end", 5.0, peak_width=100)
        assert (infos[0]["p0"], infos[0]["p1"], infos[0]["pivot"]) == (info["p0"], info["p1"], info["pivot"])
    else:
        ref = pervoxel.chain_all_device(fid, t, None, "end", 5.0, peak_width=100)["out"]
    for o in outs:
        assert torch.equal(o.cuda(), ref)


@pytest.mark.parametrize("mode,n_in,zf", [(None, 2048, None), ("single", 1024, 2048), ("single", 4096, None), ("all", 1024, None)])
def test_host_c_abi_chain_matches_device_chain(mode, n_in, zf):
    """xmr_chain_host_c64 (numpy in, numpy out, pageable memory, no torch) == the device-resident chain."""
    import torch
    from xmris_b200 import chain, hostabi, pervoxel

    batch = 2500
    fid, t = _fids("1H", batch, n_in, seed=33)
    host = fid.cpu().numpy()
    ap = None if mode is None else dict(mode=mode, peak_width=100)
    out, freqs, info = hostabi.chain_host(host.reshape(50, 50, n_in), t, zf, "end", 5.0, autophase=ap, chunk=700)
    assert out.shape == (50, 50, zf or n_in)
    if mode is None:
        ref, rfreqs, _ = chain.chain_to_spectrum(fid, t, zf, "end", 5.0)
    elif mode == "single":
        ref, rfreqs, rinfo = chain.chain_single(fid, t, zf, "end", 5.0, peak_width=100)
        assert (info["p0"], info["p1"], info["pivot"]) == (rinfo["p0"], rinfo["p1"], rinfo["pivot"])
    else:
        r = pervoxel.chain_all_device(fid, t, zf, "end", 5.0, peak_width=100)
        ref, rfreqs = r["out"], r["freqs"]
        np.testing.assert_array_equal(info["p0"].ravel(), r["p0"].cpu().numpy())
        np.testing.assert_array_equal(info["p1"].ravel(), r["p1"].cpu().numpy())
    np.testing.assert_array_equal(freqs, rfreqs)
    assert np.array_equal(out.reshape(batch, -1), ref.cpu().numpy())
    hostabi.release_workspace()


@pytest.mark.parametrize("n_in,zf", [(1024, 2048), (4096, None)])
def test_host_chain_streams_twice_when_the_fids_may_not_stay_resident(n_in, zf):
    """mode="single" on a batch above the resident limit (a data set larger than HBM): pass 1 over double-buffered chunks,
    the winning row fetched again, pass 2 over re-uploaded chunks -- bit-identical to the resident form."""
    from xmris_b200 import hostabi

    batch = 2500
    fid, t = _fids("1H", batch, n_in, seed=35)
    host = fid.cpu().numpy()
    ap = dict(mode="single", peak_width=100)
    ref, rfreqs, rinfo = hostabi.chain_host(host, t, zf, "end", 5.0, autophase=ap, chunk=300)
    hostabi.release_workspace()
    hostabi.set_resident_limit(1)
    try:
        out, freqs, info = hostabi.chain_host(host, t, zf, "end", 5.0, autophase=ap, chunk=300)
    finally:
        hostabi.set_resident_limit(0)
        hostabi.release_workspace()
    assert (info["p0"], info["p1"], info["pivot"]) == (rinfo["p0"], rinfo["p1"], rinfo["pivot"])
    assert np.array_equal(out, ref)


def test_device_chain_replays_a_cuda_graph_and_gives_the_same_result():
    import torch

    """xmr_chain_single_dev_c64 captures its launches up to the search's final kernel once per argument set (second call) and
    replays the graph afterwards: same angles and spectra as the eager first call, and the replay counter moves."""
    from xmris_b200 import _lib, chain
    from xmris_b200.synth import make_fids_torch

    lib = _lib.load()
    fid, t = make_fids_torch("1H", 3000, 1024, torch.device("cuda:0"), seed=77)
    outs, infos = [], []
    before = lib.xmr_chain_single_graph_launches()
    for _ in range(4):
        out, _, info = chain.chain_single(fid, t, 2048, "end", 5.0, peak_width=100)
        outs.append(out.clone())
        infos.append(info)
    torch.cuda.synchronize()
    assert lib.xmr_chain_single_graph_launches() - before >= 2          # calls 2..4 run the captured graph
    for o, i in zip(outs[1:], infos[1:]):
        assert i["p0"] == infos[0]["p0"] and i["p1"] == infos[0]["p1"] and i["pivot"] == infos[0]["pivot"]
        assert torch.equal(o, outs[0])


def test_sharded_chain_equals_the_one_call_chain():
    """front -> all-gather -> back (xmr_chain_single_front_c64 / _back_c64, the multi-GPU split of the chain) on two "ranks"
    emulated on one GPU (each half of the batch is a shard; the all-gather is a device copy): the same global winner, the same
    angles and spectra as the one-call chain on the whole batch, on both ranks."""
    import torch

    from xmris_b200 import chain, sharding
    from xmris_b200.synth import make_fids_torch

    dev = torch.device("cuda:0")
    fid, t = make_fids_torch("1H", 2001, 2048, dev, seed=31)
    whole, _, winfo = chain.chain_single(fid, t, None, "end", 5.0, peak_width=100)
    lo = [0, 1100]
    hi = [1100, 2001]
    slots = []

    class Capture:                       # pass 1 of both shards first (collects their slots) ...
        world_size = 2

        def __call__(self, recv, send):
            slots.append(send.clone())
            raise StopIteration

    for r in range(2):
        try:
            chain.chain_single(fid[lo[r]:hi[r]], t, None, "end", 5.0, peak_width=100, all_gather=Capture(), row_offset=lo[r])
        except StopIteration:
            pass
    gathered = torch.stack(slots)
    host = gathered.cpu().numpy().view(np.complex64).reshape(2, -1)
    win, best, grow = sharding.select_winner(host)
    assert grow == winfo["winning_row"] and abs(best - winfo["max_abs"]) <= 1e-6 * best

    class Replay:                        # ... then the full chain per shard with the gathered slots
        world_size = 2

        def __call__(self, recv, send):
            recv.copy_(gathered)

    for r in range(2):
        part, _, info = chain.chain_single(fid[lo[r]:hi[r]], t, None, "end", 5.0, peak_width=100, all_gather=Replay(),
                                           row_offset=lo[r])
        assert info["winning_row"] == winfo["winning_row"] and info["pivot"] == winfo["pivot"]
        assert abs(info["p0"] - winfo["p0"]) < 2e-3 and abs(info["p1"] - winfo["p1"]) < 6e-3
        err = (part - whole[lo[r]:hi[r]]).abs().max().item() / whole.abs().max().item()
        assert err < 1e-4, err
