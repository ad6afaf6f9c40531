#!/usr/bin/env python
"""Benchmark of the FID -> spectrum hot path (BASELINE.json metric: spectra/sec, 4096-pt, full chain).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode single|all]

One "step" = one pass of the full chain ``zero_fill -> apodize_exp -> to_spectrum -> autophase`` over one batch of
synthetic Lorentzian FIDs (config C5: 2^20 voxels x 4096 points per GPU, lb = 5 Hz).  ``value`` is whole-job
throughput with the FIDs resident in HBM; ``e2e`` is the same chain with pinned HOST buffers and the H2D / D2H
copies inside the timed region.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 4096
LB = 5.0
PEAK_WIDTH = 100          # accessor default (accessor.py:634)
METRIC = "spectra/sec (4096-pt, full chain)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="single", choices=["single", "all"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="voxels per GPU")
    ap.add_argument("--e2e-batch", type=int, default=1 << 16, help="voxels per GPU for the host-buffer (e2e) leg")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="voxels of the CPU baseline sample")
    ap.add_argument("--ref-sample", type=int, default=65536, help="voxels per step of the --impl reference arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-voxel", action="store_true")
    ap.add_argument("--per-voxel-batch", type=int, default=1 << 17, help="voxels per GPU for the mode=all side measurement")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs: NVML polled from a thread every 2 ms
    (`nvidia-smi -lms` through a pipe is block-buffered and loses the samples of a sub-second region); falls back to
    single `nvidia-smi` queries when NVML cannot be loaded."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.rows, self.first, self.alive, self.thread, self.how = index, [], 0, False, None, None

    def _nvml_handle(self):
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(self.index).uuid)
        return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))

            def poll():
                while self.alive:
                    try:
                        self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), mx, int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.how = "nvml"
        except Exception:
            def poll():
                while self.alive:
                    try:
                        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                        r = [c.strip() for c in out.strip().split(",")]
                        mask = sum(bit for (bit, _), col in zip(self.REASONS, (4, 5, 6, 7))
                                   if len(r) > col and r[col].lower().startswith("active"))
                        self.rows.append((float(r[1]), float(r[2]), mask))
                    except Exception:
                        time.sleep(0.05)

            self.how = "nvidia-smi"
        self.alive = True
        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()

    def mark(self):
        """Index of the next sample: call at the start of the timed region (the sampler is started before warm-up)."""
        self.first = len(self.rows)

    def stop(self):
        self.alive = False
        if self.thread is not None:
            self.thread.join(timeout=6)
        rows = self.rows[self.first:] or self.rows[-1:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"],
                    "samples": 0}
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": max(r[1] for r in rows),
                "reasons": sorted(name for bit, name in self.REASONS if mask & bit), "samples": len(rows), "source": self.how}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (numpy/scipy restatement of the reference's own path)
# ---------------------------------------------------------------------------------------------------------

def _cpu_worker(conn, seed, rows, n_points):
    """One host process of the reference arm: owns a block of voxels and serves the chain's whole-array steps."""
    from oracle import xmris_oracle as orc
    from xmris_b200.synth import make_fids_numpy

    fid, t, _ = make_fids_numpy("1H", rows, n_points, seed=seed)
    spec = freqs = None
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd[0] == "pass_a":      # zero_fill -> apodize_exp -> to_spectrum + local argmax (phasing.py:229-231)
            spec, freqs = orc.chain_to_spectrum(fid, 1, t, None, "end", LB)
            a = np.abs(spec)
            flat = int(np.argmax(a))
            conn.send((float(a.ravel()[flat]), flat))
        elif cmd[0] == "row":
            conn.send((spec[cmd[1]].copy(), freqs))
        elif cmd[0] == "pass_b":    # phase() on the whole block (phasing.py:290)
            out, _ = orc.phase(spec, 1, freqs, cmd[1], cmd[2], cmd[3])
            conn.send(float(np.abs(out[0, 0])))
        else:
            conn.close()
            return


def cpu_chain_single_process(n_spectra, n_points):
    """The reference path exactly as it runs: one process, whole-array numpy ops, one DE search."""
    from oracle import xmris_oracle as orc
    from xmris_b200.synth import make_fids_numpy

    fid, t, _ = make_fids_numpy("1H", n_spectra, n_points, seed=4321)
    t0 = time.perf_counter()
    out, freqs, info = orc.chain(fid, 1, t, None, "end", LB, mode="single", peak_width=PEAK_WIDTH)
    dt = time.perf_counter() - t0
    return n_spectra / dt, dt, info


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path with all host cores (voxel blocks in worker
    processes; the single DE search of mode="single" runs once on the winning spectrum, as in phasing.py:276-290)."""
    import multiprocessing as mp

    from oracle import xmris_oracle as orc

    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, 64))
    rows = max(64, args.ref_sample // workers)
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for w in range(workers):
        parent, child = ctx.Pipe()
        pr = ctx.Process(target=_cpu_worker, args=(child, 1000 + w, rows, N_POINTS), daemon=True)
        pr.start()
        conns.append(parent)
        procs.append(pr)
    for c in conns:
        assert c.recv() == "ready"

    def step():
        for c in conns:
            c.send(("pass_a",))
        res = [c.recv() for c in conns]
        win = max(range(workers), key=lambda i: (res[i][0], -i))
        row, idx = divmod(res[win][1], N_POINTS)
        conns[win].send(("row", row))
        spec1d, freqs = conns[win].recv()
        pivot = float(freqs[idx])
        p0, p1, _ = orc.autophase_search(spec1d, freqs, pivot, "acme", idx, 1, False)
        for c in conns:
            c.send(("pass_b", p0, p1, pivot))
        for c in conns:
            c.recv()

    times = []
    for _ in range(args.warmup):
        step()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    for c in conns:
        c.send(("quit",))
    for pr in procs:
        pr.join(timeout=10)
    total = rows * workers
    ms = 1e3 * float(np.mean(times))
    value = total / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": f"C5: {args.batch} voxels x {N_POINTS}-pt FID per GPU -> {N_POINTS}-pt spectrum, lb={LB}, full chain "
                               f"zero_fill(no-op)->apodize_exp->to_spectrum->autophase(mode=single, acme)",
                   "l2": "inputs (32 GiB/GPU) far exceed the 126 MB L2; no explicit flush", "autophase_mode": "single",
                   "sample_voxels_per_step": total,
                   "impl": "oracle port of the reference (numpy pocketfft + scipy differential_evolution); each step is a bounded "
                           "sample of the workload (same chain, same per-step search), sized to finish within minutes"},
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": workers, "kind": "port",
                         "sample": f"{total} voxels x {N_POINTS} pts per step, process pool over voxel blocks"},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------


def bind_host_to_gpu(torch, local_rank):
    """Multi-GPU runs: restrict this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned
    e2e buffers are first-touched on that GPU's NUMA node.  Returns the CPU list used (None: left unchanged)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = local & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return sorted(use)
    except Exception:   # no NVML / no affinity information: leave the scheduler alone
        pass
    return None



def run_ours(args):
    import torch

    from xmris_b200 import _lib, chain, sharding
    from xmris_b200.synth import make_fids_torch, time_coord

    _lib.load()   # fail loudly when the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cpus = bind_host_to_gpu(torch, local_rank) if world > 1 else None
    batch, n = args.batch, N_POINTS
    t = time_coord(n, 5000.0)

    fid = torch.empty((batch, n), dtype=torch.complex64, device=dev)
    make_fids_torch("1H", batch, n, dev, seed=1234 + rank, out=fid)
    out = torch.empty((batch, n), dtype=torch.complex64, device=dev)
    # multi-GPU: ONE device-side all-gather of the ranks' candidate rows per step (sharding.SlotAllGather); no host round trip
    gather = sharding.SlotAllGather(dist) if dist is not None else None

    launches = {"n": 0}
    k2_ms = []
    parts = []
    from xmris_b200 import device as D

    def step(record=False):
        if args.mode == "single":
            # the public device-resident entry point: ONE call = pass 1 -> winner -> search -> pass 2 (graph-replayed front)
            _, _, info = chain.chain_single(fid, t, None, "end", LB, peak_width=PEAK_WIDTH, out=out, all_gather=gather,
                                            row_offset=rank * batch)
            if record:
                parts.append(D.chain_single_last_timing())     # CUDA events on the launching stream, recorded by the C side
            # our kernels per step: K1-max, argmax x2, pack, select, K1 (1 row), geometry, coarse, zoom, polish, fine A, fine B,
            # finalize, K1 store+phase
            launches["n"] += 14
            return info["p0"], info["p1"], info["pivot"], info
        from xmris_b200 import pervoxel

        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        info = pervoxel.chain_all_device(fid, t, None, "end", LB, out=out)
        if record:
            e1.record()
            k2_ms.append((e0, e1))
        launches["n"] += info["launches"]
        return info

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark()
    launches["n"] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    last = None
    for _ in range(args.steps):
        last = step(record=True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = world * batch / (ms_step / 1e3)
    breakdown = None
    if parts:
        dom_ms = float(np.mean([b for _, b in parts]))
        breakdown = {"pass1_stats_argmax_search_ms": float(np.mean([a for a, _ in parts])), "pass2_store_phase_ms": dom_ms}
    else:
        dom_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_ms]))

    # ---- parity of what the timed region produced (rank 0's shard): 256 random output rows and the angles against the oracle
    parity = None
    worst = None
    if args.mode == "single":
        from oracle import xmris_oracle as orc

        p0_g, p1_g, piv_g, info_g = last
        if rank == 0:
            rng = np.random.default_rng(99)
            rows = np.unique(rng.integers(0, batch, size=256))
            fid_rows = fid[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.complex128)
            ref_spec, ref_freqs = orc.chain_to_spectrum(fid_rows, 1, t, None, "end", LB)
            want, _ = orc.phase(ref_spec, 1, ref_freqs, p0_g, p1_g, piv_g)
            got = out[torch.from_numpy(rows).to(dev)].cpu().numpy()
            rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
            parity = {"rows_checked": int(len(rows)), "max_rel_l2": float(rel.max()), "tolerance_rel_l2": 1e-5}
            wr = int(info_g.get("winning_row", -1)) - rank * batch
            if 0 <= wr < batch:
                # the reference's optimiser (seeded DE + polish, phasing.py:276-284) on the winning spectrum
                w_spec, _ = orc.chain_to_spectrum(fid[wr:wr + 1].cpu().numpy().astype(np.complex128), 1, t, None, "end", LB)
                _, rinfo = orc.autophase(w_spec[0], 0, ref_freqs, peak_width=PEAK_WIDTH)
                f_gpu = float(orc.acme_score([p0_g, p1_g], w_spec[0], ref_freqs, piv_g))
                parity.update({"dp0_deg": abs(((p0_g - rinfo["p0"] + 180.0) % 360.0) - 180.0), "dp1_deg": abs(p1_g - rinfo["p1"]),
                               "pivot_equal": bool(piv_g == rinfo["pivot"]), "objective_gpu": f_gpu,
                               "objective_reference": float(rinfo["fun"]), "tolerance_deg": 0.1,
                               "winning_row": wr})
            else:
                parity["angles"] = "winning row lives on another rank"
        # ---- pass 1 in its worst case: amplitudes ascending with the row index defeat the branch-and-bound pruning ------------
        geo = chain.chain_geometry(n, t, None, "end", LB)
        _, amax, _ = D.fid_to_spectrum(fid, n_out=n, window=chain._win(geo, dev), store=False, want_stats=True, want_index=False)
        order = torch.argsort(amax.reshape(-1))                 # rows by ascending max |S|: every row is a new record
        for s0 in range(0, batch, 1 << 16):
            torch.index_select(fid, 0, order[s0:s0 + (1 << 16)], out=out[s0:s0 + (1 << 16)])
        del amax, order
        chain.local_stats(out, geo)
        torch.cuda.synchronize()
        wa, wb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wa.record()
        for _ in range(3):
            chain.local_stats(out, geo)
        wb.record()
        torch.cuda.synchronize()
        worst = wa.elapsed_time(wb) / 3

    # ---- the per-voxel chain (autophase mode="all", kernel K2) on a bounded slice of the same workload -----------------
    per_voxel = None
    if args.mode == "single" and not args.no_per_voxel:
        from xmris_b200 import pervoxel

        nb = min(batch, args.per_voxel_batch)
        pervoxel.chain_all_device(fid[:nb], t, None, "end", LB, out=out[:nb], peak_width=PEAK_WIDTH)
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(2):
            pervoxel.chain_all_device(fid[:nb], t, None, "end", LB, out=out[:nb], peak_width=PEAK_WIDTH)
        pe1.record()
        barrier()
        pv_ms = pe0.elapsed_time(pe1) / 2
        if dist is not None:
            tm = torch.tensor([pv_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            pv_ms = float(tm.item())
        per_voxel = {"value": world * nb / (pv_ms / 1e3), "unit": "spectra/s", "voxels_per_gpu": nb, "ms_per_step": pv_ms,
                     "what": "same chain with autophase mode='all' (one fused kernel per voxel: FFT + in-CTA (p0,p1) search + "
                             "phase); SFU/FP32-bound, see DESIGN.md K2"}

    # ---- the step after the chain in the reference's pipeline: baseline_als on the phased spectra (kernel K3) ----------
    als = None
    if args.mode == "single" and not args.no_per_voxel and rank == 0:
        from xmris_b200 import device as D

        nb = min(batch, args.per_voxel_batch)
        als_out = torch.empty((nb, n), dtype=torch.float32, device=dev)
        D.baseline_als(out[:nb], out=als_out)
        torch.cuda.synchronize()
        ae0, ae1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ae0.record()
        for _ in range(2):
            D.baseline_als(out[:nb], out=als_out)
        ae1.record()
        torch.cuda.synchronize()
        als_ms = ae0.elapsed_time(ae1) / 2
        als_bytes = (58.0 * n * 10 + 20.0 * n) * nb
        als = {"value": nb / (als_ms / 1e3), "unit": "spectra/s", "voxels": nb, "ms": als_ms, "n_iter": 10,
               "gbs": als_bytes / (als_ms / 1e3) / 1e9, "algorithmic_bytes": als_bytes,
               "what": "baseline_als (lam=1e5, p=0.001, 10 iterations) on the phased spectra of one GPU: batched banded "
                       "LDL^T in float64, HBM-bound on 58 B/point/iteration of factor scratch; see DESIGN.md K3"}
        del als_out
        torch.cuda.empty_cache()

    # ---- e2e: pinned host buffers, H2D + chain + D2H inside the timed region ---------------------------------
    e2e = None
    if not args.no_e2e:
        from xmris_b200 import hostpipe

        eb = min(args.e2e_batch, batch)
        h_in = torch.empty((eb, n), dtype=torch.complex64, pin_memory=True)
        h_in.copy_(fid[:eb])
        h_out = torch.empty((eb, n), dtype=torch.complex64, pin_memory=True)
        pipe = hostpipe.HostChain(dev, eb, n, t, None, "end", LB, mode=args.mode, peak_width=PEAK_WIDTH,
                                  exchange=None if dist is None else sharding.make_exchange(dist, dev, n, rank * eb * n))
        for _ in range(2):
            pipe.run(h_in, h_out)
        barrier()
        t0 = time.perf_counter()
        reps = min(max(10, 4 * args.steps), 40)   # long enough a stream that the one-batch pipeline fill is amortised
        # a stream of `reps` batches through the public host API; every batch is uploaded, processed and written back
        pipe.run_many([(h_in, h_out)] * reps)
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            tm = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        e2e = {"value": world * eb / dt, "unit": "spectra/s", "h2d_bytes_per_step": int(eb * n * 8),
               "d2h_bytes_per_step": int(eb * n * 8), "voxels_per_gpu": eb, "ms_per_step": dt * 1e3, "batches_timed": reps,
               "api": "hostpipe.HostChain.run_many: pinned host batches, H2D of batch i+1 overlaps D2H of batch i"}
        if host_cpus is not None:
            e2e["host_cpus_bound_to_gpu"] = len(host_cpus)

        # the PCIe ceiling this leg runs against: the same pinned buffers copied with nothing else going on
        # (one direction at a time, then both at once), every rank at the same time
        def copy_rate(h2d, d2h):
            sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            barrier()
            t1 = time.perf_counter()
            for _ in range(4):
                if h2d:
                    with torch.cuda.stream(sa):
                        pipe.d_ins[0].copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(sb):
                        h_out.copy_(pipe.d_ins[1], non_blocking=True)
            sa.synchronize()
            sb.synchronize()
            return 4 * eb * n * 8 / (time.perf_counter() - t1) / 1e9

        up, down, both = copy_rate(True, False), copy_rate(False, True), copy_rate(True, True)
        e2e["pcie"] = {"h2d_gbs": up, "d2h_gbs": down, "duplex_each_way_gbs": both,
                       "what": "plain pinned copies of the same buffers, GB/s per direction (rank 0%s)"
                               % ("" if dist is None else ", all ranks copying at once")}
        e2e["frac_of_pcie_duplex"] = (eb * n * 8 / dt / 1e9) / both

        # the same chain through ONE C-ABI call on the same pinned host buffers (numpy view, no torch on the path)
        if dist is None:
            from xmris_b200 import hostabi

            np_in, np_out = h_in.numpy(), h_out.numpy()
            ap = dict(mode=args.mode, peak_width=PEAK_WIDTH)
            hostabi.chain_host(np_in, t, None, "end", LB, autophase=ap, out=np_out)
            t0 = time.perf_counter()
            for _ in range(3):
                hostabi.chain_host(np_in, t, None, "end", LB, autophase=ap, out=np_out)
            dtc = (time.perf_counter() - t0) / 3
            e2e["c_abi_single_call"] = {"value": eb / dtc, "unit": "spectra/s", "ms_per_call": dtc * 1e3,
                                        "api": "xmr_chain_host_c64 (one synchronous call per batch, no overlap across batches)"}
            hostabi.release_workspace()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_alg = 8.0 * (n + n) * batch                       # FID read once + spectrum written once
    achieved = b_alg / (dom_ms / 1e3) / 1e9
    chain_achieved = b_alg / (ms_step / 1e3) / 1e9
    dominant = "k1_kernel<4096> store+phase (pass 2)" if args.mode == "single" else "k2 per-voxel chain kernel"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if args.mode == "single" and os.path.isfile(tpath):
        tj = json.load(open(tpath))    # bytes/spectrum from the committed ncu --set full capture, scaled to this launch
        traffic = (tj["dram_bytes_read_per_spectrum"] + tj["dram_bytes_write_per_spectrum"]) * batch
    if als is not None:
        als["frac_of_hbm_peak"] = als["gbs"] / peak
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": b_alg, "kernel": dominant, "kernel_ms": dom_ms, "peak_source": peak_src,
                "chain_achieved": chain_achieved, "chain_frac": chain_achieved / peak,
                "note": "frac = the dominant kernel alone; chain_frac = the WHOLE step (the number the metric names): "
                        "8*(n_in+n_out)*batch / step time; mode=single must read the FID twice (compulsory 8*(2*n_in+n_out)), "
                        "so chain_frac <= 2/3"}
    if breakdown is not None and worst is not None:
        breakdown["pass1_worst_case_ms"] = worst
        breakdown["pass1_worst_case_what"] = ("pass 1 + argmax alone on the same FIDs re-ordered by ascending max |S| (every row "
                                              "beats the running maximum): nothing is pruned, every spectrum is transformed in full")
        breakdown["chain_frac_worst_case"] = b_alg / ((ms_step + max(0.0, worst - breakdown["pass1_stats_argmax_search_ms"])) / 1e3) / 1e9 / peak

    cpu = None
    if not args.no_cpu:
        v, dt, info = cpu_chain_single_process(args.cpu_sample, n)
        cpu = {"value": v, "unit": "spectra/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_sample} voxels x {n} pts, full chain mode=single, {dt:.1f} s, host has "
                         f"{len(os.sched_getaffinity(0))} usable cores"}

    line = {
        "metric": METRIC, "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": f"C5: {batch} voxels x {n}-pt FID per GPU -> {n}-pt spectrum, lb={LB}, full chain "
                               f"zero_fill(no-op)->apodize_exp->to_spectrum->autophase(mode={args.mode}, acme)",
                   "l2": "inputs (32 GiB/GPU) far exceed the 126 MB L2; no explicit flush", "autophase_mode": args.mode},
        "chain_roofline_frac": roofline["chain_frac"],
        "roofline": roofline, "breakdown_ms": breakdown, "parity": parity, "per_voxel": per_voxel, "baseline_als": als, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches["n"], "clocks": clocks,
        "result": {"p0": last[0], "p1": last[1], "pivot": last[2]} if args.mode == "single" else None,
        "exchange": None if world == 1 else "one all_gather of the ranks' candidate rows per step (device-side winner selection, "
                                            "redundant search); no host round trip between the passes",
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    """Print the ONE JSON line on the real stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse()
    # stdout carries the JSON line and nothing else: file descriptor 1 is pointed at stderr for the rest of the run, so
    # that native libraries which print to stdout (the NCCL version banner) cannot get in front of it
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference_arm(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
