"""End-to-end chain on HOST buffers: pinned input -> H2D -> fused chain -> D2H -> pinned output.

This is the call a user with host-resident data makes (and what ``bench.py`` times as ``e2e``).  Copies and kernels
are pipelined over voxel chunks on three CUDA streams so that the PCIe transfers overlap the compute:

  mode="single"  chunk c: H2D(c) || pass-1 statistics(c-1);  global argmax + one search;  then
                 chunk c: pass-2 store+phase(c) || D2H(c-1).  The device keeps the whole shard's FIDs (read twice).
  mode="all"     chunk c: H2D(c) || per-voxel chain(c-1) || D2H(c-2).
``run_many`` software-pipelines consecutive batches: the upload of batch i+1 overlaps the write-back of batch i.
"""

from __future__ import annotations

import numpy as np

from . import chain
from . import device as D


class HostChain:
    def __init__(self, device, batch, n_in, time_coord, target_points=None, position="end", lb=None, mode="single",
                 method="acme", peak_width=0.5, target_coord=None, p0_only=False, chunk=8192, exchange=None):
        import torch

        self.torch = torch
        self.dev = device
        self.batch, self.n_in = batch, n_in
        self.geo = chain.chain_geometry(n_in, time_coord, target_points, position, lb)
        self.n_out = self.geo["n_out"]
        self.mode, self.method, self.peak_width = mode, method, peak_width
        self.target_coord, self.p0_only, self.exchange = target_coord, p0_only, exchange
        self.chunk = min(chunk, batch)
        self.s_in, self.s_cmp, self.s_out, self.s_a = (torch.cuda.Stream(device) for _ in range(4))
        self.stats_done = [None, None]
        # two input sets: while batch i is being written back, batch i+1 is already being uploaded (run_many)
        self.d_ins = [torch.empty((batch, n_in), dtype=torch.complex64, device=device) for _ in range(2)]
        self.absmaxs = [torch.empty(batch, dtype=torch.float32, device=device) for _ in range(2)]
        self.d_in, self.absmax = self.d_ins[0], self.absmaxs[0]
        self.running = [torch.zeros(1, dtype=torch.float32, device=device) for _ in range(2)]
        self.d_out = [torch.empty((self.chunk, self.n_out), dtype=torch.complex64, device=device) for _ in range(2)]
        self.h2d_bytes = batch * n_in * 8
        self.d2h_bytes = batch * self.n_out * 8
        self.info = None

    def _chunks(self):
        return [(lo, min(lo + self.chunk, self.batch)) for lo in range(0, self.batch, self.chunk)]

    def run(self, h_in, h_out):
        """``h_in`` [batch, n_in] / ``h_out`` [batch, n_out]: pinned complex64 host tensors.  Blocks until done."""
        return self.run_many([(h_in, h_out)])[0]

    # ---- a stream of batches: upload of batch i+1 overlaps the write-back of batch i (PCIe is full duplex) -----------
    def _stage_a(self, h_in, k):
        """Enqueue H2D of a whole batch into input set k and (mode="single") the statistics pass trailing the copies."""
        torch = self.torch
        geo, win = self.geo, chain._win(self.geo, self.dev)
        d_in, absmax = self.d_ins[k], self.absmaxs[k]
        done = []
        first = True
        for lo, hi in self._chunks():
            with torch.cuda.stream(self.s_in):
                d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            done.append(ev)
            if self.mode == "single":
                # statistics run on their own stream so that batch i+1's pass 1 never queues in front of batch i's pass 2
                with torch.cuda.stream(self.s_a):
                    self.s_a.wait_event(ev)
                    if self.n_out in D.SUPPORTED_N:
                        # the running maximum is shared by the chunks of this batch (reset on the first one)
                        D.fid_absmax_pruned(d_in[lo:hi], n_out=self.n_out, pad_left=geo["pad_left"], window=win,
                                            absmax=absmax[lo:hi], running=self.running[k], reset=first)
                    else:
                        _, am, _ = D.fid_to_spectrum(d_in[lo:hi], n_out=self.n_out, pad_left=geo["pad_left"], window=win,
                                                     store=False, want_stats=True, want_index=False)
                        absmax[lo:hi].copy_(am, non_blocking=True)
                    first = False
        if self.mode == "single":
            ev = torch.cuda.Event()
            ev.record(self.s_a)
            self.stats_done[k] = ev
        return done

    def _stage_b(self, h_out, k, in_done, out_free):
        """Search (mode="single": one host round trip) + output pass per chunk with the D2H copies trailing it."""
        torch = self.torch
        geo, win = self.geo, chain._win(self.geo, self.dev)
        d_in = self.d_ins[k]
        a = b = None
        if self.mode == "single":
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.stats_done[k])
                vmax, findex = D.global_argmax(self.absmaxs[k], None, self.n_out)

                def search():
                    return chain.search_on_row(d_in[findex // self.n_out], geo, findex, self.method, self.peak_width,
                                               self.target_coord, self.p0_only, 0.0)

                p0, p1, pivot, fun = search() if self.exchange is None else self.exchange(vmax, findex, search)
                self.info = dict(p0=p0, p1=p1, pivot=pivot, fun=fun)
                a, b, _, _ = chain.phase_turns(geo["freqs"], p0, p1, pivot)
        for i, (lo, hi) in enumerate(self._chunks()):
            buf = self.d_out[i % 2][: hi - lo]
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(in_done[i])
                if out_free[i % 2] is not None:
                    self.s_cmp.wait_event(out_free[i % 2])
                if self.mode == "single":
                    D.fid_to_spectrum(d_in[lo:hi], n_out=self.n_out, pad_left=geo["pad_left"], window=win,
                                      phase_turns=(a, b), out=buf)
                else:
                    from . import pervoxel

                    pervoxel.chain_all_device(d_in[lo:hi], None, None, "end", None, out=buf, geo=geo, method=self.method,
                                              peak_width=self.peak_width, target_coord=self.target_coord,
                                              p0_only=self.p0_only)
                done = torch.cuda.Event()
                done.record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                h_out[lo:hi].copy_(buf, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_out)
                out_free[i % 2] = ev
        return out_free

    def run_many(self, pairs):
        """Process a sequence of ``(h_in, h_out)`` pinned batches (each ``[batch, n]``).  Every batch is copied in,
        transformed + phased and copied out; consecutive batches are software-pipelined.  Blocks until all are done."""
        torch = self.torch
        pairs = list(pairs)
        for h_in, h_out in pairs:
            if not (h_in.is_pinned() and h_out.is_pinned()):
                raise ValueError("HostChain needs pinned host tensors")
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out, self.s_a):
            s.wait_stream(cur)
        out_free = [None, None]
        infos = []
        in_done = self._stage_a(pairs[0][0], 0) if pairs else None
        last_b_done = [None, None]    # input set k may be overwritten once batch (i-2)'s output pass has been enqueued+run
        for i, (h_in, h_out) in enumerate(pairs):
            nxt = None
            if i + 1 < len(pairs):
                k2 = (i + 1) % 2
                if last_b_done[k2] is not None:
                    self.s_in.wait_event(last_b_done[k2])      # set k2 was last read by batch i-1's output pass
                nxt = self._stage_a(pairs[i + 1][0], k2)
            out_free = self._stage_b(h_out, i % 2, in_done, out_free)
            ev = torch.cuda.Event()
            ev.record(self.s_cmp)
            last_b_done[i % 2] = ev
            infos.append(self.info)
            in_done = nxt
        self.s_out.synchronize()
        self.s_cmp.synchronize()
        self.s_a.synchronize()
        return infos
