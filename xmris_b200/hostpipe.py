"""End-to-end chain on HOST buffers: pinned input -> H2D -> fused chain -> D2H -> pinned output.

This is the call a user with host-resident data makes (and what ``bench.py`` times as ``e2e``).  Copies and kernels
are pipelined over voxel chunks on three CUDA streams so that the PCIe transfers overlap the compute:

  mode="single"  chunk c: H2D(c) || pass-1 statistics(c-1);  global argmax + one search;  then
                 chunk c: pass-2 store+phase(c) || D2H(c-1).  The device keeps the whole shard's FIDs (read twice).
  mode="all"     chunk c: H2D(c) || per-voxel chain(c-1) || D2H(c-2).
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
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.d_in = torch.empty((batch, n_in), dtype=torch.complex64, device=device)
        self.d_out = [torch.empty((self.chunk, self.n_out), dtype=torch.complex64, device=device) for _ in range(2)]
        self.absmax = torch.empty(batch, dtype=torch.float32, device=device)
        self.argmax = torch.empty(batch, dtype=torch.int32, device=device)
        self.h2d_bytes = batch * n_in * 8
        self.d2h_bytes = batch * self.n_out * 8
        self.info = None

    def _chunks(self):
        return [(lo, min(lo + self.chunk, self.batch)) for lo in range(0, self.batch, self.chunk)]

    def run(self, h_in, h_out):
        """``h_in`` [batch, n_in] / ``h_out`` [batch, n_out]: pinned complex64 host tensors.  Blocks until done."""
        torch = self.torch
        if not (h_in.is_pinned() and h_out.is_pinned()):
            raise ValueError("HostChain needs pinned host tensors")
        geo, win = self.geo, chain._win(self.geo, self.dev)
        chunks = self._chunks()
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        in_done = []
        for lo, hi in chunks:                                   # H2D of every chunk, back to back
            with torch.cuda.stream(self.s_in):
                self.d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                in_done.append(ev)
        if self.mode == "single":
            with torch.cuda.stream(self.s_cmp):
                for (lo, hi), ev in zip(chunks, in_done):       # pass 1 trails the copies chunk by chunk
                    self.s_cmp.wait_event(ev)
                    _, am, _ = D.fid_to_spectrum(self.d_in[lo:hi], n_out=self.n_out, pad_left=geo["pad_left"], window=win,
                                                 store=False, want_stats=True, want_index=False)
                    self.absmax[lo:hi].copy_(am, non_blocking=True)
                vmax, findex = D.global_argmax(self.absmax, None, self.n_out)

                def search():
                    return chain.search_on_row(self.d_in[findex // self.n_out], geo, findex, self.method, self.peak_width,
                                               self.target_coord, self.p0_only, 0.0)

                p0, p1, pivot, fun = search() if self.exchange is None else self.exchange(vmax, findex, search)
                self.info = dict(p0=p0, p1=p1, pivot=pivot, fun=fun)
                a, b, _, _ = chain.phase_turns(geo["freqs"], p0, p1, pivot)
            out_free = [None, None]
            for i, (lo, hi) in enumerate(chunks):
                buf = self.d_out[i % 2][: hi - lo]
                with torch.cuda.stream(self.s_cmp):
                    if out_free[i % 2] is not None:
                        self.s_cmp.wait_event(out_free[i % 2])
                    D.fid_to_spectrum(self.d_in[lo:hi], n_out=self.n_out, pad_left=geo["pad_left"], window=win,
                                      phase_turns=(a, b), out=buf)
                    done = torch.cuda.Event()
                    done.record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(done)
                    h_out[lo:hi].copy_(buf, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.s_out)
                    out_free[i % 2] = ev
        else:
            from . import pervoxel

            out_free = [None, None]
            for i, ((lo, hi), ev_in) in enumerate(zip(chunks, in_done)):
                buf = self.d_out[i % 2][: hi - lo]
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(ev_in)
                    if out_free[i % 2] is not None:
                        self.s_cmp.wait_event(out_free[i % 2])
                    pervoxel.chain_all_device(self.d_in[lo:hi], None, None, "end", None, out=buf, geo=geo)
                    done = torch.cuda.Event()
                    done.record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(done)
                    h_out[lo:hi].copy_(buf, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.s_out)
                    out_free[i % 2] = ev
        self.s_out.synchronize()
        self.s_cmp.synchronize()
        return self.info
