"""Pass 2 with its phase parameters from kernel parameters (constant bank) against the device-memory variant the chain uses
(K1_FAST_PHDEV: a 128-byte shared table) -- development probe, C5 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, device as D
from xmris_b200.synth import make_fids_torch

dev = torch.device("cuda:0")
batch, n = 1 << 20, 4096
fid, t = make_fids_torch("1H", batch, n, dev, seed=1)
out = torch.empty((batch, n), dtype=torch.complex64, device=dev)
geo = chain.chain_geometry(n, t, None, "end", 5.0, None)
w = chain._win(geo, dev)
for rep in range(4):
    ts = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D.fid_to_spectrum(fid, n_out=n, window=w, phase_turns=(0.123, 0.000321), out=out)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    host = min(ts[1:])
    backs = []
    for i in range(6):
        chain.chain_single(fid, t, None, "end", 5.0, peak_width=100, out=out)
        torch.cuda.synchronize()
        backs.append(D.chain_single_last_timing()[1])
    print(f"host-parameter variant {host:.3f} ms | device-parameter variant (chain pass 2) {min(backs[1:]):.3f} ms", flush=True)
