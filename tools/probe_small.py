"""Where the time of the small configurations (C2, C4) goes in chain mode=single (development probe): the C call's own CUDA
events (front = pass 1 .. search, back = pass 2) against the wall time of the whole Python call and of its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, device as D
from xmris_b200.synth import make_fids_torch

dev = torch.device("cuda:0")
for name, fam, batch, n_in, zf, lb in [("C2 64x64x2048", "1H", 4096, 2048, None, 5.0), ("C4 65536x1024 13C", "13C", 65536, 1024, None, 10.0),
                                       ("C3 32^3 4096->8192", "1H", 32768, 4096, 8192, 5.0)]:
    fid, t = make_fids_torch(fam, batch, n_in, dev, seed=1)
    n_out = zf or n_in
    out = torch.empty((batch, n_out), dtype=torch.complex64, device=dev)
    for _ in range(4):
        chain.chain_single(fid, t, zf, "end", lb, peak_width=100, out=out)
    torch.cuda.synchronize()
    walls, evs, fronts, backs = [], [], [], []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        e0.record()
        chain.chain_single(fid, t, zf, "end", lb, peak_width=100, out=out)
        e1.record()
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - w0) * 1e3)
        evs.append(e0.elapsed_time(e1))
        f, b = D.chain_single_last_timing()
        fronts.append(f); backs.append(b)
    w0 = time.perf_counter()
    for _ in range(50):
        geo = chain.chain_geometry(n_in, t, zf, "end", lb, None)
        chain._win(geo, dev)
    tg = (time.perf_counter() - w0) / 50 * 1e3
    print(f"{name:22s} wall {np.median(walls):.3f} ms | events around the call {np.median(evs):.3f} (min {min(evs):.3f}) | C-side front {np.median(fronts):.3f} "
          f"+ pass 2 {np.median(backs):.3f} = {np.median(fronts)+np.median(backs):.3f} | host geometry+window {tg:.3f} ms", flush=True)
