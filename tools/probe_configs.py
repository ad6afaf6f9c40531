"""Per-config device throughput (development probe): to_spectrum-only, chain mode=single, chain mode=all."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, pervoxel
from xmris_b200.synth import make_fids_torch

dev = torch.device("cuda:0")
def timeit(fn, iters=5):
    fn(); fn(); fn(); torch.cuda.synchronize()     # (eager call, graph capture, first replay)
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

CONFIGS = [("C2 64x64x2048", "1H", 4096, 2048, None, 5.0), ("C2x64 (262144x2048)", "1H", 262144, 2048, None, 5.0),
           ("C3 32^3 4096->8192", "1H", 32768, 4096, 8192, 5.0), ("C4 65536x1024 13C", "13C", 65536, 1024, None, 10.0),
           ("C5/8 131072x4096", "1H", 131072, 4096, None, 5.0)]
only = sys.argv[1:] 
for name, fam, batch, n_in, zf, lb in CONFIGS:
    fid, t = make_fids_torch(fam, batch, n_in, dev, seed=1)
    n_out = zf or n_in
    out = torch.empty((batch, n_out), dtype=torch.complex64, device=dev)
    balg = 8.0 * (n_in + n_out) * batch
    t1 = timeit(lambda: chain.chain_to_spectrum(fid, t, zf, "end", lb, out=out))
    t2 = timeit(lambda: chain.chain_single(fid, t, zf, "end", lb, peak_width=100, out=out))
    nb = min(batch, 16384)
    t3 = timeit(lambda: pervoxel.chain_all_device(fid[:nb], t, zf, "end", lb, out=out[:nb], peak_width=100), iters=2)
    print(f"{name:24s} to_spectrum {t1:8.3f} ms {batch/t1/1e3:8.2f} Mspec/s {balg/t1/1e6:7.0f} GB/s | single {t2:8.3f} ms "
          f"{batch/t2/1e3:8.2f} Mspec/s ({balg/t2/1e6/6550.1:.2f} of roofline) | all ({nb}) {t3:8.2f} ms {nb/t3:8.1f} kspec/s", flush=True)
    del fid, out
