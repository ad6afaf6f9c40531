"""Throughput of baseline_als on the device (development probe): spectra/s and scratch-traffic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import device as D
dev = torch.device("cuda:0")
CASES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]      # "batch,n,n_iter"
for batch, n, it in CASES or [(65536, 1024, 10), (65536, 2048, 10), (65536, 4096, 10), (262144, 4096, 10), (1048576, 1024, 10)]:
    x = torch.randn(batch, n, device=dev).cumsum(dim=1).to(torch.complex64)
    out = torch.empty(batch, n, dtype=torch.float32, device=dev)
    ts = []
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.baseline_als(x, n_iter=it, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    gb = (58.0 * n * it + 12.0 * n + 8.0 * n) * batch / 1e9
    print(f"baseline_als {batch} x {n}, {it} iterations: {ms:8.2f} ms  {batch/ms:8.1f} kspec/s  {gb/ms*1e3:6.0f} GB/s "
          f"({gb/ms*1e3/6550.1:.2f} of the HBM roofline on 58 B/point/iteration)", flush=True)
    del x, out
