"""Quick device timing of K1 variants (development probe; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import device as D

dev = torch.device("cuda:0")
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), float(np.median(ts))

for (batch, n_in, n_out) in [(1 << 17, 4096, 4096), (1 << 16, 4096, 8192), (1 << 18, 2048, 2048), (1 << 19, 1024, 1024)]:
    x = torch.randn((batch, n_in, 2), device=dev, dtype=torch.float32)
    x = torch.view_as_complex(x)
    out = torch.empty((batch, n_out), dtype=torch.complex64, device=dev)
    t = np.arange(n_out) / 5000.0
    w = np.exp(-np.pi * 5.0 * t) / np.sqrt(n_out)
    bytes_alg = 8.0 * (n_in + n_out) * batch
    for name, fn in [
        ("store", lambda: D.fid_to_spectrum(x, n_out=n_out, window=w, out=out)),
        ("store+stats", lambda: D.fid_to_spectrum(x, n_out=n_out, window=w, out=out, want_stats=True)),
        ("stats only", lambda: D.fid_to_spectrum(x, n_out=n_out, window=w, store=False, want_stats=True)),
        ("store+phase", lambda: D.fid_to_spectrum(x, n_out=n_out, window=w, out=out, phase_turns=(0.1, 0.001))),
    ]:
        best, med = timeit(fn)
        b = bytes_alg if name != "stats only" else 8.0 * n_in * batch
        print(f"{n_in}->{n_out} x{batch} {name:12s} best {best:8.3f} ms  med {med:8.3f} ms  {b/best/1e6:8.1f} GB/s  "
              f"{batch/best/1e3:8.2f} Mspec/s", flush=True)
    # copy reference
    a = torch.empty(batch * n_in * 2, device=dev); b2 = torch.empty_like(a)
    best, med = timeit(lambda: b2.copy_(a))
    print(f"   torch copy same bytes-in: {2*a.numel()*4/best/1e6:.1f} GB/s", flush=True)
    del x, out, a, b2
