"""to_spectrum throughput of zero-filled geometries at large batch (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import chain
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
CASES = [tuple(int(v) if v != "None" else None for v in a.split(",")) for a in sys.argv[1:]]   # "n_in,n_out|None,batch"
for n_in, zf, batch in CASES or [(4096, None, 262144), (2048, 4096, 262144), (1024, 4096, 262144), (2048, None, 524288), (1024, 2048, 524288),
                        (4096, 8192, 131072), (8192, None, 131072), (2048, 8192, 131072), (1024, None, 1 << 20), (512, 1024, 1 << 20)]:
    fid, t = make_fids_torch("1H", batch, n_in, dev, seed=1)
    n_out = zf or n_in
    out = torch.empty((batch, n_out), dtype=torch.complex64, device=dev)
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); chain.chain_to_spectrum(fid, t, zf, "end", 5.0, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    gb = 8.0 * (n_in + n_out) * batch / 1e9
    tc = []
    for i in range(8):     # (call 1 runs eagerly, call 2 captures the CUDA graph, later calls replay it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); chain.chain_single(fid, t, zf, "end", 5.0, peak_width=100, out=out); e1.record(); torch.cuda.synchronize()
        tc.append(e0.elapsed_time(e1))
    mc = min(tc[3:])
    print(f"{n_in:5d} -> {n_out:5d} x {batch:8d}: to_spectrum {ms:7.3f} ms  {gb/ms*1e3:6.0f} GB/s  ({gb/ms*1e3/6550.1:.2f} of roofline)"
          f" | chain mode=single {mc:7.3f} ms ({gb/mc*1e3/6550.1:.2f} of roofline; compulsory 8*(2*n_in+n_out): "
          f"{8.0*(2*n_in+n_out)*batch/1e9/mc*1e3/6550.1:.2f})", flush=True)
    del fid, out
