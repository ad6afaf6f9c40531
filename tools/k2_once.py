"""One per-voxel chain (K2) call on `n batch [family]` -- the command behind the ncu captures of K2 at other lengths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import pervoxel
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
n, batch = int(sys.argv[1]), int(sys.argv[2])
fam = sys.argv[3] if len(sys.argv) > 3 else "1H"
fid, t = make_fids_torch(fam, batch, n, dev, seed=1)
out = torch.empty_like(fid)
for i in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pervoxel.chain_all_device(fid, t, None, "end", 5.0, out=out, peak_width=100); e1.record(); torch.cuda.synchronize()
    print(f"n={n} batch={batch}: {e0.elapsed_time(e1):.2f} ms = {batch / e0.elapsed_time(e1):.1f} k spectra/s")
