"""Small invocation of every hot kernel for compute-sanitizer (racecheck / memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, pervoxel, device as D
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
for n_in, zf, batch in [(4096, None, 700), (2048, None, 300), (1024, 2048, 300), (4096, 8192, 100), (512, None, 200), (64, 128, 50)]:
    fid, t = make_fids_torch("1H", batch, n_in, dev, seed=3)
    spec, freqs, info = chain.chain_single(fid, t, zf, "end", 5.0, peak_width=100)
    if (zf or n_in) >= 512:
        r = pervoxel.chain_all_device(fid[:40], t, zf, "end", 5.0, peak_width=100)
    back, _, _ = D.fid_to_spectrum(spec, inverse=True, in_shift=(zf or n_in) // 2, out_shift=0)
    torch.cuda.synchronize()
    print(n_in, zf, "ok", info["p0"], info["p1"], flush=True)
