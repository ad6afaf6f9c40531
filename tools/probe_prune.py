import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, device as D
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
batch, n = 1 << 18, 4096
fid, t = make_fids_torch("1H", batch, n, dev, seed=1234)
geo = chain.chain_geometry(n, t, None, "end", 5.0)
win = chain._win(geo, dev)
for rep in range(3):
    absmax, run = D.fid_absmax_pruned(fid, n_out=n, window=win)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); absmax, run = D.fid_absmax_pruned(fid, n_out=n, window=win); e1.record(); torch.cuda.synchronize()
    _, am, _ = D.fid_to_spectrum(fid, n_out=n, window=win, store=False, want_stats=True, want_index=False)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(); D.fid_to_spectrum(fid, n_out=n, window=win, store=False, want_stats=True, want_index=False); e3.record(); torch.cuda.synchronize()
    print(f"zero-entry (pruned) fraction {(absmax == 0).float().mean().item():.4f}  pruned pass {e0.elapsed_time(e1):.3f} ms  plain pass {e2.elapsed_time(e3):.3f} ms  "
          f"max equal: {bool(absmax.max() == am.max())} argmax equal: {int(absmax.argmax()) == int(am.argmax())}  running {run.item()**0.5:.4f} vs {am.max().item():.4f}")
