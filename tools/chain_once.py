"""Run the mode="single" chain a few times on one BASELINE config (for an ncu launch list of its kernels).
usage: python tools/chain_once.py C2|C3|C4|C5 [calls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import chain
from xmris_b200.synth import make_fids_torch
CFG = {"C2": ("1H", 4096, 2048, None, 5.0), "C3": ("1H", 32768, 4096, 8192, 5.0), "C4": ("13C", 65536, 1024, None, 10.0),
       "C5": ("1H", 131072, 4096, None, 5.0)}
fam, batch, n_in, zf, lb = CFG[sys.argv[1]]
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
fid, t = make_fids_torch(fam, batch, n_in, dev, seed=1)
out = torch.empty((batch, zf or n_in), dtype=torch.complex64, device=dev)
for _ in range(calls):
    _, _, info = chain.chain_single(fid, t, zf, "end", lb, peak_width=100, out=out)
torch.cuda.synchronize()
print(info)
