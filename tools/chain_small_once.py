import os, sys
sys.path.insert(0, "/root/repo")
import torch
from xmris_b200 import chain
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
n = int(sys.argv[1]); batch = int(sys.argv[2]); fam = sys.argv[3]
fid, t = make_fids_torch(fam, batch, n, dev, seed=1)
out = torch.empty_like(fid)
chain.chain_single(fid, t, None, "end", 5.0, peak_width=100, out=out)
torch.cuda.synchronize()
