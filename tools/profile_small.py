"""Host-side overhead of the chain on a small batch (C2: 4096 x 2048): cProfile over repeated calls."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import chain
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
fid, t = make_fids_torch("1H", 4096, 2048, dev, seed=1)
out = torch.empty_like(fid)
for _ in range(5):
    chain.chain_single(fid, t, None, "end", 5.0, peak_width=100, out=out)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    chain.chain_single(fid, t, None, "end", 5.0, peak_width=100, out=out)
torch.cuda.synchronize()
print("chain_single per call: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
t0 = time.perf_counter()
for _ in range(50):
    chain.chain_to_spectrum(fid, t, None, "end", 5.0, out=out)
torch.cuda.synchronize()
print("chain_to_spectrum per call: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    chain.chain_single(fid, t, None, "end", 5.0, peak_width=100, out=out)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
