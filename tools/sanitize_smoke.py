"""Small run of every hot kernel (round 2), e.g. under compute-sanitizer where that is available:
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
K1 (store, store+phase from host and device parameters), K1-max (odd/even tile mix: the level-1 race of ADVICE r1), the
one-spectrum search incl. the Newton polish, the device chain (eager call, then the captured graph), K2-ACME and the ROI K2.
(compute-sanitizer is closed on the round-2 GPU pool; the shared-memory protocol of k2_acme.cuh was reviewed barrier by
barrier instead, and the GPU tests compare every path with the oracle.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmris_b200 import chain, pervoxel
from xmris_b200.synth import make_fids_torch
dev = torch.device("cuda:0")
for fam, batch, n, zf, lb in (("1H", 301, 1024, 2048, 5.0), ("1H", 67, 4096, None, 5.0), ("13C", 150, 512, None, 10.0)):
    fid, t = make_fids_torch(fam, batch, n, dev, seed=3)
    for _ in range(3):                      # eager, capture, replay
        spec, freqs, info = chain.chain_single(fid, t, zf, "end", lb, peak_width=100)
    r = pervoxel.chain_all_device(fid[:24], t, zf, "end", lb, peak_width=100)
    r2 = pervoxel.chain_all_device(fid[:24], t, zf, "end", lb, peak_width=100, method="positivity")
    r3 = pervoxel.chain_all_device(fid[:24], t, zf, "end", lb, peak_width=100, p0_only=True)
    torch.cuda.synchronize()
    print(fam, n, zf, "single", round(info["p0"], 3), round(info["p1"], 3), "all", float(r["p0"][0]), float(r["p1"][0]), flush=True)
print("done")
