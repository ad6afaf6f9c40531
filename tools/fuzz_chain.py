"""Randomised equivalence sweep of the mode="single" chain's code paths: device-resident one-call chain == host-buffer chain
(FIDs resident between the passes) == host-buffer chain streaming the chunks twice, bit for bit (to an ulp at 8192 points), over random lengths,
zero-fill factors, batch and chunk sizes (`python tools/fuzz_chain.py [cases] [seed]`; development stress run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import chain, hostabi
from xmris_b200.synth import make_fids_torch

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
for c in range(cases):
    n_out = int(2 ** rng.integers(6, 14))
    zf = int(rng.choice([1, 1, 2, 4]))
    n_in = max(n_out // zf, 16)
    batch = int(rng.choice([1, 5, 149, 297, int(rng.integers(2, 3000))]))
    chunk = int(rng.choice([0, 64, 257, 1000]))
    fam = "1H" if rng.random() < 0.7 else "13C"
    fid, t = make_fids_torch(fam, batch, n_in, dev, seed=int(rng.integers(1, 1 << 30)))
    tp = None if n_out == n_in else n_out
    ref, rfreqs, rinfo = chain.chain_single(fid, t, tp, "end", 5.0, peak_width=100)
    host = fid.cpu().numpy()
    ap = dict(mode="single", peak_width=100)
    outs = []
    for limit in (0, 1):
        hostabi.set_resident_limit(limit)
        o, f, info = hostabi.chain_host(host, t, tp, "end", 5.0, autophase=ap, chunk=chunk)
        outs.append((o, info))
    hostabi.set_resident_limit(0)
    hostabi.release_workspace()
    r = ref.cpu().numpy()
    # (N = 8192: the twiddles come from float32 power chains which the compiler contracts per kernel instantiation -- the
    #  host-parameter and device-parameter variants of pass 2 agree to an ulp, not bit for bit; up to 4096 points: identical)
    same = (lambda o: np.array_equal(o, r)) if n_out < 8192 else (lambda o: np.abs(o - r).max() <= 3e-7 * np.abs(r).max())
    ok = all(same(o) and (i["p0"], i["p1"], i["pivot"]) == (rinfo["p0"], rinfo["p1"], rinfo["pivot"]) for o, i in outs)
    if not ok or c % 10 == 0:
        print(f"{c:3d} {fam} n_in={n_in} n_out={n_out} batch={batch} chunk={chunk}: p0={rinfo['p0']:.3f} p1={rinfo['p1']:.3f} {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        sys.exit(1)
print(f"{cases} cases ok")
