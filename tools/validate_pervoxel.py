"""Quality of the per-voxel search (K2) against the reference's per-spectrum DE on larger samples than the tests use.

    python tools/validate_pervoxel.py [nvox]      (on the GPU box; the oracle runs on all host cores)
"""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import xmris_oracle as orc


def _ref_one(args):
    spec, freqs = args
    _, info = orc.autophase(spec, 0, freqs, peak_width=100)
    return info["p0"], info["p1"], info["pivot"], info["fun"]


def main():
    import torch
    from xmris_b200 import pervoxel
    from xmris_b200.synth import make_fids_numpy

    nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfgs = [("C2_2048", "1H", 2048, None, 5.0), ("C3_4096_zf8192", "1H", 4096, 8192, 5.0), ("C4_13C_1024", "13C", 1024, None, 10.0),
            ("C5_4096", "1H", 4096, None, 5.0)]
    workers = len(os.sched_getaffinity(0))
    for name, fam, n_in, zf, lb in cfgs:
        fid, t, _ = make_fids_numpy(fam, nvox, n_in, seed=900 + n_in)
        fid = fid.astype(np.complex64)
        ref_spec, freqs = orc.chain_to_spectrum(fid.astype(np.complex128), 1, t, zf, "end", lb)
        t0 = time.time()
        with mp.get_context("fork").Pool(workers) as pool:
            ref = np.array(pool.map(_ref_one, [(ref_spec[i], freqs) for i in range(nvox)]))
        t_ref = time.time() - t0
        x = torch.from_numpy(fid).cuda()
        pervoxel.chain_all(x, t, zf, "end", lb, peak_width=100)
        torch.cuda.synchronize()
        t0 = time.time()
        _, _, info = pervoxel.chain_all(x, t, zf, "end", lb, peak_width=100)
        t_gpu = time.time() - t0
        match = better = worse = ill = 0
        rel = []
        for i in range(nvox):
            if ref[i, 3] < 0:
                ill += 1
                continue
            f = orc.acme_score([info["p0"][i], info["p1"][i]], ref_spec[i], freqs, info["pivot"][i])
            rel.append((f - ref[i, 3]) / abs(ref[i, 3]))
            if abs(info["p0"][i] - ref[i, 0]) <= 0.1 and abs(info["p1"][i] - ref[i, 1]) <= 0.1:
                match += 1
            elif f <= ref[i, 3] * (1 + 1e-5):
                better += 1
            else:
                worse += 1
        rel = np.array(rel)
        if os.environ.get("XMR_DUMP"):
            os.makedirs("gpurun_out", exist_ok=True)
            np.savez_compressed(f"gpurun_out/pervoxel_{name}.npz", ref=ref, p0=info["p0"], p1=info["p1"], pivot=info["pivot"],
                                fun=info["fun"], seed=900 + n_in)
        print(f"{name:16s} n={nvox} match {match} better-or-equal {better} worse {worse} ill {ill} | worst rel excess "
              f"{rel.max():.2e} | oracle {t_ref:.1f}s on {workers} cores, gpu {t_gpu*1e3:.0f} ms", flush=True)


if __name__ == "__main__":
    main()
