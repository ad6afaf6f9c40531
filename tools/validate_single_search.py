"""Robustness of the one-spectrum search (mode="single" optimiser) on many spectra, each treated as its own problem."""
import multiprocessing as mp, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xmris_oracle as orc

def _ref_one(args):
    spec, freqs = args
    _, info = orc.autophase(spec, 0, freqs, peak_width=100)
    return info["p0"], info["p1"], info["pivot"], info["fun"]

def main():
    import torch
    from xmris_b200 import device as D, chain
    from xmris_b200.synth import make_fids_numpy
    nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for name, fam, n_in, zf, lb in [("C2_2048", "1H", 2048, None, 5.0), ("C4_13C_1024", "13C", 1024, None, 10.0), ("C5_4096", "1H", 4096, None, 5.0)]:
        fid, t, _ = make_fids_numpy(fam, nvox, n_in, seed=700 + n_in)
        fid = fid.astype(np.complex64)
        ref_spec, freqs = orc.chain_to_spectrum(fid.astype(np.complex128), 1, t, zf, "end", lb)
        with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
            ref = np.array(pool.map(_ref_one, [(ref_spec[i], freqs) for i in range(nvox)]))
        spec_t, _, geo = chain.chain_to_spectrum(torch.from_numpy(fid).cuda(), t, zf, "end", lb)
        match = better = worse = 0
        worst = 0.0
        for i in range(nvox):
            if ref[i, 3] < 0:
                continue
            idx = int(np.argmax(np.abs(ref_spec[i])))
            _, _, u0, du = chain.phase_turns(freqs, 0.0, 0.0, float(freqs[idx]))
            r = D.autophase_search(spec_t[i].contiguous(), u0, du, "acme", idx, 1, False).cpu().numpy()
            f = orc.acme_score([r[0], r[1]], ref_spec[i], freqs, float(freqs[idx]))
            if abs(r[0] - ref[i, 0]) <= 0.1 and abs(r[1] - ref[i, 1]) <= 0.1:
                match += 1
            elif f <= ref[i, 3] * (1 + 1e-5):
                better += 1
            else:
                worse += 1
                worst = max(worst, (f - ref[i, 3]) / abs(ref[i, 3]))
        print(f"{name}: n={nvox} match {match} better-or-equal {better} worse {worse} (worst rel excess {worst:.2e})", flush=True)

if __name__ == "__main__":
    main()
