"""Quality / cost of the one-spectrum search for several search geometries (xmr_autophase_search_tuning).

Reference answers: the oracle's differential evolution (one run per spectrum, all host cores).  Per geometry: how many
spectra land within 0.1 deg, how many elsewhere with an equal-or-better objective, how many worse, and the search time.
usage: python tools/tune_single_search.py [nvox] ["p0step,p1step,starts,levels,f32_levels,late_starts,first_ratio" ...]
"""
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
    from xmris_b200 import _lib, device as D, chain
    from xmris_b200.synth import make_fids_numpy
    nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfgs = [tuple(float(x) for x in a.split(",")) for a in sys.argv[2:]] or [(6, 15, 4, 4, 2, 2, 2.5)]
    lib = _lib.load()
    for name, fam, n_in, zf, lb in [("C2_2048", "1H", 2048, None, 5.0), ("C4_13C_1024", "13C", 1024, None, 10.0),
                                    ("C5_4096", "1H", 4096, None, 5.0)]:
        fid, t, _ = make_fids_numpy(fam, nvox, n_in, seed=700 + n_in)
        fid = fid.astype(np.complex64)
        ref_spec, freqs = orc.chain_to_spectrum(fid.astype(np.complex128), 1, t, zf, "end", lb)
        with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
            ref = np.array(pool.map(_ref_one, [(ref_spec[i], freqs) for i in range(nvox)]))
        spec_t, _, geo = chain.chain_to_spectrum(torch.from_numpy(fid).cuda(), t, zf, "end", lb)
        for cfg in cfgs:
            _lib.check(lib.xmr_autophase_search_tuning(cfg[0], cfg[1], *[int(c) for c in cfg[2:6]], cfg[6]))
            match = better = worse = 0
            worst = 0.0
            ms = []
            for i in range(nvox):
                if ref[i, 3] < 0:
                    continue
                idx = int(np.argmax(np.abs(ref_spec[i])))
                _, _, u0, du = chain.phase_turns(freqs, 0.0, 0.0, float(freqs[idx]))
                row = spec_t[i].contiguous()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = D.autophase_search(row, u0, du, "acme", idx, 1, False)
                e1.record()
                r = r.cpu().numpy()
                ms.append(e0.elapsed_time(e1))
                f = orc.acme_score([r[0], r[1]], ref_spec[i], freqs, float(freqs[idx]))
                if abs(r[0] - ref[i, 0]) <= 0.1 and abs(r[1] - ref[i, 1]) <= 0.1:
                    match += 1
                elif f <= ref[i, 3] * (1 + 1e-5):
                    better += 1
                else:
                    worse += 1
                    worst = max(worst, (f - ref[i, 3]) / abs(ref[i, 3]))
            print(f"{name} cfg={cfg}: n={nvox} match {match} better-or-equal {better} worse {worse} "
                  f"(worst rel excess {worst:.2e})  search {np.median(ms):.3f} ms", flush=True)


if __name__ == "__main__":
    main()
