"""Angle parity of BOTH autophase searches on 1024 seeded spectra per BASELINE shape (C2 / C3 / C4 / C5), as a measured fact.

north_star: "phi0/phi1 within 0.1 deg".  Fixtures (tests/golden/angle_parity_{ref,tight}.npz, made by tools/angle_parity.py):
  ref    the reference's optimiser call verbatim (differential_evolution best1bin, tol=0.01, seed=42 + L-BFGS-B polish,
         phasing.py:276-284) on every spectrum -- through the oracle, a bit-exact restatement of the reference;
  tight  the SAME optimiser with tol=1e-6, popsize=60, best of seeds 1..5: the adjudicator.  The reference's answer is not
         always converged (piecewise-smooth objective): where the tight run lands on the GPU's angles with an
         objective <= the reference's, the mismatch is the reference's non-convergence and counts as a match.
Floors below are set from profiles/parity_r2.json (measured on a B200), a little under the measured values.
"""

import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_l2
from oracle import xmris_oracle as orc

sys.path.insert(0, os.path.join(ROOT, "tools"))
import angle_parity as ap  # noqa: E402

pytestmark = pytest.mark.gpu

# shape -> (floor of match % incl. adjudicated [single, all], ceiling of worse % [single, all])
FLOORS = {
    "C2_2048": ((98.5, 98.0), (1.5, 2.0)),
    "C3_4096_zf8192": ((98.5, 98.5), (1.5, 1.5)),
    "C4_13C_1024": ((99.0, 99.5), (1.0, 0.5)),
    "C5_4096": ((98.5, 98.5), (1.5, 1.5)),
}
N = 1024


@pytest.mark.parametrize("shape", ap.SHAPES, ids=[s[0] for s in ap.SHAPES])
def test_angles_match_reference_or_its_converged_optimum(shape):
    import torch

    from xmris_b200 import chain, device as D, pervoxel

    name, fam, n_in, zf, lb, seed = shape
    ref = np.load(os.path.join(ap.GOLD, "angle_parity_ref.npz"))[name][:N]
    tight = np.load(os.path.join(ap.GOLD, "angle_parity_tight.npz"))
    tb, nseeds = ap.tight_best_of(tight, name, N)
    assert nseeds == len(ap.TIGHT_SEEDS)
    fid, t, spec, freqs = ap.spectra(shape, N)
    x = torch.from_numpy(fid).cuda()

    # ---- per-voxel kernel (mode="all") -------------------------------------------------------------------------------
    out_t, _, info = pervoxel.chain_all(x, t, zf, "end", lb, peak_width=100)
    got = out_t.cpu().numpy()
    assert np.array_equal(info["pivot"], ref[:, 2])                       # each voxel pivots on its own |S| maximum
    for i in range(0, N, 64):                                             # spectra: the reference's phase() at the GPU's angles
        same, _ = orc.phase(spec[i], 0, freqs, info["p0"][i], info["p1"][i], info["pivot"][i])
        assert rel_l2(got[i], same) < 1e-5, (i, rel_l2(got[i], same))
    # ---- one-spectrum search (mode="single" with each spectrum as its own problem) --------------------------------------
    spec_t, _, _ = chain.chain_to_spectrum(x, t, zf, "end", lb)
    _, argmax = D.row_absmax(spec_t)
    argmax = argmax.cpu().numpy()
    single = np.zeros((N, 2))
    for i in range(N):
        idx = int(argmax[i])
        assert freqs[idx] == ref[i, 2]
        _, _, u0, du = chain.phase_turns(freqs, 0.0, 0.0, float(freqs[idx]))
        single[i] = D.autophase_search(spec_t[i].contiguous(), u0, du, "acme", idx, 1, False).cpu().numpy()[:2]

    for k, (kern, p0, p1) in enumerate((("single", single[:, 0], single[:, 1]), ("all", info["p0"], info["p1"]))):
        f_gpu = np.array([orc.acme_score([p0[i], p1[i]], spec[i], freqs, ref[i, 2]) for i in range(N)])
        code = ap.classify(np.asarray(p0), np.asarray(p1), f_gpu, ref, tb)
        well = int((code != 4).sum())
        match = 100.0 * ((code == 0) | (code == 1)).sum() / well
        worse = 100.0 * (code == 3).sum() / well
        print(f"{name} {kern}: within 0.1 deg of the reference {100.0 * (code == 0).sum() / well:.2f} %, reference unconverged "
              f"{100.0 * (code == 1).sum() / well:.2f} %, better {100.0 * (code == 2).sum() / well:.2f} %, worse {worse:.2f} %")
        assert match >= FLOORS[name][0][k], (name, kern, match)
        assert worse <= FLOORS[name][1][k], (name, kern, worse)
