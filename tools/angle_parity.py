"""Angle parity of the autophase searches against the reference's optimiser on >= 1024 seeded spectra per shape.

north_star: "phi0/phi1 within 0.1 deg".  The reference's answer is ``differential_evolution(..., tol=0.01, seed=42)`` + an
L-BFGS-B polish on a piecewise-smooth objective (phasing.py:276-284): it is not always converged.  This
tool makes the comparison a committed, adjudicated fact ("next round" item 1):

    python tools/angle_parity.py ref    [n]   CPU: the oracle's reference DE per spectrum       -> tests/golden/angle_parity_ref.npz
    python tools/angle_parity.py tight  [n]   CPU: DE with tol=1e-6, popsize=60, seeds 1..5      -> tests/golden/angle_parity_tight.npz
    python tools/angle_parity.py gpu          GPU box: single search + per-voxel kernel on the same spectra -> gpurun_out/angle_parity_gpu.npz
    python tools/angle_parity.py report       combine -> profiles/parity_r2.json (+ a markdown table on stdout)

Every spectrum is its own 1-D problem (the reference's mode="single" on a 1-D input; pivot = its own |S| maximum).
Classification per spectrum and kernel (well-posed = the reference's own objective value is positive, SURVEY finding 5):
    match          |dp0| (mod 360) <= 0.1 deg and |dp1| <= 0.1 deg against the reference's answer
    ref_unconverged the GPU's angles are within 0.1 deg of the tight DE's best answer AND that answer's objective is <= the
                   reference's: the reference stopped short of (or in another basin than) the minimum the tighter run of ITS OWN
                   optimiser converges to -- counted as a match
    better         elsewhere, objective (float64, oracle) <= the reference's and <= the tight optimum's (nobody found it)
    worse          everything else
The oracle is the checker only (tests / tools); nothing here is imported by the product.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

SHAPES = [  # name, family, n_in, zero-fill target, lb, data seed
    ("C2_2048", "1H", 2048, None, 5.0, 42021),
    ("C3_4096_zf8192", "1H", 4096, 8192, 5.0, 42022),
    ("C4_13C_1024", "13C", 1024, None, 10.0, 42023),
    ("C5_4096", "1H", 4096, None, 5.0, 42024),
]
GOLD = os.path.join(ROOT, "tests", "golden")
TIGHT_SEEDS = (1, 2, 3, 4, 5)
N_SET = 1024           # spectra per shape in the committed fixtures
TOL = 0.1


def spectra(shape, n):
    """The seeded problem set: complex64-rounded FIDs and the oracle's float64 spectra of them."""
    from oracle import xmris_oracle as orc
    from xmris_b200.synth import make_fids_numpy

    name, fam, n_in, zf, lb, seed = shape
    assert n <= N_SET
    fid, t, _ = make_fids_numpy(fam, N_SET, n_in, seed=seed)     # the generator's draws depend on the batch size: always the
    fid = fid[:n].astype(np.complex64)                             # full set, then the first n
    spec, freqs = orc.chain_to_spectrum(fid.astype(np.complex128), 1, t, zf, "end", lb)
    return fid, t, spec, freqs


def _ref_one(args):
    from oracle import xmris_oracle as orc

    spec, freqs = args
    _, info = orc.autophase(spec, 0, freqs, peak_width=100)
    return info["p0"], info["p1"], info["pivot"], info["fun"], info["nfev"], info["target_idx"]


def _fast_acme(ph, S, u):
    """The reference's ACME objective (phasing.py:100-122) without the per-call coordinate bookkeeping (same arithmetic)."""
    d = (S * np.exp(1j * (np.radians(ph[0]) + np.radians(ph[1]) * u))).real
    g = np.abs((d[1:] - d[:-1]) / 2)
    q = g / g.sum()
    q[q == 0] = 1
    h = -(q * np.log(q)).sum()
    a = d - np.abs(d)
    pf = ((a / 2) ** 2).sum() if a.sum() < 0 else 0.0
    return (h + 1000 * pf) / d.shape[-1] / d.max()


def _tight_one(args):
    import scipy.optimize as so

    spec, freqs, pivot, seed = args
    u = (freqs - pivot) / (freqs.max() - freqs.min())
    r = so.differential_evolution(_fast_acme, bounds=[(-180.0, 180.0), (-4000.0, 4000.0)], args=(spec, u), strategy="best1bin",
                                  tol=1e-6, popsize=60, seed=seed)
    return r.x[0], r.x[1], r.fun, r.nfev


def cmd_ref(n, workers):
    out = {}
    for shape in SHAPES:
        t0 = time.time()
        _, _, spec, freqs = spectra(shape, n)
        with mp.get_context("fork").Pool(workers) as pool:
            ref = np.array(pool.map(_ref_one, [(spec[i], freqs) for i in range(n)], chunksize=4))
        out[shape[0]] = ref
        print(f"{shape[0]}: {n} reference DE runs in {time.time() - t0:.0f} s on {workers} workers, median nfev {np.median(ref[:, 4]):.0f}, "
              f"ill-posed (fun < 0): {(ref[:, 3] < 0).sum()}", flush=True)
        np.savez_compressed(os.path.join(GOLD, "angle_parity_ref.npz"), **out)


def cmd_tight(n, workers):
    ref = np.load(os.path.join(GOLD, "angle_parity_ref.npz"))
    path = os.path.join(GOLD, "angle_parity_tight.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for seed in TIGHT_SEEDS:
        for shape in SHAPES:
            key = f"{shape[0]}_seed{seed}"
            if key in out and len(out[key]) >= n:
                continue
            t0 = time.time()
            _, _, spec, freqs = spectra(shape, n)
            piv = ref[shape[0]][:n, 2]
            with mp.get_context("fork").Pool(workers) as pool:
                res = np.array(pool.map(_tight_one, [(spec[i], freqs, piv[i], seed) for i in range(n)], chunksize=4))
            out[key] = res
            np.savez_compressed(path, **out)
            print(f"{key}: {n} tight DE runs in {time.time() - t0:.0f} s, median nfev {np.median(res[:, 3]):.0f}", flush=True)


def cmd_gpu(n):
    import torch

    from xmris_b200 import chain, device as D, pervoxel

    ref = np.load(os.path.join(GOLD, "angle_parity_ref.npz"))
    out = {}
    if os.environ.get("XMR_POLISH"):           # experiment knob: "starts,f32_levels,fine_b"
        from xmris_b200 import _lib

        _lib.check(_lib.load().xmr_autophase_search_polish(*[int(v) for v in os.environ["XMR_POLISH"].split(",")]))
    if os.environ.get("XMR_TUNING"):           # experiment knob: "p0_step,p1_step,starts,levels,f32_levels,late_starts,first_ratio"
        from xmris_b200 import _lib

        v = os.environ["XMR_TUNING"].split(",")
        _lib.check(_lib.load().xmr_autophase_search_tuning(float(v[0]), float(v[1]), int(v[2]), int(v[3]), int(v[4]), int(v[5]), float(v[6])))
    for shape in SHAPES:
        name, fam, n_in, zf, lb, seed = shape
        m = min(n, len(ref[name]))
        fid, t, spec, freqs = spectra(shape, m)
        x = torch.from_numpy(fid).cuda()
        # per-voxel kernel: FID -> phased spectrum + angles
        pervoxel.chain_all(x[:8], t, zf, "end", lb, peak_width=100)
        torch.cuda.synchronize()
        t0 = time.time()
        _, _, info = pervoxel.chain_all(x, t, zf, "end", lb, peak_width=100)
        torch.cuda.synchronize()
        t_all = time.time() - t0
        # one-spectrum search (mode="single" on each spectrum as its own problem)
        spec_t, _, geo = chain.chain_to_spectrum(x, t, zf, "end", lb)
        _, argmax = D.row_absmax(spec_t)
        argmax = argmax.cpu().numpy()
        single = np.zeros((m, 3))
        ms = []
        for i in range(m):
            idx = int(argmax[i])
            _, _, u0, du = chain.phase_turns(freqs, 0.0, 0.0, float(freqs[idx]))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = D.autophase_search(spec_t[i].contiguous(), u0, du, "acme", idx, 1, False)
            e1.record()
            single[i] = r.cpu().numpy()[:3]
            ms.append(e0.elapsed_time(e1))
        out[name + "_single"] = single
        out[name + "_single_pivot"] = argmax[:m].astype(np.int64)
        out[name + "_all"] = np.stack([info["p0"], info["p1"], info["fun"]], axis=1)
        out[name + "_all_pivot"] = np.asarray(info["pivot"], dtype=np.float64)
        out[name + "_timing"] = np.array([np.median(ms), t_all * 1e3])
        print(f"{name}: {m} spectra, single search median {np.median(ms):.3f} ms, per-voxel kernel {t_all * 1e3:.1f} ms total", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "angle_parity_gpu.npz"), **out)


def _wrap(d):
    return np.abs(((d + 180.0) % 360.0) - 180.0)


def classify(p0, p1, f_gpu, ref, tight_best):
    """Per-spectrum class codes: 0 match, 1 ref_unconverged, 2 better, 3 worse, 4 ill-posed (reference fun < 0)."""
    n = len(p0)
    code = np.full(n, 3)
    near_ref = (_wrap(p0 - ref[:, 0]) <= TOL) & (np.abs(p1 - ref[:, 1]) <= TOL)
    near_tight = (_wrap(p0 - tight_best[:, 0]) <= TOL) & (np.abs(p1 - tight_best[:, 1]) <= TOL)
    tight_ok = tight_best[:, 2] <= ref[:, 3] * (1 + 1e-9)
    code[(f_gpu <= ref[:, 3] * (1 + 1e-9)) & (f_gpu <= tight_best[:, 2] * (1 + 1e-9))] = 2
    code[near_tight & tight_ok] = 1
    code[near_ref] = 0
    code[ref[:, 3] < 0] = 4
    return code


def tight_best_of(tight, name, n):
    runs = [tight[f"{name}_seed{s}"][:n] for s in TIGHT_SEEDS if f"{name}_seed{s}" in tight]
    if not runs:
        return None, 0
    runs = np.stack(runs)                       # (seeds, n, 4)
    fvals = np.where(runs[:, :, 2] > 0, runs[:, :, 2], np.inf)      # a run that dived into the ACME pole (f < 0) is no adjudicator
    k = np.argmin(fvals, axis=0)
    return runs[k, np.arange(runs.shape[1])], len(runs)


def cmd_report(gpu_path=None):
    from oracle import xmris_oracle as orc

    ref = np.load(os.path.join(GOLD, "angle_parity_ref.npz"))
    tight = np.load(os.path.join(GOLD, "angle_parity_tight.npz"))
    gpu = np.load(gpu_path or os.path.join(ROOT, "gpurun_out", "angle_parity_gpu.npz"))
    report = {"tolerance_deg": TOL, "reference": "scipy differential_evolution(best1bin, tol=0.01, seed=42) + L-BFGS-B polish (phasing.py:276-284)",
              "adjudicator": f"the same optimiser with tol=1e-6, popsize=60, best of seeds {list(TIGHT_SEEDS)}", "shapes": {}}
    rows = []
    for shape in SHAPES:
        name = shape[0]
        n = len(gpu[name + "_single"])
        _, _, spec, freqs = spectra(shape, n)
        r = ref[name][:n]
        tb, nseeds = tight_best_of(tight, name, n)
        entry = {"n": int(n), "tight_seeds": nseeds, "ill_posed": int((r[:, 3] < 0).sum())}
        for kern in ("single", "all"):
            g = gpu[f"{name}_{kern}"]
            piv = r[:, 2]
            f_gpu = np.array([orc.acme_score([g[i, 0], g[i, 1]], spec[i], freqs, piv[i]) for i in range(n)])
            code = classify(g[:, 0], g[:, 1], f_gpu, r, tb)
            ok = code != 4
            tot = int(ok.sum())
            d0, d1 = _wrap(g[:, 0] - r[:, 0])[ok], np.abs(g[:, 1] - r[:, 1])[ok]
            excess = ((f_gpu - np.minimum(r[:, 3], tb[:, 2])) / np.abs(np.minimum(r[:, 3], tb[:, 2])))[code == 3]
            entry[kern] = {
                "within_0.1deg_of_reference_pct": round(100.0 * (code == 0).sum() / tot, 2),
                "reference_unconverged_pct": round(100.0 * (code == 1).sum() / tot, 2),
                "better_objective_pct": round(100.0 * (code == 2).sum() / tot, 2),
                "worse_pct": round(100.0 * (code == 3).sum() / tot, 2),
                "match_incl_adjudicated_pct": round(100.0 * ((code == 0) | (code == 1)).sum() / tot, 2),
                "worst_rel_excess": float(excess.max()) if len(excess) else 0.0,
                "dp0_deg_p50_p90_p99": [round(float(np.percentile(d0, q)), 4) for q in (50, 90, 99)],
                "dp1_deg_p50_p90_p99": [round(float(np.percentile(d1, q)), 4) for q in (50, 90, 99)],
            }
            rows.append((name, kern, entry[kern]))
        entry["timing_ms"] = {"single_search_median": float(gpu[name + "_timing"][0]), "per_voxel_total": float(gpu[name + "_timing"][1])}
        report["shapes"][name] = entry
    with open(os.path.join(ROOT, "profiles", "parity_r2.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("| shape | kernel | within 0.1 deg of reference | + reference unconverged (tight DE agrees with GPU) | = match | better objective | worse |")
    print("|---|---|---:|---:|---:|---:|---:|")
    for name, kern, e in rows:
        print(f"| {name} | {kern} | {e['within_0.1deg_of_reference_pct']} % | {e['reference_unconverged_pct']} % | "
              f"**{e['match_incl_adjudicated_pct']} %** | {e['better_objective_pct']} % | {e['worse_pct']} % |")


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "report"
    n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1024
    workers = int(os.environ.get("XMR_WORKERS", max(1, len(os.sched_getaffinity(0)) - 2)))
    if cmd == "ref":
        cmd_ref(n, workers)
    elif cmd == "tight":
        cmd_tight(n, workers)
    elif cmd == "gpu":
        cmd_gpu(n)
    else:
        cmd_report(sys.argv[2] if len(sys.argv) > 2 else None)
