"""Randomised sweep of K1 / K1-max geometries against numpy (float64): every transform length, zero-fill position, window
kind, batch size around the residency boundaries, with statistics, phase and the branch-and-bound pass -- a development
stress run for the GPU box (`python tools/fuzz_k1.py [cases] [seed]`), not part of the test suite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmris_b200 import device as D

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
worst = 0.0
for c in range(cases):
    n_out = int(2 ** rng.integers(4, 14))
    kind = rng.integers(0, 4)
    n_in = n_out if kind == 0 else (n_out // 2 if kind == 1 else (n_out // 4 if kind == 2 and n_out >= 64 else int(rng.integers(1, n_out + 1))))
    pad_left = 0 if rng.random() < 0.7 else int(rng.integers(0, n_out - n_in + 1))
    batch = int(rng.choice([1, 2, 3, 147, 148, 149, 295, 296, 297, 300, 593, int(rng.integers(1, 700))]))
    x = (rng.standard_normal((batch, n_in)) + 1j * rng.standard_normal((batch, n_in))) * np.exp(-np.arange(n_in) / max(n_in / 3.0, 1.0))
    x *= rng.uniform(0.2, 5.0, size=(batch, 1))
    t = (np.arange(n_out) - pad_left) / 4000.0
    wk = rng.integers(0, 3)
    w = None if wk == 0 else np.exp(-np.pi * 4.0 * np.abs(t)) / np.sqrt(n_out)
    if wk == 2:
        w = w * (1.0 + 0.25 * np.cos(0.31 * np.arange(n_out)))       # does not factor: table window
    padded = np.zeros((batch, n_out), dtype=np.complex128)
    padded[:, pad_left:pad_left + n_in] = x
    ww = np.full(n_out, 1.0 / np.sqrt(n_out)) if w is None else w
    ref = np.fft.fftshift(np.fft.fft(padded * ww, axis=1), axes=1)
    xd = torch.from_numpy(x.astype(np.complex64)).to(dev)
    spec, amax, imax = D.fid_to_spectrum(xd, n_out=n_out, pad_left=pad_left, window=w, want_stats=True)
    got = spec.cpu().numpy()
    den = np.maximum(np.linalg.norm(ref, axis=1), 1e-30)
    err = float(np.max(np.linalg.norm(got - ref, axis=1) / den))
    am = np.abs(ref).max(axis=1)
    e2 = float(np.max(np.abs(amax.cpu().numpy() - am) / np.maximum(am, 1e-30)))
    a, b = float(rng.uniform(-1, 1)), float(rng.uniform(-2e-3, 2e-3))
    specp, _, _ = D.fid_to_spectrum(xd, n_out=n_out, pad_left=pad_left, window=w, phase_turns=(a, b))
    refp = ref * np.exp(2j * np.pi * (a + b * np.arange(n_out)))[None, :]
    e3 = float(np.max(np.linalg.norm(specp.cpu().numpy() - refp, axis=1) / den))
    pruned, running = D.fid_absmax_pruned(xd, n_out=n_out, pad_left=pad_left, window=w)
    pr = pruned.cpu().numpy()
    win = int(np.argmax(am))
    ok_win = int(np.argmax(pr)) == int(np.argmax(amax.cpu().numpy())) and abs(pr.max() - am[win]) <= 3e-5 * am[win]
    worst = max(worst, err, e3)
    flag = "" if (err < 1e-5 and e2 < 3e-5 and e3 < 1e-5 and ok_win) else "   <-- FAIL"
    if flag or c % 25 == 0:
        print(f"{c:4d} n_in={n_in:5d} n_out={n_out:5d} pad_left={pad_left:5d} batch={batch:4d} window={wk}: fft {err:.2e} max {e2:.2e} phase {e3:.2e} winner {ok_win}{flag}", flush=True)
    if flag:
        sys.exit(1)
print(f"{cases} cases ok, worst rel. L2 {worst:.2e}")
