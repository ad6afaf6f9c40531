"""CPU: the banded LDL^T recurrence that ``csrc/baseline_als.cu`` runs per thread, restated in numpy float64 (vectorised over
spectra, the same operations in the same order), against the oracle's ``scipy.sparse`` solves (``processing/baseline.py:10-39``).
Guards the band formulas of ``D'D`` at the array ends (every ``n >= 3``) and the fused forward / backward sweeps."""

import numpy as np
import pytest

from conftest import load_golden
from oracle import xmris_oracle as orc


def dtd_bands(n):
    """Bands of D'D for the (n-2) x n second-difference operator: rows k = i, i-1, i-2 of D touch column i."""
    i = np.arange(n)
    v0 = (i <= n - 3).astype(float)
    v1 = ((i >= 1) & (i <= n - 2)).astype(float)
    v2 = (i >= 2).astype(float)
    diag = v0 + 4.0 * v1 + v2
    off1 = np.where(i <= n - 2, -2.0 * (v0 + v1), 0.0)      # (i, i+1)
    off2 = v0                                               # (i, i+2)
    return diag, off1, off2


def als_ldlt(y, lam, p, n_iter):
    """``als_kernel`` on the host: per iteration a factorisation fused with the forward substitution, then the backward
    substitution fused with the re-weighting; Q, P, V are what the kernel streams through its HBM scratch."""
    batch, n = y.shape
    dg, o1, o2 = dtd_bands(n)
    w = np.ones((batch, n))
    z = np.zeros((batch, n))
    Q, P, V = (np.zeros((batch, n)) for _ in range(3))
    for _ in range(n_iter):
        d1 = d2 = l1 = l2 = l2n = u1 = u2 = np.zeros(batch)
        for i in range(n):
            d = (w[:, i] + lam * dg[i]) - l1 * l1 * d1 - l2 * l2 * d2
            u = w[:, i] * y[:, i] - l1 * u1 - l2 * u2
            dinv = 1.0 / d
            q = (lam * o1[i] - l2n * l1 * d1) * dinv          # l1_{i+1}
            pp = (lam * o2[i]) * dinv                          # l2_{i+2}
            Q[:, i], P[:, i], V[:, i] = q, pp, u * dinv
            d2, d1, u2, u1 = d1, d, u1, u
            l2, l1, l2n = l2n, q, pp
        z1 = z2 = np.zeros(batch)
        for i in range(n - 1, -1, -1):
            zi = V[:, i] - Q[:, i] * z1 - P[:, i] * z2
            z[:, i] = zi
            z2, z1 = z1, zi
        w = p * (y > z) + (1 - p) * (y < z)                   # baseline.py:37
    return z


def test_bands_equal_the_reference_operator():
    from scipy import sparse

    for n in (3, 4, 5, 6, 17):
        D = sparse.diags([1, -2, 1], [0, 1, 2], shape=(n - 2, n), dtype=float)
        full = (D.T @ D).toarray()
        dg, o1, o2 = dtd_bands(n)
        want = np.diag(dg) + np.diag(o1[: n - 1], 1) + np.diag(o1[: n - 1], -1) + np.diag(o2[: n - 2], 2) + np.diag(o2[: n - 2], -2)
        assert np.array_equal(full, want), n


@pytest.mark.parametrize("n,lam,p,n_iter", [(3, 10.0, 0.1, 3), (4, 1e2, 0.05, 5), (5, 1e3, 0.01, 5), (6, 1e5, 0.001, 10),
                                           (50, 1e3, 0.02, 6), (257, 1e5, 0.01, 10)])
def test_recurrence_matches_sparse_solves(n, lam, p, n_iter):
    rng = np.random.default_rng(n)
    y = rng.standard_normal((7, n)).cumsum(axis=1) + 3.0 * rng.random((7, n))
    want = np.stack([orc.als_core(row, lam, p, n_iter) for row in y])
    got = als_ldlt(y, lam, p, n_iter)
    assert np.linalg.norm(got - want) <= 1e-8 * max(np.linalg.norm(want), 1.0)


def test_recurrence_matches_reference_golden():
    g = load_golden("baseline")
    spec = g["spec"].reshape(-1, 1024).real
    for tag, kw in [("a", dict(lam=1e5, p=0.01, n_iter=10)), ("c", dict(lam=1e7, p=0.05, n_iter=4))]:
        corrected = spec - als_ldlt(spec, **kw)
        ref = g[f"corr_{tag}"].reshape(-1, 1024)
        assert np.linalg.norm(corrected - ref) <= 1e-7 * np.linalg.norm(ref)
