"""Shared helpers of the test-suite: small synthetic inputs and the parity metrics named in BASELINE.json."""
import numpy as np


def small_matrix(n, m, kt, seed, density_scale=0.5):
    """noisy low-rank nonnegative matrix, sparse-ish like normalised log counts (genes x cells)."""
    rng = np.random.default_rng(seed)
    Wt = rng.gamma(0.3, 1.0, size=(n, kt))
    Ht = rng.gamma(0.3, 1.0, size=(kt, m))
    lam = density_scale * (Wt @ Ht)
    X = rng.poisson(rng.gamma(2.0, lam / 2.0))
    cs = np.maximum(X.sum(axis=0), 1)
    return np.asfortranarray(np.log1p(X / cs[None, :] * np.median(cs)))


def seeded_init(A, k, seed):
    """NNDSVD + zero replacement, computed by the oracle so both sides start identically."""
    from oracle import nnsvd_init_ref, zero_replace_ref
    rng = np.random.default_rng(seed)
    W0, H0 = nnsvd_init_ref(A, k)
    return zero_replace_ref(W0, H0, A.mean(), rng)


def rel_fro(X, Y):
    return float(np.linalg.norm(np.asarray(X) - np.asarray(Y)) / max(np.linalg.norm(np.asarray(Y)), 1e-300))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


MEASURED = []   # (test id, quantity, measured value, bar) of every recorded parity check; conftest writes them out


def within(name, value, bar):
    """assert value < bar, and keep the measured value next to its bar (-> gpurun_out/parity_measured.json)"""
    import os
    value = float(value)
    MEASURED.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": name, "measured": value, "bar": float(bar)})
    assert value < bar, "%s: measured %.3e, bar %.3e" % (name, value, bar)
