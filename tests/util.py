"""Shared helpers for the parity tests (golden fixtures + oracle models)."""
import os

import numpy as np

from oracle.lopq_oracle import OracleModel
from tests.golden.make_golden import CASES, case_inputs  # noqa: F401  (re-exported)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    return z, OracleModel.from_npz(z)


def searches(z):
    for si in range(int(z["n_searches"])):
        limit = int(z["s%d_limit" % si])
        yield si, int(z["s%d_quota" % si]), (None if limit < 0 else limit)


def random_model_params(D, V, M, K, seed, coarse_f32=False):
    """Untrained but well-formed LOPQ parameters (random centroids, random orthogonal local
    rotations): parity tests only need *a* model, and these need no k-means."""
    rng = np.random.RandomState(seed)
    h, ds, m = D // 2, D // M, M // 2
    Cs = tuple((rng.randn(V, h) * 0.5).astype(np.float32 if coarse_f32 else np.float64) for _ in range(2))
    Rs = tuple(np.stack([np.linalg.qr(rng.randn(h, h))[0] for _ in range(V)]) for _ in range(2))
    mus = tuple(rng.randn(V, h) * 0.05 for _ in range(2))
    subs = tuple([rng.randn(K, ds) * 0.3 for _ in range(m)] for _ in range(2))
    return Cs, Rs, mus, subs


def random_data(params, n, seed, dtype=np.float32, dup_frac=0.02):
    """Points scattered around the coarse centroids, with a few exact duplicates (=> exact ties)."""
    Cs = params[0]
    rng = np.random.RandomState(seed)
    V = Cs[0].shape[0]
    X = np.concatenate([Cs[s][rng.randint(0, V, size=n)] + 0.35 * rng.randn(n, Cs[s].shape[1]) for s in (0, 1)], axis=1)
    k = int(n * dup_frac)
    if k:
        X[rng.randint(0, n, size=k)] = X[rng.randint(0, n, size=k)]
    return X.astype(dtype)
