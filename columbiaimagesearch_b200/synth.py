"""Seeded synthetic inputs for the LOPQ hot path (SURVEY.md section 8d / BASELINE.md section 3).

Two families:

* ``lattice_gmm``  -- parity fixtures.  Every value is (integer / 1024) so the float32 arrays are
  bit-identical on any CPU (no normalisation, no libm), and only the seed has to be committed.
* ``dlib_style`` / ``sentibank_style`` -- the BASELINE recipe: 4096-centre Gaussian mixture,
  sigma 0.35, L2-normalised float32 (2048-d variant ReLU-clamped); near-duplicate queries
  ``normalise(db[i] + rho * u)``.
"""
import numpy as np


def lattice_gmm(n, D, seed, centres=64, spread=300, noise=90, dtype=np.float32, dup_frac=0.0):
    """n x D vectors on the 1/1024 lattice: integer centre + integer noise, exact in float32."""
    rng = np.random.RandomState(seed)
    C = rng.randint(-spread, spread + 1, size=(centres, D))
    a = rng.randint(0, centres, size=n)
    X = C[a] + rng.randint(-noise, noise + 1, size=(n, D))
    if dup_frac > 0:  # exact duplicates (identical codes in one cell => exactly tied distances)
        k = int(n * dup_frac)
        src = rng.randint(0, n, size=k)
        dst = rng.randint(0, n, size=k)
        X[dst] = X[src]
    return (X.astype(np.float64) / 1024.0).astype(dtype)


def lattice_queries(X, nq, seed, jitter=12):
    """Near-duplicate queries of lattice rows, still on the lattice (exact)."""
    rng = np.random.RandomState(seed)
    idx = rng.randint(0, X.shape[0], size=nq)
    Xi = np.rint(X[idx].astype(np.float64) * 1024.0).astype(np.int64)
    Q = Xi + rng.randint(-jitter, jitter + 1, size=Xi.shape)
    return (Q.astype(np.float64) / 1024.0).astype(X.dtype), idx


def _normalise(X):
    return X / np.linalg.norm(X, axis=1, keepdims=True)


def dlib_style(n, D=128, seed=1234, centres=4096, sigma=0.35, relu=False, centre_seed=None, chunk=1 << 18):
    """BASELINE recipe, NumPy (host) version; returns float32 [n, D]."""
    crng = np.random.RandomState(seed if centre_seed is None else centre_seed)
    C = crng.randn(centres, D)
    rng = np.random.RandomState(seed + 1)
    out = np.empty((n, D), np.float32)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        X = C[rng.randint(0, centres, size=b - a)] + sigma * rng.randn(b - a, D)
        if relu:
            X = np.maximum(X, 0.0)
            X[:, 0] += 1e-3  # keep norms non-zero
        out[a:b] = _normalise(X)
    return out


def sentibank_style(n, D=2048, seed=1234, **kw):
    return dlib_style(n, D=D, seed=seed, relu=True, **kw)


def near_duplicate_queries(X, nq, rho=0.1, seed=4321):
    """q = normalise(db[i] + rho * u), u uniform on the sphere.  Returns (Q float32, i)."""
    rng = np.random.RandomState(seed)
    idx = rng.randint(0, X.shape[0], size=nq)
    U = _normalise(rng.randn(nq, X.shape[1]))
    return _normalise(X[idx].astype(np.float64) + rho * U).astype(np.float32), idx


# ---- device-side (torch) versions of the same recipe, for 1e7-row benchmark databases ---------------
def dlib_style_torch(n, D=128, seed=1234, device="cuda", centres=4096, sigma=0.35, relu=False, chunk=1 << 20):
    """Same mixture as dlib_style (identical centres: NumPy RandomState(seed)), points drawn with the
    torch CUDA generator; returns a float32 [n, D] tensor on `device`."""
    import torch
    C = torch.from_numpy(np.random.RandomState(seed).randn(centres, D)).to(device=device, dtype=torch.float32)
    g = torch.Generator(device=device)
    g.manual_seed(seed + 1)
    out = torch.empty((n, D), dtype=torch.float32, device=device)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        idx = torch.randint(0, centres, (b - a,), generator=g, device=device)
        X = C[idx] + sigma * torch.randn((b - a, D), generator=g, device=device, dtype=torch.float32)
        if relu:
            X = torch.clamp_min(X, 0.0)
            X[:, 0] += 1e-3
        out[a:b] = X / X.norm(dim=1, keepdim=True)
    return out


def near_duplicate_queries_torch(X, nq, rho=0.1, seed=4321):
    """q = normalise(db[i] + rho * u), u uniform on the sphere; returns (Q float32 [nq, D], i int64 [nq])."""
    import torch
    g = torch.Generator(device=X.device)
    g.manual_seed(seed)
    idx = torch.randint(0, X.shape[0], (nq,), generator=g, device=X.device)
    U = torch.randn((nq, X.shape[1]), generator=g, device=X.device, dtype=torch.float32)
    U = U / U.norm(dim=1, keepdim=True)
    Q = X[idx] + rho * U
    return Q / Q.norm(dim=1, keepdim=True), idx


def exact_nn_torch(X, Q, chunk=1 << 20):
    """Exact L2 nearest neighbour of every query by brute force (eval.py:7-38 compute_all_neighbors), on the
    device: ground truth for recall.  Returns int64 [nq] row indices."""
    import torch
    best = torch.full((Q.shape[0],), float("inf"), device=X.device)
    arg = torch.zeros((Q.shape[0],), dtype=torch.int64, device=X.device)
    qn = (Q * Q).sum(1)
    for a in range(0, X.shape[0], chunk):
        xb = X[a:a + chunk]
        d = qn[:, None] - 2.0 * (Q @ xb.T) + (xb * xb).sum(1)[None, :]
        v, i = d.min(dim=1)
        upd = v < best
        best = torch.where(upd, v, best)
        arg = torch.where(upd, i + a, arg)
    return arg
