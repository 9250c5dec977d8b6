"""LOPQ training (mirror of the training half of lopq/lopq/model.py:19-437) -- a "next" row of the
scope table: models are *inputs* of the hot path, so parity here is statistical (distortion /
recall), not bit-exact.  Same algorithm as the reference -- coarse k-means per split, per-cluster
residual covariance -> eigenvectors -> eigenvalue allocation (OPQ balancing) -> local rotation,
k-means sub-quantizers on the locally projected residuals -- but batched: covariances are GEMMs
instead of a per-point Python loop (model.py:142-155), k-means is a seeded Lloyd iteration in
NumPy instead of sklearn's MiniBatchKMeans (model.py:312,330).
"""
import numpy as np


def eigenvalue_allocation(num_buckets, eigenvalues):
    """model.py:19-71 -- greedy assignment of eigen-directions (descending eigenvalue) to the not-full
    bucket with the smallest log-product, to balance variance across sub-vectors."""
    D = len(eigenvalues)
    per = D // num_buckets
    nz = np.abs(eigenvalues[np.nonzero(eigenvalues)])
    ev = eigenvalues / (nz.min() if nz.size else 1.0)
    logs = np.log2(np.maximum(np.abs(ev), 1e-300))
    prod = np.zeros(num_buckets)
    size = np.zeros(num_buckets, dtype=int)
    perm = np.zeros((num_buckets, per), dtype=int)
    for ind in np.argsort(ev)[::-1]:
        open_ = np.nonzero(size < per)[0]
        b = open_[prod[open_].argmin()]
        prod[b] += logs[ind]
        perm[b, size[b]] = ind
        size[b] += 1
    return perm.reshape(D)


def _assign_device(X, C):
    """Nearest centroid of every row on the GPU: utils.predict_cluster (utils.py:33-53, direct-form squared distances,
    first minimum, float64) over all rows in one b2l_encode call -- the O(n k d) step of every Lloyd iteration and of
    compute_residuals (model.py:236-239).  The centroids are wrapped as the first coarse split of a throw-away model."""
    from .model import _cluster_handle
    C = np.ascontiguousarray(C, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    h = _cluster_handle(C)
    coarse, _ = h.encode(np.concatenate([X, X], axis=1), want_fine=False)
    return coarse[:, 0].astype(np.int64)


_KM = {}


def _kmeans_handle():
    """one bare library handle per process for b2l_kmeans (needs no model)"""
    import os
    from .. import _native
    ent = _KM.get(os.getpid())
    if ent is None:
        _KM.clear()
        ent = _native.Handle()
        _KM[os.getpid()] = ent
    return ent


def _use_device(device):
    """device=None means the GPU, like every other arithmetic step of this package (no silent host fallback: without a
    CUDA device the library fails loudly).  device=False keeps the assignments on the host (NumPy) -- training has no
    parity contract (models are inputs of the hot path), and the flag exists for comparisons."""
    return True if device is None else bool(device)


def _assign(X, C, chunk=65536, device=False):
    if device:
        return _assign_device(X, C)
    out = np.empty(X.shape[0], dtype=np.int64)
    cn = (C * C).sum(1)
    for a in range(0, X.shape[0], chunk):
        xb = X[a:a + chunk]
        out[a:a + chunk] = (cn[None, :] - 2.0 * (xb @ C.T)).argmin(1)
    return out


def kmeans(X, k, iters, rng, n_init=1, device=False):
    """Seeded Lloyd k-means (random distinct initial points; empty clusters re-seeded).  device=True: the assignment step
    runs on the GPU (_assign_device); the centroid update is a segmented sum on the host."""
    X = np.asarray(X, dtype=np.float64)
    best, best_cost = None, np.inf
    if device:
        # the whole Lloyd loop on the GPU (b2l_kmeans): assignments in NumPy's float64 order, centroid sums by atomics
        from .. import _native
        h = _kmeans_handle()
        for _ in range(max(1, n_init)):
            C0 = X[rng.choice(X.shape[0], size=k, replace=X.shape[0] < k)].copy()
            reseed = rng.randint(0, X.shape[0], size=(max(1, iters), k))
            C, _, cost = h.kmeans(X, C0, iters, reseed)
            if cost < best_cost:
                best, best_cost = C, cost
        return best
    for _ in range(max(1, n_init)):
        C = X[rng.choice(X.shape[0], size=k, replace=X.shape[0] < k)].copy()
        for _it in range(iters):
            a = _assign(X, C, device=device)
            cnt = np.bincount(a, minlength=k)
            S = _segment_sum(X, a, k)
            empty = cnt == 0
            C = np.where(empty[:, None], X[rng.randint(0, X.shape[0], size=k)], S / np.maximum(cnt, 1)[:, None])
        a = _assign(X, C, device=device)
        cost = ((X - C[a]) ** 2).sum()
        if cost < best_cost:
            best, best_cost = C, cost
    return best


def _segment_sum(X, a, k):
    order = np.argsort(a, kind="stable")
    bounds = np.searchsorted(a[order], np.arange(k + 1))
    S = np.zeros((k, X.shape[1]))
    Xs = X[order]
    for c in range(k):
        if bounds[c + 1] > bounds[c]:
            S[c] = Xs[bounds[c]:bounds[c + 1]].sum(0)
    return S


def compute_local_rotations(data, C, num_buckets, device=False):
    """model.py:74-206 -- per-cluster residual mean, covariance, eigenvectors permuted by
    eigenvalue_allocation.  Returns (R [V,D,D], mu [V,D], assignments, residuals)."""
    V, D = C.shape
    a = _assign(data, C, device=device)
    residuals = data - C[a]
    R = np.zeros((V, D, D))
    mu = np.zeros((V, D))
    for c in range(V):
        r = residuals[a == c]
        n = r.shape[0]
        if n:
            mu[c] = r.mean(0)
        if n < D:
            ev, vecs = np.ones(D), np.eye(D)
        else:
            cov = (r.T @ r) / (n - 1) - np.outer(mu[c], mu[c])
            ev, vecs = np.linalg.eigh((cov + cov.T) / 2)
        R[c] = vecs[:, eigenvalue_allocation(num_buckets, ev)].T      # rows = permuted eigenvectors (model.py:204)
    return R, mu, a, residuals


def project_residuals_to_local(residuals, assignments, Rs, mu):
    """model.py:209-234, batched per cluster."""
    out = np.zeros(residuals.shape)
    for c in range(Rs.shape[0]):
        sel = assignments == c
        if sel.any():
            out[sel] = (residuals[sel] - mu[c]) @ Rs[c].T
    return out


def train(data, V=8, M=4, subquantizer_clusters=256, parameters=None, kmeans_coarse_iters=10, kmeans_local_iters=20,
          n_init=10, subquantizer_sample_ratio=1.0, random_state=None, verbose=False, device=None):
    """model.py:339-437 -- returns (Cs, Rs, mus, subquantizers); existing parameters are kept.  device: True / False / None
    (= the GPU when one is present) for the nearest-centroid assignments."""
    rng = np.random.RandomState(random_state)
    device = _use_device(device)
    data = np.asarray(data, dtype=np.float64)
    Cs, Rs, mus, subs = parameters if parameters is not None else (None, None, None, None)
    h = data.shape[1] // 2
    halves = (data[:, :h], data[:, h:2 * h])
    if Cs is None:
        Cs = tuple(kmeans(x, V, kmeans_coarse_iters, rng, n_init, device) for x in halves)
        if verbose:
            print("coarse quantizers trained")
    m = M // 2
    if Rs is None or mus is None or subs is None:
        rot = [compute_local_rotations(x, C, m, device) for x, C in zip(halves, Cs)]
        if Rs is None or mus is None:
            Rs, mus = tuple(r[0] for r in rot), tuple(r[1] for r in rot)
        if subs is None:
            out = []
            for (R_, mu_, a, res), R, mu in zip(rot, Rs, mus):
                proj = project_residuals_to_local(res, a, R, mu)
                if subquantizer_sample_ratio != 1.0:
                    n = int(proj.shape[0] * subquantizer_sample_ratio)
                    proj = proj[rng.choice(proj.shape[0], size=n, replace=False)]
                ds = h // m
                out.append([kmeans(proj[:, j * ds:(j + 1) * ds], subquantizer_clusters, kmeans_local_iters, rng, n_init, device)
                            for j in range(m)])
                if verbose:
                    print("subquantizers of one split trained")
            subs = tuple(out)
    return Cs, Rs, mus, subs


def train_pca(data, dims=256, subsample=None):
    """model.py:242-287 -- PCA by eigen-decomposition of the covariance of the first `subsample` rows; the kept
    eigen-directions are permuted by eigenvalue_allocation(2, E) so that the two coarse halves carry balanced variance
    (model.py:276-278).  Returns (P [D0, dims], mu)."""
    X = np.asarray(data, dtype=np.float64)
    if subsample:
        X = X[:min(int(subsample), X.shape[0])]
    dims = min(int(dims), X.shape[1])
    mu = X.mean(0)
    cov = (X.T @ X) / max(1, X.shape[0] - 1) - np.outer(mu, mu)          # the reference's estimator (model.py:266-269)
    ev, vecs = np.linalg.eigh(cov)                                       # ascending eigenvalues
    ev, vecs = ev[-dims:], vecs[:, -dims:]
    return vecs[:, eigenvalue_allocation(2, ev)], mu
