"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference LOPQ hot path.

This module is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``columbiaimagesearch_b200``) never does; it fails loudly when its CUDA library is missing.

Parity status: **pinned against the real reference run in the build container** --
``tests/test_oracle_vs_reference.py`` executes the reference's own code (via
``oracle/ref_loader.py``) next to this restatement and requires bit-identical codes, cell
orders, LUTs (same ufunc order => identical float64 bits) and search results;
``tests/golden/*.npz`` holds outputs *of the reference itself* (generator:
``tests/golden/make_golden.py``) for the GPU box, where /root/reference does not exist.
The reference ships no golden vectors / known-answer tests of its own for this path
(SURVEY.md section 4 and 8c).

Every function cites the reference lines it restates (paths relative to
/root/reference/lopq/lopq/).  Arithmetic follows NumPy promotion rules exactly as the reference
does: float32 query against float64 parameters computes in float64; float32 query against
float32 coarse centroids (what MiniBatchKMeans returns for float32 training data) computes the
coarse distances in float32.
"""
import heapq
from collections import defaultdict, namedtuple
from itertools import count

import numpy as np

LOPQCode = namedtuple("LOPQCode", ["coarse", "fine"])  # model.py:444


class OracleModel(object):
    """Parameter holder with the reference attribute names (model.py:463-493, 857-875)."""

    def __init__(self, Cs, Rs, mus, subquantizers, pca_P=None, pca_mu=None, renorm=False):
        self.Cs, self.Rs, self.mus, self.subquantizers = Cs, Rs, mus, subquantizers
        self.pca_P, self.pca_mu, self.renorm = pca_P, pca_mu, renorm
        self.V = Cs[0].shape[0]
        self.num_coarse_splits = len(Cs)
        self.num_fine_splits = len(subquantizers[0])
        self.M = self.num_fine_splits * self.num_coarse_splits
        self.subquantizer_clusters = subquantizers[0][0].shape[0]

    @property
    def is_pca(self):
        return self.pca_P is not None

    @classmethod
    def from_npz(cls, z, prefix=""):
        g = lambda k: z[prefix + k]
        m = int(g("M"))
        subs = g("subs")
        subquantizers = ([subs[j] for j in range(m // 2)], [subs[j] for j in range(m // 2, m)])
        Cs = (g("C0"), g("C1"))
        Rs = (g("Rs")[0], g("Rs")[1])
        mus = (g("mus")[0], g("mus")[1])
        if prefix + "pca_P" in z:
            return cls(Cs, Rs, mus, subquantizers, g("pca_P"), g("pca_mu"), bool(g("renorm")))
        return cls(Cs, Rs, mus, subquantizers)


def model_to_npz_dict(model, prefix=""):
    """Flatten a reference / oracle / product model into arrays (fixture format)."""
    d = {
        "C0": np.asarray(model.Cs[0]), "C1": np.asarray(model.Cs[1]),
        "Rs": np.stack([np.asarray(model.Rs[0]), np.asarray(model.Rs[1])]),
        "mus": np.stack([np.asarray(model.mus[0]), np.asarray(model.mus[1])]),
        "subs": np.stack([np.asarray(s) for s in list(model.subquantizers[0]) + list(model.subquantizers[1])]),
        "M": np.int64(model.M),
    }
    if getattr(model, "pca_P", None) is not None:
        d["pca_P"] = np.asarray(model.pca_P)
        d["pca_mu"] = np.asarray(model.pca_mu)
        d["renorm"] = np.bool_(model.renorm)
    return {prefix + k: v for k, v in d.items()}


# ----------------------------------------------------------------------------- primitives
def iterate_splits(x, splits):
    """utils.py:8-22 -- equal contiguous sub-vectors (py2 integer division)."""
    size = len(x) // splits
    for s in range(splits):
        yield x[s * size:(s + 1) * size], s


def cluster_distances(x, centroids):
    """utils.py:47 / search.py:39 -- direct-form squared L2, summed along the last axis
    (NumPy pairwise summation), dtype by NumPy promotion."""
    return ((x - centroids) ** 2).sum(axis=1)


def predict_cluster(x, centroids):
    """utils.py:33-53 -- first-minimum argmin, smallest unsigned type that fits."""
    cid = cluster_distances(x, centroids).argmin(axis=0)
    n = centroids.shape[0]
    if n <= 256:
        return np.uint8(cid)
    if n <= 65536:
        return np.uint16(cid)
    return np.uint32(cid)


def apply_pca(model, x, dtype=np.float32):
    """model.py:961-978 -- (x - mu) . P, optional L2 re-normalisation, cast."""
    y = np.dot(x - model.pca_mu, model.pca_P)
    if model.renorm:
        if y.ndim > 1:
            nrm = np.linalg.norm(y, axis=1)
            y = y / np.tile(nrm[:, np.newaxis], (1, model.pca_P.shape[1]))
        else:
            y = y / np.linalg.norm(y)
    return y.astype(dtype=dtype)


def predict_coarse(model, x):
    """model.py:563-573."""
    return tuple(predict_cluster(cx, model.Cs[s]) for cx, s in iterate_splits(x, model.num_coarse_splits))


def project(model, x, coarse, coarse_split=None):
    """model.py:604-641 -- r = cx - C[c]; R[c] . (r - mu[c]); no transpose."""
    if coarse_split is None:
        it = iterate_splits(x, model.num_coarse_splits)
    else:
        it = [(np.split(x, model.num_coarse_splits)[coarse_split], coarse_split)]
    out = []
    for cx, s in it:
        c = coarse[s]
        r = cx - model.Cs[s][c]
        out.append(np.dot(model.Rs[s][c], r - model.mus[s][c]))
    return np.concatenate(out)


def predict_fine(model, x, coarse=None):
    """model.py:575-602."""
    if coarse is None:
        coarse = predict_coarse(model, x)
    px = project(model, x, coarse)
    fine = []
    for cx, s in iterate_splits(px, model.num_coarse_splits):
        subC = model.subquantizers[s]
        fine += [predict_cluster(fx, subC[j]) for fx, j in iterate_splits(cx, model.num_fine_splits)]
    return tuple(fine)


def predict(model, x):
    """model.py:543-561 (PCA variant 980-1003: apply_PCA first)."""
    if model.is_pca:
        x = apply_pca(model, x)
    coarse = predict_coarse(model, x)
    return LOPQCode(coarse, predict_fine(model, x, coarse))


def compute_codes(model, data):
    """utils.py:203-218 -- [predict(d) for d in data]."""
    return [predict(model, d) for d in data]


def subquantizer_distances(model, x, coarse, coarse_split=None):
    """model.py:673-704 -- LUT: project (both splits), then per sub-vector squared distances to
    all K sub-centroids.  Returns a list of m (or M) float64 arrays of length K."""
    px = project(model, x, coarse)
    if coarse_split is None:
        it = iterate_splits(px, model.num_coarse_splits)
    else:
        it = [(np.split(px, model.num_coarse_splits)[coarse_split], coarse_split)]
    out = []
    for cx, s in it:
        subC = model.subquantizers[s]
        out += [cluster_distances(fx, subC[j]) for fx, j in iterate_splits(cx, model.num_fine_splits)]
    return out


def reconstruct(model, code):
    """model.py:643-671 -- R[c]^T . concat(sub-centroids) + mu[c] + C[c] per split."""
    coarse, fine = code
    xs = []
    for fc, s in iterate_splits(fine, model.num_coarse_splits):
        subC = model.subquantizers[s]
        sx = np.concatenate([subC[j][f] for j, f in enumerate(fc)])
        c = coarse[s]
        xs.append(np.dot(model.Rs[s][c].transpose(), sx) + model.mus[s][c] + model.Cs[s][c])
    return np.concatenate(xs)


# ----------------------------------------------------------------------------- multi-index
def multisequence(x, centroids):
    """search.py:13-82 -- multi-sequence traversal of the V x V grid.

    Yields (dist, (c0, c1)) in non-decreasing dist; heap entries are (dist, (i0, i1)) with
    i* positions in the per-split argsort, so ties are broken by that tuple.  A cell is
    pushed once both of its predecessors (i0-1,i1), (i0,i1-1) have been popped."""
    V = centroids[0].shape[0]
    dists, order = [], []
    for cx, s in iterate_splits(x, len(centroids)):
        d = cluster_distances(cx, centroids[s])
        dists.append(d)
        order.append(np.argsort(d))

    def cell_dist(i0, i1):
        # python sum([...]) starts from int 0: 0 + d0 + d1
        return 0 + dists[0][order[0][i0]] + dists[1][order[1][i1]]

    heap = [(cell_dist(0, 0), (0, 0))]
    done = set()
    while heap:
        d, (i0, i1) = heapq.heappop(heap)
        yield d, (order[0][i0], order[1][i1])
        done.add((i0, i1))
        if (i1 == 0 or (i0 + 1, i1 - 1) in done) and i0 + 1 < V:
            heapq.heappush(heap, (cell_dist(i0 + 1, i1), (i0 + 1, i1)))
        if (i0 == 0 or (i0 - 1, i1 + 1) in done) and i1 + 1 < V:
            heapq.heappush(heap, (cell_dist(i0, i1 + 1), (i0, i1 + 1)))


class OracleSearcher(object):
    """search.py:85-224 + 310-382 -- in-RAM dict index, quota retrieval, ADC, stable sort."""

    def __init__(self, model):
        self.model = model
        self.index = defaultdict(list)
        self.nb_indexed = 0

    def add_codes(self, codes, ids=None):
        """search.py:325-369 -- append (id, code) to its cell unless that id is already there."""
        if ids is None:
            ids = count()
        seen = {}
        for item_id, code in zip(ids, codes):
            cell = code[0]
            if cell not in seen:
                seen[cell] = set(i for i, _ in self.index[cell]) if cell in self.index else set()
            if item_id not in seen[cell]:
                self.index[cell].append((item_id, code))
                seen[cell].add(item_id)
                self.nb_indexed += 1

    def add_data(self, data, ids=None):
        """search.py:94-108."""
        self.add_codes(compute_codes(self.model, data), ids)

    def get_cell(self, cell):
        return self.index[cell]  # search.py:372-382 (defaultdict: creates empty cells)

    def get_result_quota(self, x, quota=10):
        """search.py:110-135 -- whole cells in multisequence order until len >= quota."""
        retrieved, visited = [], 0
        for _, cell in multisequence(x, self.model.Cs):
            retrieved += self.get_cell(cell)
            visited += 1
            if len(retrieved) >= quota:
                break
        return retrieved, visited

    def compute_distances(self, x, items):
        """search.py:137-177 -- LUT halves memoised per c0 / per c1; dist = left-to-right
        python sum over the M LUT entries (float64)."""
        memo = [{}, {}]
        out = []
        for item in items:
            coarse, fine = item[1]
            for s in (0, 1):
                if coarse[s] not in memo[s]:
                    memo[s][coarse[s]] = subquantizer_distances(self.model, x, coarse, coarse_split=s)
            lut = memo[0][coarse[0]] + memo[1][coarse[1]]
            out.append((sum([lut[i][fc] for i, fc in enumerate(fine)]), item))
        return out

    def search(self, x, quota=10, limit=None, with_dists=False):
        """search.py:179-224.  Returns (list of (id, code[, dist]) tuples, visited)."""
        if self.model.is_pca:
            x = apply_pca(self.model, x)
        retrieved, visited = self.get_result_quota(x, quota)
        results = sorted(self.compute_distances(x, retrieved), key=lambda d: d[0])  # stable
        if limit is None:
            limit = quota
        results = results[:limit]
        if with_dists:
            R = namedtuple("Result", ["id", "code", "dist"])
            return [R(d[1][0], d[1][1], d[0]) for d in results], visited
        R = namedtuple("Result", ["id", "code"])
        return [R(d[1][0], d[1][1]) for d in results], visited


# ----------------------------------------------------------------------------- vectorised forms
# Same arithmetic, batched over rows so that 1e5..1e7-row ground truth finishes in seconds.
# `((X[:, None, :] - C[None]) ** 2).sum(axis=2)` reduces along the contiguous last axis with the
# same pairwise routine as the per-row form, so results are bit-identical (tested); the local
# rotation uses one matmul per coarse code instead of one dgemv per row, which may differ in
# the last ulp of float64 (BLAS kernel order) -- never enough to move an argmin on tested data.
def encode_batch(model, X, chunk=8192):
    """Vectorised model.py:543-602 over rows.  Returns (coarse [n,2] int32, fine [n,M] uint8)."""
    X = np.asarray(X)
    if model.is_pca:
        X = apply_pca(model, X)
    n, D = X.shape
    h, m, ds = D // 2, model.num_fine_splits, D // model.M
    coarse = np.empty((n, 2), np.int32)
    fine = np.empty((n, model.M), np.uint8 if model.subquantizer_clusters <= 256 else np.int32)
    for a in range(0, n, chunk):
        xb = X[a:a + chunk]
        for s in (0, 1):
            cx = xb[:, s * h:(s + 1) * h]
            C = model.Cs[s]
            c = ((cx[:, None, :] - C[None]) ** 2).sum(axis=2).argmin(axis=1)
            coarse[a:a + chunk, s] = c
            r = cx - C[c]
            v = r - model.mus[s][c]
            px = np.empty(v.shape, np.result_type(v.dtype, model.Rs[s].dtype))
            for cc in np.unique(c):
                sel = c == cc
                px[sel] = v[sel] @ model.Rs[s][cc].T
            for j in range(m):
                fx = px[:, j * ds:(j + 1) * ds]
                sub = model.subquantizers[s][j]
                d = ((fx[:, None, :] - sub[None]) ** 2).sum(axis=2)
                fine[a:a + chunk, s * m + j] = d.argmin(axis=1)
    return coarse, fine


class ArrayIndex(object):
    """Array form of LOPQSearcher's dict index: per cell, rows in insertion order."""

    def __init__(self, V, coarse, fine, ids=None):
        n = coarse.shape[0]
        self.V = V
        self.ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids)
        cell = coarse[:, 0].astype(np.int64) * V + coarse[:, 1]
        # search.py:356 -- an id is kept once per cell (first occurrence wins)
        _, first = np.unique(np.stack([cell, self.ids.astype(np.int64)], 1), axis=0, return_index=True)
        keep = np.sort(first)
        cell, self.ids, fine, coarse = cell[keep], self.ids[keep], fine[keep], coarse[keep]
        order = np.argsort(cell, kind="stable")
        self.rows = order
        self.fine = fine[order]
        self.coarse = coarse[order]
        self.row_ids = self.ids[order]
        self.sizes = np.bincount(cell, minlength=V * V)
        self.starts = np.concatenate([[0], np.cumsum(self.sizes)])


def search_arrays(model, index, x, quota=10, limit=None):
    """Vectorised search.py:179-224 against an ArrayIndex.

    Returns (ids[k], dists[k] float64, coarse[k,2], fine[k,M], visited)."""
    if model.is_pca:
        x = apply_pca(model, x)
    M, m = model.M, model.num_fine_splits
    segs, visited, got = [], 0, 0
    for _, cell in multisequence(x, model.Cs):
        cid = int(cell[0]) * index.V + int(cell[1])
        segs.append((cell, int(index.starts[cid]), int(index.starts[cid + 1])))
        got += index.sizes[cid]
        visited += 1
        if got >= quota:
            break
    memo = [{}, {}]
    dparts, rparts = [], []
    for cell, a, b in segs:
        if a == b:
            continue
        for s in (0, 1):
            if cell[s] not in memo[s]:
                memo[s][cell[s]] = subquantizer_distances(model, x, cell, coarse_split=s)
        lut = memo[0][cell[0]] + memo[1][cell[1]]
        f = index.fine[a:b]
        d = 0 + lut[0][f[:, 0]]
        for j in range(1, M):
            d = d + lut[j][f[:, j]]          # left-to-right, as python sum() in search.py:173
        dparts.append(d)
        rparts.append(np.arange(a, b))
    if not dparts:
        e = np.empty(0)
        return e.astype(np.int64), e, np.empty((0, 2), np.int32), np.empty((0, M), np.uint8), visited
    d = np.concatenate(dparts)
    rows = np.concatenate(rparts)
    order = np.argsort(d, kind="stable")    # sorted() is stable: ties keep retrieval order
    if limit is None:
        limit = quota
    order = order[:limit]
    rows = rows[order]
    return index.row_ids[rows], d[order], index.coarse[rows], index.fine[rows], visited


def recall_at(searcher_fn, queries, nns, thresholds=(1, 10, 100, 1000)):
    """eval.py:92-142 -- recall@T: the single true nearest neighbour appears in the first T
    results of search(q, quota=thresholds[-1])."""
    rec = np.zeros(len(thresholds))
    for i, q in enumerate(queries):
        ids = searcher_fn(q, thresholds[-1])
        for j, rid in enumerate(ids):
            if rid == nns[i]:
                for k, t in enumerate(thresholds):
                    if j < t:
                        rec[k] += 1
    return rec / len(queries)
