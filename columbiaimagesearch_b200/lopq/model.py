"""Mirror of lopq/lopq/model.py (inference half): LOPQModel / LOPQModelPCA with the reference's
constructor, attributes and method names; every arithmetic step of predict / project /
get_subquantizer_distances / apply_PCA runs in libb200lopq (CUDA).  Parameter containers stay plain
NumPy arrays so pickles written by the reference (storer/local.py:58) load into these classes.
"""
import os
from collections import namedtuple

import numpy as np

from .. import _native
from .utils import iterate_splits

LOPQCode = namedtuple("LOPQCode", ["coarse", "fine"])   # model.py:444

_NATIVE_ATTR = "_b2l_handle"


def _uint_type(n):
    return np.uint8 if n <= 256 else (np.uint16 if n <= 65536 else np.uint32)


_cluster_cache = {}


def _cluster_handle(centroids):
    """Throw-away model whose first coarse split is `centroids` (for utils.predict_cluster and
    search.multisequence, which take bare centroid arrays in the reference API)."""
    key = (id(centroids), centroids.shape, centroids.dtype.str, os.getpid())
    ent = _cluster_cache.get(key)
    if ent is not None and ent[0] is centroids:
        return ent[1]
    n, d = centroids.shape
    h = _native.Handle()
    z = np.zeros((n, d, d))
    h.set_model((centroids, centroids), (z, z), (np.zeros((n, d)), np.zeros((n, d))),
                ([np.zeros((1, d))], [np.zeros((1, d))]))
    if len(_cluster_cache) > 8:
        _cluster_cache.clear()
    _cluster_cache[key] = (centroids, h)
    return h


# ---- module-level training helpers of lopq/lopq/model.py (imported by eval.py:146 and by user scripts) ------------------
def eigenvalue_allocation(num_buckets, eigenvalues):
    """model.py:19-71."""
    from .train import eigenvalue_allocation as f
    return f(num_buckets, eigenvalues)


def compute_residuals(data, C, device=None):
    """model.py:236-239 -- (data - C[assignments], assignments); the nearest-centroid assignments (utils.predict_cluster over
    rows) run on the GPU unless device=False."""
    from .train import _assign, _use_device
    data = np.asarray(data, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    assignments = _assign(data, C, device=_use_device(device))
    return data - C[assignments], assignments


def project_residuals_to_local(residuals, assignments, Rs, mu):
    """model.py:209-234."""
    from .train import project_residuals_to_local as f
    return f(np.asarray(residuals), np.asarray(assignments), np.asarray(Rs), np.asarray(mu))


def compute_local_rotations(data, C, num_buckets, device=None):
    """model.py:74-206 -> (Rs, mus, assignments, residuals)."""
    from .train import compute_local_rotations as f, _use_device
    return f(np.asarray(data, dtype=np.float64), np.asarray(C, dtype=np.float64), num_buckets, _use_device(device))


def train_coarse(data, V=8, kmeans_coarse_iters=10, n_init=10, random_state=None, device=None):
    """model.py:290-317 -- centroids [V, D] of one coarse quantizer (Lloyd k-means on the device; the reference uses
    sklearn's MiniBatchKMeans: statistical parity only)."""
    from .train import kmeans, _use_device
    return kmeans(np.asarray(data, dtype=np.float64), V, kmeans_coarse_iters, np.random.RandomState(random_state), n_init, _use_device(device))


def train_subquantizers(data, num_buckets, subquantizer_clusters=256, kmeans_local_iters=20, n_init=10, random_state=None, device=None):
    """model.py:320-336 -- one k-means codebook per sub-vector of the (projected) data."""
    from .train import kmeans, _use_device
    rng = np.random.RandomState(random_state)
    return [kmeans(np.asarray(d, dtype=np.float64), subquantizer_clusters, kmeans_local_iters, rng, n_init, _use_device(device))
            for d in np.split(np.asarray(data), num_buckets, axis=1)]


def train(data, V=8, M=4, subquantizer_clusters=256, parameters=None, kmeans_coarse_iters=10, kmeans_local_iters=20, n_init=10,
          subquantizer_sample_ratio=1.0, random_state=None, verbose=False, device=None):
    """model.py:339-437 -> (Cs, Rs, mus, subquantizers)."""
    from .train import train as f
    return f(data, V, M, subquantizer_clusters, parameters, kmeans_coarse_iters, kmeans_local_iters, n_init,
             subquantizer_sample_ratio, random_state, verbose, device)


def train_pca(data, dims=256, subsample=None):
    """model.py:242-287 -> (P, mu)."""
    from .train import train_pca as f
    return f(data, dims, subsample)


class LOPQModel(object):
    def __init__(self, V=8, M=4, subquantizer_clusters=256, parameters=None):
        """model.py:448-493.  parameters = ((C1, C2), (Rs1, Rs2), (mu1, mu2), (subquantizers1, subquantizers2))."""
        self.Cs, self.Rs, self.mus, self.subquantizers = parameters if parameters is not None else (None, None, None, None)
        self._init_shape(V, M, subquantizer_clusters)

    def _init_shape(self, V, M, subquantizer_clusters):
        if self.Cs is not None:
            self.V = self.Cs[0].shape[0]
            self.num_coarse_splits = len(self.Cs)
        else:
            self.V = V
            self.num_coarse_splits = 2
        if self.subquantizers is not None:
            self.num_fine_splits = len(self.subquantizers[0])
            self.M = self.num_fine_splits * self.num_coarse_splits
            self.subquantizer_clusters = self.subquantizers[0][0].shape[0]
        else:
            self.num_fine_splits = M // 2
            self.M = M
            self.subquantizer_clusters = subquantizer_clusters

    # ---- native handle (never pickled) --------------------------------------------------------
    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop(_NATIVE_ATTR, None)
        return d

    def _pca_params(self):
        return None, None, False

    def _native(self):
        """The model's own library handle (encode / project / LUT probes), created lazily so that a
        model unpickled before fork() initialises CUDA in the worker that first uses it."""
        ent = self.__dict__.get(_NATIVE_ATTR)
        pid = os.getpid()
        if ent is not None and ent[1] != pid:      # inherited through fork(): the parent's CUDA context is not usable here
            ent[0].h = None
            ent = None
        if ent is None:
            ent = (self._new_handle(), pid)
            self.__dict__[_NATIVE_ATTR] = ent
        return ent[0]

    def _new_handle(self, device=None):
        if self.Cs is None or self.Rs is None or self.mus is None or self.subquantizers is None:
            raise ValueError("model parameters are not set (fit or pass `parameters`)")
        P, mu, renorm = self._pca_params()
        h = _native.Handle(device)
        h.set_model(self.Cs, self.Rs, self.mus, self.subquantizers, P, mu, renorm)
        return h

    def _invalidate(self):
        self.__dict__.pop(_NATIVE_ATTR, None)

    # ---- training ("next" row; statistical parity only) -------------------------------------------
    def fit(self, data, kmeans_coarse_iters=10, kmeans_local_iters=20, n_init=10, subquantizer_sample_ratio=1.0,
            random_state=None, verbose=False, device=None):
        """model.py:495-519 -> train (model.py:339-437).  The nearest-centroid assignments (every k-means iteration,
        compute_residuals) run on the GPU; `device=False` keeps them on the host."""
        from .train import train
        self.Cs, self.Rs, self.mus, self.subquantizers = train(
            data, self.V, self.M, self.subquantizer_clusters, (self.Cs, self.Rs, self.mus, self.subquantizers),
            kmeans_coarse_iters, kmeans_local_iters, n_init, subquantizer_sample_ratio, random_state, verbose, device)
        self._invalidate()

    def get_split_parameters(self, split):
        """model.py:521-541."""
        return (self.Cs[split] if self.Cs is not None else None,
                self.Rs[split] if self.Rs is not None else None,
                self.mus[split] if self.mus is not None else None,
                self.subquantizers[split] if self.subquantizers is not None else None)

    # ---- encode --------------------------------------------------------------------------------------
    def predict(self, x):
        """model.py:543-561 -- LOPQCode(coarse, fine) of one vector."""
        coarse, fine = self._native().encode(np.asarray(x)[None, :])
        ct, ft = _uint_type(self.V), _uint_type(self.subquantizer_clusters)
        return LOPQCode(tuple(ct(c) for c in coarse[0]), tuple(ft(f) for f in fine[0]))

    def _post_pca(self, x):
        return np.asarray(x)

    def predict_coarse(self, x):
        """model.py:563-573 (x is the D-dim, post-PCA vector)."""
        if self.V > 64:
            cells, _, _ = self._native().cell_order_prefix(np.asarray(x)[None, :], quota=0, max_cells=1)
        else:
            cells, _, _ = self._native().cell_order(np.asarray(x)[None, :], quota=0)
        ct = _uint_type(self.V)
        return (ct(cells[0, 0] // self.V), ct(cells[0, 0] % self.V))

    def predict_fine(self, x, coarse=None):
        """model.py:575-602 -- fine codes under the given (default: nearest) coarse codes."""
        if coarse is None:
            coarse = self.predict_coarse(x)
        _, lut = self._native().project_lut(np.asarray(x)[None, :], [[int(coarse[0]), int(coarse[1])]], want_px=False)
        ft = _uint_type(self.subquantizer_clusters)
        return tuple(ft(k) for k in lut[0].argmin(axis=1))

    def project(self, x, coarse, coarse_split=None):
        """model.py:604-641 -- R[c] . (x_s - C[c] - mu[c]) per split (float64)."""
        px, _ = self._native().project_lut(np.asarray(x)[None, :], [[int(coarse[0]), int(coarse[1])]], want_lut=False)
        if coarse_split is None:
            return px[0]
        return np.split(px[0], self.num_coarse_splits)[coarse_split]

    def get_subquantizer_distances(self, x, coarse, coarse_split=None):
        """model.py:673-704 -- list of M (or M/2) float64 arrays of K squared distances."""
        _, lut = self._native().project_lut(np.asarray(x)[None, :], [[int(coarse[0]), int(coarse[1])]], want_px=False)
        m = self.num_fine_splits
        if coarse_split is None:
            return [lut[0, j] for j in range(self.M)]
        return [lut[0, coarse_split * m + j] for j in range(m)]

    def reconstruct(self, codes):
        """model.py:643-671 -- R[c]^T . concat(sub-centroids) + mu[c] + C[c] per split.  Not on the
        query path (used by the eval.py invariants); plain NumPy on the host."""
        coarse, fine = codes
        out = []
        for fc, split in iterate_splits(fine, self.num_coarse_splits):
            C, R, mu, subC = self.get_split_parameters(split)
            sx = np.concatenate([subC[j][int(f)] for j, f in enumerate(fc)])
            c = int(coarse[split])
            out.append(np.dot(R[c].transpose(), sx) + mu[c] + C[c])
        return np.concatenate(out)

    def get_cell_id_for_coarse_codes(self, coarse_codes):
        """model.py:706-707 (computed in Python ints: the reference's uint8 arithmetic overflows for V > 16)."""
        return int(coarse_codes[1]) + int(coarse_codes[0]) * self.V

    def get_coarse_codes_for_cell_id(self, cell_id):
        """model.py:709-710."""
        return (int(cell_id // self.V), int(cell_id % self.V))

    # ---- the reference's own file formats ----------------------------------------------------------
    def export_mat(self, filename):
        """model.py:712-728 -- Cs [2,V,h], Rs [2,V,h,h], mus [2,V,h], subs [2,M/2,K,ds], V, M in a .mat file."""
        from scipy.io import savemat
        savemat(filename, {"Cs": np.stack([np.asarray(c) for c in self.Cs]), "Rs": np.stack([np.asarray(r) for r in self.Rs]),
                           "mus": np.stack([np.asarray(m) for m in self.mus]),
                           "subs": np.stack([np.stack([np.asarray(s) for s in half]) for half in self.subquantizers]),
                           "V": self.V, "M": self.M})

    @staticmethod
    def load_mat(filename):
        """model.py:730-746."""
        from scipy.io import loadmat
        d = loadmat(filename)
        return LOPQModel(parameters=(tuple(d["Cs"]), tuple(d["Rs"]), tuple(d["mus"]),
                                     tuple([sub for sub in half] for half in d["subs"])))

    def export_proto(self, f):
        """model.py:748-786 -- LOPQModelParams (lopq_model_pb2.py) with float32 values; `f` is a path or a binary file
        object (closed afterwards, as in the reference).  Missing parameter groups are left out."""
        from . import proto
        D = 2 * self.Cs[0].shape[1] if self.Cs is not None else 0
        buf = proto.encode_model(D, self.V, self.M, self.subquantizer_clusters, self.Cs, self.Rs, self.mus, self.subquantizers)
        if isinstance(f, str):
            f = open(f, "wb")
        f.write(buf)
        f.close()

    @staticmethod
    def load_proto(filename):
        """model.py:788-820 -- a model from the protobuf format; None (and a message) when the file cannot be opened."""
        from . import proto
        try:
            with open(filename, "rb") as fh:
                p = proto.decode_model(fh.read())
        except IOError:
            print(str(filename) + ": Could not open file.")
            return None
        halves = lambda a: [a[:len(a) // 2], a[len(a) // 2:]]
        Cs = Rs = mus = subs = None
        if p["Cs"]:
            Cs = p["Cs"]
        if p["Rs"]:
            Rs = [np.stack(h) for h in halves(p["Rs"])]
        if p["mus"]:
            mus = [np.stack(h) for h in halves(p["mus"])]
        if p["subs"]:
            subs = halves(p["subs"])
        return LOPQModel(V=p["V"] or 8, M=p["M"] or 4, subquantizer_clusters=p["num_subquantizers"] or 256,
                         parameters=(Cs, Rs, mus, subs))

    # ---- flat persistence (the product pickles models; this is the fixture format) -----------------
    def to_npz_dict(self):
        d = {"C0": np.asarray(self.Cs[0]), "C1": np.asarray(self.Cs[1]),
             "Rs": np.stack([np.asarray(self.Rs[0]), np.asarray(self.Rs[1])]),
             "mus": np.stack([np.asarray(self.mus[0]), np.asarray(self.mus[1])]),
             "subs": np.stack([np.asarray(s) for s in list(self.subquantizers[0]) + list(self.subquantizers[1])]),
             "M": np.int64(self.M)}
        P, mu, renorm = self._pca_params()
        if P is not None:
            d.update(pca_P=np.asarray(P), pca_mu=np.asarray(mu), renorm=np.bool_(renorm))
        return d

    @staticmethod
    def from_npz(z, prefix=""):
        g = lambda k: z[prefix + k]
        m = int(g("M"))
        subs = g("subs")
        params = ((g("C0"), g("C1")), (g("Rs")[0], g("Rs")[1]), (g("mus")[0], g("mus")[1]),
                  ([subs[j] for j in range(m // 2)], [subs[j] for j in range(m // 2, m)]))
        if prefix + "pca_P" in z:
            return LOPQModelPCA(renorm=bool(g("renorm")), parameters=params + (g("pca_P"), g("pca_mu")))
        return LOPQModel(parameters=params)


class LOPQModelPCA(LOPQModel):
    def __init__(self, V=8, M=4, subquantizer_clusters=256, renorm=False, parameters=None):
        """model.py:826-875.  parameters = (Cs, Rs, mus, subquantizers, P, mu)."""
        (self.Cs, self.Rs, self.mus, self.subquantizers, self.pca_P, self.pca_mu) = \
            parameters if parameters is not None else (None, None, None, None, None, None)
        self.renorm = renorm
        self._init_shape(V, M, subquantizer_clusters)

    def _pca_params(self):
        return self.pca_P, self.pca_mu, self.renorm

    def fit_pca(self, data, pca_dims=256, pca_subsample=None):
        """model.py:878-886 -> train_pca (model.py:242-287).  Retraining an existing PCA is an error, as in the reference."""
        from .train import train_pca
        if self.pca_P is None or self.pca_mu is None:
            self.pca_P, self.pca_mu = train_pca(data, pca_dims, pca_subsample)
            self._invalidate()
        else:
            raise ValueError("You are trying to retrain PCA...")

    def fit(self, data, pca_dims=256, kmeans_coarse_iters=10, kmeans_local_iters=20, n_init=10,
            subquantizer_sample_ratio=1.0, random_state=None, verbose=False, pca_subsample=None,
            apply_pca=True, train_pca=True):
        """model.py:888-931 -- PCA (when `train_pca`), projection of the training data (when `apply_pca`), then LOPQ
        training.  The product trains on features it has already projected:
        ``fit(train_np, verbose=True, apply_pca=False, train_pca=False)`` (searcher_lopqhbase.py:462)."""
        if train_pca:
            self.fit_pca(data, pca_dims, pca_subsample)
        proj = self.apply_PCA(np.asarray(data)) if apply_pca else np.asarray(data)
        LOPQModel.fit(self, proj, kmeans_coarse_iters, kmeans_local_iters, n_init, subquantizer_sample_ratio,
                      random_state, verbose)

    def apply_PCA(self, x, dtype=np.float32):
        """model.py:961-978 -- (x - mu) . P, optional L2 renorm, cast to `dtype`.  float32 (the default, and what search
        and predict use) is produced by the device kernel; any other dtype takes the float64 result of the same kernel
        so that it is not rounded through float32 first."""
        x = np.asarray(x)
        h = _pca_only_handle(self) if self.Cs is None else self._native()      # model not trained yet: PCA-only handle
        x2 = x if x.ndim > 1 else x[None, :]
        y = h.apply_pca(x2) if dtype == np.float32 else h.apply_pca(x2, f64_out=True).astype(dtype, copy=False)
        return y if x.ndim > 1 else y[0]


def _pca_only_handle(model):
    D = model.pca_P.shape[1]
    h = _native.Handle()
    z = np.zeros((1, D // 2))
    h.set_model((z, z), (np.zeros((1, D // 2, D // 2)),) * 2, (z, z), ([np.zeros((1, D // 2))], [np.zeros((1, D // 2))]),
                model.pca_P, model.pca_mu, model.renorm)
    return h
