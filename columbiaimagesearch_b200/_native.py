"""ctypes binding of libb200lopq.so (include/b200lopq.h).

The library is the only compute path of this package: there is no CPU fallback.  Importing this
module only loads the shared object (works without a GPU, so symbol checks can run anywhere);
creating a handle needs a CUDA device and raises ``NativeError`` otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lopq.so")
ABI_VERSION = 3


class NativeError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("scan_ms", C.c_double), ("plan_ms", C.c_double), ("select_ms", C.c_double), ("total_ms", C.c_double),
                ("codes_scanned", C.c_int64), ("scan_bytes", C.c_int64), ("work_items", C.c_int64),
                ("lut_slots", C.c_int64), ("kernel_launches", C.c_int64), ("exact_queries", C.c_int64),
                ("acc_calls", C.c_int64), ("acc_scan_ms", C.c_double), ("acc_plan_ms", C.c_double), ("acc_select_ms", C.c_double),
                ("acc_total_ms", C.c_double), ("acc_codes_scanned", C.c_int64), ("acc_scan_bytes", C.c_int64),
                ("acc_work_items", C.c_int64), ("acc_kernel_launches", C.c_int64), ("acc_exact_queries", C.c_int64),
                ("packed", C.c_int64), ("rescan_queries", C.c_int64), ("acc_rescan_queries", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_vp, _i, _i64, _h = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
# name -> (restype, argtypes); must list every symbol include/b200lopq.h declares
SIGNATURES = {
    "b2l_create": (_i, [_i, C.POINTER(_h)]),
    "b2l_destroy": (_i, [_h]),
    "b2l_create_sibling": (_i, [_h, C.POINTER(_h)]),
    "b2l_last_error": (C.c_char_p, [_h]),
    "b2l_version": (_i, []),
    "b2l_set_model": (_i, [_h, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b2l_set_pca": (_i, [_h, _i, _vp, _vp, _i]),
    "b2l_encode": (_i, [_h, _vp, _i, _i64, _i, _vp, _vp]),
    "b2l_set_fine_mode": (_i, [_h, _i]),
    "b2l_encode_guard_count": (_i64, [_h, _i]),
    "b2l_debug_fine_scores": (_i, [_h, _vp, _i, _i64, _i, _vp, _vp]),
    "b2l_apply_pca": (_i, [_h, _vp, _i, _i64, _i, _vp]),
    "b2l_apply_pca64": (_i, [_h, _vp, _i, _i64, _i, _vp]),
    "b2l_project_lut": (_i, [_h, _vp, _i, _i64, _vp, _vp, _vp]),
    "b2l_index_add": (_i, [_h, _vp, _vp, _i64, _vp, _i]),
    "b2l_index_clear": (_i, [_h]),
    "b2l_index_size": (_i64, [_h]),
    "b2l_index_cell_sizes": (_i, [_h, _vp]),
    "b2l_index_set_global_cell_sizes": (_i, [_h, _vp]),
    "b2l_index_get_cell": (_i64, [_h, _i, _i, _i64, _vp, _vp]),
    "b2l_cell_order": (_i, [_h, _vp, _i, _i, _i64, _vp, _vp, _vp]),
    "b2l_cell_order_prefix": (_i, [_h, _vp, _i, _i, _i64, _i, _vp, _vp, _vp]),
    "b2l_search": (_i, [_h, _vp, _i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2l_records_bytes": (_i64, [_h, _i, _i]),
    "b2l_search_local": (_i, [_h, _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "b2l_search_merge": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2l_merge_block_bytes": (_i64, [_h, _i, _i]),
    "b2l_search_merge_block": (_i, [_h, _vp, _i, _i, _i, _vp, _i]),
    "b2l_comm_init": (_i, [_h, _i, _i, _i, _i, _i]),
    "b2l_comm_handle_bytes": (_i, []),
    "b2l_comm_get_handle": (_i, [_h, _vp, C.POINTER(_vp)]),
    "b2l_comm_connect": (_i, [_h, _vp, _i]),
    "b2l_sharded_block_bytes": (_i64, [_h, _i, _i]),
    "b2l_search_sharded": (_i, [_h, _vp, _i, _i, _i, _i64, _i, _vp, _i]),
    "b2l_comm_error": (_i, [_h]),
    "b2l_kmeans": (_i, [_h, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b2l_get_stats": (_i, [_h, C.POINTER(Stats)]),
    "b2l_reset_stats": (_i, [_h]),
    "b2l_set_scan_mode": (_i, [_h, _i]),
    "b2l_set_preselect": (_i, [_h, _i]),
    "b2l_set_async": (_i, [_h, _i]),
    "b2l_sync": (_i, [_h]),
    "b2l_debug_force_redo": (_i, [_h, _i]),
    "b2l_debug_candidates": (_i, [_h, _i, _vp, _vp]),
    "b2l_stream": (_vp, [_h]),
}

_lib = None


def load_library():
    """Load libb200lopq.so (built by columbiaimagesearch_b200/build.py).  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError("%s is missing: build it with `python -m columbiaimagesearch_b200.build` "
                          "(nvcc, sm_100a).  This package has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.b2l_version() != ABI_VERSION:
        raise NativeError("libb200lopq.so ABI version %d != expected %d: rebuild" % (lib.b2l_version(), ABI_VERSION))
    _lib = lib
    return lib


def _ptr(a):
    """Host ndarray -> void*, int (device pointer) passes through, None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def _as_queries(X):
    """C-contiguous 2-D float32/float64 view of the input and its is_f64 flag."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[None, :]
    if X.dtype == np.float64:
        return np.ascontiguousarray(X), 1
    if X.dtype != np.float32:
        # NumPy promotes any other dtype to float64 against float64 parameters
        return np.ascontiguousarray(X, dtype=np.float64), 1
    return np.ascontiguousarray(X), 0


class Handle(object):
    """One b2l handle = one CUDA stream + one model + one inverted-index shard."""

    def __init__(self, device=None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "B2L_DEVICE" not in os.environ else int(os.environ["B2L_DEVICE"])
        h = _h()
        rc = self.lib.b2l_create(int(device), C.byref(h))
        if rc != 0:
            raise NativeError(self.lib.b2l_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self.D = self.V = self.M = self.K = self.D0 = None

    def create_sibling(self):
        """A handle sharing this one's model and index with its own stream and workspaces (b2l_create_sibling)."""
        h = _h()
        self._check(self.lib.b2l_create_sibling(self.h, C.byref(h)))
        s = Handle.__new__(Handle)
        s.lib, s.h, s.device = self.lib, h, self.device
        s.D, s.V, s.M, s.K, s.D0 = self.D, self.V, self.M, self.K, self.D0
        s._parent = self                        # keeps the parent alive for as long as the sibling lives
        return s

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2l_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise NativeError("libb200lopq: %s (code %d)" % (self.lib.b2l_last_error(self.h).decode(), rc))
        return rc

    # ---- model -------------------------------------------------------------------------------
    def set_model(self, Cs, Rs, mus, subquantizers, pca_P=None, pca_mu=None, renorm=False):
        V, h = np.asarray(Cs[0]).shape
        D = 2 * h
        subs = list(subquantizers[0]) + list(subquantizers[1])
        M = len(subs)
        K, ds = np.asarray(subs[0]).shape
        coarse_f32 = int(np.asarray(Cs[0]).dtype == np.float32 and np.asarray(Cs[1]).dtype == np.float32)
        f = lambda a: np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in a]))
        cs, rs, ms, sb = f(Cs), f(Rs), f(mus), f(subs)
        assert cs.shape == (2, V, h) and rs.shape == (2, V, h, h) and ms.shape == (2, V, h) and sb.shape == (M, K, ds)
        self._check(self.lib.b2l_set_model(self.h, D, V, M, K, coarse_f32, _ptr(cs), _ptr(rs), _ptr(ms), _ptr(sb)))
        self.D, self.V, self.M, self.K, self.D0 = D, V, M, K, D
        if pca_P is not None:
            P = np.ascontiguousarray(pca_P, dtype=np.float64)
            mu = np.ascontiguousarray(pca_mu, dtype=np.float64)
            assert P.shape[1] == D and mu.shape[0] == P.shape[0]
            self._check(self.lib.b2l_set_pca(self.h, P.shape[0], _ptr(P), _ptr(mu), int(bool(renorm))))
            self.D0 = P.shape[0]

    # ---- encode ---------------------------------------------------------------------------------
    def encode(self, X, want_fine=True):
        X, f64 = _as_queries(X)
        n = X.shape[0]
        assert X.shape[1] == self.D0, "expected %d-d vectors, got %d" % (self.D0, X.shape[1])
        coarse = np.empty((n, 2), np.int32)
        fine = np.empty((n, self.M), np.uint8) if want_fine else None
        if n:
            self._check(self.lib.b2l_encode(self.h, _ptr(X), f64, n, 0, _ptr(coarse), _ptr(fine)))
        return coarse, fine

    def encode_device(self, x_ptr, n, coarse_ptr, fine_ptr, f64=False):
        """Device-resident encode: raw device pointers (e.g. torch tensor .data_ptr())."""
        self._check(self.lib.b2l_encode(self.h, _ptr(int(x_ptr)), int(f64), int(n), 1, _ptr(int(coarse_ptr)), _ptr(int(fine_ptr))))

    def set_fine_mode(self, mode):
        self._check(self.lib.b2l_set_fine_mode(self.h, int(mode)))

    def debug_fine_scores(self, X, j):
        """Tensor-core scores [128][256] of sub-quantizer j for the first 128 rows, and their float64 projections."""
        X, f64 = _as_queries(X)
        scores = np.empty((128, 256), np.float32)
        px = np.empty((128, self.D), np.float64)
        self._check(self.lib.b2l_debug_fine_scores(self.h, _ptr(X), f64, X.shape[0], int(j), _ptr(scores), _ptr(px)))
        return scores, px

    def encode_guard_count(self, reset=False):
        return int(self._check(self.lib.b2l_encode_guard_count(self.h, int(bool(reset)))))

    def apply_pca(self, X, f64_out=False):
        X, f64 = _as_queries(X)
        Y = np.empty((X.shape[0], self.D), np.float64 if f64_out else np.float32)
        fn = self.lib.b2l_apply_pca64 if f64_out else self.lib.b2l_apply_pca
        self._check(fn(self.h, _ptr(X), f64, X.shape[0], 0, _ptr(Y)))
        return Y

    def project_lut(self, X, coarse, want_px=True, want_lut=True):
        X, f64 = _as_queries(X)
        n = X.shape[0]
        co = np.ascontiguousarray(np.asarray(coarse, dtype=np.int32).reshape(n, 2))
        px = np.empty((n, self.D), np.float64) if want_px else None
        lut = np.empty((n, self.M, self.K), np.float64) if want_lut else None
        self._check(self.lib.b2l_project_lut(self.h, _ptr(X), f64, n, _ptr(co), _ptr(px), _ptr(lut)))
        return px, lut

    # ---- index ----------------------------------------------------------------------------------
    def index_add(self, coarse, fine, rowids=None):
        coarse = np.ascontiguousarray(coarse, dtype=np.int32).reshape(-1, 2)
        fine = np.ascontiguousarray(fine, dtype=np.uint8).reshape(coarse.shape[0], self.M)
        if rowids is not None:
            rowids = np.ascontiguousarray(rowids, dtype=np.int64)
            assert rowids.shape[0] == coarse.shape[0]
        self._check(self.lib.b2l_index_add(self.h, _ptr(coarse), _ptr(fine), coarse.shape[0], _ptr(rowids), 0))

    def index_add_device(self, coarse_ptr, fine_ptr, n, rowids_ptr=None):
        self._check(self.lib.b2l_index_add(self.h, _ptr(int(coarse_ptr)), _ptr(int(fine_ptr)), int(n),
                                           None if rowids_ptr is None else _ptr(int(rowids_ptr)), 1))

    def index_clear(self):
        self._check(self.lib.b2l_index_clear(self.h))

    def index_size(self):
        return int(self.lib.b2l_index_size(self.h))

    def cell_sizes(self):
        s = np.zeros(self.V * self.V, np.int64)
        self._check(self.lib.b2l_index_cell_sizes(self.h, _ptr(s)))
        return s

    def set_global_cell_sizes(self, sizes):
        s = np.ascontiguousarray(sizes, dtype=np.int64)
        assert s.shape[0] == self.V * self.V
        self._check(self.lib.b2l_index_set_global_cell_sizes(self.h, _ptr(s)))

    def get_cell(self, c0, c1):
        n = self._check(self.lib.b2l_index_get_cell(self.h, int(c0), int(c1), 0, None, None))
        rowids = np.empty(n, np.int64)
        fine = np.empty((n, self.M), np.uint8)
        if n:
            self._check(self.lib.b2l_index_get_cell(self.h, int(c0), int(c1), n, _ptr(rowids), _ptr(fine)))
        return rowids, fine

    # ---- search ---------------------------------------------------------------------------------
    def cell_order(self, X, quota=None):
        X, f64 = _as_queries(X)
        nq = X.shape[0]
        assert X.shape[1] == self.D
        vv = self.V * self.V
        cells = np.empty((nq, vv), np.int32)
        dists = np.empty((nq, vv), np.float64)
        nvis = np.empty(nq, np.int32)
        q = np.iinfo(np.int64).max if quota is None else int(quota)
        self._check(self.lib.b2l_cell_order(self.h, _ptr(X), f64, nq, q, _ptr(cells), _ptr(dists), _ptr(nvis)))
        return cells, dists, nvis

    def cell_order_prefix(self, X, quota=None, max_cells=1024):
        """V > 64: the first min(visited, max_cells) cells of the multi-sequence traversal (quota cut applied)."""
        X, f64 = _as_queries(X)
        nq = X.shape[0]
        assert X.shape[1] == self.D
        cells = np.empty((nq, max_cells), np.int32)
        dists = np.empty((nq, max_cells), np.float64)
        nvis = np.empty(nq, np.int32)
        q = np.iinfo(np.int64).max if quota is None else int(quota)
        self._check(self.lib.b2l_cell_order_prefix(self.h, _ptr(X), f64, nq, q, int(max_cells), _ptr(cells), _ptr(dists), _ptr(nvis)))
        return cells, dists, nvis

    def search(self, Q, quota, k):
        Q, f64 = _as_queries(Q)
        nq = Q.shape[0]
        assert Q.shape[1] == self.D0, "expected %d-d queries, got %d" % (self.D0, Q.shape[1])
        k = int(k)
        # the library writes every element (rows beyond count[q] come back zero-filled)
        out = dict(rowid=np.empty((nq, k), np.int64), dist=np.empty((nq, k), np.float64), coarse=np.empty((nq, k, 2), np.int32),
                   fine=np.empty((nq, k, self.M), np.uint8), count=np.empty(nq, np.int32), visited=np.empty(nq, np.int32))
        self._check(self.lib.b2l_search(self.h, _ptr(Q), f64, nq, 0, int(quota), k, _ptr(out["rowid"]), _ptr(out["dist"]),
                                        _ptr(out["coarse"]), _ptr(out["fine"]), _ptr(out["count"]), _ptr(out["visited"])))
        if nq and int(out["count"].min()) < k:
            pad = np.arange(k)[None, :] >= out["count"][:, None]
            out["rowid"][pad] = -1
            out["dist"][pad] = np.nan
        return out

    def search_device(self, q_ptr, nq, quota, k, rowid_ptr, dist_ptr, coarse_ptr, fine_ptr, count_ptr, visited_ptr, f64=False):
        p = lambda v: None if v is None else _ptr(int(v))
        self._check(self.lib.b2l_search(self.h, p(q_ptr), int(f64), int(nq), 1, int(quota), int(k), p(rowid_ptr), p(dist_ptr),
                                        p(coarse_ptr), p(fine_ptr), p(count_ptr), p(visited_ptr)))

    def records_bytes(self, nq, k):
        return int(self._check(self.lib.b2l_records_bytes(self.h, int(nq), int(k))))

    def search_local(self, Q, quota, k, records_ptr, exact=False, on_device=False, nq=None, f64=False):
        if on_device:
            self._check(self.lib.b2l_search_local(self.h, _ptr(int(Q)), int(f64), int(nq), 1, int(quota), int(k), int(exact),
                                                  _ptr(int(records_ptr))))
            return
        Q, f64 = _as_queries(Q)
        self._check(self.lib.b2l_search_local(self.h, _ptr(Q), f64, Q.shape[0], 0, int(quota), int(k), int(exact),
                                              _ptr(int(records_ptr))))

    def search_merge(self, records_all_ptr, nranks, nq, k):
        k = int(k)
        out = dict(rowid=np.full((nq, k), -1, np.int64), dist=np.full((nq, k), np.nan), coarse=np.zeros((nq, k, 2), np.int32),
                   fine=np.zeros((nq, k, self.M), np.uint8), count=np.zeros(nq, np.int32), visited=np.zeros(nq, np.int32),
                   certified=np.zeros(nq, np.uint8))
        self._check(self.lib.b2l_search_merge(self.h, _ptr(int(records_all_ptr)), int(nranks), int(nq), k, 0, _ptr(out["rowid"]),
                                              _ptr(out["dist"]), _ptr(out["coarse"]), _ptr(out["fine"]), _ptr(out["count"]),
                                              _ptr(out["visited"]), _ptr(out["certified"])))
        return out

    def merge_block_bytes(self, nq, k):
        return int(self._check(self.lib.b2l_merge_block_bytes(self.h, int(nq), int(k))))

    def search_merge_block(self, records_all_ptr, nranks, nq, k, block_ptr, on_device=False):
        """Merge into one packed block (layout: include/b200lopq.h); asynchronous when set_async(True)."""
        self._check(self.lib.b2l_search_merge_block(self.h, _ptr(int(records_all_ptr)), int(nranks), int(nq), int(k),
                                                    _ptr(int(block_ptr)), int(on_device)))

    # ---- multi-GPU exchange inside the library ---------------------------------------------------------
    def comm_init(self, world, rank, max_nq_home, max_k, f64=False):
        self._check(self.lib.b2l_comm_init(self.h, int(world), int(rank), int(max_nq_home), int(max_k), int(bool(f64))))

    def comm_handle(self):
        """(64-byte IPC handle of this rank's window, its local device address)."""
        buf = C.create_string_buffer(self.lib.b2l_comm_handle_bytes())
        ptr = _vp()
        self._check(self.lib.b2l_comm_get_handle(self.h, buf, C.byref(ptr)))
        return buf.raw, int(ptr.value)

    def comm_local_ptr(self):
        ptr = _vp()
        self._check(self.lib.b2l_comm_get_handle(self.h, None, C.byref(ptr)))
        return int(ptr.value)

    def comm_connect(self, handles=None, pointers=None):
        """handles: list of `world` 64-byte IPC handles (one process per GPU); pointers: list of `world` window addresses
        (ranks that are handles of this process)."""
        if pointers is not None:
            arr = (C.c_void_p * len(pointers))(*[C.c_void_p(int(p)) for p in pointers])
            self._check(self.lib.b2l_comm_connect(self.h, C.cast(arr, _vp), 1))
        else:
            blob = b"".join(handles)
            self._check(self.lib.b2l_comm_connect(self.h, C.c_char_p(blob), 0))

    def sharded_block_bytes(self, nq_home, k):
        return int(self._check(self.lib.b2l_sharded_block_bytes(self.h, int(nq_home), int(k))))

    def search_sharded(self, Qhome, nq_home, quota, k, block_ptr, on_device=False, block_on_device=False, f64=False):
        """Qhome: host ndarray (float32, C-contiguous, pinned when asynchronous) or a device pointer (on_device)."""
        q = _ptr(int(Qhome)) if on_device else _ptr(Qhome)
        self._check(self.lib.b2l_search_sharded(self.h, q, int(f64), int(nq_home), int(on_device), int(quota), int(k),
                                                _ptr(int(block_ptr)), int(block_on_device)))

    def comm_error(self):
        return int(self._check(self.lib.b2l_comm_error(self.h)))

    def kmeans(self, X, C0, iters, reseed):
        """Lloyd k-means on the device (b2l_kmeans).  Returns (centroids [k,d] float64, assignments [n] int32, cost)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        Cc = np.ascontiguousarray(C0, dtype=np.float64).copy()
        n, d = X.shape
        k = Cc.shape[0]
        rs = np.ascontiguousarray(reseed, dtype=np.int64).reshape(max(1, iters), k) if iters > 0 else None
        assign = np.empty(n, np.int32)
        cost = np.zeros(1, np.float64)
        self._check(self.lib.b2l_kmeans(self.h, _ptr(X), n, d, k, int(iters), _ptr(Cc), _ptr(rs), _ptr(assign), _ptr(cost)))
        return Cc, assign, float(cost[0])

    def stats(self):
        s = Stats()
        self._check(self.lib.b2l_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self.lib.b2l_reset_stats(self.h))

    def set_scan_mode(self, mode):
        """0: 16-bit packed tables first (default); 1: float32 tables only."""
        self._check(self.lib.b2l_set_scan_mode(self.h, int(mode)))

    def set_preselect(self, kp_min):
        """at least kp_min candidates per query re-ranked in float64 (0: default k + 8 rounded up to a power of two)"""
        self._check(self.lib.b2l_set_preselect(self.h, int(kp_min)))

    def set_async(self, enabled):
        self._check(self.lib.b2l_set_async(self.h, int(bool(enabled))))

    def sync(self):
        self._check(self.lib.b2l_sync(self.h))

    def search_merge_ptrs(self, records_all_ptr, nranks, nq, k, rowid_ptr, dist_ptr, coarse_ptr, fine_ptr, count_ptr, visited_ptr,
                          certified_ptr, on_device):
        """Merge into caller-owned buffers given as raw pointers (device, or pinned host); asynchronous when set_async(True)."""
        p = lambda v: None if v is None else _ptr(int(v))
        self._check(self.lib.b2l_search_merge(self.h, _ptr(int(records_all_ptr)), int(nranks), int(nq), int(k), int(on_device), p(rowid_ptr),
                                              p(dist_ptr), p(coarse_ptr), p(fine_ptr), p(count_ptr), p(visited_ptr), p(certified_ptr)))

    def debug_force_redo(self, mask):
        self._check(self.lib.b2l_debug_force_redo(self.h, int(mask)))

    def debug_candidates(self, nq):
        app = np.zeros(nq, np.uint32)
        bnd = np.zeros(nq, np.uint32)
        self._check(self.lib.b2l_debug_candidates(self.h, int(nq), _ptr(app), _ptr(bnd)))
        return app, bnd.view(np.float32)

    def stream(self):
        return int(self.lib.b2l_stream(self.h) or 0)
