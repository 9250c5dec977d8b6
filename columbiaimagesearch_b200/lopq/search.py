"""Mirror of lopq/lopq/search.py: multisequence, LOPQSearcherBase / LOPQSearcher with the
reference's method names, arguments and return shapes.  The inverted index lives on the GPU
(cell-major code array); search = cell order + quota cut + LUT build + ADC scan + top-k in
libb200lopq.  Host code here only keeps ids (arbitrary Python objects in the reference), applies
the per-cell id de-duplication of add_codes (search.py:325-369) and wraps results.
"""
import os
from collections import namedtuple
from itertools import count

import numpy as np

from .. import _native
from .model import LOPQCode, LOPQModelPCA, _cluster_handle
from .utils import compute_codes_arrays

_ID_BITS = 39       # fast de-dup key = cell << 39 | id  (non-negative integer ids below 2**39; cell < 2**24 at V = 4096)
_DENSE_V = 64       # up to this V the library materialises the whole V x V cell order; above it, prefixes of the traversal


def multisequence(x, centroids):
    """search.py:13-82 -- generator of (dist, (c0, c1)) over the V x V cells in non-decreasing
    d0[c0] + d1[c1], the multi-sequence order.  Evaluated on the device (whole order at once)."""
    x = np.asarray(x)
    c0, c1 = np.asarray(centroids[0]), np.asarray(centroids[1])
    V = c0.shape[0]
    h = _pair_handle(c0, c1)
    f32 = x.dtype == np.float32 and c0.dtype == np.float32 and c1.dtype == np.float32
    if V > _DENSE_V:
        # large V: the traversal is produced in growing prefixes (the generator is usually abandoned after a few cells)
        done, cap = 0, 1024
        while done < V * V:
            cap = min(cap, V * V)
            cells, dists, nvis = h.cell_order_prefix(x[None, :], quota=None, max_cells=cap)
            n = int(nvis[0])
            for i in range(done, n):
                d = np.float32(dists[0, i]) if f32 else np.float64(dists[0, i])
                yield d, (int(cells[0, i]) // V, int(cells[0, i]) % V)
            done = n
            if n < cap:
                break
            cap *= 8
        return
    cells, dists, nvis = h.cell_order(x[None, :])
    for i in range(int(nvis[0])):
        d = np.float32(dists[0, i]) if f32 else np.float64(dists[0, i])
        yield d, (int(cells[0, i]) // V, int(cells[0, i]) % V)


_pair_cache = {}


def _pair_handle(c0, c1):
    key = (id(c0), id(c1), os.getpid())
    ent = _pair_cache.get(key)
    if ent is not None and ent[0] is c0 and ent[1] is c1:
        return ent[2]
    V, d = c0.shape
    h = _native.Handle()
    z = np.zeros((V, d, d))
    h.set_model((c0, c1), (z, z), (np.zeros((V, d)), np.zeros((V, d))), ([np.zeros((1, d))], [np.zeros((1, d))]))
    if len(_pair_cache) > 8:
        _pair_cache.clear()
    _pair_cache[key] = (c0, c1, h)
    return h


def codes_to_arrays(codes, M):
    """Iterable of LOPQCode / (coarse, fine) tuples / [coarse, fine] lists, or an (ndarray, ndarray)
    pair, -> (coarse [n,2] int32, fine [n,M] uint8)."""
    if isinstance(codes, tuple) and len(codes) == 2 and isinstance(codes[0], np.ndarray) and codes[0].ndim == 2:
        return np.ascontiguousarray(codes[0], dtype=np.int32), np.ascontiguousarray(codes[1], dtype=np.uint8)
    codes = list(codes)
    n = len(codes)
    if n == 0:
        return np.zeros((0, 2), np.int32), np.zeros((0, M), np.uint8)
    # bulk conversion (the per-update code dicts the product pickles hold ~10^4..10^6 entries, searcher_lopqhbase.py:506-522)
    coarse = np.asarray([c[0] for c in codes], dtype=np.int32).reshape(n, 2)
    fine = np.asarray([c[1] for c in codes], dtype=np.uint8).reshape(n, M)
    return coarse, fine


class _IndexView(object):
    """Read-only stand-in for the reference's ``defaultdict(list)`` index (search.py:322)."""

    def __init__(self, searcher):
        self._s = searcher

    def __getitem__(self, cell):
        return self._s.get_cell(cell)

    def __contains__(self, cell):
        sizes = self._s._handle.cell_sizes()
        V = self._s.model.V
        return 0 <= cell[0] < V and 0 <= cell[1] < V and sizes[int(cell[0]) * V + int(cell[1])] > 0

    def keys(self):
        sizes = self._s._handle.cell_sizes()
        V = self._s.model.V
        return [(int(c) // V, int(c) % V) for c in np.nonzero(sizes)[0]]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return int(np.count_nonzero(self._s._handle.cell_sizes()))


class LOPQSearcherBase(object):
    """search.py:85-307."""

    def __init__(self):
        self.nb_indexed = 0
        self.verbose = 0

    def get_nb_indexed(self):
        return self.nb_indexed

    def add_data(self, data, ids=None, num_procs=1):
        """search.py:94-108 -- encode on the GPU, then add_codes."""
        self.add_codes(compute_codes_arrays(data, self.model), ids)

    def add_codes_from_dict(self, codes_dict):
        """search.py:275-283 -- {id: [coarse, fine]} as pickled per update by the product."""
        ids = list(codes_dict.keys())
        self.add_codes([codes_dict[k] for k in ids], ids)

    def add_codes_from_local(self, path):
        """search.py:227-243 -- TSV lines `id\\t[[c0, c1], [f...]]`."""
        import ast
        ids, codes = [], []
        with open(path) as f:
            for line in f:
                if not line.strip():
                    continue
                a, b = line.rstrip("\n").split("\t")
                ids.append(a)
                codes.append(ast.literal_eval(b))
        self.add_codes(codes, ids)

    def _query_vector(self, x):
        # search.py:198-200: the searcher applies the PCA; here the library does it inside the call
        return np.asarray(x)

    def get_result_quota(self, x, quota=10):
        """search.py:110-135 -- (items of whole cells in multisequence order until >= quota, visited).
        x is the D-dim (post-PCA) vector, as in the reference."""
        V = self.model.V
        if V > _DENSE_V:
            cap = 4096
            while True:
                cells, _, nvis = self._handle.cell_order_prefix(np.asarray(x)[None, :], quota=quota, max_cells=cap)
                if int(nvis[0]) < cap or cap >= V * V:
                    break
                cap = min(V * V, cap * 8)
        else:
            cells, _, nvis = self._handle.cell_order(np.asarray(x)[None, :], quota=quota)
        retrieved = []
        for i in range(int(nvis[0])):
            retrieved += self.get_cell((int(cells[0, i]) // V, int(cells[0, i]) % V))
        return retrieved, int(nvis[0])

    def compute_distances(self, x, items):
        """search.py:137-177 -- [(dist, item)] with the ADC distance of every item; LUT halves from
        the device (one probe per distinct coarse pair), the per-item gather-sum on the host.  The
        production path never calls this: search() ranks on the GPU."""
        x = np.asarray(x)
        pairs = sorted(set((int(it[1][0][0]), int(it[1][0][1])) for it in items))
        luts = {}
        if pairs:
            _, lut = self.model._native().project_lut(np.repeat(x[None, :], len(pairs), 0), pairs, want_px=False)
            luts = {p: lut[i] for i, p in enumerate(pairs)}
        out = []
        M = self.model.M
        for it in items:
            coarse, fine = it[1]
            t = luts[(int(coarse[0]), int(coarse[1]))]
            d = t[0][int(fine[0])]
            for j in range(1, M):
                d = d + t[j][int(fine[j])]
            out.append((d, it))
        return out

    def search(self, x, quota=10, limit=None, with_dists=False):
        """search.py:179-224 -- (list of Result(id, code[, dist]), visited)."""
        out = self.search_batch(np.asarray(x)[None, :], quota=quota, limit=limit)
        n = int(out["count"][0])
        ids, V = out["ids"], self.model.V
        res = []
        if with_dists:
            R = namedtuple("Result", ["id", "code", "dist"])
        else:
            R = namedtuple("Result", ["id", "code"])
        for j in range(n):
            code = LOPQCode(coarse=(int(out["coarse"][0, j, 0]), int(out["coarse"][0, j, 1])),
                            fine=tuple(int(f) for f in out["fine"][0, j]))
            i = ids[0][j]
            i = i.item() if isinstance(i, np.generic) else i
            res.append(R(i, code, float(out["dist"][0, j])) if with_dists else R(i, code))
        return res, int(out["visited"][0])


class LOPQSearcher(LOPQSearcherBase):
    """search.py:310-382 with a GPU-resident index.  ``device`` selects the GPU (default: LOCAL_RANK
    or 0).  Extra, array-level entry points: add_codes((coarse, fine) arrays), search_batch."""

    def __init__(self, model, device=None, keep_host_copy=True):
        """The library handle (CUDA context, device index) is created by the PROCESS that first needs the device, not
        here: the product builds its searcher in the gunicorn master (`--preload`) and serves from forked workers
        (setup/components/search/docker-compose.yml:67), and a CUDA context does not survive fork().  add_codes only
        records the rows on the host (the reference's in-RAM index is copied by fork the same way); each process
        uploads them when it first searches.  keep_host_copy=False drops the host rows after upload (not fork-safe)."""
        super(LOPQSearcher, self).__init__()
        self.model = model
        self._device = device
        self._keep_host = bool(keep_host_copy)
        self._h, self._h_pid = None, None
        self._host_rows = []               # [(coarse, fine, rowids)] of everything added (replayed after fork)
        self._uploaded = 0                 # entries of _host_rows already on this process's device
        self.index = _IndexView(self)
        self._numeric = True               # ids seen so far are all non-negative ints < 2**40
        self._row_ids = []                 # per add call: ndarray of ids (numeric) in row order
        self._row_cells = []               # per add call: ndarray of cell ids in row order
        self._row_ids_flat = None
        self._id2num, self._num2id = None, None      # generic ids: dense numbering
        self._keys = None                  # sorted int64 keys (cell << 40 | idnum) of everything indexed
        self._pending_keys = []

    # ---- device handle (per process) ---------------------------------------------------------------
    @property
    def _handle(self):
        pid = os.getpid()
        if self._h is None or self._h_pid != pid:
            if self._h is not None:
                if not self._keep_host:
                    raise RuntimeError("searcher built with keep_host_copy=False cannot be used from a forked process")
                self._h.h = None           # the parent's handle: its context is not ours to destroy
            self._h = self.model._new_handle(self._device)
            self._h_pid = pid
            self._uploaded = 0
        while self._uploaded < len(self._host_rows):
            coarse, fine, rows = self._host_rows[self._uploaded]
            self._h.index_add(coarse, fine, rows)
            self._uploaded += 1
            if not self._keep_host:
                self._host_rows[self._uploaded - 1] = None
        if not self._keep_host and self._host_rows:
            self._host_rows, self._uploaded = [], 0
        return self._h

    def _device_add(self, coarse, fine, rows):
        self._host_rows.append((coarse, fine, rows))
        if self._h is not None and self._h_pid == os.getpid():
            self._handle                   # already on a device in this process: upload now  # noqa: B018

    def _device_clear(self):
        self._host_rows, self._uploaded = [], 0
        if self._h is not None and self._h_pid == os.getpid():
            self._h.index_clear()

    # ---- id bookkeeping --------------------------------------------------------------------------
    def _to_numeric(self, ids, n):
        """ids -> int64 array usable in de-dup keys; switches to dense numbering for non-integer ids."""
        if ids is None:
            return np.arange(n, dtype=np.int64), True       # `count()` restarts at every call (search.py:337)
        if isinstance(ids, count):
            start = next(ids)
            return np.arange(start, start + n, dtype=np.int64), True
        if isinstance(ids, range):
            ids = np.arange(ids.start, ids.stop, ids.step, dtype=np.int64)[:n]
        if not isinstance(ids, np.ndarray):
            ids = list(ids)[:n] if not hasattr(ids, "__len__") else ids
            arr = np.asarray(ids[:n]) if len(ids) else np.zeros(0, np.int64)
        else:
            arr = ids[:n]
        if self._numeric and arr.dtype.kind in "iu" and (arr.size == 0 or (arr.min() >= 0 and arr.max() < (1 << _ID_BITS))):
            return arr.astype(np.int64), False
        # generic ids (sha1 strings ...): number them densely in first-seen order
        if self._numeric:
            self._numeric = False
            self._id2num, self._num2id = {}, []
            if self._row_ids:          # renumber what was indexed under integer ids
                old_ids, old_cells = np.concatenate(self._row_ids), np.concatenate(self._row_cells)
                uniq, inv = np.unique(old_ids, return_inverse=True)
                self._num2id = [int(v) for v in uniq]
                self._id2num = {v: i for i, v in enumerate(self._num2id)}
                m = inv.astype(np.int64)
                self._row_ids, self._row_cells, self._row_ids_flat = [m], [old_cells], None
                self._keys, self._pending_keys = None, [(old_cells << _ID_BITS) | m]
        seq = ids if not isinstance(ids, np.ndarray) else ids.tolist()
        nums = np.empty(n, np.int64)
        for i in range(n):
            v = seq[i]
            v = v.item() if isinstance(v, np.generic) else v
            j = self._id2num.get(v)
            if j is None:
                j = len(self._num2id)
                self._id2num[v] = j
                self._num2id.append(v)
            nums[i] = j
        return nums, False

    def _existing_keys(self):
        if self._pending_keys:
            parts = ([self._keys] if self._keys is not None else []) + self._pending_keys
            self._keys = np.sort(np.concatenate(parts))
            self._pending_keys = []
        return self._keys

    def add_codes(self, codes, ids=None):
        """search.py:325-369 -- append codes to their cells; an id already present in a cell is skipped."""
        coarse, fine = codes_to_arrays(codes, self.model.M)
        n = coarse.shape[0]
        if n == 0:
            return
        V = self.model.V
        bad = ((coarse < 0) | (coarse >= V)).any(axis=1)
        if bad.any():
            # a coarse code outside [0, V) names a cell no query can ever visit; the reference logs such an item and goes
            # on (search.py:343-367).  It is dropped here before any bookkeeping, so host and device stay in step.
            if self.verbose > 0:
                print("Discarding %d codes with a coarse code outside [0, %d)" % (int(bad.sum()), V))
            good = np.nonzero(~bad)[0]
            if ids is not None and not isinstance(ids, (count, range)):
                seq = ids if hasattr(ids, "__getitem__") else list(ids)
                ids = [seq[int(i)] for i in good if int(i) < len(seq)]
            elif ids is not None:
                seq = np.asarray([next(ids) for _ in range(n)]) if isinstance(ids, count) else np.asarray(ids)[:n]
                ids = seq[good[good < seq.shape[0]]]
            else:
                ids = good.astype(np.int64)
            coarse, fine = coarse[good], fine[good]
            n = coarse.shape[0]
            if n == 0:
                return
        nums, known_unique = self._to_numeric(ids, n)
        n = min(n, nums.shape[0])
        coarse, fine, nums = coarse[:n], fine[:n], nums[:n]
        cell = coarse[:, 0].astype(np.int64) * V + coarse[:, 1]
        key = (cell << _ID_BITS) | nums
        keep = None
        if not known_unique:
            _, first = np.unique(key, return_index=True)
            if first.shape[0] != n:
                keep = np.zeros(n, bool)
                keep[first] = True
        if self.nb_indexed > 0:
            dup = np.isin(key, self._existing_keys())
            if dup.any():
                keep = ~dup if keep is None else (keep & ~dup)
        if keep is not None:
            if self.verbose > 0:
                print("Discarding %d duplicate samples" % int(n - keep.sum()))
            coarse, fine, nums, key, cell = coarse[keep], fine[keep], nums[keep], key[keep], cell[keep]
        if coarse.shape[0] == 0:
            return
        base = self.nb_indexed
        self._device_add(np.ascontiguousarray(coarse), np.ascontiguousarray(fine), np.arange(base, base + coarse.shape[0], dtype=np.int64))
        self._row_ids.append(nums)
        self._row_cells.append(cell)
        self._row_ids_flat = None
        self._pending_keys.append(key)
        self.nb_indexed += coarse.shape[0]

    def _ids_of_rows(self, rowids):
        if self._row_ids_flat is None:
            self._row_ids_flat = np.concatenate(self._row_ids) if self._row_ids else np.zeros(0, np.int64)
        nums = np.take(self._row_ids_flat, rowids, mode="clip") if self._row_ids_flat.size else np.zeros_like(rowids)
        if self._numeric:
            return nums
        out = np.empty(nums.shape, dtype=object)
        flat = out.reshape(-1)
        for i, v in enumerate(nums.reshape(-1)):
            flat[i] = self._num2id[int(v)]
        return out

    def get_cell(self, cell):
        """search.py:372-382 -- list of (id, LOPQCode) of the cell, in insertion order."""
        rowids, fine = self._handle.get_cell(int(cell[0]), int(cell[1]))
        ids = self._ids_of_rows(rowids)
        co = (int(cell[0]), int(cell[1]))
        return [((i.item() if isinstance(i, np.generic) else i), LOPQCode(co, tuple(int(v) for v in f))) for i, f in zip(ids, fine)]

    # ---- batched search (array level) ---------------------------------------------------------------
    def search_batch(self, X, quota=10, limit=None):
        """Batch form of search(): dict(ids [nq,k], dist [nq,k] float64, coarse [nq,k,2], fine [nq,k,M],
        count [nq], visited [nq]); rows beyond count[q] are padding (ids -1 / None)."""
        if limit is None:
            limit = quota
        k = int(max(1, min(int(limit), max(1, self.nb_indexed))))
        out = self._handle.search(self._query_vector(X), quota, k)
        ids = self._ids_of_rows(out["rowid"])
        if out["count"].size and int(out["count"].min()) < k:
            pad = np.arange(k)[None, :] >= out["count"][:, None]
            if self._numeric:
                ids = np.where(pad, -1, ids)
            else:
                ids[pad] = None
        out["ids"] = ids
        return out

    def stats(self):
        return self._handle.stats()

    # ---- flat persistence of the index (restart without re-encoding) -------------------
    def export_arrays(self):
        """(coarse [n,2] int32, fine [n,M] uint8, ids) of everything indexed, cell by cell in in-cell order: adding them
        back in this order to an empty searcher rebuilds the same index (same retrieval order, same results)."""
        V = self.model.V
        sizes = self._handle.cell_sizes()
        co, fi, rows = [], [], []
        for c in np.nonzero(sizes)[0]:
            r, f = self._handle.get_cell(int(c) // V, int(c) % V)
            co.append(np.tile(np.array([int(c) // V, int(c) % V], np.int32), (r.shape[0], 1)))
            fi.append(f)
            rows.append(r)
        if not rows:
            return np.zeros((0, 2), np.int32), np.zeros((0, self.model.M), np.uint8), np.zeros(0, np.int64)
        return np.concatenate(co), np.concatenate(fi), self._ids_of_rows(np.concatenate(rows))

    def save_index(self, path):
        coarse, fine, ids = self.export_arrays()
        np.savez(path, coarse=coarse, fine=fine, ids=ids, allow_pickle=True)

    def load_index(self, path):
        z = np.load(path, allow_pickle=True)
        self.add_codes((z["coarse"], z["fine"]), z["ids"])


class LOPQSearcherLMDB(LOPQSearcher):
    """search.py:385-499 -- the product's default searcher (searcher_lopqhbase.py:198-206): the index is persisted in an
    LMDB database (key = cell as 2 x native uint16 + str(id) bytes, value = M fine-code bytes, search.py:425-470) and a
    cell is returned in KEY order (search.py:486-496), so ties in distance are ordered by the bytes of str(id) -- not by
    insertion as in LOPQSearcher -- and re-adding an id in a cell overwrites it (txn.put).

    Here LMDB stays the persistent store (when the `lmdb` module is present and a path is given; it is opened exactly as
    the reference does) and the search index is the GPU-resident one, rebuilt in key order after every change.
    `lmdb_path=None` keeps the same semantics in memory only.  A path without the `lmdb` module is an ImportError."""

    def __init__(self, model, lmdb_path=None, id_lambda=int, device=None):
        super(LOPQSearcherLMDB, self).__init__(model, device)
        self.lmdb_path, self.id_lambda = lmdb_path, id_lambda
        self.env = self.index_db = None
        self._items = {}                  # key bytes -> fine codes (uint8 array)
        self._stale = False
        if lmdb_path is not None:
            import lmdb                   # fails loudly when the module is missing
            self.env = lmdb.open(self.lmdb_path, map_size=1024 * 1000000 * 32, max_dbs=1)
            self.index_db = self.env.open_db(b"index")
            with self.env.begin(db=self.index_db) as txn:
                for key, value in txn.cursor():
                    self._items[bytes(key)] = np.frombuffer(bytes(value), np.uint8)
            self._stale = bool(self._items)
        self.nb_indexed = len(self._items)

    # wire format of search.py:425-443
    @staticmethod
    def encode_cell(cell):
        return np.asarray(cell, dtype=np.uint16).tobytes()

    @staticmethod
    def decode_cell(cell_bytes):
        return tuple(int(v) for v in np.frombuffer(cell_bytes, np.uint16))

    @staticmethod
    def encode_fine_codes(fine):
        return np.asarray(fine, dtype=np.uint8).tobytes()

    @staticmethod
    def decode_fine_codes(fine_bytes):
        return tuple(int(v) for v in np.frombuffer(fine_bytes, np.uint8))

    def get_nb_indexed(self):
        self.nb_indexed = len(self._items)
        return self.nb_indexed

    def add_codes(self, codes, ids=None):
        """search.py:445-470."""
        coarse, fine = codes_to_arrays(codes, self.model.M)
        n = coarse.shape[0]
        if ids is None:
            ids = count()
        txn = self.env.begin(db=self.index_db, write=True) if self.env is not None else None
        try:
            for i, item_id in zip(range(n), ids):
                item_id = item_id.item() if isinstance(item_id, np.generic) else item_id
                # py2 `bytes(item_id)` (search.py:463): an id that is already bytes is the key suffix as it is
                key = self.encode_cell(coarse[i]) + (bytes(item_id) if isinstance(item_id, (bytes, bytearray)) else str(item_id).encode())
                self._items[key] = fine[i].copy()
                if txn is not None:
                    txn.put(key, self.encode_fine_codes(fine[i]))
        finally:
            if txn is not None:
                txn.commit()
                self.env.sync()
        self.nb_indexed = len(self._items)
        self._stale = True

    def _rebuild(self):
        """Device index in key order: rows sorted by the full LMDB key (cell bytes, then str(id) bytes)."""
        self._device_clear()
        keys = sorted(self._items)
        n = len(keys)
        self._row_ids, self._row_cells, self._row_ids_flat = [], [], None
        self._numeric, self._id2num, self._num2id = False, {}, []
        if n:
            coarse = np.frombuffer(b"".join(k[:4] for k in keys), np.uint16).reshape(n, 2).astype(np.int32)
            fine = np.stack([self._items[k] for k in keys])
            # the reference is Python 2: the key suffix handed to id_lambda is a `str` (search.py:486-496), and the
            # product passes id_lambda=str (searcher_lopqhbase.py:204-206) -- decode, or str(b'..') would quote it
            self._num2id = [self.id_lambda(k[4:].decode()) for k in keys]
            self._device_add(coarse, fine, np.arange(n, dtype=np.int64))
            self._row_ids = [np.arange(n, dtype=np.int64)]
            self._row_cells = [coarse[:, 0].astype(np.int64) * self.model.V + coarse[:, 1]]
        self._stale = False

    def get_cell(self, cell):
        if self._stale:
            self._rebuild()
        return super(LOPQSearcherLMDB, self).get_cell(cell)

    def search_batch(self, X, quota=10, limit=None):
        if self._stale:
            self._rebuild()
        return super(LOPQSearcherLMDB, self).search_batch(X, quota, limit)

    def get_result_quota(self, x, quota=10):
        if self._stale:
            self._rebuild()
        return super(LOPQSearcherLMDB, self).get_result_quota(x, quota)


# name the plugin accepts in conf key `lopq_searcher` (searcher_lopqhbase.py:198-222)
LOPQSearcherGPU = LOPQSearcher
