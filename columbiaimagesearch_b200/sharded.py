"""Multi-GPU searcher: the inverted lists are sharded by coarse cell over the ranks of one
torch.distributed group (one process per GPU); queries are replicated; every rank ranks the visited
cells it owns; ONE all-gather of the per-rank top-k record buffers (NCCL over NVLink) is followed by
the merge kernel on every rank.  SURVEY.md section 8(e).

Host logic only (ownership map, id bookkeeping, the collective); all arithmetic is in libb200lopq.
`handle` is injectable so the CPU test-suite can drive this class over gloo with a stand-in.
"""
import numpy as np


def cell_owner(V, world):
    """cell -> rank.  Anti-diagonal assignment: cells adjacent in multi-sequence order share c0 or c1,
    so (c0 + c1) mod world spreads a query's first few cells over different ranks."""
    c0, c1 = np.divmod(np.arange(V * V), V)
    return ((c0 + c1) % world).astype(np.int64)


class ShardedLOPQSearcher(object):
    def __init__(self, model, group=None, device=None, handle=None, backend_device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.model = model
        self._handle = handle if handle is not None else model._new_handle(device)
        self.owner = cell_owner(model.V, self.world)
        self.nb_indexed = 0             # global
        self.nb_local = 0
        self._tdev = backend_device if backend_device is not None else "cuda:%d" % self._handle.device
        self._dirty = False

    # ---- index ------------------------------------------------------------------------------------
    def add_codes_arrays(self, coarse, fine, row_base=None):
        """Every rank passes the SAME (coarse [n,2], fine [n,M]) arrays (replicated input); each keeps the
        rows of the cells it owns.  Row ids are global insertion indices, so results are rank-independent."""
        coarse = np.asarray(coarse)
        n = coarse.shape[0]
        base = self.nb_indexed if row_base is None else int(row_base)
        cell = coarse[:, 0].astype(np.int64) * self.model.V + coarse[:, 1]
        mine = self.owner[cell] == self.rank
        rows = np.nonzero(mine)[0]
        if rows.size:
            self._handle.index_add(coarse[rows], np.asarray(fine)[rows], rows.astype(np.int64) + base)
        self.nb_local += int(rows.size)
        self.nb_indexed = base + n
        self._dirty = True

    def add_codes_device(self, coarse_t, fine_t, row_base=None):
        """Same with torch CUDA tensors (coarse int32 [n,2], fine uint8 [n,M]); the filter runs on the device."""
        import torch
        n = coarse_t.shape[0]
        base = self.nb_indexed if row_base is None else int(row_base)
        owner = torch.from_numpy(self.owner).to(coarse_t.device)
        cell = coarse_t[:, 0].long() * self.model.V + coarse_t[:, 1].long()
        rows = torch.nonzero(owner[cell] == self.rank).squeeze(1)
        if rows.numel():
            co = coarse_t[rows].contiguous()
            fi = fine_t[rows].contiguous()
            ids = (rows + base).contiguous()
            torch.cuda.synchronize(coarse_t.device)
            self._handle.index_add_device(co.data_ptr(), fi.data_ptr(), rows.numel(), ids.data_ptr())
        self.nb_local += int(rows.numel())
        self.nb_indexed = base + n
        self._dirty = True

    def finalize(self):
        """Exchange cell sizes: the quota cut needs the GLOBAL size of every cell on every rank."""
        import torch
        local = self._handle.cell_sizes()
        t = torch.from_numpy(local.copy()).to(self._tdev)
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        self._handle.set_global_cell_sizes(t.cpu().numpy())
        self._dirty = False

    # ---- search -----------------------------------------------------------------------------------
    def _gather_merge(self, Q, quota, k, exact, nq, q_ptr=None):
        import torch
        h = self._handle
        nbytes = h.records_bytes(nq, k)
        rec = torch.empty(nbytes, dtype=torch.uint8, device=self._tdev)
        if q_ptr is not None:
            h.search_local(q_ptr, quota, k, rec.data_ptr(), exact=exact, on_device=True, nq=nq)
        else:
            h.search_local(Q, quota, k, rec.data_ptr(), exact=exact)
        if self.world > 1:
            allrec = torch.empty(nbytes * self.world, dtype=torch.uint8, device=self._tdev)
            self.dist.all_gather_into_tensor(allrec, rec, group=self.group)
            if allrec.is_cuda:
                torch.cuda.current_stream(allrec.device).synchronize()
        else:
            allrec = rec
        return h.search_merge(allrec.data_ptr(), self.world, nq, k)

    def search_batch(self, X, quota=10, limit=None):
        """X: ndarray (host) or torch CUDA tensor [nq, D0] float32.  Returns dict(ids = global insertion indices
        [nq,k], dist, coarse, fine, count, visited); identical on every rank."""
        if self._dirty:
            self.finalize()
        if limit is None:
            limit = quota
        k = int(max(1, min(int(limit), max(1, self.nb_indexed))))
        on_dev = hasattr(X, "data_ptr")
        if on_dev:
            assert X.is_contiguous() and X.dim() == 2
            nq = X.shape[0]
            out = self._gather_merge(None, quota, k, False, nq, X.data_ptr())
        else:
            X = np.asarray(X)
            X = X[None, :] if X.ndim == 1 else X
            nq = X.shape[0]
            out = self._gather_merge(X, quota, k, False, nq)
        redo = np.nonzero(out["certified"] == 0)[0]
        if redo.size:           # same set on every rank (the flags are computed from the gathered buffers)
            Xr = X[redo].cpu().numpy() if on_dev else X[redo]
            sub = self._gather_merge(np.ascontiguousarray(Xr), quota, k, True, int(redo.size))
            for key in ("rowid", "dist", "coarse", "fine", "count"):
                out[key][redo] = sub[key]
        out["exact_queries"] = int(redo.size)
        ids = out["rowid"].copy()
        ids[np.arange(k)[None, :] >= out["count"][:, None]] = -1
        out["ids"] = ids
        return out

    def stats(self):
        return self._handle.stats()
