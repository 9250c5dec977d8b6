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
    def __init__(self, model, group=None, device=None, handle=None, backend_device=None, emulate=None):
        """emulate = (world, rank): profiling aid -- behave as that rank of a `world`-way sharded index inside a single
        process (no collective; the caller supplies the global cell sizes through finalize(global_sizes=...))."""
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.model = model
        self._handle = handle if handle is not None else model._new_handle(device)
        self.owner = cell_owner(model.V, self.world)
        if emulate is not None:
            self.owner = cell_owner(model.V, int(emulate[0]))
            self.rank = int(emulate[1])
        self.nb_indexed = 0             # global
        self.nb_local = 0
        self._tdev = backend_device if backend_device is not None else "cuda:%d" % self._handle.device
        self._dirty = False
        # the CUDA library handle gets the stream-ordered pipeline; an injected stand-in (CPU tests) the plain one
        self._pipelined = handle is None
        self._stream = None
        self._pool = {}
        self._pool_next = 0
        # lanes: the handle and its siblings (own stream + workspaces, shared model and index); asynchronous batches go
        # round-robin over them, so the kernels and exchange waits of one batch overlap the scan of another
        self._lanes = []
        self._rr = 0
        self.active_lanes = 0           # 0 = all lanes; n = only the first n (measurement passes that need un-overlapped kernels)

    def _lane(self):
        """(handle, torch stream) of the next asynchronous batch."""
        import torch
        if not self._lanes:
            self._lanes = [[self._handle, None]]
        n = len(self._lanes) if not self.active_lanes else max(1, min(len(self._lanes), int(self.active_lanes)))
        lane = self._lanes[self._rr % n]
        self._rr += 1
        if lane[1] is None:
            lane[1] = torch.cuda.ExternalStream(lane[0].stream(), device=self._tdev)
            lane[0].set_async(True)
        if self._stream is None:
            self._stream = lane[1]                # (also the flag "the asynchronous pipeline has been used")
        return lane

    def enable_pipelining(self, lanes=2):
        """Create `lanes - 1` sibling handles (b2l_create_sibling).  Call after the index is complete (finalize())."""
        if self._dirty:
            self.finalize()
        if not self._lanes:
            self._lanes = [[self._handle, None]]
        while len(self._lanes) < lanes:
            self._lanes.append([self._handle.create_sibling(), None])

    def _sync_lanes(self):
        for h, _ in (self._lanes or [[self._handle, None]]):
            h.sync()

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

    def finalize(self, global_sizes=None):
        """Exchange cell sizes: the quota cut needs the GLOBAL size of every cell on every rank."""
        import torch
        if global_sizes is not None:
            self._handle.set_global_cell_sizes(np.asarray(global_sizes, dtype=np.int64))
            self._dirty = False
            return
        local = self._handle.cell_sizes()
        t = torch.from_numpy(local.copy()).to(self._tdev)
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        self._handle.set_global_cell_sizes(t.cpu().numpy())
        self._dirty = False

    # ---- exchange inside the library (peer-mapped windows, csrc/comm.cuh) ---------------------------------
    def enable_peer_exchange(self, max_nq_home, max_k=16, peers=None):
        """Switch the search to b2l_search_sharded: queries and per-rank top-k records travel through peer-mapped device
        windows (CUDA IPC between the processes of the group), the record all-to-all is fused into the selection kernel,
        and no NCCL call or host round trip remains on the search path.  Collective over the group (the 64-byte IPC
        handles are exchanged once, on the host).  `peers`: the other ShardedLOPQSearcher objects when the ranks are
        handles of ONE process (emulation / tests), listed in rank order; then call it on every one of them."""
        world = self.world if peers is None else len(peers)
        if self._dirty:
            self.finalize()
        if not self._lanes:
            self._lanes = [[self._handle, None]]
        if peers is not None:                 # one process: every window must exist before connecting (single lane)
            for p in peers:
                if not getattr(p, "_peer_init", False):
                    p._handle.comm_init(world, p.rank, int(max_nq_home), int(max_k))
                    p._peer_init = True
            self._handle.comm_connect(pointers=[p._handle.comm_local_ptr() for p in peers])
        else:
            mine = []
            for h, _ in self._lanes:
                h.comm_init(world, self.rank, int(max_nq_home), int(max_k))
                mine.append(h.comm_handle()[0])
            if world > 1:
                allh = [None] * world
                self.dist.all_gather_object(allh, mine, group=self.group)
                for li, (h, _) in enumerate(self._lanes):
                    h.comm_connect(handles=[allh[r][li] for r in range(world)])
                self.dist.barrier(group=self.group)
        self._peer = True
        self._peer_world = world
        for h, _ in self._lanes:
            h.set_async(True)

    def _home_buffers(self, nq_home, k, D):
        import torch
        key = ("home", nq_home, k, D)
        sets = self._pool.get(key)
        if sets is None:
            nbytes = self._handle.sharded_block_bytes(nq_home, k)
            sets = [dict(out_h=torch.zeros(nbytes, dtype=torch.uint8).pin_memory(),
                         q_h=torch.empty((nq_home, D), dtype=torch.float32).pin_memory(),
                         event=torch.cuda.Event()) for _ in range(3)]
            self._pool[key] = sets
        self._pool_next = (self._pool_next + 1) % 3
        return sets[self._pool_next]

    def search_home_async(self, Xhome, quota=10, limit=None):
        """One batch through the in-library exchange.  Xhome [nq_home, D0] float32 (host ndarray or CUDA tensor) is THIS
        rank's slice of the global batch (global query index = rank * nq_home + i; the same nq_home on every rank).
        Returns a pending object whose result() holds the final top-k of the home queries."""
        import torch
        assert getattr(self, "_peer", False), "enable_peer_exchange() first"
        if self._dirty:
            self.finalize()
        if limit is None:
            limit = quota
        k = int(max(1, min(int(limit), max(1, self.nb_indexed))))
        h, stream = self._lane()
        on_dev = hasattr(Xhome, "data_ptr")
        nq_home, D = int(Xhome.shape[0]), int(Xhome.shape[1])
        b = self._home_buffers(nq_home, k, D)
        if on_dev:
            assert Xhome.is_contiguous() and Xhome.dtype == torch.float32
            stream.wait_stream(torch.cuda.current_stream(Xhome.device))
            h.search_sharded(Xhome.data_ptr(), nq_home, quota, k, b["out_h"].data_ptr(), on_device=True)
        else:
            qh = b["q_h"].numpy()
            np.copyto(qh, Xhome, casting="same_kind")
            h.search_sharded(qh, nq_home, quota, k, b["out_h"].data_ptr())
        b["event"].record(stream)
        return _PendingHome(self, b, Xhome, nq_home, k, quota)

    def _redo_home(self, out, Xhome, quota, k, mine):
        """Fallback chain for the queries of a batch that some rank could not certify.  Collective: every rank brings the
        home rows it could not certify (possibly none); all ranks re-run the union with float32 tables, then exactly, over
        the host-driven all-gather protocol, and each rank patches its own rows."""
        D = int(Xhome.shape[1])
        rows = (Xhome[mine].cpu().numpy() if hasattr(Xhome, "data_ptr") else np.asarray(Xhome)[mine]).astype(np.float32).reshape(-1, D)
        if self._peer_world > 1:
            parts = [None] * self._peer_world
            self.dist.all_gather_object(parts, (self.rank, mine.tolist(), rows), group=self.group)
            parts = sorted(parts, key=lambda p: p[0])
        else:
            parts = [(self.rank, mine.tolist(), rows)]
        owner = np.concatenate([np.full(len(p[1]), p[0], np.int64) for p in parts])
        local = np.concatenate([np.asarray(p[1], np.int64) for p in parts])
        Xall = np.concatenate([np.asarray(p[2], np.float32).reshape(-1, D) for p in parts])
        if not local.size:
            return 0, 0
        h = self._handle
        self._sync_lanes()
        h.set_async(False)
        try:
            n32 = nex = 0
            todo = np.arange(local.size)
            for mode in (2, 1):
                if not todo.size:
                    break
                sub = self._gather_merge(np.ascontiguousarray(Xall[todo]), quota, k, mode, int(todo.size))
                sel = owner[todo] == self.rank
                for key in ("rowid", "dist", "coarse", "fine", "count"):
                    out[key][local[todo][sel]] = sub[key][sel]
                if mode == 2:
                    n32 += int(todo.size)
                    todo = todo[sub["certified"] == 0]
                else:
                    nex += int(todo.size)
                    todo = todo[:0]
        finally:
            h.set_async(True)
        return n32, nex

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

    # ---- stream-ordered pipeline (CUDA handle) ---------------------------------------------------------
    def _buffers(self, nq, k, D):
        """One of three rotating buffer sets for (nq, k): records, gathered records, one packed output block on the
        device and its pinned host mirror.  Three sets let two batches be in flight while a third is being read."""
        import torch
        key = (nq, k, D)
        sets = self._pool.get(key)
        if sets is None:
            M = self.model.M
            nk = nq * k
            a256 = lambda n: (n + 255) & ~255
            offs, off = {}, 0
            for name, nbytes in (("rowid", nk * 8), ("dist", nk * 8), ("coarse", nk * 8), ("fine", nk * M), ("count", nq * 4),
                                 ("visited", nq * 4), ("certified", nq)):
                offs[name] = off
                off += a256(nbytes)
            assert off == self._handle.merge_block_bytes(nq, k)
            nbytes = self._handle.records_bytes(nq, k)
            sets = []
            for _ in range(3):
                sets.append(dict(rec=torch.empty(nbytes, dtype=torch.uint8, device=self._tdev),
                                 allrec=torch.empty(nbytes * self.world, dtype=torch.uint8, device=self._tdev) if self.world > 1 else None,
                                 out_h=torch.zeros(off, dtype=torch.uint8).pin_memory(),
                                 q_h=torch.empty((nq, D), dtype=torch.float32).pin_memory(),
                                 event=torch.cuda.Event(), offs=offs))
            self._pool[key] = sets
        self._pool_next = (self._pool_next + 1) % 3
        return sets[self._pool_next]

    def search_batch_async(self, X, quota=10, limit=None):
        """Enqueue one batch on the handle's stream (local search -> all-gather -> merge -> one device-to-host copy)
        and return a pending object; ``.result()`` waits for it (``result(copy=False)``: zero-copy views of the pinned
        result block, valid until two more batches have been enqueued).  Up to two batches may be pending at a time, so the
        host-side launch work of batch i+1 overlaps the device work of batch i.  Every rank must call this (and
        ``result``) in the same order."""
        import torch
        if self._dirty:
            self.finalize()
        if limit is None:
            limit = quota
        k = int(max(1, min(int(limit), max(1, self.nb_indexed))))
        h, stream = self._lane()
        on_dev = hasattr(X, "data_ptr")
        nq, D = int(X.shape[0]), int(X.shape[1])
        b = self._buffers(nq, k, D)
        o = b["offs"]
        base = b["out_h"].data_ptr()
        # All copies from / to the pinned host buffers are enqueued by the library on its own stream (raw pointers):
        # torch only sees that stream for the all-gather and the completion event.
        if on_dev:
            assert X.is_contiguous() and X.dtype == torch.float32
            stream.wait_stream(torch.cuda.current_stream(X.device))
            h.search_local(X.data_ptr(), quota, k, b["rec"].data_ptr(), exact=False, on_device=True, nq=nq)
        else:
            qh = b["q_h"].numpy()
            np.copyto(qh, X, casting="same_kind")
            h.search_local(qh, quota, k, b["rec"].data_ptr(), exact=False)
        if self.world > 1:
            with torch.cuda.stream(stream):
                self.dist.all_gather_into_tensor(b["allrec"], b["rec"], group=self.group)
            allrec = b["allrec"]
        else:
            allrec = b["rec"]
        h.search_merge_block(allrec.data_ptr(), self.world, nq, k, base, on_device=False)      # one copy, same layout as `offs`
        b["event"].record(stream)
        return _PendingSearch(self, b, X, nq, k, quota)

    def search_batch(self, X, quota=10, limit=None):
        """X: ndarray (host) or torch CUDA tensor [nq, D0] float32.  Returns dict(ids = global insertion indices
        [nq,k], dist, coarse, fine, count, visited); identical on every rank."""
        if self._pipelined:
            if not hasattr(X, "data_ptr"):
                X = np.asarray(X)
                X = X[None, :] if X.ndim == 1 else X
                if X.dtype != np.float32:            # float64 queries: the synchronous path keeps their precision
                    return self._search_batch_sync(X, quota, limit)
            return self.search_batch_async(X, quota, limit).result()
        return self._search_batch_sync(X, quota, limit)

    def _search_batch_sync(self, X, quota=10, limit=None):
        if self._dirty:
            self.finalize()
        if limit is None:
            limit = quota
        k = int(max(1, min(int(limit), max(1, self.nb_indexed))))
        # the handle may be in asynchronous mode (pipelined batches, the in-library exchange): drain every lane and switch
        # the main handle to synchronous calls for the host-driven protocol below
        was_async = self._stream is not None or getattr(self, "_peer", False)
        if was_async:
            self._sync_lanes()
            self._handle.set_async(False)
        try:
            return self._search_batch_sync_impl(X, quota, k)
        finally:
            if was_async:
                self._handle.set_async(True)

    def _search_batch_sync_impl(self, X, quota, k):
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
        n32, nex = self._redo_chain(out, (lambda idx: X[idx].cpu().numpy()) if on_dev else (lambda idx: X[idx]), redo, quota, k)
        out["exact_queries"], out["rescan_queries"] = nex, n32
        ids = out["rowid"].copy()
        ids[np.arange(k)[None, :] >= out["count"][:, None]] = -1
        out["ids"] = ids
        return out

    def stats(self):
        return self._handle.stats()

    def lane_stats(self):
        """statistics of every lane (the handle and its siblings)"""
        return [h.stats() for h, _ in (self._lanes or [[self._handle, None]])]

    def reset_stats(self):
        for h, _ in (self._lanes or [[self._handle, None]]):
            h.reset_stats()

    def _redo_chain(self, out, Xsel, redo, quota, k):
        """Uncertified queries (the same set on every rank: the flags come from the gathered buffers) are re-run with
        float32 tables, and what is still uncertified with the float64 full sort; rows of `out` are patched.  `Xsel(idx)`
        returns the host rows of the given query indices.  Returns (#float32 re-runs, #exact re-runs)."""
        n32 = nex = 0
        for mode in (2, 1):
            if not redo.size:
                break
            sub = self._gather_merge(np.ascontiguousarray(Xsel(redo)), quota, k, mode, int(redo.size))
            for key in ("rowid", "dist", "coarse", "fine", "count"):
                out[key][redo] = sub[key]
            if mode == 2:
                n32 += int(redo.size)
                redo = redo[sub["certified"] == 0]
            else:
                nex += int(redo.size)
                redo = redo[:0]
        return n32, nex

    def close(self):
        """Wait for work in flight, then release the buffer pool before the library handle (and its stream) goes away."""
        h = getattr(self, "_handle", None)
        if h is not None and self._stream is not None:
            try:
                self._sync_lanes()
            except Exception:
                pass
        self._pool = {}
        self._stream = None
        for sib, _ in getattr(self, "_lanes", [])[1:]:          # siblings go before the parent handle
            try:
                sib.close()
            except Exception:
                pass
        self._lanes = self._lanes[:1] if getattr(self, "_lanes", None) else []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PendingSearch(object):
    """A batch enqueued by ShardedLOPQSearcher.search_batch_async."""

    def __init__(self, searcher, bufs, X, nq, k, quota):
        self.s, self.b, self.X, self.nq, self.k, self.quota = searcher, bufs, X, nq, k, quota
        self._out = None

    def result(self, copy=True):
        """copy=False returns views of the pinned result block: valid only until two more batches have been enqueued."""
        if self._out is not None:
            return self._out
        s, b, nq, k = self.s, self.b, self.nq, self.k
        M = s.model.M
        b["event"].synchronize()
        raw = b["out_h"].numpy()
        o = b["offs"]
        # zero-copy views of the pinned block: valid until this buffer set comes round again (two more batches enqueued)
        view = lambda name, dt, shape: raw[o[name]:o[name] + int(np.prod(shape)) * np.dtype(dt).itemsize].view(dt).reshape(shape)
        out = dict(rowid=view("rowid", np.int64, (nq, k)), dist=view("dist", np.float64, (nq, k)),
                   coarse=view("coarse", np.int32, (nq, k, 2)), fine=view("fine", np.uint8, (nq, k, M)),
                   count=view("count", np.int32, (nq,)), visited=view("visited", np.int32, (nq,)),
                   certified=view("certified", np.uint8, (nq,)))
        redo = np.nonzero(out["certified"] == 0)[0]
        n32 = nex = 0
        if redo.size:
            X = self.X
            sel = (lambda idx: X[idx].cpu().numpy()) if hasattr(X, "data_ptr") else (lambda idx: np.asarray(X)[idx])
            s._sync_lanes()
            s._handle.set_async(False)
            try:
                n32, nex = s._redo_chain(out, sel, redo, self.quota, k)
            finally:
                s._handle.set_async(True)
        out["exact_queries"], out["rescan_queries"] = nex, n32
        ids = out["rowid"]
        if nq and int(out["count"].min()) < k:
            pad = np.arange(k)[None, :] >= out["count"][:, None]
            ids = ids.copy()
            ids[pad] = -1
            out["dist"][pad] = np.nan
        out["ids"] = ids
        if copy:
            out = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in out.items()}
        self._out = out
        return out


class _PendingHome(object):
    """A batch enqueued by ShardedLOPQSearcher.search_home_async."""

    def __init__(self, searcher, bufs, X, nq, k, quota):
        self.s, self.b, self.X, self.nq, self.k, self.quota = searcher, bufs, X, nq, k, quota
        self._out = None

    def result(self, copy=True):
        if self._out is not None:
            return self._out
        s, b, nq, k = self.s, self.b, self.nq, self.k
        M = s.model.M
        b["event"].synchronize()
        raw = b["out_h"].numpy()
        a256 = lambda n: (n + 255) & ~255
        nk = nq * k
        o, off = {}, 0
        for name, nbytes in (("rowid", nk * 8), ("dist", nk * 8), ("coarse", nk * 8), ("fine", nk * M), ("count", nq * 4),
                             ("visited", nq * 4), ("certified", nq)):
            o[name] = off
            off += a256(nbytes)
        view = lambda name, dt, shape: raw[o[name]:o[name] + int(np.prod(shape)) * np.dtype(dt).itemsize].view(dt).reshape(shape)
        out = dict(rowid=view("rowid", np.int64, (nq, k)), dist=view("dist", np.float64, (nq, k)),
                   coarse=view("coarse", np.int32, (nq, k, 2)), fine=view("fine", np.uint8, (nq, k, M)),
                   count=view("count", np.int32, (nq,)), visited=view("visited", np.int32, (nq,)),
                   certified=view("certified", np.uint8, (nq,)))
        tail = raw[off:off + 36].view(np.int32)
        if int(tail[8]) != 0:
            raise RuntimeError("multi-GPU exchange: the wait for rank %d timed out" % (int(tail[8]) - 1))
        unc = tail[:s._peer_world]
        n32 = nex = 0
        if int(unc.sum()) > 0:                               # somebody needs the fallback chain: every rank takes part
            mine = np.nonzero(out["certified"] == 0)[0]
            n32, nex = s._redo_home(out, self.X, self.quota, k, mine)
        out["exact_queries"], out["rescan_queries"] = nex, n32
        ids = out["rowid"]
        if nq and int(out["count"].min()) < k:
            pad = np.arange(k)[None, :] >= out["count"][:, None]
            ids = ids.copy()
            ids[pad] = -1
            out["dist"][pad] = np.nan
        out["ids"] = ids
        if copy:
            out = {k_: (v.copy() if isinstance(v, np.ndarray) else v) for k_, v in out.items()}
        self._out = out
        return out
