"""CPU, world_size 2 over gloo: host logic of the cell-sharded searcher (columbiaimagesearch_b200/sharded.py).

The CUDA library cannot run here, so the per-rank handle is a STAND-IN that ranks its shard with the oracle
(test infrastructure; the product never does this).  What is under test is everything around the handle:
the cell -> rank ownership map, global row ids, the all-reduce of per-cell sizes (the quota cut of
search.py:128-133 counts GLOBAL cell sizes), the single all-gather of the per-rank top-k record buffers, the
rank-major merge, and the uncertified -> exact re-run protocol.  Expected results: the single-process oracle
over the whole database (ids, distances, visited counts identical on every rank).
"""
import ctypes
import os
import socket

import numpy as np
import pytest

from oracle import lopq_oracle as orc
from tests.util import load_case, case_inputs

REC_Q = 16      # per-query header bytes: count int32, visited int32, uncertified int32, pad
REC_E = 24      # per-entry bytes: dist f64, pos i64, rowid i64


class OracleShardHandle(object):
    """Stand-in for _native.Handle on one rank: same method names/arguments as sharded.py uses."""

    device = 0

    def __init__(self, omodel, force_uncertified=()):
        self.m = omodel
        self.V, self.M = omodel.V, omodel.M
        self.cells = {}
        self.gsize = None
        self.force_uncertified = set(force_uncertified)
        self.calls = []

    def index_add(self, coarse, fine, rowids):
        for c, f, r in zip(np.asarray(coarse), np.asarray(fine), np.asarray(rowids)):
            self.cells.setdefault(int(c[0]) * self.V + int(c[1]), []).append((int(r), tuple(int(v) for v in f)))

    def cell_sizes(self):
        s = np.zeros(self.V * self.V, np.int64)
        for c, rows in self.cells.items():
            s[c] = len(rows)
        return s

    def set_global_cell_sizes(self, sizes):
        self.gsize = np.asarray(sizes, dtype=np.int64).copy()

    def records_bytes(self, nq, k):
        return nq * (REC_Q + k * REC_E)

    def stats(self):
        return {}

    def search_local(self, Q, quota, k, records_ptr, exact=False, on_device=False, nq=None):
        assert not on_device
        Q = np.asarray(Q)
        nq = Q.shape[0]
        self.calls.append(("local", nq, int(exact)))
        buf = (ctypes.c_uint8 * self.records_bytes(nq, k)).from_address(records_ptr)
        raw = np.frombuffer(buf, dtype=np.uint8)
        raw[:] = 0
        for qi, x in enumerate(Q):
            got, visited, ents = 0, 0, []
            memo = [{}, {}]
            for _, cell in orc.multisequence(x, self.m.Cs):
                cid = int(cell[0]) * self.V + int(cell[1])
                rows = self.cells.get(cid, [])
                if rows:
                    for s in (0, 1):
                        if cell[s] not in memo[s]:
                            memo[s][cell[s]] = orc.subquantizer_distances(self.m, x, cell, coarse_split=s)
                    lut = memo[0][cell[0]] + memo[1][cell[1]]
                    for i, (rid, f) in enumerate(rows):
                        d = 0 + lut[0][f[0]]
                        for j in range(1, self.M):
                            d = d + lut[j][f[j]]
                        ents.append((float(d), got + i, rid))
                got += int(self.gsize[cid])
                visited += 1
                if got >= quota:
                    break
            ents.sort(key=lambda e: (e[0], e[1]))
            ents = ents[:k]
            o = qi * (REC_Q + k * REC_E)
            # stage 0 (default scan): the forced queries come back uncertified; stage 2 (float32-table re-run): the second
            # query of the re-run still does; stage 1 (exact) certifies everything
            unc = (int(exact) == 0 and qi in self.force_uncertified) or (int(exact) == 2 and qi == 1)
            hdr = np.array([len(ents), visited, int(unc), 0], np.int32)
            raw[o:o + REC_Q] = hdr.view(np.uint8)
            if ents:
                e = np.zeros(len(ents), dtype=[("d", "<f8"), ("p", "<i8"), ("r", "<i8")])
                e["d"], e["p"], e["r"] = zip(*ents)
                raw[o + REC_Q:o + REC_Q + len(ents) * REC_E] = e.view(np.uint8)

    def search_merge(self, records_all_ptr, nranks, nq, k):
        self.calls.append(("merge", nranks, nq))
        rb = self.records_bytes(nq, k)
        buf = (ctypes.c_uint8 * (rb * nranks)).from_address(records_all_ptr)
        raw = np.frombuffer(buf, dtype=np.uint8)
        out = dict(rowid=np.full((nq, k), -1, np.int64), dist=np.full((nq, k), np.nan), coarse=np.zeros((nq, k, 2), np.int32),
                   fine=np.zeros((nq, k, self.M), np.uint8), count=np.zeros(nq, np.int32), visited=np.zeros(nq, np.int32),
                   certified=np.ones(nq, np.uint8))
        for qi in range(nq):
            ents = []
            for r in range(nranks):
                o = r * rb + qi * (REC_Q + k * REC_E)
                cnt, vis, unc, _ = raw[o:o + REC_Q].view(np.int32)
                e = raw[o + REC_Q:o + REC_Q + cnt * REC_E].view([("d", "<f8"), ("p", "<i8"), ("r", "<i8")])
                ents += [(float(a), int(b), int(c)) for a, b, c in zip(e["d"], e["p"], e["r"])]
                out["visited"][qi] = vis
                if unc:
                    out["certified"][qi] = 0
            ents.sort(key=lambda t: (t[0], t[1]))
            ents = ents[:k]
            out["count"][qi] = len(ents)
            for j, (d, _, rid) in enumerate(ents):
                out["rowid"][qi, j], out["dist"][qi, j] = rid, d
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, retq):
    import torch.distributed as dist
    from columbiaimagesearch_b200.sharded import ShardedLOPQSearcher, cell_owner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, omodel = load_case("B")
        _, db, Q, _ = case_inputs("B")
        coarse, fine = z["db_coarse"], z["db_fine"]
        n = coarse.shape[0]
        h = OracleShardHandle(omodel, force_uncertified=(1, 5))
        s = ShardedLOPQSearcher(omodel, handle=h, backend_device="cpu")
        # two add calls: row ids must stay global insertion indices
        s.add_codes_arrays(coarse[:n // 2], fine[:n // 2])
        s.add_codes_arrays(coarse[n // 2:], fine[n // 2:])
        owner = cell_owner(omodel.V, world)
        cell = coarse[:, 0].astype(np.int64) * omodel.V + coarse[:, 1]
        assert s.nb_indexed == n and s.nb_local == int((owner[cell] == rank).sum())
        assert sorted(set(owner.tolist())) == list(range(world))
        nq = 12
        out = s.search_batch(Q[:nq], quota=700, limit=10)
        gs = np.bincount(cell, minlength=omodel.V ** 2)
        assert np.array_equal(h.gsize, gs), "global cell sizes after the all-reduce"
        retq.put((rank, out["ids"], out["dist"], out["visited"], out["count"], (out["exact_queries"], out["rescan_queries"]), h.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_search_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    retq = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, retq)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = retq.get(timeout=240)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    z, omodel = load_case("B")
    _, db, Q, _ = case_inputs("B")
    index = orc.ArrayIndex(omodel.V, z["db_coarse"], z["db_fine"], np.arange(z["db_coarse"].shape[0], dtype=np.int64))
    nq = 12
    for rank in range(world):
        ids, dist_, visited, count, exact_q, calls = res[rank]
        for i in range(nq):
            e_ids, e_d, _, _, e_vis = orc.search_arrays(omodel, index, Q[i], 700, 10)
            assert count[i] == len(e_ids)
            assert np.array_equal(ids[i][:len(e_ids)], e_ids), "rank %d query %d ids" % (rank, i)
            assert np.array_equal(dist_[i][:len(e_ids)], e_d), "rank %d query %d dists" % (rank, i)
            assert visited[i] == e_vis
        # the two queries flagged uncertified were re-run with float32 tables, the one still uncertified exactly
        assert exact_q == (1, 2)
        assert calls == [("local", nq, 0), ("merge", world, nq), ("local", 2, 2), ("merge", world, 2), ("local", 1, 1), ("merge", world, 1)]
    # identical on every rank
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
