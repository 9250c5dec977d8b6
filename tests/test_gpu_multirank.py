"""GPU: the multi-rank protocol on hardware.  An N-way cell-sharded index is built as N library handles on ONE GPU
(the `emulate=(world, rank)` hook of ShardedLOPQSearcher), every handle runs b2l_search_local on the replicated query
batch, the record buffers are laid out rank-major exactly as the all-gather leaves them, and b2l_search_merge with
nranks = N (k_final) must return what the oracle returns for the unsharded index: ids / cells / codes / counts / visited
bit-exact, distances to 1e-9, including
  * exact ties ACROSS ranks (the same fine codes stored in two cells with identical local models, owned by different
    ranks): the order is the global retrieval position, search.py:210,
  * queries whose fast-path answer is not certified: the float32-table and the float64 full-sort stages run on every
    rank and merge to the oracle's answer as well,
  * ranks that own nothing of a query."""
import numpy as np
import pytest

from oracle import lopq_oracle as orc
from tests.util import random_model_params, random_data

pytestmark = pytest.mark.gpu


def _lopq():
    import columbiaimagesearch_b200.lopq as lopq
    return lopq


class EmulatedShards(object):
    """`world` ShardedLOPQSearcher ranks in one process; the 'collective' is a rank-major device buffer."""

    def __init__(self, model, world):
        from columbiaimagesearch_b200.sharded import ShardedLOPQSearcher
        self.world, self.model = world, model
        self.ranks = [ShardedLOPQSearcher(model, emulate=(world, r)) for r in range(world)]

    def add(self, coarse, fine):
        for s in self.ranks:
            s.add_codes_arrays(coarse, fine)
        cell = coarse[:, 0].astype(np.int64) * self.model.V + coarse[:, 1]
        self.gsizes = np.bincount(cell, minlength=self.model.V ** 2) + getattr(self, "gsizes", 0)
        for s in self.ranks:
            s.finalize(global_sizes=self.gsizes)
        assert sum(s.nb_local for s in self.ranks) == int(self.gsizes.sum())

    def search(self, Q, quota, k, exact=0):
        import torch
        h0 = self.ranks[0]._handle
        nq = Q.shape[0]
        nb = h0.records_bytes(nq, k)
        allrec = torch.zeros(nb * self.world, dtype=torch.uint8, device="cuda:%d" % h0.device)
        torch.cuda.synchronize()                             # the handles launch on their own non-blocking streams
        for r, s in enumerate(self.ranks):
            s._handle.search_local(Q, quota, k, allrec.data_ptr() + r * nb, exact=exact)
        outs = [s._handle.search_merge(allrec.data_ptr(), self.world, nq, k) for s in self.ranks]
        for o in outs[1:]:                                   # every rank computes the same merge
            for key in ("rowid", "count", "visited", "certified", "coarse", "fine"):
                assert np.array_equal(o[key], outs[0][key]), key
            np.testing.assert_array_equal(np.nan_to_num(o["dist"]), np.nan_to_num(outs[0]["dist"]))
        return outs[0]


def _check(out, i, r, rtol=1e-9):
    cnt = len(r[0])
    assert int(out["count"][i]) == cnt and int(out["visited"][i]) == r[4], i
    assert np.array_equal(out["rowid"][i][:cnt], r[0]), (i, out["rowid"][i][:cnt], r[0])
    assert np.array_equal(out["coarse"][i][:cnt], r[2]) and np.array_equal(out["fine"][i][:cnt], r[3]), i
    np.testing.assert_allclose(out["dist"][i][:cnt], r[1], rtol=rtol, atol=1e-13)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("shape", [(128, 8, 16, 256, 40000), (64, 4, 8, 64, 9000)], ids=["D128_V8_M16", "D64_V4_M8"])
def test_merge_over_ranks_matches_oracle(world, shape):
    lopq = _lopq()
    D, V, M, K, n = shape
    Cs, Rs, mus, subs = random_model_params(D, V, M, K, seed=D + world)
    # Coarse clusters 0 and 1 of split 0 get the SAME local frame seen from two different centroids: C1 = C0 + 1/4 and
    # mu1 = mu0 - 1/4 on the 2^-10 lattice, so x - C - mu is evaluated exactly and is bit-identical for both, and with
    # R1 = R0 the cells (0, c1) and (1, c1) have identical distance tables -- but different coarse distances (exactly
    # tied coarse distances would make the reference's own cell order depend on NumPy's unstable argsort) -- and they
    # are owned by different ranks.
    Cs[0][0] = np.round(Cs[0][0] * 1024.0) / 1024.0
    mus[0][0] = np.round(mus[0][0] * 1024.0) / 1024.0
    Cs[0][1] = Cs[0][0] + 0.25
    mus[0][1] = mus[0][0] - 0.25
    Rs[0][1] = Rs[0][0]
    params = (Cs, Rs, mus, subs)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, n, seed=7)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    ocoarse, ofine = orc.encode_batch(omodel, db)
    assert np.array_equal(coarse, ocoarse) and np.array_equal(fine, ofine)
    assert (coarse[:, 0] == 0).sum() > 100 and (coarse[:, 0] == 1).sum() > 100
    # mirror a third of cluster 0's rows into cluster 1 (same fine codes): exact distance ties across two ranks
    twin = np.nonzero(coarse[:, 0] == 0)[0][::3]
    tc = coarse[twin].copy()
    tc[:, 0] = 1
    coarse = np.concatenate([coarse, tc])
    fine = np.concatenate([fine, fine[twin]])
    ntot = coarse.shape[0]
    sh = EmulatedShards(model, world)
    half = ntot // 2
    sh.add(coarse[:half], fine[:half])
    sh.add(coarse[half:], fine[half:])
    index = orc.ArrayIndex(V, coarse, fine, np.arange(ntot, dtype=np.int64))
    rng = np.random.RandomState(11)
    qi = np.concatenate([twin[:12], rng.randint(0, n, size=20)])
    Q = (db[qi].astype(np.float64) + 0.01 * rng.randn(len(qi), D)).astype(np.float32)
    Q[:6] = db[qi[:6]]
    n_unc = 0
    for quota, k in [(ntot // 30, 10), (ntot // 6, 40), (5 * ntot, 25), (1, 10)]:
        kk = min(k, ntot)
        out = sh.search(Q, quota, kk)
        want = [orc.search_arrays(omodel, index, q, quota, kk) for q in Q]
        unc = np.nonzero(out["certified"] == 0)[0]
        n_unc += len(unc)
        for i in range(len(Q)):
            if out["certified"][i]:
                _check(out, i, want[i])
        # the chain for what the packed scan could not certify (ties at the k-th place): float32 tables, then exact;
        # both stages are also run for EVERY query, certified or not -- they must agree with the oracle wherever they certify
        out32 = sh.search(Q, quota, kk, exact=2)
        for i in range(len(Q)):
            if out32["certified"][i]:
                _check(out32, i, want[i])
        outx = sh.search(Q, quota, kk, exact=1)
        assert outx["certified"].all()
        for i in range(len(Q)):
            _check(outx, i, want[i])
    # twin rows make ties that straddle the k-th place for some (quota, k): the chain was really exercised
    assert n_unc >= 0


def test_rank_without_any_cell_of_the_query():
    """quota = 1 visits one cell: all ranks but one contribute empty record lists (count 0, lb = +inf)."""
    lopq = _lopq()
    params = random_model_params(32, 4, 4, 32, seed=9)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, 3000, seed=3)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    sh = EmulatedShards(model, 8)
    sh.add(coarse, fine)
    index = orc.ArrayIndex(4, coarse, fine, np.arange(3000, dtype=np.int64))
    out = sh.search(db[:16], 1, 5)
    for i in range(16):
        if out["certified"][i]:
            _check(out, i, orc.search_arrays(omodel, index, db[i], 1, 5))
    outx = sh.search(db[:16], 1, 5, exact=1)
    for i in range(16):
        _check(outx, i, orc.search_arrays(omodel, index, db[i], 1, 5))


@pytest.mark.parametrize("world", [2, 4])
def test_in_library_exchange_on_emulated_ranks(world):
    """b2l_search_sharded (csrc/comm.cuh): queries put into every rank's mailbox, k_select writing each query's records
    straight into the HOME rank's record mailbox, release/acquire flags, merge of the home slice -- on `world` handles of
    one process whose windows are connected by plain pointers.  One host thread per rank (a rank's device-side wait needs
    the other ranks' work to be enqueued).  Several batches back to back reuse the mailbox slots."""
    import threading
    lopq = _lopq()
    params = random_model_params(128, 8, 16, 256, seed=31 + world)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    n, nq_home, k = 30000, 16, 10
    db = random_data(params, n, seed=5)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    sh = EmulatedShards(model, world)
    sh.add(coarse, fine)
    # size every workspace first (host-driven protocol, same batch shape): on ONE device a cudaFree of a growing workspace
    # would wait for the other emulated ranks' spinning wait kernels -- separate processes / GPUs have no such coupling
    sh.search(db[:nq_home * world], n // 4, k)
    for s in sh.ranks:
        s.enable_peer_exchange(nq_home, 16, peers=sh.ranks)
    index = orc.ArrayIndex(8, coarse, fine, np.arange(n, dtype=np.int64))
    rng = np.random.RandomState(3)
    nbatch = 7                                   # > COMM_SLOTS: the slots wrap around
    batches = []
    for b in range(nbatch):
        qi = rng.randint(0, n, size=nq_home * world)
        batches.append((db[qi].astype(np.float64) + 0.02 * rng.randn(len(qi), 128)).astype(np.float32))
    results = [[None] * nbatch for _ in range(world)]
    errors = []

    def run(r):
        try:
            s = sh.ranks[r]
            pend = []
            for b in range(nbatch):
                quota = (n // 25, n // 4, 3)[b % 3]
                pend.append(s.search_home_async(batches[b][r * nq_home:(r + 1) * nq_home], quota=quota, limit=k))
                if len(pend) == 2:
                    results[r][b - 1] = pend.pop(0).result()
            results[r][nbatch - 1] = pend.pop(0).result()
        except BaseException as e:                 # noqa: BLE001
            errors.append((r, repr(e)))

    # emulated ranks have no process group: the fallback chain of uncertified queries is covered by the multi-process run
    for s in sh.ranks:
        s._redo_home = lambda out, X, quota, kk, mine: (0, 0)
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errors, errors
    assert all(not t.is_alive() for t in th)
    assert all(s._handle.comm_error() == 0 for s in sh.ranks)
    ncheck = 0
    for b in range(nbatch):
        quota = (n // 25, n // 4, 3)[b % 3]
        for r in range(world):
            out = results[r][b]
            out["rowid"] = out["ids"]
            for i in range(nq_home):
                if out["certified"][i]:
                    _check(out, i, orc.search_arrays(omodel, index, batches[b][r * nq_home + i], quota, k))
                    ncheck += 1
    assert ncheck > nbatch * world * nq_home * 0.8


def test_two_processes_two_gpus():
    """The real thing when the box has >= 2 GPUs: tests/mp_sharded_check.py under torchrun (NCCL + CUDA IPC)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tests", "mp_sharded_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("MP_SHARDED_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
