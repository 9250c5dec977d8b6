"""Multi-process, multi-GPU check of the sharded searcher (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mp_sharded_check.py

Every rank builds its cell shard of the same seeded database, then answers batches through
  (a) the in-library exchange (CUDA-IPC windows, b2l_search_sharded: search_home_async) and
  (b) the host-driven protocol (NCCL all-gather of record buffers: search_batch),
and compares ids / cells / codes / counts / visited / distances of ITS home queries (a) and of all queries (b) with the
oracle on the unsharded index -- including duplicated rows, whose exact ties at the k-th place are not certifiable by
the fast scan and go through the collective fallback chain.  Prints `MP_SHARDED_OK <rank>` per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import columbiaimagesearch_b200.lopq as lopq
    from columbiaimagesearch_b200.sharded import ShardedLOPQSearcher
    from oracle import lopq_oracle as orc
    from tests.util import random_model_params, random_data
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    params = random_model_params(128, 8, 16, 256, seed=77)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    n, nq_home, k = 40000, 32, 10
    db = random_data(params, n, seed=5, dup_frac=0.2)            # many exact duplicates: ties straddling the k-th place
    enc = model._new_handle(local)
    coarse, fine = enc.encode(db)
    s = ShardedLOPQSearcher(model, device=local)
    s.add_codes_arrays(coarse, fine)
    s.finalize()
    index = orc.ArrayIndex(8, coarse, fine, np.arange(n, dtype=np.int64))
    rng = np.random.RandomState(3)
    batches = []
    for b in range(6):
        qi = rng.randint(0, n, size=nq_home * world)
        Q = (db[qi].astype(np.float64) + 0.02 * rng.randn(len(qi), 128)).astype(np.float32)
        Q[::3] = db[qi[::3]]                                       # exact database points (often duplicated)
        batches.append(Q)

    def check(out, i, q, quota):
        r = orc.search_arrays(omodel, index, q, quota, k)
        cnt = len(r[0])
        assert int(out["count"][i]) == cnt and int(out["visited"][i]) == r[4], (i, out["count"][i], cnt, out["visited"][i], r[4])
        assert np.array_equal(out["ids"][i][:cnt], r[0]), (i, out["ids"][i][:cnt], r[0])
        assert np.array_equal(out["coarse"][i][:cnt], r[2]) and np.array_equal(out["fine"][i][:cnt], r[3])
        np.testing.assert_allclose(out["dist"][i][:cnt], r[1], rtol=1e-9, atol=1e-13)

    quotas = (n // 25, n // 4, 3)
    # (b) host-driven protocol first (also sizes the workspaces)
    redo_b = 0
    for b, Q in enumerate(batches[:3]):
        out = s.search_batch(Q, quota=quotas[b % 3], limit=k)
        redo_b += out["exact_queries"] + out["rescan_queries"]
        for i in range(Q.shape[0]):
            check(out, i, Q[i], quotas[b % 3])
    # (a) in-library exchange over two lanes (sibling handles), two batches in flight, host and device inputs alternating;
    #     the last batches with every fifth query forced down the collective fallback chain
    s.enable_pipelining(2)
    s.enable_peer_exchange(nq_home, 16)
    pend, outs = [], []
    for b, Q in enumerate(batches):
        if b == 3:
            for h, _ in s._lanes:
                h.debug_force_redo(4)
        home = Q[rank * nq_home:(rank + 1) * nq_home]
        x = torch.from_numpy(home).cuda() if b % 2 else home
        pend.append(s.search_home_async(x, quota=quotas[b % 3], limit=k))
        if len(pend) == 2:
            outs.append(pend.pop(0).result())
    outs += [p.result() for p in pend]
    redo_a = 0
    for b, (Q, out) in enumerate(zip(batches, outs)):
        redo_a += out["exact_queries"] + out["rescan_queries"]
        for i in range(nq_home):
            check(out, i, Q[rank * nq_home + i], quotas[b % 3])
    assert s._handle.comm_error() == 0
    assert redo_a > 0, "the forced fallback chain did not run"
    for h, _ in s._lanes:
        h.debug_force_redo(0)
    t = torch.tensor([redo_a, redo_b], device="cuda:%d" % local)
    dist.all_reduce(t)
    dist.barrier()
    print("MP_SHARDED_OK %d world=%d fallback-chain queries: exchange %d, all-gather %d" % (rank, world, int(t[0]), int(t[1])), flush=True)
    s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
