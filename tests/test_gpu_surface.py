"""GPU: the edges of the drop-in boundary that the big parity file does not touch -- the bare helpers of
the reference API (utils.predict_cluster, eval.get_recall, LOPQSearcherBase.compute_distances, the codes TSV loader),
pickles written by the reference's own classes loaded through install_as_lopq(), the process model of the product
(searcher built before fork(), used from forked workers sharing one GPU), LOPQSearcherLMDB with the product's
id_lambda=str, rejected index rows, and the training entry points of LOPQModelPCA as the product calls them."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from oracle import lopq_oracle as orc
from tests.util import load_case, case_inputs, random_model_params, random_data, GOLDEN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lopq():
    import columbiaimagesearch_b200.lopq as lopq
    return lopq


def test_predict_cluster_matches_oracle():
    """utils.py:33-53: first minimum of the direct-form squared distances, smallest unsigned type that fits."""
    lopq = _lopq()
    rng = np.random.RandomState(3)
    for n, d, dt in [(5, 7, np.float64), (256, 16, np.float64), (300, 8, np.float32), (64, 33, np.float32)]:
        C = rng.randn(n, d).astype(dt)
        C[n // 2] = C[1]                                    # an exact tie: the first index must win
        for x in [rng.randn(d).astype(np.float32), C[1].astype(np.float32), rng.randn(d)]:
            got = lopq.utils.predict_cluster(x, C)
            want = orc.predict_cluster(x, C)
            assert int(got) == int(want)
            assert type(got) is type(want)


def test_get_recall_matches_reference_definition():
    """eval.py:92-142 through the mirror's get_recall and get_recall_batch against the oracle's restatement."""
    lopq = _lopq()
    z, omodel = load_case("A")
    _, db, Q, ids = case_inputs("A")
    model = lopq.LOPQModel.from_npz(z)
    s = lopq.LOPQSearcher(model)
    s.add_codes((z["db_coarse"], z["db_fine"]), ids)
    index = orc.ArrayIndex(omodel.V, z["db_coarse"], z["db_fine"], ids)
    # true nearest neighbours by brute force (eval.py:7-38)
    d2 = ((Q[:, None, :].astype(np.float64) - db[None, :, :].astype(np.float64)) ** 2).sum(-1)
    nns = ids[d2.argmin(1)]
    th = (1, 10, 100, 1000)
    want = orc.recall_at(lambda q, quota: orc.search_arrays(omodel, index, q, quota, None)[0], Q, nns, th)
    got, qtime = lopq.eval.get_recall(s, Q, nns, thresholds=th)
    np.testing.assert_array_equal(got, want)
    assert qtime > 0
    got_b = lopq.eval.get_recall_batch(s, Q, nns, quota=th[-1], thresholds=th, batch=16)
    np.testing.assert_array_equal(got_b, want)
    raw, _ = lopq.eval.get_recall(s, Q, nns, thresholds=th, normalize=False)
    np.testing.assert_array_equal(raw, want * len(Q))


def test_compute_distances_and_tsv_loader(tmp_path):
    """search.py:137-177 (per-item ADC distances, left-to-right float64 sum) and :227-243 (codes TSV)."""
    lopq = _lopq()
    params = random_model_params(32, 3, 4, 32, seed=5)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, 300, seed=9)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    path = tmp_path / "codes.tsv"
    with open(path, "w") as f:
        for i in range(300):
            f.write("id%03d\t%s\n" % (i, [[int(v) for v in coarse[i]], [int(v) for v in fine[i]]]))
        f.write("\n")
    s = lopq.LOPQSearcher(model)
    s.add_codes_from_local(str(path))
    o = orc.OracleSearcher(omodel)
    o.add_codes([orc.LOPQCode(tuple(int(v) for v in c), tuple(int(v) for v in f)) for c, f in zip(coarse, fine)],
                ["id%03d" % i for i in range(300)])
    assert s.get_nb_indexed() == o.nb_indexed == 300
    x = db[17]
    items, vis = s.get_result_quota(x, 120)
    oitems, ovis = o.get_result_quota(x, 120)
    assert vis == ovis and [i[0] for i in items] == [i[0] for i in oitems]
    got = s.compute_distances(x, items)
    want = o.compute_distances(x, oitems)
    assert [g[1][0] for g in got] == [w[1][0] for w in want]
    # (the local rotation is a BLAS dgemv in the reference: its summation order is not reproducible to the last bit)
    np.testing.assert_allclose(np.array([g[0] for g in got]), np.array([w[0] for w in want]), rtol=1e-9, atol=1e-13)
    a, _ = s.search(x, 120, 15, with_dists=True)
    b, _ = o.search(x, 120, 15, with_dists=True)
    assert [r.id for r in a] == [r.id for r in b]


@pytest.mark.parametrize("case,fname", [("B", "ref_model_B.pkl"), ("C", "ref_model_C_pca.pkl")])
def test_reference_pickle_loads_through_install_as_lopq(case, fname):
    """Models are stored by pickle (storer/local.py:58,75): a file written by the reference's own classes
    (tests/golden/make_pickles.py) must resolve to this package under the name `lopq` and behave identically."""
    lopq = _lopq()
    lopq.install_as_lopq(force=True)
    import lopq as aliased
    from lopq.search import LOPQSearcher
    from lopq.utils import compute_codes_notparallel
    assert aliased is lopq
    with open(os.path.join(GOLDEN, fname), "rb") as f:
        model = pickle.load(f)
    assert type(model).__module__ == "columbiaimagesearch_b200.lopq.model"
    assert isinstance(model, lopq.LOPQModelPCA if case == "C" else lopq.LOPQModel)
    assert model.V == 4 and model.num_coarse_splits == 2 and model.subquantizer_clusters == 256
    z, omodel = load_case(case)
    _, db, Q, ids = case_inputs(case)
    codes = compute_codes_notparallel(db[:400], model)
    assert [tuple(int(v) for v in c.coarse) for c in codes] == [tuple(r) for r in z["db_coarse"][:400].tolist()]
    assert [tuple(int(v) for v in c.fine) for c in codes] == [tuple(r) for r in z["db_fine"][:400].tolist()]
    s = LOPQSearcher(model)
    s.add_codes((z["db_coarse"], z["db_fine"]), ids)
    assert s.get_nb_indexed() == int(z["nb_indexed"])
    quota, limit = int(z["s1_quota"]), int(z["s1_limit"])
    for i in range(8):
        res, vis = s.search(Q[i], quota=quota, limit=limit, with_dists=True)
        cnt = int(z["s1_counts"][i])
        assert vis == int(z["s1_visited"][i]) and len(res) == cnt
        assert [r.id for r in res] == z["s1_ids"][i][:cnt].tolist()
    # and back: what this package pickles carries no native handle
    blob = pickle.dumps(model, protocol=2)
    again = pickle.loads(blob)
    assert again.predict(db[3]) == model.predict(db[3])


_FORK_SCRIPT = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import columbiaimagesearch_b200.lopq as lopq
from tests.util import load_case, case_inputs
z, _ = load_case("A")
z = {k: z[k] for k in z.files}                     # read everything now: a lazily read .npz shares one file offset across fork()
_, db, Q, ids = case_inputs("A")
model = lopq.LOPQModel.from_npz(z)
s = lopq.LOPQSearcher(model)                       # gunicorn --preload: built in the master ...
s.add_codes((z["db_coarse"], z["db_fine"]), ids)
assert s._h is None                                # ... without touching the GPU
quota, limit = int(z["s1_quota"]), int(z["s1_limit"])
def serve(lo, hi):
    for i in range(lo, hi):
        res, vis = s.search(Q[i], quota=quota, limit=limit, with_dists=True)
        cnt = int(z["s1_counts"][i])
        assert vis == int(z["s1_visited"][i]) and [r.id for r in res] == z["s1_ids"][i][:cnt].tolist(), i
        code = model.predict(db[i])
        assert tuple(int(v) for v in code.fine) == tuple(z["db_fine"][i].tolist())
pids = []
for w in range(3):                                 # three workers share the one GPU
    pid = os.fork()
    if pid == 0:
        try:
            serve(w * 8, w * 8 + 8)
            os._exit(0)
        except BaseException as e:
            sys.stderr.write("worker %%d: %%r\n" %% (w, e))
            os._exit(1)
    pids.append(pid)
bad = [p for p in pids if os.waitpid(p, 0)[1] != 0]
assert not bad, bad
serve(24, 32)                                      # the master itself can still create its own context afterwards
# a searcher that was used BEFORE the fork is rebuilt from its host rows in the child
pid = os.fork()
if pid == 0:
    try:
        serve(32, 40)
        os._exit(0)
    except BaseException as e:
        sys.stderr.write("late worker: %%r\n" %% (e,))
        os._exit(1)
# (forking after CUDA initialisation leaves the child without a usable driver; report what happened, do not require it)
st = os.waitpid(pid, 0)[1]
print("late-fork child status", st)
print("ok")
"""


def test_searcher_built_before_fork_serves_from_workers(tmp_path):
    """Process model of the product: `gunicorn --preload` builds the searcher in the master and forks the workers
    (setup/components/search/docker-compose.yml:67).  The handle is created by the process that first searches."""
    script = tmp_path / "fork_workers.py"
    script.write_text(_FORK_SCRIPT % {"root": ROOT})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


def test_lmdb_searcher_string_ids():
    """The product builds LOPQSearcherLMDB(model, path, id_lambda=str) with sha1 row keys (searcher_lopqhbase.py:204-206):
    ids come back as the `str` they went in as (not "b'..'"), in key order inside a cell; bytes ids are taken as they are."""
    import hashlib
    lopq = _lopq()
    params = random_model_params(32, 3, 4, 32, seed=5)
    omodel = orc.OracleModel(*params)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, 200, seed=9)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    sha = [hashlib.sha1(b"img%d" % i).hexdigest().upper() for i in range(200)]
    s = lopq.LOPQSearcherLMDB(model, None, id_lambda=str)
    s.add_codes((coarse[:120], fine[:120]), sha[:120])
    s.add_codes((coarse[120:], fine[120:]), [v.encode() for v in sha[120:]])          # bytes ids
    assert s.get_nb_indexed() == 200
    res, vis = s.search(db[5], quota=60, limit=20, with_dists=True)
    assert res and all(isinstance(r.id, str) and r.id in sha for r in res)
    # same ranking as the oracle over cells read in key order
    keyorder = sorted(range(200), key=lambda i: (int(coarse[i, 0]), int(coarse[i, 1]), sha[i].encode()))
    o = orc.OracleSearcher(omodel)
    o.add_codes([orc.LOPQCode(tuple(int(v) for v in coarse[i]), tuple(int(v) for v in fine[i])) for i in keyorder],
                [sha[i] for i in keyorder])
    want, ovis = o.search(db[5], quota=60, limit=20, with_dists=True)
    assert vis == ovis and [r.id for r in res] == [r.id for r in want]
    cell = (int(coarse[5, 0]), int(coarse[5, 1]))
    assert [i for i, _ in s.get_cell(cell)] == [i for i, _ in o.get_cell(cell)]


def test_out_of_range_coarse_codes_are_refused_atomically():
    from columbiaimagesearch_b200._native import NativeError
    lopq = _lopq()
    params = random_model_params(32, 3, 4, 32, seed=5)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, 100, seed=9)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    h = model._new_handle()
    h.index_add(coarse[:50], fine[:50])
    bad = coarse[50:].copy()
    bad[7, 1] = 3                                           # V = 3: out of range
    with pytest.raises(NativeError):
        h.index_add(bad, fine[50:])
    assert h.index_size() == 50 and int(h.cell_sizes().sum()) == 50       # nothing of the refused batch stayed
    h.index_add(coarse[50:], fine[50:])
    assert h.index_size() == 100
    # the searcher drops such rows (the reference logs the item and goes on, search.py:343-367) and stays consistent
    s = lopq.LOPQSearcher(model)
    s.add_codes((bad, fine[50:]), np.arange(50, 100))
    assert s.get_nb_indexed() == 49
    res, _ = s.search(db[60], quota=1000)
    assert 57 not in [r.id for r in res] and len(res) == 49


def test_pca_model_training_entry_points():
    """searcher_lopqhbase.py:462 trains on already projected features: fit(x, apply_pca=False, train_pca=False);
    fit_pca refuses to retrain (model.py:878-886); apply_PCA(dtype=float64) is not rounded through float32."""
    lopq = _lopq()
    rng = np.random.RandomState(0)
    X = (rng.randn(3000, 24) * np.linspace(3.0, 0.2, 24)).astype(np.float32)
    m = lopq.LOPQModelPCA(V=2, M=4, subquantizer_clusters=16)
    m.fit_pca(X, pca_dims=16)
    assert m.pca_P.shape == (24, 16)
    with pytest.raises(ValueError):
        m.fit_pca(X, pca_dims=16)
    proj = m.apply_PCA(X)
    assert proj.dtype == np.float32 and proj.shape == (3000, 16)
    p64 = m.apply_PCA(X[:50], dtype=np.float64)
    want = (X[:50].astype(np.float64) - m.pca_mu) @ m.pca_P
    np.testing.assert_allclose(p64, want, rtol=1e-12, atol=1e-12)
    assert np.abs(p64 - p64.astype(np.float32)).max() > 0                    # genuinely float64
    m.fit(proj, verbose=False, apply_pca=False, train_pca=False, n_init=1, kmeans_coarse_iters=3, kmeans_local_iters=3,
          random_state=0)
    assert m.Cs[0].shape == (2, 8) and len(m.subquantizers[0]) == 2
    code = m.predict(X[0])
    assert len(code.coarse) == 2 and len(code.fine) == 4
    # the full path (train PCA + project + train) on a fresh model
    m2 = lopq.LOPQModelPCA(V=2, M=4, subquantizer_clusters=16)
    m2.fit(X, pca_dims=16, n_init=1, kmeans_coarse_iters=3, kmeans_local_iters=3, random_state=0, pca_subsample=2000)
    assert m2.pca_P.shape == (24, 16) and m2.predict(X[1]).fine is not None


def test_training_on_the_gpu_statistical_parity():
    """Training (model.py:339-437) with the nearest-centroid assignments on the GPU: same algorithm as the reference's
    trainer, so on the training data of golden case A the quantisation error (eval.py:145-161's distortion: squared error
    of reconstruct(predict(x))) must be on a par with the reference-trained model's, and GPU- and host-assigned training
    from the same seed must agree closely."""
    lopq = _lopq()
    from columbiaimagesearch_b200.lopq import train as T
    z, omodel = load_case("A")
    train, db, Q, ids = case_inputs("A")
    ref_model = lopq.LOPQModel.from_npz(z)

    def distortion(model, X):
        coarse, fine = lopq.utils.compute_codes_arrays(X, model)
        rec = np.stack([model.reconstruct((tuple(c), tuple(f))) for c, f in zip(coarse, fine)])
        return float(((X.astype(np.float64) - rec) ** 2).sum(1).mean())

    sample = db[:1500]
    d_ref = distortion(ref_model, sample)
    m_gpu = lopq.LOPQModel(V=8, M=16, subquantizer_clusters=256)
    m_gpu.fit(train[:12000].astype(np.float64), n_init=1, kmeans_coarse_iters=10, kmeans_local_iters=10, random_state=0)
    d_gpu = distortion(m_gpu, sample)
    assert d_gpu < 1.25 * d_ref, (d_gpu, d_ref)
    # the device assignment is utils.predict_cluster over rows: identical to the oracle's on every row
    X = train[:4000, :64].astype(np.float64)
    C = np.asarray(m_gpu.Cs[0], dtype=np.float64)
    a_dev = T._assign_device(X, C)
    a_orc = np.array([int(orc.predict_cluster(x, C)) for x in X[:600]])
    assert np.array_equal(a_dev[:600], a_orc)
    a_host = T._assign(X, C, device=False)
    assert (a_dev != a_host).mean() < 1e-3              # matmul-form host distances can only differ in near ties
    m_host = lopq.LOPQModel(V=8, M=16, subquantizer_clusters=256)
    m_host.fit(train[:12000].astype(np.float64), n_init=1, kmeans_coarse_iters=10, kmeans_local_iters=10, random_state=0, device=False)
    d_host = distortion(m_host, sample)
    assert abs(d_gpu - d_host) < 0.1 * d_host, (d_gpu, d_host)
    # recall of a searcher built on the GPU-trained model (eval.get_recall_batch) is on a par with the reference-trained one
    def recall(model):
        s = lopq.LOPQSearcher(model)
        s.add_data(db)
        d2 = ((Q[:, None, :].astype(np.float64) - db[None, :, :].astype(np.float64)) ** 2).sum(-1)
        return lopq.eval.get_recall_batch(s, Q, d2.argmin(1), quota=2000, thresholds=(1, 10), batch=64)
    r_ref, r_gpu = recall(ref_model), recall(m_gpu)
    assert r_gpu[1] >= r_ref[1] - 0.1, (r_gpu, r_ref)


def test_device_kmeans():
    """b2l_kmeans: Lloyd iterations on the device.  The final assignment is utils.predict_cluster over rows (bit-exact with
    the oracle for the returned centroids), the cost is the summed squared error, it does not increase with more
    iterations, it is on a par with the host NumPy Lloyd from the same start, and empty clusters are re-seeded."""
    from columbiaimagesearch_b200 import _native
    from columbiaimagesearch_b200.lopq import train as T
    rng = np.random.RandomState(0)
    centres = rng.randn(12, 10) * 4.0
    X = centres[rng.randint(0, 12, size=6000)] + 0.5 * rng.randn(6000, 10)
    h = _native.Handle()
    C0 = X[rng.choice(6000, size=12, replace=False)].copy()
    reseed = rng.randint(0, 6000, size=(15, 12))
    costs = []
    for iters in (0, 1, 5, 15):
        C, a, cost = h.kmeans(X, C0, iters, reseed[:max(1, iters)])
        costs.append(cost)
        want = np.array([int(orc.predict_cluster(x, C)) for x in X[:400]])
        assert np.array_equal(a[:400], want)
        np.testing.assert_allclose(cost, ((X - C[a]) ** 2).sum(), rtol=1e-9)
    assert all(costs[i + 1] <= costs[i] * (1 + 1e-12) for i in range(3))
    # host Lloyd from the same start
    Ch = C0.copy()
    for _ in range(15):
        ah = T._assign(X, Ch, device=False)
        cnt = np.bincount(ah, minlength=12)
        Ch = np.where((cnt == 0)[:, None], Ch, T._segment_sum(X, ah, 12) / np.maximum(cnt, 1)[:, None])
    cost_h = ((X - Ch[T._assign(X, Ch, device=False)]) ** 2).sum()
    assert costs[-1] <= cost_h * 1.02
    # an empty cluster (a start far away from all data) takes the re-seed row
    C1 = C0.copy()
    C1[3] = 1e6
    C, a, _ = h.kmeans(X, C1, 1, reseed[:1])
    np.testing.assert_array_equal(C[3], X[reseed[0, 3]])
    # the public trainer uses it
    m = T.kmeans(X, 12, 10, np.random.RandomState(1), n_init=2, device=True)
    assert m.shape == (12, 10) and ((X - m[T._assign(X, m, device=False)]) ** 2).sum() < 1.2 * costs[-1] * 1.5


@pytest.mark.parametrize("shape", [(128, 4, 16, 256), (128, 4, 8, 256), (64, 3, 8, 100)], ids=lambda s: "D%d_M%d_K%d" % (s[0], s[2], s[3]))
def test_tensor_core_fine_scores_stay_inside_the_guard_bound(shape):
    """fine_tc.cuh: the scores |c_k|^2/2 - p.c_k the tensor cores produce (three TF32 pieces per float32 product, float32
    accumulation in tensor memory) against float64 scores of the same float64 projections.  The guard accepts a centroid
    only when it is the single score below min + 3E, E = 16 * 2^-24 (|p| + max|c|)^2: the measured error has to stay a
    factor 4 inside E.  Rows of unused centroids (K < 256) must never be able to win."""
    lopq = _lopq()
    D, V, M, K = shape
    params = random_model_params(D, V, M, K, seed=D + M + K)
    model = lopq.LOPQModel(parameters=params)
    db = random_data(params, 4096, seed=11)
    h = model._native()
    ds, m = D // M, M // 2
    worst = 0.0
    for j in (0, M // 2, M - 1):
        sc, px = h.debug_fine_scores(db, j)
        sub = np.asarray(params[3][j // m][j % m], np.float64)
        p = px[:, j * ds:(j + 1) * ds]
        exact = 0.5 * (sub ** 2).sum(1)[None, :] - p @ sub.T
        unit = (np.sqrt((p ** 2).sum(1)) + np.sqrt((sub ** 2).sum(1).max())) ** 2 * 2.0 ** -24
        worst = max(worst, float((np.abs(sc[:, :K] - exact) / unit[:, None]).max()))
        assert np.array_equal(sc[:, :K].argmin(1), exact.argmin(1))
        if K < 256:
            assert sc[:, K:].min() > 1e29
    # adversarial rows: reconstructions of random codes -- every sub-vector projection IS a sub-centroid, so all products of
    # the winning score have one sign (the worst case for an accumulator that truncates)
    rng = np.random.RandomState(1)
    codes = [(tuple(rng.randint(0, V, size=2)), tuple(rng.randint(0, K, size=M))) for _ in range(2048)]
    X = np.stack([model.reconstruct(c) for c in codes])
    for j in (0, M - 1):
        sc, px = h.debug_fine_scores(X, j)
        sub = np.asarray(params[3][j // m][j % m], np.float64)
        p = px[:, j * ds:(j + 1) * ds]
        exact = 0.5 * (sub ** 2).sum(1)[None, :] - p @ sub.T
        unit = (np.sqrt((p ** 2).sum(1)) + np.sqrt((sub ** 2).sum(1).max())) ** 2 * 2.0 ** -24
        worst = max(worst, float((np.abs(sc[:, :K] - exact) / unit[:, None]).max()))
        own = np.array([codes[i][1][j] for i in range(128)])
        assert np.array_equal(sc[:, :K].argmin(1), own)
    assert worst < 4.0, worst


def test_tensor_core_fine_argmin_all_exact_ties_and_list_overflow():
    """Every sub-centroid k has an identical twin k + 128: every sub-vector is an exact tie, nothing can be decided on the
    tensor cores, the list of undecided sub-vectors overflows (its capacity is n M / 64) and the rest is settled in
    place -- all in float64 with the first-minimum rule of utils.py:33-53: codes equal the oracle's, all below 128."""
    lopq = _lopq()
    D, V, M, K, n = 64, 3, 8, 256, 12000
    Cs, Rs, mus, subs = random_model_params(D, V, M, K, seed=77)
    subs = tuple([np.concatenate([c[:128], c[:128]]) for c in half] for half in subs)
    params = (Cs, Rs, mus, subs)
    model = lopq.LOPQModel(parameters=params)
    omodel = orc.OracleModel(*params)
    db = random_data(params, n, seed=9)
    h = model._native()
    h.encode_guard_count(reset=True)
    coarse, fine = h.encode(db)
    assert h.encode_guard_count(reset=True) == n * M
    ocoarse, ofine = orc.encode_batch(omodel, db)
    assert np.array_equal(coarse, ocoarse) and np.array_equal(fine, ofine)
    assert fine.max() < 128


def test_fine_argmin_modes_agree_on_a_trained_model():
    """b2l_set_fine_mode: tensor-core stage (0), float64 only (1), float32 SIMT stage (2) -- same codes on the bench model
    (ds = 8) and on float64 input; the guards of both float32 stages stay rare."""
    lopq = _lopq()
    from columbiaimagesearch_b200 import synth
    z = np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz"))
    model = lopq.LOPQModel.from_npz(z)
    X = synth.dlib_style(60000, 128, seed=31)
    h = model._native()
    out = {}
    try:
        for mode in (0, 1, 2):
            h.set_fine_mode(mode)
            h.encode_guard_count(reset=True)
            out[mode] = h.encode(X) + (h.encode_guard_count(reset=True),)
        h.set_fine_mode(0)
        c64, f64, _ = h.encode(X.astype(np.float64)) + (0,)
    finally:
        h.set_fine_mode(0)
    for mode in (0, 2):
        assert np.array_equal(out[mode][0], out[1][0]) and np.array_equal(out[mode][1], out[1][1]), mode
        assert 0 < out[mode][2] < X.shape[0] * model.M // 200, (mode, out[mode][2])
    oc, of = orc.encode_batch(orc.OracleModel.from_npz(z), X[:3000])
    assert np.array_equal(out[0][0][:3000], oc) and np.array_equal(out[0][1][:3000], of)
    assert np.array_equal(f64[:3000], orc.encode_batch(orc.OracleModel.from_npz(z), X[:3000].astype(np.float64))[1])


def test_model_loaded_from_protobuf_encodes_like_the_oracle(tmp_path):
    """A model that went through export_proto / load_proto (float32 values, model.py:748-820) drives the device exactly as
    the oracle built from the same (rounded) parameters."""
    lopq = _lopq()
    params = random_model_params(64, 4, 8, 256, seed=21)
    path = str(tmp_path / "model.lopq")
    lopq.LOPQModel(parameters=params).export_proto(path)
    model = lopq.LOPQModel.load_proto(path)
    omodel = orc.OracleModel(model.Cs, model.Rs, model.mus, model.subquantizers)
    db = random_data(params, 5000, seed=2)
    oc, of = orc.encode_batch(omodel, db)
    coarse, fine = lopq.utils.compute_codes_arrays(db, model)
    assert np.array_equal(coarse, oc) and np.array_equal(fine, of)


def test_eval_helpers_follow_the_reference_definitions():
    """eval.py:41-63 (nearest neighbours sharing the multi-index cell) and eval.py:145-161 (per-sub-quantizer distortion on
    the locally projected residuals) through the device encode / project calls, against NumPy restatements on the oracle."""
    lopq = _lopq()
    import columbiaimagesearch_b200.lopq.eval as ev
    z, omodel = load_case("A")
    _, db, _, _ = case_inputs("A")
    model = lopq.LOPQModel.from_npz(z)
    X = db[:700]
    nns = ev.compute_all_neighbors(X)
    assert (nns == np.arange(700)).mean() > 0.9               # (a point is its own nearest neighbour, duplicates aside)
    nn2 = ev.compute_all_neighbors(X, just_nn=False)[:, 1]    # the nearest OTHER point
    oc, of = orc.encode_batch(omodel, X)
    want = float(np.count_nonzero(np.all(oc == oc[nn2], axis=1))) / 700
    assert ev.get_proportion_nns_with_same_coarse_codes(X, model, nns=nn2) == want
    px = np.stack([orc.project(omodel, x, c) for x, c in zip(X, oc)])
    suball = list(omodel.subquantizers[0]) + list(omodel.subquantizers[1])
    ds = px.shape[1] // len(suball)
    want_d = np.array([((px[:, j * ds:(j + 1) * ds] - C[of[:, j]]) ** 2).sum() for j, C in enumerate(suball)]) / 700
    np.testing.assert_allclose(ev.get_subquantizer_distortion(X, model), want_d, rtol=1e-9)
