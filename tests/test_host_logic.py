"""CPU: host-side pieces that need no GPU -- split / chunk helpers against the oracle and the reference semantics, the
cell -> rank ownership map, the plugin quota rule, and the LMDB wire format of LOPQSearcherLMDB (search.py:425-443)."""
import array

import numpy as np

from oracle import lopq_oracle as orc
from columbiaimagesearch_b200.lopq import utils as gutils
from columbiaimagesearch_b200.lopq.search import LOPQSearcherLMDB, codes_to_arrays
from columbiaimagesearch_b200.lopq.model import LOPQCode
from columbiaimagesearch_b200.sharded import cell_owner
from columbiaimagesearch_b200 import plugin


def test_iterate_splits_matches_oracle():
    x = np.arange(24.0)
    for splits in (1, 2, 3, 4, 8):
        a = [(v.tolist(), s) for v, s in gutils.iterate_splits(x, splits)]
        b = [(v.tolist(), s) for v, s in orc.iterate_splits(x, splits)]
        assert a == b


def test_chunk_ranges_cover_everything():
    # utils.py:164-175: contiguous ranges, remainder goes to the first worker
    for n, p in [(10, 4), (7, 7), (100000, 8), (3, 1)]:
        r = gutils.get_chunk_ranges(n, p)
        assert r[0][0] == 0 and r[-1][1] == n and len(r) == p
        assert all(r[i][1] == r[i + 1][0] for i in range(p - 1))
        assert r[0][1] - r[0][0] == n // p + (n - p * (n // p))


def test_cell_owner_spreads_neighbours():
    for V, world in [(8, 1), (8, 2), (8, 4), (8, 8), (4, 3)]:
        own = cell_owner(V, world).reshape(V, V)
        assert set(np.unique(own).tolist()) == set(range(world))
        if world > 1:
            # cells adjacent in multi-sequence order share c0 or c1: their owners differ
            assert (own[:, :-1] != own[:, 1:]).all() and (own[:-1, :] != own[1:, :]).all()
        counts = np.bincount(own.ravel(), minlength=world)
        assert counts.max() - counts.min() <= max(1, V)


def test_plugin_quota_rule():
    # searcher_lopqhbase.py:838: quota = min(1000 * max_returned, 10000)
    assert plugin.plugin_quota(1) == 1000 and plugin.plugin_quota(5) == 5000
    assert plugin.plugin_quota(10) == 10000 and plugin.plugin_quota(100) == 10000


def test_lmdb_wire_format_matches_reference_encoding():
    # search.py:425-443: cell = array('H').tostring(), fine = array('B').tostring(); key = cell bytes + str(id)
    cell, fine = (3, 517), (0, 255, 17, 4)
    assert LOPQSearcherLMDB.encode_cell(cell) == array.array("H", cell).tobytes()
    assert LOPQSearcherLMDB.encode_fine_codes(fine) == array.array("B", fine).tobytes()
    assert LOPQSearcherLMDB.decode_cell(array.array("H", cell).tobytes()) == cell
    assert LOPQSearcherLMDB.decode_fine_codes(array.array("B", fine).tobytes()) == fine


def test_codes_to_arrays_accepts_reference_shapes():
    codes = [LOPQCode((1, 2), (3, 4, 5, 6)), ((0, 7), (9, 8, 7, 6)), [[5, 5], [1, 1, 1, 1]]]
    coarse, fine = codes_to_arrays(codes, 4)
    assert coarse.dtype == np.int32 and fine.dtype == np.uint8
    assert coarse.tolist() == [[1, 2], [0, 7], [5, 5]] and fine.tolist() == [[3, 4, 5, 6], [9, 8, 7, 6], [1, 1, 1, 1]]
    c2, f2 = codes_to_arrays((coarse, fine), 4)
    assert c2 is not None and np.array_equal(c2, coarse) and np.array_equal(f2, fine)


def test_reconstruct_matches_reference_golden():
    """LOPQModel.reconstruct (model.py:643-671; host NumPy in the mirror package) against the reference's own output."""
    import columbiaimagesearch_b200.lopq as lopq
    from tests.util import load_case
    for name in ("A", "B", "C"):
        z, _ = load_case(name)
        model = lopq.LOPQModel.from_npz(z)
        for i in range(16):
            code = (tuple(int(v) for v in z["db_coarse"][i]), tuple(int(v) for v in z["db_fine"][i]))
            np.testing.assert_allclose(model.reconstruct(code), z["probe_recon"][i], rtol=1e-12, atol=1e-13)
        assert model.get_cell_id_for_coarse_codes((3, 2)) == 2 + 3 * model.V
        assert model.get_coarse_codes_for_cell_id(2 + 3 * model.V) == (3, 2)


def test_train_pca_balances_the_two_halves():
    """model.py:242-287: the kept eigen-directions are permuted by eigenvalue_allocation(2, E), so the two coarse halves
    of the projected space carry comparable variance (plain descending order would put most of it in the first half);
    pca_dims is clamped to D; pca_subsample takes the FIRST rows."""
    from columbiaimagesearch_b200.lopq.train import train_pca, eigenvalue_allocation
    rng = np.random.RandomState(0)
    scales = np.array([9.0, 7.0, 5.0, 4.0, 3.0, 2.5, 2.0, 1.5, 1.2, 1.0, 0.8, 0.5])
    X = rng.randn(20000, 12) * scales
    P, mu = train_pca(X, 8)
    assert P.shape == (12, 8) and mu.shape == (12,)
    Y = (X - mu) @ P
    v = Y.var(0)
    first, second = v[:4].sum(), v[4:].sum()
    # log-products are balanced: both halves hold large and small directions
    assert 0.4 < np.log(v[:4]).sum() / np.log(v[4:]).sum() < 2.5
    assert max(first, second) / min(first, second) < 2.0
    ev = np.sort(np.linalg.eigvalsh(np.cov(X.T)))[-8:]
    np.testing.assert_allclose(np.sort(v), ev, rtol=1e-2)
    assert sorted(eigenvalue_allocation(2, ev).tolist()) == list(range(8))
    # matches the reference's estimator and permutation (oracle port of the same lines)
    P2, _ = train_pca(X, 64)
    assert P2.shape == (12, 12)
    Pa, mua = train_pca(X, 8, subsample=5000)
    Pb, mub = train_pca(X[:5000], 8)
    np.testing.assert_array_equal(Pa, Pb)
    np.testing.assert_array_equal(mua, mub)


def test_lmdb_searcher_persistence_branch_with_a_stand_in_module(monkeypatch):
    """The `lmdb`-backed branch of LOPQSearcherLMDB (the product's default searcher, searcher_lopqhbase.py:198-206) against
    a stand-in `lmdb` module (tests/fake_lmdb.py; the real one is absent in this image): what add_codes writes is the
    reference's wire format (key = 2 x native uint16 cell + str(id) bytes, value = M fine-code bytes, search.py:425-470),
    and a second searcher opened on the same path finds it again, in key order.  No GPU: nothing here searches."""
    import sys
    from tests import fake_lmdb

    class _Model(object):
        V, M = 4, 8
    monkeypatch.setitem(sys.modules, "lmdb", fake_lmdb)
    s = LOPQSearcherLMDB(_Model(), "/fake/lmdb_index_x", id_lambda=str)
    rng = np.random.RandomState(0)
    coarse = rng.randint(0, 4, size=(50, 2)).astype(np.int32)
    fine = rng.randint(0, 256, size=(50, 8)).astype(np.uint8)
    ids = ["sha1_%02d" % i for i in range(50)]
    s.add_codes((coarse, fine), ids)
    s.add_codes((coarse[:5], fine[:5][:, ::-1]), ids[:5])              # re-adding an id in its cell overwrites (txn.put)
    assert s.get_nb_indexed() == 50
    store = fake_lmdb._STORES["/fake/lmdb_index_x"][b"index"]
    key0 = array.array("H", [int(coarse[7, 0]), int(coarse[7, 1])]).tobytes() + b"sha1_07"
    assert store[key0] == array.array("B", fine[7].tolist()).tobytes()   # the reference's encoders, byte for byte
    s2 = LOPQSearcherLMDB(_Model(), "/fake/lmdb_index_x", id_lambda=str)
    assert s2.get_nb_indexed() == 50 and s2._stale
    assert sorted(s2._items) == sorted(store)
    k5 = array.array("H", [int(coarse[2, 0]), int(coarse[2, 1])]).tobytes() + b"sha1_02"
    assert s2._items[k5].tolist() == fine[2][::-1].tolist()


def test_model_protobuf_and_mat_formats(tmp_path):
    """export_proto / load_proto (model.py:748-820) through the package's own wire codec: byte-identical to what the
    protobuf runtime writes with the reference's schema (committed fixture; regenerated live when the reference is
    present), float32 values back as float64 arrays in the reference's container layout; a missing file gives None;
    export_mat / load_mat (model.py:712-746) round-trip."""
    import os
    import columbiaimagesearch_b200.lopq as lopq
    from tests.golden import make_proto
    params = make_proto.small_params()
    model = lopq.LOPQModel(parameters=params)
    golden = open(os.path.join(os.path.dirname(make_proto.__file__), "ref_model_proto.pb"), "rb").read()
    path = str(tmp_path / "m.lopq")
    model.export_proto(path)
    assert open(path, "rb").read() == golden
    model.export_proto(open(path, "wb"))                      # a file object works too
    assert open(path, "rb").read() == golden
    if os.path.exists(make_proto.REF_PB2):
        assert make_proto.serialize_like_reference(params, 3, 4, 16) == golden
    back = lopq.LOPQModel.load_proto(path)
    assert (back.V, back.M, back.subquantizer_clusters, back.num_fine_splits) == (3, 4, 16, 2)
    f32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
    for s in (0, 1):
        assert back.Cs[s].dtype == np.float64 and np.array_equal(back.Cs[s], f32(params[0][s]))
        assert back.Rs[s].shape == (3, 4, 4) and np.array_equal(back.Rs[s], f32(params[1][s]))
        assert back.mus[s].shape == (3, 4) and np.array_equal(back.mus[s], f32(params[2][s]))
        assert len(back.subquantizers[s]) == 2
        for j in (0, 1):
            assert np.array_equal(back.subquantizers[s][j], f32(params[3][s][j]))
    assert lopq.LOPQModel.load_proto(str(tmp_path / "missing.lopq")) is None
    # a partial model (coarse centroids only), as the reference allows while training in stages
    part = lopq.LOPQModel(V=3, M=4, subquantizer_clusters=16, parameters=(params[0], None, None, None))
    part.export_proto(path)
    back = lopq.LOPQModel.load_proto(path)
    assert back.Rs is None and back.subquantizers is None and np.array_equal(back.Cs[1], f32(params[0][1])) and back.M == 4
    mat = str(tmp_path / "m.mat")
    model.export_mat(mat)
    back = lopq.LOPQModel.load_mat(mat)
    assert back.M == 4 and back.V == 3
    for s in (0, 1):
        assert np.array_equal(back.Cs[s], params[0][s]) and np.array_equal(back.Rs[s], params[1][s])
        assert np.array_equal(back.mus[s], params[2][s]) and np.array_equal(back.subquantizers[s][1], params[3][s][1])


def test_model_module_training_helpers_host_path():
    """The module-level helpers of lopq/lopq/model.py that user code imports (eval.py:146: compute_residuals,
    project_residuals_to_local; train_coarse / train_subquantizers / train / train_pca) exist with the reference's
    signatures; device=False keeps the assignments on the host so this runs without a GPU."""
    import inspect
    import columbiaimagesearch_b200.lopq.model as mm
    rng = np.random.RandomState(4)
    X = np.concatenate([rng.randn(300, 6) + 4.0, rng.randn(300, 6) - 4.0])
    C = mm.train_coarse(X, V=2, kmeans_coarse_iters=8, n_init=2, random_state=0, device=False)
    assert C.shape == (2, 6) and abs(abs(C[:, 0]).mean() - 4.0) < 0.5
    res, a = mm.compute_residuals(X, C, device=False)
    assert res.shape == X.shape and set(a.tolist()) == {0, 1} and np.allclose(res, X - C[a])
    assert abs(np.bincount(a)[0] - 300) <= 2
    R, mu, a2, res2 = mm.compute_local_rotations(X, C, 1, device=False)
    assert R.shape == (2, 6, 6) and np.array_equal(a, a2)
    proj = mm.project_residuals_to_local(res, a, R, mu)
    assert np.allclose(np.linalg.norm(proj, axis=1), np.linalg.norm(res - mu[a], axis=1))      # rotations are orthogonal
    subs = mm.train_subquantizers(proj, 2, subquantizer_clusters=8, kmeans_local_iters=5, n_init=1, random_state=1, device=False)
    assert len(subs) == 2 and subs[0].shape == (8, 3)
    assert list(inspect.signature(mm.train_coarse).parameters)[:5] == ["data", "V", "kmeans_coarse_iters", "n_init", "random_state"]
    assert list(inspect.signature(mm.train_subquantizers).parameters)[:6] == ["data", "num_buckets", "subquantizer_clusters", "kmeans_local_iters", "n_init", "random_state"]
    assert mm.eigenvalue_allocation(2, np.array([4.0, 3.0, 2.0, 1.0])).shape == (4,)
