"""CPU, build container only: pin the oracle against the reference's own code on fresh inputs."""
import contextlib
import io

import numpy as np
import pytest

from oracle import lopq_oracle as orc
from oracle import ref_loader
from columbiaimagesearch_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def setup():
    ref = ref_loader.load()
    train = synth.lattice_gmm(4000, 32, 101, dtype=np.float32)
    model = ref.LOPQModel(V=4, M=4, subquantizer_clusters=64)
    model.fit(train, n_init=1, random_state=1)
    om = orc.OracleModel(model.Cs, model.Rs, model.mus, model.subquantizers)
    db = synth.lattice_gmm(3000, 32, 102, dup_frac=0.05)
    Q, _ = synth.lattice_queries(db, 20, 103)
    return ref, model, om, db, Q


def test_codes_luts_cells_bit_identical(setup):
    ref, model, om, db, Q = setup
    rc = ref.utils.compute_codes_notparallel(db[:800], model)
    oc = orc.compute_codes(om, db[:800])
    assert [tuple(map(int, c.coarse)) + tuple(map(int, c.fine)) for c in rc] == \
           [tuple(map(int, c.coarse)) + tuple(map(int, c.fine)) for c in oc]
    for q in Q[:6]:
        rs = list(ref.multisequence(q, model.Cs))
        os_ = list(orc.multisequence(q, om.Cs))
        assert [(float(d), tuple(map(int, c))) for d, c in rs] == [(float(d), tuple(map(int, c))) for d, c in os_]
        assert type(rs[0][0]) is type(os_[0][0])  # float32 cell distances under float32 centroids
        c = model.predict_coarse(q)
        for split in (None, 0, 1):
            a = np.stack(model.get_subquantizer_distances(q, c, coarse_split=split))
            b = np.stack(orc.subquantizer_distances(om, q, c, coarse_split=split))
            assert np.array_equal(a, b)  # same ufunc order => same float64 bits
        assert np.array_equal(model.reconstruct(rc[0]), orc.reconstruct(om, oc[0]))


def test_search_identical(setup):
    ref, model, om, db, Q = setup
    ids = np.arange(db.shape[0]) % 2800
    rcodes = ref.utils.compute_codes_notparallel(db, model)
    rs = ref.LOPQSearcher(model)
    rs.add_codes(rcodes, ids)
    os_ = orc.OracleSearcher(om)
    os_.add_codes(orc.compute_codes(om, db), ids)
    assert rs.get_nb_indexed() == os_.nb_indexed
    coarse = np.array([c.coarse for c in rcodes], np.int32)
    fine = np.array([c.fine for c in rcodes], np.uint8)
    index = orc.ArrayIndex(om.V, coarse, fine, ids)
    for quota, limit in [(1, None), (50, 7), (700, None), (10 ** 6, 40)]:
        for q in Q:
            with contextlib.redirect_stdout(io.StringIO()):
                a, va = rs.search(q, quota, limit, with_dists=True)
            b, vb = os_.search(q, quota, limit, with_dists=True)
            assert va == vb and len(a) == len(b)
            assert [(x.id, float(x.dist)) for x in a] == [(x.id, float(x.dist)) for x in b]
            r = orc.search_arrays(om, index, q, quota, limit)
            assert vb == r[4] and [x.id for x in b] == r[0].tolist()
            assert np.array_equal(np.array([x.dist for x in b]), r[1])


def test_pca_model_identical(setup):
    ref = setup[0]
    train = synth.lattice_gmm(3000, 24, 111, dtype=np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.LOPQModelPCA(V=2, M=2, subquantizer_clusters=32, renorm=True)
        m.fit(train, pca_dims=16, n_init=1, random_state=2)
    om = orc.OracleModel(m.Cs, m.Rs, m.mus, m.subquantizers, m.pca_P, m.pca_mu, m.renorm)
    x = train[:50]
    assert np.array_equal(m.apply_PCA(x), orc.apply_pca(om, x))
    for row in x[:20]:
        a, b = m.predict(row), orc.predict(om, row)
        assert tuple(map(int, a.coarse + a.fine)) == tuple(map(int, b.coarse + b.fine))


def test_train_pca_matches_reference():
    """The training mirror's PCA (host NumPy, not on the hot path) against the reference's train_pca (model.py:242-287):
    same subspace, same eigenvalue_allocation permutation of the columns, same first-rows subsample."""
    from columbiaimagesearch_b200.lopq.train import train_pca, eigenvalue_allocation
    ref = ref_loader.load()
    rng = np.random.RandomState(5)
    X = rng.randn(3000, 20) * np.linspace(4.0, 0.3, 20)
    with contextlib.redirect_stdout(io.StringIO()):
        rp, rdims = ref.model.train_pca(X, 12, 2000)
    P, mu = train_pca(X, 12, 2000)
    assert rdims == 12 and P.shape == rp["P"].shape
    np.testing.assert_allclose(mu, rp["mu"], rtol=0, atol=1e-12)
    sign = np.sign((P * rp["P"]).sum(0))
    np.testing.assert_allclose(P * sign, rp["P"], rtol=0, atol=1e-8)
    assert eigenvalue_allocation(2, rp["E"]).tolist() == ref.model.eigenvalue_allocation(2, rp["E"]).tolist()


def test_mirror_compute_all_neighbors_matches_reference():
    """eval.py:7-38 (the ground-truth helper of the recall harness): the mirror's chunked evaluation returns the
    reference's indices, nearest-only and fully ranked."""
    import columbiaimagesearch_b200.lopq.eval as ev
    ref = ref_loader.load()
    rng = np.random.RandomState(0)
    a, b = rng.randn(300, 16), rng.randn(500, 16)
    b[17] = b[3]                                              # a tie: argmin / argsort order of the reference
    assert np.array_equal(ev.compute_all_neighbors(a, b, chunk=64), ref.eval.compute_all_neighbors(a, b))
    assert np.array_equal(ev.compute_all_neighbors(a, chunk=77), ref.eval.compute_all_neighbors(a))
    assert np.array_equal(ev.compute_all_neighbors(a[:40], b[:60], just_nn=False, chunk=7),
                          ref.eval.compute_all_neighbors(a[:40], b[:60], just_nn=False))


def test_mirror_xvecs_io_matches_reference(tmp_path):
    """utils.py:64-131 (.fvecs / .ivecs / .bvecs of corpus-texmex.irisa.fr): files written by either side read back
    identically by both, max_num and the squeeze of single-column data included."""
    import columbiaimagesearch_b200.lopq.utils as mu
    ref = ref_loader.load()
    rng = np.random.RandomState(2)
    cases = {"f": rng.randn(7, 5).astype(np.float32), "i": rng.randint(0, 2 ** 31 - 1, size=(6, 3)), "b": rng.randint(0, 256, size=(9, 4))}
    for bt, data in cases.items():
        a, b = str(tmp_path / ("ref." + bt)), str(tmp_path / ("mir." + bt))
        ref.utils.save_xvecs(data, a, base_type=bt)
        mu.save_xvecs(data, b, base_type=bt)
        assert open(a, "rb").read() == open(b, "rb").read()
        for mx in (None, 3):
            rb = mu.load_xvecs(b, base_type=bt, max_num=mx)
            if bt == "b":             # (the reference's own 'b' reader unpacks the 4-byte dimension with format 'B' and raises)
                assert rb.dtype == np.float64 and np.array_equal(rb, data[:mx].astype(np.float64))
                continue
            ra = ref.utils.load_xvecs(a, base_type=bt, max_num=mx)
            assert ra.dtype == rb.dtype and ra.shape == rb.shape and np.array_equal(ra, rb), (bt, mx)
    one = str(tmp_path / "one.f")
    mu.save_xvecs(np.arange(5, dtype=np.float32), one)          # scalar rows: vectors of length 1
    assert np.array_equal(mu.load_xvecs(one), ref.utils.load_xvecs(one)) and mu.load_xvecs(one).shape == (5,)
    assert np.array_equal(mu.concat_new_first([np.ones((2, 3)), np.zeros((2, 3))]), ref.utils.concat_new_first([np.ones((2, 3)), np.zeros((2, 3))]))
