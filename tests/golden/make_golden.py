"""Generate the committed golden vectors by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs tests/golden/case_{A,B,C}.npz.  Inputs are regenerated from seeds by
``columbiaimagesearch_b200.synth.lattice_gmm`` (exact 1/1024-lattice values, bit-identical on any
CPU), so only seeds, the trained model parameters and the reference's outputs are stored.

  case A  D=128 V=8 M=16 K=256, float64 training data  -> all-float64 parameters
  case B  D=32  V=4 M=8  K=256, float32 training data  -> float32 coarse centroids (coarse
          distances and cell ordering computed in float32 by NumPy promotion), duplicate ids
  case C  LOPQModelPCA D0=48 -> 32, V=4 M=4, renorm=True
"""
import io
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from oracle.lopq_oracle import model_to_npz_dict   # noqa: E402
from columbiaimagesearch_b200 import synth         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "A": dict(D=128, V=8, M=16, K=256, train=(20000, 11, np.float64), db=(20000, 12, 0.02), nq=48, qseed=13,
              searches=[(10, None), (2000, 10), (6000, 100), (30000, 20)], pca=None, id_mod=None),
    "B": dict(D=32, V=4, M=8, K=256, train=(8000, 21, np.float32), db=(6000, 22, 0.05), nq=32, qseed=23,
              searches=[(1, None), (500, 10), (3000, 64), (100000, 300)], pca=None, id_mod=5500),
    "C": dict(D=48, V=4, M=4, K=256, train=(8000, 31, np.float32), db=(5000, 32, 0.02), nq=24, qseed=33,
              searches=[(10, None), (800, 10), (100000, 50)], pca=dict(dims=32, renorm=True), id_mod=None),
}


def case_inputs(name):
    """Deterministic inputs of a case (shared with the tests)."""
    c = CASES[name]
    n, seed, dt = c["train"]
    train = synth.lattice_gmm(n, c["D"], seed, dtype=dt)
    n, seed, dup = c["db"]
    db = synth.lattice_gmm(n, c["D"], seed, dup_frac=dup)
    Q, _ = synth.lattice_queries(db, c["nq"], c["qseed"])
    ids = np.arange(db.shape[0], dtype=np.int64)
    if c["id_mod"]:
        ids = ids % c["id_mod"]
    return train, db, Q, ids


def pack_results(all_res, M, k):
    nq = len(all_res)
    ids = -np.ones((nq, k), np.int64)
    dists = np.full((nq, k), np.nan)
    coarse = np.zeros((nq, k, 2), np.int32)
    fine = np.zeros((nq, k, M), np.uint8)
    counts = np.zeros(nq, np.int32)
    visited = np.zeros(nq, np.int32)
    for i, (res, vis) in enumerate(all_res):
        counts[i], visited[i] = len(res), vis
        for j, r in enumerate(res):
            ids[i, j], dists[i, j] = r.id, r.dist
            coarse[i, j], fine[i, j] = r.code[0], r.code[1]
    return ids, dists, coarse, fine, counts, visited


def make(name):
    ref = ref_loader.load()
    c = CASES[name]
    train, db, Q, ids = case_inputs(name)
    with contextlib.redirect_stdout(io.StringIO()):
        if c["pca"]:
            model = ref.LOPQModelPCA(V=c["V"], M=c["M"], subquantizer_clusters=c["K"], renorm=c["pca"]["renorm"])
            model.fit(train, pca_dims=c["pca"]["dims"], n_init=1, random_state=0)
        else:
            model = ref.LOPQModel(V=c["V"], M=c["M"], subquantizer_clusters=c["K"])
            model.fit(train, n_init=1, random_state=0)
    out = model_to_npz_dict(model)
    M = model.M

    codes = ref.utils.compute_codes_notparallel(db, model)
    out["db_coarse"] = np.array([cd.coarse for cd in codes], np.int32)
    out["db_fine"] = np.array([cd.fine for cd in codes], np.uint8)

    searcher = ref.LOPQSearcher(model)
    searcher.add_codes(codes, ids)
    out["nb_indexed"] = np.int64(searcher.get_nb_indexed())
    with contextlib.redirect_stdout(io.StringIO()):
        for si, (quota, limit) in enumerate(c["searches"]):
            res = []
            for q in Q:
                r, vis = searcher.search(q, quota=quota, limit=limit, with_dists=True)
                res.append((list(r), vis))
            k = max(1, max(len(r) for r, _ in res))
            for key, arr in zip(("ids", "dists", "coarse", "fine", "counts", "visited"), pack_results(res, M, k)):
                out["s%d_%s" % (si, key)] = arr
            out["s%d_quota" % si] = np.int64(quota)
            out["s%d_limit" % si] = np.int64(-1 if limit is None else limit)
    out["n_searches"] = np.int64(len(c["searches"]))

    # intermediate quantities for the first few queries (post-PCA vectors for case C)
    nprobe = 8
    xs = [model.apply_PCA(q) if c["pca"] else q for q in Q[:nprobe]]
    if c["pca"]:
        out["pca_out"] = np.stack(xs)
    ms_d, ms_c = [], []
    for x in xs:
        seq = list(ref.multisequence(x, model.Cs))
        ms_d.append([float(d) for d, _ in seq])
        ms_c.append([[int(cc[0]), int(cc[1])] for _, cc in seq])
    out["ms_dists"] = np.array(ms_d, np.float64)
    out["ms_cells"] = np.array(ms_c, np.int32)
    probe_coarse = np.array([[int(v) for v in model.predict_coarse(x)] for x in xs], np.int32)
    probe_coarse[1::2] = (probe_coarse[1::2] + 1) % model.V      # also non-nearest local frames
    out["probe_coarse"] = probe_coarse
    out["probe_project"] = np.stack([model.project(x, tuple(pc)) for x, pc in zip(xs, probe_coarse)])
    out["probe_lut"] = np.stack([np.stack(model.get_subquantizer_distances(x, tuple(pc)))
                                 for x, pc in zip(xs, probe_coarse)])
    out["probe_recon"] = np.stack([model.reconstruct(codes[i]) for i in range(16)])
    # float64 queries (same values) -- with float32 coarse centroids this switches the coarse
    # arithmetic to float64 by NumPy promotion
    if not c["pca"]:
        res = []
        with contextlib.redirect_stdout(io.StringIO()):
            quota, limit = c["searches"][1]
            for q in Q[:16].astype(np.float64):
                r, vis = searcher.search(q, quota=quota, limit=limit, with_dists=True)
                res.append((list(r), vis))
        for key, arr in zip(("ids", "dists", "coarse", "fine", "counts", "visited"), pack_results(res, M, limit)):
            out["f64_%s" % key] = arr
    path = os.path.join(HERE, "case_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024.0), "nb_indexed", int(out["nb_indexed"]))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or sorted(CASES)):
        make(nm)
