"""Mirror of lopq/lopq/eval.py: the recall harness that defines the headline metric."""
import time

import numpy as np


def compute_all_neighbors(data1, data2=None, just_nn=True, chunk=2048):
    """eval.py:7-38 -- for each row of data1 the index of its nearest row of data2 (``just_nn``), or all indices of data2
    sorted by distance.  Host ground truth (scipy ``cdist`` + ``argmin`` / ``argsort`` row by row as the reference does),
    evaluated in row chunks so that the m1 x m2 distance matrix never has to exist at once."""
    from scipy.spatial.distance import cdist
    data1 = np.asarray(data1)
    data2 = data1 if data2 is None else np.asarray(data2)
    nns = np.zeros(data1.shape[0] if just_nn else (data1.shape[0], data2.shape[0]), dtype=int)
    for a in range(0, data1.shape[0], chunk):
        dists = cdist(data1[a:a + chunk], data2)
        if just_nn:
            nns[a:a + chunk] = np.argmin(dists, axis=1)
        else:
            for i in range(dists.shape[0]):
                nns[a + i] = np.argsort(dists[i])
    return nns


def get_proportion_nns_with_same_coarse_codes(data, model, nns=None):
    """eval.py:41-63 -- share of points whose nearest neighbour falls into the same multi-index cell; the coarse codes of
    all points come from one batched device call (predict_coarse over rows)."""
    data = np.asarray(data)
    if nns is None:
        nns = compute_all_neighbors(data)
    coarse, _ = model._native().encode(data, want_fine=False)
    same = np.all(coarse == coarse[np.asarray(nns)], axis=1)
    return float(np.count_nonzero(same)) / data.shape[0]


def get_subquantizer_distortion(data, model):
    """eval.py:145-161 -- mean squared quantisation error of every sub-quantizer on the locally projected residuals.  Codes
    and projections come from the device (b2l_encode, b2l_project_lut); the reference splits the projection into 8 parts
    whatever M is (its `np.split(pall, 8, axis=1)`), which is M for the models it ships -- M parts here."""
    data = np.asarray(data)
    h = model._native()
    coarse, fine = h.encode(data)
    px, _ = h.project_lut(data, coarse, want_px=True, want_lut=False)
    suball = list(model.subquantizers[0]) + list(model.subquantizers[1])
    ds = px.shape[1] // len(suball)
    out = np.empty(len(suball))
    for j, C in enumerate(suball):
        r = px[:, j * ds:(j + 1) * ds] - np.asarray(C, np.float64)[fine[:, j]]
        out[j] = (r * r).sum()
    return out / data.shape[0]


def get_recall(searcher, queries, indices, thresholds=(1, 10, 100, 1000), normalize=True, verbose=False):
    """eval.py:92-142 -- recall@T: the true nearest neighbour ``indices[i]`` appears among the
    first T results of ``searcher.search(q, quota=thresholds[-1])``; also the mean query time."""
    recall = np.zeros(len(thresholds))
    query_time = 0.0
    for i, d in enumerate(queries):
        nn = indices[i]
        t0 = time.perf_counter()
        results, _ = searcher.search(d, thresholds[-1])
        query_time += time.perf_counter() - t0
        if verbose and i % 50 == 0:
            print("%d/%d queries" % (i, len(queries)))
        for j, res in enumerate(results):
            rid = res[0]
            if rid == nn:
                for k, t in enumerate(thresholds):
                    if j < t:
                        recall[k] += 1
    if normalize:
        N = len(queries)
        return recall / N, query_time / N
    return recall, query_time


def get_recall_batch(searcher, queries, indices, quota, thresholds=(1, 10, 100), batch=1024):
    """Batched form of get_recall for large runs: same definition, one ``search_batch`` call per
    `batch` queries, results cut at thresholds[-1]."""
    k = int(thresholds[-1])
    recall = np.zeros(len(thresholds))
    indices = np.asarray(indices)
    for a in range(0, len(queries), batch):
        out = searcher.search_batch(queries[a:a + batch], quota=quota, limit=k)
        ids = out["ids"]
        hit = ids == indices[a:a + batch, None]
        hit &= np.arange(k)[None, :] < out["count"][:, None]
        rank = np.where(hit.any(axis=1), hit.argmax(axis=1), k)
        for j, t in enumerate(thresholds):
            recall[j] += np.count_nonzero(rank < t)
    return recall / len(queries)


def get_cell_histogram(data, model):
    """eval.py:66-74 -- number of points per multi-index cell."""
    from .utils import compute_codes_arrays
    coarse, _ = compute_codes_arrays(data, model)
    hist = np.zeros(model.V ** 2, dtype=np.int64)
    np.add.at(hist, coarse[:, 0].astype(np.int64) * model.V + coarse[:, 1], 1)
    return hist


def get_proportion_of_reconstructions_with_same_codes(data, model):
    """eval.py:77-89 -- encode -> reconstruct -> encode consistency."""
    from .utils import compute_codes_arrays
    coarse, fine = compute_codes_arrays(data, model)
    recon = np.stack([model.reconstruct((tuple(c), tuple(f))) for c, f in zip(coarse, fine)])
    c2, f2 = compute_codes_arrays(recon, model)
    same = np.all(coarse == c2, axis=1) & np.all(fine == f2, axis=1)
    return float(np.count_nonzero(same)) / len(data)
