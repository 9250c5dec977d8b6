"""Batched form of the search loop of the cufacesearch plugin (cufacesearch/cufacesearch/searcher/
searcher_lopqhbase.py:838-857 and :964-970): the plugin normalises one feature at a time and calls
``self.searcher.search(normed_feat, quota=quota, limit=max_returned, with_dists=True)`` in a Python loop; here all the
features of a request go to the GPU as one batch.  Same quota rule, same normalisation, same per-feature return value
``(results, visited)`` with ``Result(id, code, dist)`` items."""
from collections import namedtuple

import numpy as np

from .lopq.model import LOPQCode

Result = namedtuple("Result", ["id", "code", "dist"])


def plugin_quota(max_returned):
    """searcher_lopqhbase.py:838."""
    return min(1000 * max_returned, 10000)


def search_from_feats_batch(searcher, feats, max_returned=100, quota=None):
    """feats: iterable of 1-D feature vectors (any float dtype, not normalised).  Returns one ``(results, visited)`` per
    feature, exactly what the per-feature ``searcher.search(..., with_dists=True)`` loop of the plugin produces."""
    feats = [np.asarray(f) for f in feats]
    if not feats:
        return []
    quota = plugin_quota(max_returned) if quota is None else quota
    # searcher_lopqhbase.py:854-855: feat / ||feat||, squeezed
    X = np.stack([np.squeeze(f / np.linalg.norm(f)) for f in feats])
    out = searcher.search_batch(X, quota=quota, limit=max_returned)
    res = []
    for i in range(X.shape[0]):
        n = int(out["count"][i])
        items = []
        for j in range(n):
            rid = out["ids"][i][j]
            rid = rid.item() if isinstance(rid, np.generic) else rid
            code = LOPQCode(coarse=(int(out["coarse"][i, j, 0]), int(out["coarse"][i, j, 1])),
                            fine=tuple(int(v) for v in out["fine"][i, j]))
            items.append(Result(rid, code, float(out["dist"][i, j])))
        res.append((items, int(out["visited"][i])))
    return res
