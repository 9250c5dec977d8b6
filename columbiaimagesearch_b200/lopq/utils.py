"""Mirror of lopq/lopq/utils.py (the parts on the hot path): iterate_splits, predict_cluster,
compute_codes_parallel / compute_codes_notparallel.  The encode helpers run one batched CUDA
encode instead of a per-row Python loop (utils.py:203-218) or a process pool (utils.py:178-200).
"""
import numpy as np


def iterate_splits(x, splits):
    """utils.py:8-22 -- equal contiguous sub-vectors (py2 integer division); pure slicing."""
    split_size = len(x) // splits
    for split in range(splits):
        start = split * split_size
        yield x[start:start + split_size], split


def concat_new_first(arrs):
    """utils.py:24-29 -- arrays stacked along a new first dimension."""
    return np.concatenate([np.asarray(a)[np.newaxis, ...] for a in arrs], axis=0)


_XVECS = {"f": ("<f4", float), "i": ("<u4", int), "b": ("u1", float)}


def load_xvecs(filename, base_type="f", max_num=None):
    """utils.py:64-99 -- the .fvecs / .ivecs / .bvecs files of corpus-texmex.irisa.fr (every vector: uint32 dimension, then
    D components): an N x D array (float64 for 'f' and 'b', int for 'i', squeezed), read in one pass instead of one
    struct.unpack per component."""
    import os
    code, py_type = _XVECS[base_type]
    isz = np.dtype(code).itemsize
    with open(filename, "rb") as f:
        D = int(np.frombuffer(f.read(4), dtype="<u4")[0])
    rec = 4 + D * isz
    N = os.path.getsize(filename) // rec
    if max_num is None:
        max_num = N
    raw = np.fromfile(filename, dtype=np.uint8, count=max_num * rec).reshape(max_num, rec)
    A = np.ascontiguousarray(raw[:, 4:]).view(code).reshape(max_num, D).astype(py_type)
    return np.squeeze(A)


def save_xvecs(data, filename, base_type="f"):
    """utils.py:102-131 -- the inverse of load_xvecs (rows of any length; a scalar row is a vector of length 1)."""
    code, _ = _XVECS[base_type]
    with open(filename, "wb") as f:
        for d in data:
            d = np.atleast_1d(np.asarray(d))
            f.write(np.array([d.shape[0]], dtype="<u4").tobytes())
            f.write(d.astype(code).tobytes())


def predict_cluster(x, centroids):
    """utils.py:33-53 -- index of the nearest centroid (direct-form squared L2, first minimum),
    returned as the smallest unsigned NumPy integer type that fits.  Evaluated on the device by
    wrapping the centroids as the first coarse split of a throw-away model."""
    from .model import _cluster_handle
    x = np.asarray(x)
    centroids = np.asarray(centroids)
    cid = _cluster_handle(centroids).encode(np.concatenate([x, x])[None, :], want_fine=False)[0][0, 0]
    n = centroids.shape[0]
    if n <= 256:
        return np.uint8(cid)
    if n <= 65536:
        return np.uint16(cid)
    return np.uint32(cid)


def get_chunk_ranges(N, num_procs):
    """utils.py:164-175 -- contiguous row ranges, one per worker (kept for API compatibility)."""
    per_thread = N // num_procs
    allocation = [per_thread] * num_procs
    allocation[0] += N - num_procs * per_thread
    data_ranges = [0]
    for a in allocation:
        data_ranges.append(data_ranges[-1] + a)
    return [(data_ranges[i], data_ranges[i + 1]) for i in range(len(allocation))]


def compute_codes_arrays(data, model):
    """Batched encode: (coarse [n,2] int32, fine [n,M] uint8) -- the array form of compute_codes_*."""
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[None, :]
    return model._native().encode(data)


def codes_from_arrays(coarse, fine):
    """Wrap code arrays as the list of LOPQCode tuples the reference returns (model.py:444, 561)."""
    from .model import LOPQCode
    out = []
    for c, f in zip(coarse.tolist(), fine):
        out.append(LOPQCode(coarse=(np.uint8(c[0]), np.uint8(c[1])) if max(c) < 256 else (np.uint16(c[0]), np.uint16(c[1])),
                            fine=tuple(f)))
    return out


def compute_codes_notparallel(data, model):
    """utils.py:203-218 -- list of LOPQCode in input order."""
    coarse, fine = compute_codes_arrays(data, model)
    return codes_from_arrays(coarse, fine)


def compute_codes_parallel(data, model, num_procs=4):
    """utils.py:178-200 -- same result as compute_codes_notparallel; `num_procs` is accepted for
    signature compatibility (the GPU encodes the whole batch at once).  Returns an iterable."""
    return iter(compute_codes_notparallel(data, model))
