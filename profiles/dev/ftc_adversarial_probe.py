"""Developer probe: error of the tensor-core fine scores on ADVERSARIAL rows -- reconstructions of random codes, so every
sub-vector projection equals a sub-centroid and all products of the winning score have the same sign (the worst case
for an accumulator that truncates)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200.lopq as lopq
from tests.util import random_model_params
out = []
for name, model in [("dlib128_M16", lopq.LOPQModel.from_npz(np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz")))),
                    ("random_D128_M8", lopq.LOPQModel(parameters=random_model_params(128, 4, 8, 256, seed=9)))]:
    rng = np.random.RandomState(1)
    M, V = model.M, model.V
    n = 2048
    codes = [(tuple(rng.randint(0, V, size=2)), tuple(rng.randint(0, 256, size=M))) for _ in range(n)]
    X = np.stack([model.reconstruct(c) for c in codes])
    h = model._native()
    m = M // 2
    for j in (0, M - 1):
        sc, px = h.debug_fine_scores(X, j)
        sub = np.asarray(model.subquantizers[j // m][j % m], np.float64)
        ds = sub.shape[1]
        p = px[:, j * ds:(j + 1) * ds]
        exact = 0.5 * (sub ** 2).sum(1)[None, :] - p @ sub.T
        unit = (np.sqrt((p ** 2).sum(1)) + np.sqrt((sub ** 2).sum(1).max())) ** 2 * 2.0 ** -24
        ratio = np.abs(sc - exact) / unit[:, None]
        own = np.array([codes[i][1][j] for i in range(128)])
        print(name, "j", j, "max err / unit: all %.3f, own centroid %.3f; signed mean at own %.3f" % (
            ratio.max(), ratio[np.arange(128), own].max(), ((sc - exact) / unit[:, None])[np.arange(128), own].mean()), flush=True)
