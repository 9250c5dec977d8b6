"""Small workload that touches the kernels added at the end of round 2 -- k_coarse_big (two rows per thread) + k_coarse_redo,
k_fine_tc<16> + k_fine_redo, k_slot_prep + k_presel_emit, k_cand_dist + k_select_emit -- for compute-sanitizer:
    compute-sanitizer --tool memcheck python profiles/dev/sanitize_new_kernels.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200.lopq as lopq  # noqa: E402
from oracle import lopq_oracle as orc  # noqa: E402
from tests.util import random_model_params, random_data  # noqa: E402

D, V, M, K, n = 128, 300, 8, 256, 4096
params = random_model_params(D, V, M, K, seed=3)
model = lopq.LOPQModel(parameters=params)
omodel = orc.OracleModel(*params)
db = random_data(params, n, seed=4, dup_frac=0.05)
coarse, fine = lopq.utils.compute_codes_arrays(db, model)
oc, of = orc.encode_batch(omodel, db[:300])
assert np.array_equal(coarse[:300], oc) and np.array_equal(fine[:300], of)
s = lopq.LOPQSearcher(model)
s.add_codes((coarse, fine), None)
rng = np.random.RandomState(5)
Q = (db[rng.randint(0, n, size=300)].astype(np.float64) + 0.03 * rng.randn(300, D)).astype(np.float32)
a = s.search_batch(Q, quota=200, limit=10)
s._handle.set_scan_mode(1)
b = s.search_batch(Q, quota=200, limit=10)
s._handle.set_scan_mode(0)
for key in ("ids", "dist", "count", "visited"):
    assert np.array_equal(a[key], b[key]), key
index = orc.ArrayIndex(V, coarse, fine, np.arange(n, dtype=np.int64))
for i in range(3):
    r = orc.search_arrays(omodel, index, Q[i], 200, 10)
    assert np.array_equal(a["ids"][i][:len(r[0])], r[0])
print("new kernels ok: encode %d rows, %d queries twice" % (n, len(Q)))
