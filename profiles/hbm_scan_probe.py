#!/usr/bin/env python
"""HBM-bound regime of the ADC scan: ONE query (and 2, 4, 8) ranking the WHOLE database (quota = everything, no
cross-query reuse of a code row), SURVEY 8d's "single-query-at-a-time full scan".  Reports, per code width, the scan
kernel's CUDA-event time, algorithmic bytes (M_padded x codes ranked: every stored code byte is read from HBM once) and
the fraction of the measured HBM copy peak (MEASURED_PEAKS.json), plus the synchronous call latency.

The scan's speed does not depend on what the code bytes are, so the index is filled with uniform random codes under a
random model (no 80 GB of 2048-d vectors needed to time the 32-byte-code scan of BASELINE config 3).

usage: python profiles/hbm_scan_probe.py [n_rows]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
import columbiaimagesearch_b200.lopq as lopq                      # noqa: E402
from tests.util import random_model_params                        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    peak_src = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    peak, peak_src = 6650.0, "fallback B200_PROFILING.md"
out = {"n": n, "peak_GBps": peak, "peak_source": peak_src, "runs": {}}
for (D, V, M) in [(128, 8, 16), (256, 8, 32)]:
    params = random_model_params(D, V, M, 256, seed=1)
    model = lopq.LOPQModel(parameters=params)
    s = lopq.LOPQSearcher(model, device=0)
    h = s._handle
    g = torch.Generator(device="cuda:0")
    g.manual_seed(7)
    co = torch.randint(0, V, (n, 2), generator=g, device="cuda:0", dtype=torch.int32)
    fi = torch.randint(0, 256, (n, M), generator=g, device="cuda:0", dtype=torch.uint8)
    torch.cuda.synchronize()
    h.index_add_device(co.data_ptr(), fi.data_ptr(), n)
    s.nb_indexed = n
    s._row_ids = [np.arange(n, dtype=np.int64)]
    del co, fi
    rng = np.random.RandomState(3)
    Q = (np.concatenate([params[0][0][rng.randint(0, V, 64)], params[0][1][rng.randint(0, V, 64)]], axis=1) + 0.3 * rng.randn(64, D)).astype(np.float32)
    for nq in (1, 2, 4, 8):
        for quota, tag in ((10 * n, "full"), (n // 32, "2cells")):
            for i in range(6):                                     # warm-up + segment-length feedback
                s.search_batch(Q[i * nq:(i + 1) * nq], quota=quota, limit=10)
            reps, scan, tot, nbytes, wall, exact = 20, 0.0, 0.0, 0, 0.0, 0
            for i in range(reps):
                q = Q[(i % 6) * nq:(i % 6 + 1) * nq]
                t0 = time.perf_counter()
                s.search_batch(q, quota=quota, limit=10)
                wall += time.perf_counter() - t0
                st = s.stats()
                scan += st["scan_ms"]; tot += st["total_ms"]; nbytes += st["codes_scanned"] * M; exact += st["exact_queries"]
            gbps = nbytes / (scan * 1e-3) / 1e9
            out["runs"]["M%d_nq%d_%s" % (M, nq, tag)] = {
                "scan_ms": round(scan / reps, 4), "device_ms": round(tot / reps, 4), "call_ms": round(wall / reps * 1e3, 4),
                "codes_ranked_per_call": nbytes // M // reps, "algorithmic_GBps": round(gbps, 1), "frac_of_hbm_peak": round(gbps / peak, 4),
                "work_items": st["work_items"], "packed": st["packed"], "exact_fallback_queries": exact}
    s._handle.close()
print(json.dumps(out))
