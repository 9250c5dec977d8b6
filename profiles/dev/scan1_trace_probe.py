#!/usr/bin/env python
"""Developer probe: per-phase timestamps (%globaltimer) of k_scan1 blocks, from a library built with -DSCAN1_TRACE
(profiles/dev/libtrace.so; see profiles/dev/README).  Not part of the product or of the bench."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200._native as nat                    # noqa: E402
nat.LIB_PATH = os.path.join(ROOT, "profiles", "dev", "libtrace.so")
import torch                                                     # noqa: E402
import columbiaimagesearch_b200.lopq as lopq                      # noqa: E402
from tests.util import random_model_params                        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
res = {}
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
    lib = nat.load_library()
    lib.b2l_debug_scan1_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for quota, tag in ((10 * n, "full"),):
        for i in range(8):
            s.search_batch(Q[i:i + 1], quota=quota, limit=10)
        st = s.stats()
        buf = np.zeros(1024 * 8, np.uint64)
        lib.b2l_debug_scan1_trace(buf.ctypes.data, buf.size)
        nb = min(1024, 148 * 3)
        t = buf.reshape(1024, 8)[:nb].astype(np.int64)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        names = ["entry", "decoded+lut issued", "lut landed", "first rows landed", "first bound", "scan done", "final bound", "exit"]
        out = {"scan_ms": st["scan_ms"], "blocks": int(t.shape[0]), "work_items": st["work_items"]}
        for i, nm in enumerate(names):
            col = rel[:, i]
            out[nm] = {"min": round(float(col.min()), 2), "median": round(float(np.median(col)), 2), "max": round(float(col.max()), 2)}
        out["scan_phase_us"] = {"min": round(float((rel[:, 5] - rel[:, 4]).min()), 2), "median": round(float(np.median(rel[:, 5] - rel[:, 4])), 2),
                                "max": round(float((rel[:, 5] - rel[:, 4]).max()), 2)}
        res["M%d_%s" % (M, tag)] = out
    s._handle.close()
print(json.dumps(res, indent=1))
