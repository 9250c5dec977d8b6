#!/usr/bin/env python
"""Synchronous search latency vs batch size on the 10 M index (wall clock around LOPQSearcher.search_batch, host buffers in
and out, the call the plugin makes).  usage: python profiles/latency_probe.py"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import columbiaimagesearch_b200.lopq as lopq
from columbiaimagesearch_b200 import synth

n, k = 10_000_000, 10
z = np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz"))
model = lopq.LOPQModel.from_npz(z)
X = synth.dlib_style_torch(n, 128, seed=1234, device="cuda:0")
s = lopq.LOPQSearcher(model, device=0)
h = s._handle
co = torch.empty((n, 2), dtype=torch.int32, device="cuda:0"); fi = torch.empty((n, 16), dtype=torch.uint8, device="cuda:0")
h.encode_device(X.data_ptr(), n, co.data_ptr(), fi.data_ptr())
h.index_add_device(co.data_ptr(), fi.data_ptr(), n)
s.nb_indexed = n; s._row_ids = [np.arange(n, dtype=np.int64)]
Q, _ = synth.near_duplicate_queries_torch(X, 4096, rho=0.1, seed=5)
Qn = Q.cpu().numpy()
out = {}
for nq in (1, 4, 16, 64, 256, 1024):
    for quota in (210000,):
        reps = 40
        for i in range(8):
            s.search_batch(Qn[(i * nq) % 3072:(i * nq) % 3072 + nq], quota=quota, limit=k)
        t0 = time.perf_counter()
        for i in range(reps):
            s.search_batch(Qn[(i * nq) % 3072:(i * nq) % 3072 + nq], quota=quota, limit=k)
        dt = (time.perf_counter() - t0) / reps
        st = s.stats()
        out["nq=%d" % nq] = {"ms_per_call": round(dt * 1e3, 4), "qps": round(nq / dt, 1), "device_ms": round(st["total_ms"], 4),
                             "scan_ms": round(st["scan_ms"], 4), "work_items": st["work_items"]}
print(json.dumps(out))
