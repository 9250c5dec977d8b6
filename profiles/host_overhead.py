#!/usr/bin/env python
"""Where the end-to-end time of one search_batch call goes (wall clock, GPU box): C-ABI call with device buffers,
C-ABI call with host buffers, Python search_batch.  usage: python profiles/host_overhead.py [n_db]"""
import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import columbiaimagesearch_b200.lopq as lopq
from columbiaimagesearch_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq, k, quota = 1024, 10, 210000 * n // 10_000_000
z = np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz"))
model = lopq.LOPQModel.from_npz(z)
dev = "cuda:0"
X = synth.dlib_style_torch(n, 128, seed=1234, device=dev)
s = lopq.LOPQSearcher(model, device=0)
h = s._handle
co = torch.empty((n, 2), dtype=torch.int32, device=dev); fi = torch.empty((n, 16), dtype=torch.uint8, device=dev)
h.encode_device(X.data_ptr(), n, co.data_ptr(), fi.data_ptr())
h.index_add_device(co.data_ptr(), fi.data_ptr(), n)
s.nb_indexed = n; s._row_ids = [np.arange(n, dtype=np.int64)]
Q, _ = synth.near_duplicate_queries_torch(X, 8 * nq, rho=0.1, seed=5)
Qh = torch.empty((8 * nq, 128), dtype=torch.float32).pin_memory(); Qh.copy_(Q); Qn = Qh.numpy()
Qp = Q.cpu().numpy()          # pageable copy
outs = dict(rowid=torch.empty((nq, k), dtype=torch.int64, device=dev), dist=torch.empty((nq, k), dtype=torch.float64, device=dev),
            coarse=torch.empty((nq, k, 2), dtype=torch.int32, device=dev), fine=torch.empty((nq, k, 16), dtype=torch.uint8, device=dev),
            count=torch.empty((nq,), dtype=torch.int32, device=dev), visited=torch.empty((nq,), dtype=torch.int32, device=dev))

def t(fn, reps=24):
    for i in range(4): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(reps): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

def f_dev(i):
    q = Q[(i % 8) * nq:(i % 8 + 1) * nq]
    h.search_device(q.data_ptr(), nq, quota, k, outs["rowid"].data_ptr(), outs["dist"].data_ptr(), outs["coarse"].data_ptr(),
                    outs["fine"].data_ptr(), outs["count"].data_ptr(), outs["visited"].data_ptr())
print("C-ABI, device in/out      : %.3f ms/batch" % t(f_dev), h.stats())
print("C-ABI, pinned host in     : %.3f ms/batch" % t(lambda i: h.search(Qn[(i % 8) * nq:(i % 8 + 1) * nq], quota, k)))
print("C-ABI, pageable host in   : %.3f ms/batch" % t(lambda i: h.search(Qp[(i % 8) * nq:(i % 8 + 1) * nq], quota, k)))
print("search_batch, pinned in   : %.3f ms/batch" % t(lambda i: s.search_batch(Qn[(i % 8) * nq:(i % 8 + 1) * nq], quota=quota, limit=k)))
pr = cProfile.Profile(); pr.enable()
for i in range(16): s.search_batch(Qn[(i % 8) * nq:(i % 8 + 1) * nq], quota=quota, limit=k)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
