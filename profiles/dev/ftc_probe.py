"""Developer probe of the tensor-core fine argmin (fine_tc.cuh): score error against float64, code parity against the
float64-only mode, and the time of a device-resident encode in the three modes."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200.lopq as lopq  # noqa: E402
from columbiaimagesearch_b200 import synth  # noqa: E402
from tests.util import random_model_params, random_data  # noqa: E402

out = {}


def score_error(model, X, j):
    h = model._native()
    sc, px = h.debug_fine_scores(X, j)
    M = model.M
    m = M // 2
    sub = np.asarray(model.subquantizers[j // m][j % m], np.float64)
    ds = sub.shape[1]
    p = px[:, j * ds:(j + 1) * ds]
    exact = 0.5 * (sub ** 2).sum(1)[None, :] - p @ sub.T
    K = sub.shape[0]
    err = np.abs(sc[:, :K].astype(np.float64) - exact)
    scale = (np.sqrt((p ** 2).sum(1)) + np.sqrt((sub ** 2).sum(1).max())) ** 2 * 2.0 ** -24
    ratio = err / scale[:, None]
    am = (sc[:, :K].argmin(1) == exact.argmin(1)).mean()
    return float(ratio.max()), float(np.median(ratio)), float(am), (float(sc[:, K:].min()) if K < 256 else None)


def parity(model, X, name):
    h = model._native()
    res = {}
    h.encode_guard_count(reset=True)
    h.set_fine_mode(0)
    c0, f0 = h.encode(X)
    res["guards_tc"] = h.encode_guard_count(reset=True)
    h.set_fine_mode(2)
    c2, f2 = h.encode(X)
    res["guards_f32"] = h.encode_guard_count(reset=True)
    h.set_fine_mode(1)
    c1, f1 = h.encode(X)
    h.set_fine_mode(0)
    res["rows"] = int(X.shape[0])
    res["tc_eq_f64"] = bool(np.array_equal(f0, f1) and np.array_equal(c0, c1))
    res["f32_eq_f64"] = bool(np.array_equal(f2, f1))
    if not res["tc_eq_f64"]:
        bad = np.argwhere(f0 != f1)
        res["n_bad"] = int(len(bad))
        res["bad_head"] = bad[:10].tolist()
        res["bad_vals"] = [(int(f0[r, c]), int(f1[r, c])) for r, c in bad[:10]]
    out[name] = res
    print(name, res, flush=True)


def timing(model, n, name):
    h = model._native()
    X = synth.dlib_style_torch(n, 128, seed=77, device="cuda:0")
    co = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
    fi = torch.empty((n, model.M), dtype=torch.uint8, device="cuda:0")
    res = {}
    for mode in (0, 2):
        h.set_fine_mode(mode)
        for _ in range(3):
            h.encode_device(X.data_ptr(), n, co.data_ptr(), fi.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            h.encode_device(X.data_ptr(), n, co.data_ptr(), fi.data_ptr())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        res["mode%d_ms" % mode] = dt * 1e3
        res["mode%d_Mcodes_s" % mode] = n / dt / 1e6
        res["mode%d_fine_sum" % mode] = int(fi.sum().item())
    h.set_fine_mode(0)
    out[name] = res
    print(name, res, flush=True)


z = np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz"))
model = lopq.LOPQModel.from_npz(z)
if "--time-only" in sys.argv:
    timing(model, 1_000_000, "time_dlib128_1M")
    sys.exit(0)
Xh = synth.dlib_style_torch(300_000, 128, seed=4321, device="cuda:0").cpu().numpy()
for j in (0, 7, 15):
    r = score_error(model, Xh[:4096], j)
    out["err_dlib_j%d" % j] = r
    print("score error / (2^-24 (|p|+cmax)^2): max %.3f median %.3f, argmin agreement %.4f" % r[:3], flush=True)
parity(model, Xh, "parity_dlib128_M16")

for (D, V, M, K, n) in [(128, 4, 8, 256, 40000), (128, 4, 16, 100, 30000), (64, 3, 8, 64, 30000), (256, 2, 16, 256, 20000)]:
    params = random_model_params(D, V, M, K, seed=D + M + K)
    mdl = lopq.LOPQModel(parameters=params)
    db = random_data(params, n, seed=5)
    db[::50] = db[1::50][: db[::50].shape[0]]
    r = score_error(mdl, db[:4096], M - 1)
    out["err_D%d_M%d_K%d" % (D, M, K)] = r
    print("D%d M%d K%d: err max %.3f median %.3f agreement %.4f pad-min %s" % ((D, M, K) + r), flush=True)
    parity(mdl, db, "parity_D%d_M%d_K%d" % (D, M, K))
    parity(mdl, db.astype(np.float64) * 1e-3, "parity64_D%d_M%d_K%d" % (D, M, K))

timing(model, 1_000_000, "time_dlib128_1M")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ftc_probe.json"), "w"), indent=1)
