#!/usr/bin/env python
"""LOPQ hot-path benchmark (BASELINE.json).

    python bench.py --gpus 1 --steps 20 --warmup 3                   # headline: config 4 (10M x 128-d, V=8 M=16) on this repo
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's CPU algorithm on the host cores
    torchrun --nproc-per-node N bench.py --gpus N ...                # index sharded by coarse cell over N GPUs
    python bench.py --config c2|c3|c5                                # the other BASELINE configs (see CONFIGS)

One step = one batch of queries through the whole hot path (cell order, LUT build, ADC scan, top-k, float64 re-rank); for
--config c5 one step = one encode pass over this rank's rows.  Prints ONE JSON line (rank 0).

Multi-GPU (N > 1): the inverted lists are sharded by coarse cell; every rank brings a HOME slice of `--batch` queries per
step, so a step ranks N x batch queries and the per-GPU scan work stays what it is at N = 1 ("scaling": "weak"; the
database is fixed).  The exchange runs inside the library (peer-mapped windows, csrc/comm.cuh).  The fixed-batch figure
(the same `--batch` queries split over the ranks, strong scaling) is reported next to it as `strong`.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[3] (headline metric) and configs[1]
    "c4": dict(metric="LOPQ queries/sec @ recall@10, 10Mx128-d, V=8 M=16", n_db=10_000_000, D=128, V=8, M=16,
               model="dlib128_V8_M16.npz", quota=330_000, style="dlib"),
    # configs[0]: the reference's own CPU-runnable case (>= 100 CPU queries beside the GPU line)
    "c1": dict(metric="LOPQ queries/sec @ recall@10, 100kx128-d, V=4 M=8", n_db=100_000, D=128, V=4, M=8,
               model="dlib128_V4_M8.npz", quota=20_000, style="dlib", cpu_queries=100),
    "c2": dict(metric="LOPQ queries/sec @ recall@10, 1Mx128-d, V=8 M=16, batch=1024", n_db=1_000_000, D=128, V=8, M=16,
               model="dlib128_V8_M16.npz", quota=33_000, style="dlib"),
    # configs[2]: DeepSentibank-style 2048-d; the model is trained at start-up (134 MB of rotations: not a fixture)
    "c3": dict(metric="LOPQ queries/sec @ recall@10, 10Mx2048-d, V=8 M=32", n_db=10_000_000, D=2048, V=8, M=32,
               model=None, quota=450_000, style="sentibank"),
    # the product's shipped shape (conf/conf_search_dlibface_release.json:12-16): V=2048, M=8, 128-d, quota = min(1000 x 100, 10000)
    "pv": dict(metric="LOPQ queries/sec @ recall@10, 4Mx128-d, V=2048 M=8 (product-shaped), quota=10000 top-100", n_db=4_000_000,
               D=128, V=2048, M=8, model=None, quota=10_000, style="dlib"),
    # configs[4]: batch encode
    "c5": dict(metric="LOPQ compute_codes codes/sec, 50Mx128-d, V=8 M=16", n_db=50_000_000, D=128, V=8, M=16,
               model="dlib128_V8_M16.npz", style="dlib"),
}


def workload_string(a, cfg, name=None):
    """The workload, worded identically by both arms (`config.workload` of the b200 line and of the reference line)."""
    name = name or a.config
    if name == "c5":
        return "c5: compute_codes (LOPQModel.predict over rows), %dM x %d-d dlib-style synthetic rows, V=%d M=%d K=256" % (
            cfg["n_db"] // 1_000_000, cfg["D"], cfg["V"], cfg["M"])
    size = "%dM" % (cfg["n_db"] // 1_000_000) if cfg["n_db"] >= 1_000_000 else "%dk" % (cfg["n_db"] // 1000)
    return ("%s: %s x %d-d %s-style synthetic (4096-centre GMM, L2-normalised), V=%d M=%d K=256, near-duplicate queries (rho=%.2f), "
            "quota=%d, top-%d" % (name, size, cfg["D"], cfg["style"], cfg["V"], cfg["M"], a.rho, cfg.get("quota", 0), a.k))


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    p.add_argument("--n-db", type=int, default=0, help="override the database size of the config")
    p.add_argument("--batch", type=int, default=1024, help="queries per rank and step")
    p.add_argument("--quota", type=int, default=0)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--rho", type=float, default=0.1)
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--cpu-queries", type=int, default=6, help="queries of the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained run reported as value_sustained")
    p.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: in-library exchange or NCCL all-gather")
    p.add_argument("--emulate-shard", type=int, default=0, help="profiling aid (1 process): act as rank 0 of an N-way cell-sharded index")
    p.add_argument("--sweep", default="", help="comma-separated quotas: print recall/QPS per quota and exit")
    p.add_argument("--ntrain", type=int, default=20000, help="c3: training vectors of the 2048-d model")
    p.add_argument("--kp", type=int, default=0, help="b2l_set_preselect: minimum width of the float64 re-rank (0 = default)")
    p.add_argument("--lanes", type=int, default=2, help="handles (own stream + workspaces, shared index) the batches alternate over")
    a = p.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.n_db:
        cfg["quota"] = int(cfg.get("quota", 0) * a.n_db / cfg["n_db"]) or cfg.get("quota", 0)
        cfg["n_db"] = a.n_db
    if a.quota:
        cfg["quota"] = a.quota
    if "cpu_queries" in cfg and a.cpu_queries == 6:
        a.cpu_queries = cfg["cpu_queries"]
    a.cfg = cfg
    return a


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic(name):
    """DRAM bytes per launch of a kernel from the committed ncu --set full captures (profiles/scan_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            j = json.load(f)
        j = j.get(name, j) if isinstance(j.get(name), dict) else j
        return j["dram_bytes_per_launch"], j["source"]
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def oracle_index_for(omodel, orc, coarse, fine, sizes, starts_sorted, order, queries, quota):
    """Python dict index (the reference's LOPQSearcher layout: cell -> [(id, LOPQCode)]) restricted to the cells
    the sample queries visit -- search never touches any other cell before the quota is reached."""
    V = omodel.V
    need = set()
    for q in queries:
        got = 0
        for _, cell in orc.multisequence(q, omodel.Cs):
            need.add((int(cell[0]), int(cell[1])))
            got += int(sizes[int(cell[0]) * V + int(cell[1])])
            if got >= quota:
                break
    s = orc.OracleSearcher(omodel)
    for (c0, c1) in sorted(need):
        cid = c0 * V + c1
        rows = order[starts_sorted[cid]:starts_sorted[cid + 1]]
        co = (c0, c1)
        s.index[co] = [(int(r), orc.LOPQCode(co, tuple(f))) for r, f in zip(rows.tolist(), fine[rows].tolist())]
        s.nb_indexed += len(rows)
    return s


def time_oracle_queries(s, queries, quota, k):
    t0 = time.perf_counter()
    res = [s.search(q, quota=quota, limit=k, with_dists=True) for q in queries]
    return time.perf_counter() - t0, res


def prep_encode_fine(omodel, X, coarse):
    """DATA PREPARATION for the CPU arm (not timed, not the reference path): fine codes of rows with known
    coarse codes, matmul-form distances in float64.  Only has to produce a valid code database quickly."""
    n, D = X.shape
    h, m = D // 2, omodel.num_fine_splits
    ds = h // m
    fine = np.empty((n, omodel.M), np.uint8)
    for s in (0, 1):
        xs = X[:, s * h:(s + 1) * h].astype(np.float64)
        px = np.empty((n, h))
        for c in np.unique(coarse[:, s]):
            sel = coarse[:, s] == c
            px[sel] = (xs[sel] - omodel.Cs[s][c] - omodel.mus[s][c]) @ omodel.Rs[s][c].T
        for j in range(m):
            sub = omodel.subquantizers[s][j]
            d = (sub * sub).sum(1)[None, :] - 2.0 * px[:, j * ds:(j + 1) * ds] @ sub.T
            fine[:, s * m + j] = d.argmin(1)
    return fine


_G_SEARCHER = None          # inherited by the forked replica workers (never pickled)
_G_MODEL = None


def _worker(args):
    queries, quota, k = args
    return time_oracle_queries(_G_SEARCHER, queries, quota, k)[0]


def _enc_worker(X):
    from oracle import lopq_oracle as orc
    t0 = time.perf_counter()
    orc.compute_codes(_G_MODEL, X)
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------------------------
def run_reference(a):
    """Reference arm: the reference's algorithm (oracle port of LOPQSearcher.search: per-item Python ADC loop,
    stable sorted; c5: compute_codes_notparallel, one model.predict per row) on the host cores, pure CPU, same config.
    Replica processes (one stream per core) are the only parallelism the reference has (gunicorn workers / parmap).
    `kind` is "port": the reference's own files are Python 2 and only exist in the build container (oracle/ref_loader.py
    pins the port against them there); /root/reference is absent on the GPU box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config in ("c3", "pv"):
        print(json.dumps({"impl": "reference", "unavailable": "c3 reference arm not built: the 2048-d model is trained on the GPU at start-up; see cpu_baseline of the b200 line"}))
        return
    import multiprocessing as mp
    from oracle import lopq_oracle as orc
    cfg = a.cfg
    z = np.load(os.path.join(ROOT, "bench_models", cfg["model"]))
    omodel = orc.OracleModel.from_npz(z)
    V = omodel.V
    ncores = max(1, min(os.cpu_count() or 1, 16))
    ctx = mp.get_context("fork")
    if a.config == "c5":
        from columbiaimagesearch_b200 import synth
        global _G_MODEL
        _G_MODEL = omodel
        rows = 1500                                        # per core and step (about 0.3 s of model.predict calls)
        X = synth.dlib_style(rows * ncores * 2, 128, seed=a.seed)
        times = []
        with ctx.Pool(ncores) as pool:
            for it in range(a.warmup + a.steps):
                off = (it % 2) * rows * ncores
                jobs = [X[off + w * rows: off + (w + 1) * rows] for w in range(ncores)]
                t0 = time.perf_counter()
                pool.map(_enc_worker, jobs)
                dt = time.perf_counter() - t0
                if it >= a.warmup:
                    times.append(dt)
        total = sum(times)
        cps = a.steps * ncores * rows / total
        line = {"impl": "reference", "metric": cfg["metric"], "value": cps, "unit": "codes/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(a, cfg), "n_db": cfg["n_db"], "rows_per_step": ncores * rows},
                "cpu_baseline": {"value": cps, "unit": "codes/s", "cores": ncores, "kind": "port",
                                 "sample": "%d rows per step on %d processes (compute_codes_parallel's row-chunk fan-out, utils.py:178-200)" % (ncores * rows, ncores)},
                "e2e": {"value": cps, "unit": "codes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return
    t_prep = time.perf_counter()
    # database: same recipe (NumPy generator), coarse cells of all rows, fine codes only where needed
    rng = np.random.default_rng(a.seed + 1)
    C = np.random.RandomState(a.seed).randn(4096, 128).astype(np.float32)
    n = cfg["n_db"]
    quota = cfg["quota"]
    cell = np.empty(n, np.int32)
    chunk = 1 << 20
    h = 64
    cn = [(omodel.Cs[s].astype(np.float64) ** 2).sum(1) for s in (0, 1)]
    nq_total = max(2, a.cpu_queries) * ncores
    blocks = []
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        Xb = C[rng.integers(0, 4096, size=e0 - s0)] + 0.35 * rng.standard_normal((e0 - s0, 128), dtype=np.float32)
        Xb /= np.linalg.norm(Xb, axis=1, keepdims=True)
        cc = []
        for s in (0, 1):
            xs = Xb[:, s * h:(s + 1) * h].astype(np.float64)
            cc.append((cn[s][None, :] - 2.0 * xs @ omodel.Cs[s].astype(np.float64).T).argmin(1))
        cell[s0:e0] = cc[0] * V + cc[1]
        blocks.append(Xb)
    sizes = np.bincount(cell, minlength=V * V)
    # bounded sample: near-duplicate queries of rows of ONE cell, so only a handful of cells must be materialised
    # as Python objects (cost per query is the same as for any other query: it is linear in the codes ranked)
    home = int(np.argsort(sizes)[len(sizes) // 2])
    pool_rows = np.nonzero(cell == home)[0]
    qrows = np.sort(pool_rows[np.random.RandomState(a.seed + 7).randint(0, pool_rows.size, size=nq_total)])
    urng = np.random.RandomState(a.seed + 9)
    Q = []
    for r in qrows:
        u = urng.randn(128)
        q = blocks[r // chunk][r % chunk].astype(np.float64) + a.rho * u / np.linalg.norm(u)
        Q.append((q / np.linalg.norm(q)).astype(np.float32))
    need = set()
    for q in Q:
        got = 0
        for _, c in orc.multisequence(q, omodel.Cs):
            need.add(int(c[0]) * V + int(c[1]))
            got += int(sizes[int(c[0]) * V + int(c[1])])
            if got >= quota:
                break
    s = orc.OracleSearcher(omodel)
    for cid in sorted(need):
        rows = np.nonzero(cell == cid)[0]
        Xc = np.concatenate([blocks[b][rows[(rows >= b * chunk) & (rows < (b + 1) * chunk)] - b * chunk] for b in range(len(blocks))])
        co = (cid // V, cid % V)
        fine = prep_encode_fine(omodel, Xc, np.tile(np.array(co, np.int64), (Xc.shape[0], 1)))
        s.index[co] = [(int(r), orc.LOPQCode(co, tuple(f))) for r, f in zip(rows.tolist(), fine.tolist())]
        s.nb_indexed += len(rows)
    del blocks
    prep_s = time.perf_counter() - t_prep
    per = max(1, len(Q) // ncores)
    # a "step" = every core answers `qstep` queries
    qstep = max(1, per // max(1, a.steps + a.warmup))
    global _G_SEARCHER
    _G_SEARCHER = s
    times = []
    with ctx.Pool(ncores) as pool:
        pos = 0
        for it in range(a.warmup + a.steps):
            jobs = []
            for w in range(ncores):
                qs = [Q[(pos + w * qstep + j) % len(Q)] for j in range(qstep)]
                jobs.append((qs, quota, a.k))
            pos += ncores * qstep
            t0 = time.perf_counter()
            pool.map(_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= a.warmup:
                times.append(dt)
    total = sum(times)
    qps = a.steps * ncores * qstep / total
    line = {"impl": "reference", "metric": cfg["metric"], "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(a, cfg), "n_db": n, "quota": quota, "k": a.k,
                       "queries_per_step": ncores * qstep, "index": "python dict, visited cells only",
                       "query_sample": "near-duplicates of rows of one median-sized cell (only the visited cells are materialised as "
                                       "Python objects); cost per query is linear in the codes ranked, as for any query",
                       "prep_s": round(prep_s, 1)},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": ncores, "kind": "port",
                             "sample": "%d queries per step on %d replica processes (reference search is single-threaded); oracle port: the "
                                       "reference's Python-2 files exist only in the build container" % (ncores * qstep, ncores)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class Env(object):
    """torch / torch.distributed plumbing shared by the b200 configs."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (this implementation has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = "cuda:%d" % self.local
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device(self.dev))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def all_gather_rows(self, t, counts):
        """concatenate per-rank row blocks (rank r holds counts[r] rows) on every rank"""
        torch = self.torch
        if self.world == 1:
            return t
        mx = max(counts)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(self.world)])


# ------------------------------------------------------------------------------------------------
def train_model_c3(env, a, cfg, lopq, synth):
    """DATA PREPARATION (c3): a 2048-d V=8 M=32 model is 134 MB of float64 rotations, too large for a fixture, so it is
    trained here with the package's own trainer (lopq/train.py: the reference's algorithm, batched) on `ntrain` seeded
    vectors, identically on every rank.  Models are inputs of the hot path: not timed, not part of parity."""
    torch = env.torch
    m = lopq.LOPQModel(V=cfg["V"], M=cfg["M"], subquantizer_clusters=256)
    t0 = time.perf_counter()
    if env.rank == 0:                                          # N > 1: rank 0 trains, the parameters are broadcast
        X = synth.dlib_style_torch(a.ntrain, cfg["D"], seed=a.seed + 5, device=env.dev, relu=True).cpu().numpy().astype(np.float64)
        try:
            from threadpoolctl import threadpool_limits         # (torchrun pins OMP_NUM_THREADS=1: the eigh / GEMMs want the cores)
            with threadpool_limits(limits=max(1, (os.cpu_count() or 8) // 2)):
                m.fit(X, n_init=1, kmeans_coarse_iters=8, kmeans_local_iters=8, random_state=0)
        except ImportError:
            m.fit(X, n_init=1, kmeans_coarse_iters=8, kmeans_local_iters=8, random_state=0)
    if env.world > 1:
        V, M, h, K, ds = cfg["V"], cfg["M"], cfg["D"] // 2, 256, cfg["D"] // cfg["M"]
        shapes = [(2, V, h), (2, V, h, h), (2, V, h), (M, K, ds)]
        if env.rank == 0:
            arrs = [np.stack([np.asarray(x, np.float64) for x in m.Cs]), np.stack([np.asarray(x, np.float64) for x in m.Rs]),
                    np.stack([np.asarray(x, np.float64) for x in m.mus]),
                    np.stack([np.asarray(x, np.float64) for x in list(m.subquantizers[0]) + list(m.subquantizers[1])])]
        out = []
        for i, shp in enumerate(shapes):
            t = torch.from_numpy(arrs[i]).to(env.dev) if env.rank == 0 else torch.empty(shp, dtype=torch.float64, device=env.dev)
            env.dist.broadcast(t, 0)
            out.append(t.cpu().numpy())
        mm = M // 2
        m = lopq.LOPQModel(parameters=((out[0][0], out[0][1]), (out[1][0], out[1][1]), (out[2][0], out[2][1]),
                                       ([out[3][j] for j in range(mm)], [out[3][j] for j in range(mm, M)])))
    return m, time.perf_counter() - t0


def build_database(env, a, cfg, model, synth):
    """Synthesize the database on the device and encode it with the library, ROW-SHARDED over the ranks (each rank encodes
    its contiguous block of chunks, no communication: compute_codes_parallel's decomposition, utils.py:178-200), then
    all-gather the codes (18-40 bytes per row) so that every rank can pick the rows of its cells.
    Returns coarse_t [n,2], fine_t [n,M] (full, on every rank), the query batches, ground truth, encode statistics."""
    torch = env.torch
    n, D, M = cfg["n_db"], cfg["D"], cfg["M"]
    relu = cfg["style"] == "sentibank"
    nb = a.warmup + a.steps
    nq_tot = nb * a.batch * env.world
    nrec = min(nb, 4) * a.batch * env.world                      # recall is evaluated on the first batches
    enc = model._new_handle(env.local)
    chunk = 1 << 20 if D <= 256 else 1 << 17
    nchunks = (n + chunk - 1) // chunk
    C = torch.from_numpy(np.random.RandomState(a.seed).randn(4096, D)).to(device=env.dev, dtype=torch.float32)

    def gen_chunk(c):
        """chunk c of the database: same values on whichever rank generates it"""
        g = torch.Generator(device=env.dev)
        g.manual_seed(a.seed * 1000003 + c)
        rows = min(chunk, n - c * chunk)
        idx = torch.randint(0, 4096, (rows,), generator=g, device=env.dev)
        X = C[idx] + 0.35 * torch.randn((rows, D), generator=g, device=env.dev, dtype=torch.float32)
        if relu:
            X = torch.clamp_min(X, 0.0)
            X[:, 0] += 1e-3
        return X / X.norm(dim=1, keepdim=True)

    # queries: near-duplicates of rows of chunk 0 (so that they exist before the rest of the database is generated)
    X0 = gen_chunk(0)
    Qall, _ = synth.near_duplicate_queries_torch(X0, nq_tot, rho=a.rho, seed=a.seed + 77)
    Qrec = Qall[:nrec]
    best = torch.full((nrec,), float("inf"), device=env.dev)
    arg = torch.zeros((nrec,), dtype=torch.int64, device=env.dev)
    qn = (Qrec * Qrec).sum(1)
    # this rank's chunks: a contiguous range
    per = (nchunks + env.world - 1) // env.world
    c_lo, c_hi = min(nchunks, env.rank * per), min(nchunks, (env.rank + 1) * per)
    my_rows = max(0, min(n, c_hi * chunk) - c_lo * chunk)
    coarse_loc = torch.empty((max(my_rows, 1), 2), dtype=torch.int32, device=env.dev)
    fine_loc = torch.empty((max(my_rows, 1), M), dtype=torch.uint8, device=env.dev)
    stream = torch.cuda.ExternalStream(enc.stream(), device=env.dev)
    enc_ms = 0.0
    torch.backends.cuda.matmul.allow_tf32 = False
    for c in range(c_lo, c_hi):
        X = X0 if c == 0 else gen_chunk(c)
        rows = X.shape[0]
        o = (c - c_lo) * chunk
        if c == c_lo:                                             # warm-up call (allocations), then the timed passes
            enc.encode_device(X.data_ptr(), min(rows, 1 << 16), coarse_loc[o:].data_ptr(), fine_loc[o:].data_ptr())
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        enc.encode_device(X.data_ptr(), rows, coarse_loc[o:].data_ptr(), fine_loc[o:].data_ptr())
        ev1.record(stream)
        torch.cuda.synchronize()
        enc_ms += ev0.elapsed_time(ev1)
        # exact ground truth for the recall queries (eval.py:7-38 compute_all_neighbors), chunk by chunk
        for a0 in range(0, rows, 1 << 18):
            xb = X[a0:a0 + (1 << 18)]
            xn = (xb * xb).sum(1)
            for q0 in range(0, nrec, 4096):
                sl = slice(q0, min(nrec, q0 + 4096))
                d = qn[sl, None] - 2.0 * (Qrec[sl] @ xb.T) + xn[None, :]
                v, i = d.min(dim=1)
                upd = v < best[sl]
                best[sl] = torch.where(upd, v, best[sl])
                arg[sl] = torch.where(upd, i + (c * chunk + a0), arg[sl])
        del X
    if env.world > 1:                                             # combine the ground truth of the row blocks
        allb = [torch.empty_like(best) for _ in range(env.world)]
        alla = [torch.empty_like(arg) for _ in range(env.world)]
        env.dist.all_gather(allb, best)
        env.dist.all_gather(alla, arg)
        sb, sa = torch.stack(allb), torch.stack(alla)
        win = sb.argmin(0)
        arg = sa.gather(0, win[None, :])[0]
    counts = [max(0, min(n, min(nchunks, (r + 1) * per) * chunk) - min(nchunks, r * per) * chunk) for r in range(env.world)]
    coarse_t = env.all_gather_rows(coarse_loc[:my_rows], counts)
    fine_t = env.all_gather_rows(fine_loc[:my_rows], counts)
    assert coarse_t.shape[0] == n
    enc_s = env.max_over_ranks(enc_ms * 1e-3)[0]
    enc_stats = {"codes_per_s": n / max(enc_s, 1e-9), "n": n, "rows_per_rank": counts, "seconds_max_over_ranks": enc_s,
                 "guard_subvectors": enc.encode_guard_count(),
                 "note": "b2l_encode on device-resident float32 rows, row-sharded over the ranks (no communication), CUDA events per chunk, "
                         "first call warmed on a prefix; float64 arithmetic"}
    enc.close()
    return coarse_t, fine_t, Qall, arg.cpu().numpy(), enc_stats


def run_search(a):
    env = Env()
    torch, dist = env.torch, env.dist
    from columbiaimagesearch_b200 import synth
    import columbiaimagesearch_b200.lopq as lopq
    from columbiaimagesearch_b200.sharded import ShardedLOPQSearcher
    cfg = a.cfg
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    train_s = None
    if cfg["model"]:
        z = np.load(os.path.join(ROOT, "bench_models", cfg["model"]))
        model = lopq.LOPQModel.from_npz(z)
    else:
        model, train_s = train_model_c3(env, a, cfg, lopq, synth)
    M, D, n, quota, k, nq = model.M, cfg["D"], cfg["n_db"], cfg["quota"], a.k, a.batch
    coarse_t, fine_t, Qall, gt, enc_stats = build_database(env, a, cfg, model, synth)

    # one searcher class for every N: the inverted lists are sharded by coarse cell over the ranks (N = 1: one shard)
    if a.emulate_shard > 1 and world == 1:
        searcher = ShardedLOPQSearcher(model, device=local, emulate=(a.emulate_shard, 0))
        searcher.add_codes_device(coarse_t, fine_t)
        cell_all = coarse_t[:, 0].long() * model.V + coarse_t[:, 1].long()
        searcher.finalize(global_sizes=torch.bincount(cell_all, minlength=model.V ** 2).cpu().numpy())
    else:
        searcher = ShardedLOPQSearcher(model, device=local)
        searcher.add_codes_device(coarse_t, fine_t)
        searcher.finalize()
    handle = searcher._handle
    peer = world > 1 and a.exchange == "peer"
    kp = a.kp if a.kp else cfg.get("kp", 0)
    if kp:
        handle.set_preselect(kp)                            # (before the siblings are created: they inherit it)
    if a.lanes > 1:
        searcher.enable_pipelining(a.lanes)
    if peer:
        searcher.enable_peer_exchange(nq, max(16, k))
    nb = a.warmup + a.steps
    G = nq * world                                             # queries per step, all ranks

    def global_batch(b):
        return Qall[b * G:(b + 1) * G]

    def home_batch(b):
        return Qall[b * G + rank * nq: b * G + (rank + 1) * nq]

    def recall_of(q_quota, nbatches):
        """eval.get_recall's definition (eval.py:92-142) on whole batches, through the synchronous host-driven API"""
        rec = np.zeros(2)
        vis = cand = 0
        for b in range(nbatches):
            q = global_batch(b).cpu().numpy()
            r = lopq.eval.get_recall_batch(_Fixed(searcher, handle), q, gt[b * G:(b + 1) * G], quota=q_quota, thresholds=(1, k), batch=G)
            rec += r * G
            vis += _Fixed.last_visited
            cand += handle.stats()["codes_scanned"]
        tot = nbatches * G
        return rec[1] / tot, rec[0] / tot, vis / tot, cand / tot

    if a.sweep:
        for qv in [int(v) for v in a.sweep.split(",")]:
            r10, r1, vis, cand = recall_of(qv, min(nb, 2))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 5
            for b in range(reps):
                searcher.search_batch(global_batch(b % nb), quota=qv, limit=k)
            dt = (time.perf_counter() - t0) / reps
            handle.set_async(False)
            import torch as _t
            rec = _t.empty(handle.records_bytes(G, k), dtype=_t.uint8, device=dev)
            handle.search_local(global_batch(0).cpu().numpy(), qv, k, rec.data_ptr(), exact=0)      # one plain fast-path pass: its statistics
            st = handle.stats()
            app, bnd = handle.debug_candidates(G)
            cert = handle.search_merge(rec.data_ptr(), 1, G, k)["certified"] if world == 1 else None
            st["cand_appended"] = {"mean": float(app.mean()), "p50": float(np.median(app)), "p99": float(np.percentile(app, 99)),
                                   "max": int(app.max()), "over_cap": int((app > 8192).sum())}
            st["uncertified_first_pass"] = None if cert is None else int((cert == 0).sum())
            if rank == 0:
                print(json.dumps({"quota": qv, "recall@10": r10, "recall@1": r1, "cells_visited": vis, "codes_per_query_local": cand,
                                  "qps": G / dt, "ms_per_batch": dt * 1e3, "stats": st}))
        return

    # recall first: it runs the host-driven protocol, which also sizes every workspace before the exchange is used
    r10, r1, vis, cand = recall_of(quota, min(nb, 4))

    def enqueue(batch_of, b):
        if peer:
            return searcher.search_home_async(batch_of(b), quota=quota, limit=k)
        return searcher.search_batch_async(batch_of(b), quota=quota, limit=k)

    def run_pipelined(batch_of, first, count):
        """`count` batches through the public asynchronous API, two in flight: the host-side launch work of batch i+1
        overlaps the device work of batch i.  Every batch's results are read (and certified) on the host."""
        pend, redo, last = None, np.zeros(2, np.int64), None
        for b in range(first, first + count):
            p = enqueue(batch_of, b)
            if pend is not None:
                last = pend.result(copy=False)
                redo += (last["exact_queries"], last["rescan_queries"])
            pend = p
        last = pend.result(copy=False)
        redo += (last["exact_queries"], last["rescan_queries"])
        return redo, last

    dev_batch = home_batch if peer else global_batch
    # ---- warm-up -----------------------------------------------------------------------------------
    run_pipelined(dev_batch, 0, a.warmup)
    env.barrier()

    # ---- timed region 1: device-resident inputs ("value") ------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    stream = torch.cuda.ExternalStream(handle.stream(), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    searcher.reset_stats()
    env.barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    redo_q, last = run_pipelined(dev_batch, a.warmup, a.steps)
    exact_q, rescan_q = int(redo_q[0]), int(redo_q[1])
    searcher._sync_lanes()                                    # (the lanes have their own streams: all of them done ...
    ev1.record(stream)                                        #  ... before the closing event on lane 0's stream)
    env.barrier()
    last = {kk: (v.copy() if isinstance(v, np.ndarray) else v) for kk, v in last.items()}   # (views of a rotating pinned block)
    wall = time.perf_counter() - t0
    dev_s, wall_s = env.max_over_ranks(max(ev0.elapsed_time(ev1) * 1e-3, 0.0), wall)
    lane_st = searcher.lane_stats()
    launches_lanes = sum(s_["acc_kernel_launches"] + s_["acc_calls"] * (8 if peer else 1) for s_ in lane_st)
    # per-kernel times for the time split and the roofline: the same steps again on ONE lane, so that the CUDA events around
    # a kernel are not stretched by another batch's kernels sharing the SMs
    searcher.active_lanes = 1
    searcher.reset_stats()
    env.barrier()
    run_pipelined(dev_batch, a.warmup, a.steps)
    env.barrier()
    searcher.active_lanes = 0
    st = handle.stats()                                       # per-call CUDA-event times of those steps, summed
    scan_ms, plan_ms, sel_ms = st["acc_scan_ms"], st["acc_plan_ms"], st["acc_select_ms"]
    scan_bytes, items = st["acc_scan_bytes"], st["acc_work_items"]
    timed_calls = max(1, st["acc_calls"])
    # kernels of the library per step: the search kernels + merge (+ put / 3 signals / 3 waits of the exchange)
    launches = launches_lanes
    step_s = max(dev_s, 1e-9)
    value = a.steps * G / step_s
    # parity spot-check at N > 1 (rank 0, first queries of the last timed batch against the oracle)
    parity = None
    if rank == 0 and world > 1 and not a.no_cpu_baseline:
        parity = oracle_spot_check(a, cfg, model, coarse_t, fine_t, dev_batch(a.warmup + a.steps - 1), last, quota, k)

    # ---- sustained run (>= sustain_s seconds of back-to-back batches, device-resident inputs) --------------------
    sustained = None
    if a.sustain_s > 0:
        nsteps = max(a.steps, int(a.sustain_s / (step_s / a.steps)) + 1)
        env.barrier()
        ev0.record(stream)
        run_pipelined(lambda b: dev_batch(b % nb), 0, nsteps)
        searcher._sync_lanes()
        ev1.record(stream)
        env.barrier()
        sus_s = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)[0]
        sustained = {"value": nsteps * G / sus_s, "steps": nsteps, "seconds": sus_s}

    # ---- timed region 2: end to end through the public API with host buffers ("e2e") -----------------
    Qhost = torch.empty((a.steps * (nq if peer else G), D), dtype=torch.float32).pin_memory()
    for b in range(a.steps):
        w = nq if peer else G
        Qhost[b * w:(b + 1) * w].copy_(dev_batch(a.warmup + b))
    Qh = Qhost.numpy()
    w = nq if peer else G
    host_batch = lambda b: Qh[(b % a.steps) * w:((b % a.steps) + 1) * w]
    run_pipelined(host_batch, 0, a.warmup)                   # warm-up of the host-buffer path
    env.barrier()
    t0 = time.perf_counter()
    redo_e2e, _ = run_pipelined(host_batch, 0, a.steps)
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)[0]
    # the same, one synchronous call at a time (what the unmodified plugin loop does per request batch)
    env.barrier()
    t0 = time.perf_counter()
    for b in range(a.steps):
        if peer:
            searcher.search_home_async(host_batch(b), quota=quota, limit=k).result(copy=False)
        else:
            searcher.search_batch(host_batch(b), quota=quota, limit=k)
    env.barrier()
    e2e_sync_s = env.max_over_ranks(time.perf_counter() - t0)[0]
    clocks = sampler.stop()                                  # sampled through the timed regions
    h2d = w * D * 4                                           # per rank
    d2h = w * k * (8 + 8 + 8 + M) + w * 9 + (256 if peer else 0)

    # ---- strong scaling figure at N > 1: the same `batch` queries split over the ranks ------------------------------
    strong = None
    if peer and nq % world == 0 and (nq // world) * D * 4 % 16 == 0:
        nh = nq // world
        sb = lambda b: Qall[b * nq + rank * nh: b * nq + (rank + 1) * nh]
        run_pipelined_n = lambda first, count: _run_home(searcher, sb, first, count, quota, k)
        run_pipelined_n(0, a.warmup)
        env.barrier()
        ev0.record(stream)
        run_pipelined_n(a.warmup, a.steps)
        searcher._sync_lanes()
        ev1.record(stream)
        env.barrier()
        s_s = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)[0]
        strong = {"value": a.steps * nq / s_s, "ms_per_step": 1e3 * s_s / a.steps, "queries_per_step": nq,
                  "note": "fixed global batch of %d queries (%d per rank), device-resident inputs" % (nq, nh)}

    # ---- roofline of the dominant kernel (ADC scan) ----------------------------------------------------
    peak, peak_src = measured_peak()
    alg = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    kname = "k_scan_pk<%d>" % max(4, 1 << (M - 1).bit_length())
    traffic, traffic_src = ncu_traffic(a.config)
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    sm_hz = 1e6 * float(clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965)
    wf_min = scan_bytes / 64.0                                  # one conflict-free LDS wavefront serves 32 lanes x 2 packed queries
    g_ach = wf_min / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    g_peak = nsm * sm_hz / 1e9
    roofline = {"bound": "smem-gather", "kernel": kname, "achieved": g_ach, "peak": g_peak, "unit": "Gwavefront/s", "frac": g_ach / g_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "lookups_per_launch": scan_bytes / timed_calls, "min_wavefronts_per_launch": wf_min / timed_calls,
                "scan_ms_per_launch": scan_ms / timed_calls, "launches_timed": timed_calls,
                "hbm": {"algorithmic_GBps": alg, "algorithmic_frac_of_peak": alg / peak, "peak_GBps": peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": scan_bytes / timed_calls,
                        "dram_GBps": (traffic / (scan_ms / timed_calls * 1e-3) / 1e9) if traffic and scan_ms > 0 else None,
                        "dram_frac_of_peak": (traffic / (scan_ms / timed_calls * 1e-3) / 1e9 / peak) if traffic and scan_ms > 0 else None},
                "note": "The batched scan is bound by shared-memory LUT gathers, not HBM: one code row read from HBM/L2 serves every query "
                        "of the batch that visits the cell.  `frac` = minimum conflict-free LDS wavefronts (one per 32 lanes x 2 packed "
                        "queries = 64 look-ups) per second over the pipe's 1 wavefront/clock/SM at the sampled SM clock.  `hbm` keeps the "
                        "SURVEY 8d accounting (M x codes ranked, no credit for reuse: may exceed the HBM peak) and the measured DRAM "
                        "traffic of the committed ncu capture.  The HBM-bound regime (single query, no reuse) is measured by "
                        "--config c3 / profiles/*latency*."}

    # ---- CPU baseline (rank 0, N = 1): the reference's per-item Python loop on a bounded sample ---------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline_search(a, cfg, model, coarse_t, fine_t, Qall, searcher, quota, k)

    if rank == 0:
        line = {"metric": cfg["metric"], "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * step_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u16", "data": "synthetic",
                "config": {"workload": workload_string(a, cfg),
                           "n_db": n, "batch_per_rank": nq, "queries_per_step": G, "quota": quota, "k": k,
                           "recall@10": r10, "recall@1": r1,
                           "arithmetic": "16-bit packed table sums in the scan (float32-table and float64 fallbacks), float64 tables and re-rank",
                           "l2": "distinct query batch per step; per-step working set (%d MB codes + per-batch LUTs) exceeds the 126 MB L2"
                                 % (n * max(4, 1 << (M - 1).bit_length()) // 1_000_000),
                           "sharding": ("cells by (c0+c1) mod N; every rank brings %d home queries per step (weak scaling over a fixed "
                                        "database); exchange = %s" % (nq, "peer-mapped windows inside the library" if peer else "NCCL all-gather"))
                           if world > 1 else "single GPU",
                           "lanes": a.lanes, "preselect_kp_min": kp, "model_train_s": train_s},
                "recall@10": r10, "recall@1": r1, "cells_visited_per_query": vis,
                "codes_ranked_per_query": scan_bytes / timed_calls / M / G * world,
                "wall_s_timed_region": wall_s, "value_sustained": sustained, "strong": strong,
                "e2e": {"value": a.steps * G / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "api": "%s, 2 batches in flight, pinned host queries in / host results out every step"
                               % ("search_home_async" if peer else "search_batch_async"),
                        "sync_api_value": a.steps * G / e2e_sync_s,
                        "exact_fallback_queries": int(redo_e2e[0]), "float32_rescan_queries": int(redo_e2e[1])},
                "gpu_launches": int(launches), "exact_fallback_queries": int(exact_q), "float32_rescan_queries": int(rescan_q),
                "time_split_ms_per_step": {"plan+lut": plan_ms / timed_calls, "scan": scan_ms / timed_calls, "select": sel_ms / timed_calls,
                                           "total_local": st["acc_total_ms"] / timed_calls},
                "time_split_note": "per-kernel CUDA-event times of the same steps repeated on one lane (un-overlapped); `value` runs them "
                                   "over %d lanes, so ms_per_step is below their sum" % a.lanes,
                "work_items_per_step": items / timed_calls, "parity_spot_check": parity,
                "encode": enc_stats, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(line))
    searcher.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run_home(searcher, batch_of, first, count, quota, k):
    pend = None
    for b in range(first, first + count):
        p = searcher.search_home_async(batch_of(b), quota=quota, limit=k)
        if pend is not None:
            pend.result(copy=False)
        pend = p
    pend.result(copy=False)


class _Fixed(object):
    """adapter: get_recall_batch calls searcher.search_batch(X, quota=, limit=); route it to the synchronous entry point"""
    last_visited = 0

    def __init__(self, searcher, handle):
        self.s = searcher

    def search_batch(self, X, quota, limit):
        o = self.s._search_batch_sync(X, quota, limit)
        _Fixed.last_visited = int(o["visited"].sum())
        return o


def _host_index(model, coarse_t, fine_t):
    co = coarse_t.cpu().numpy()
    fi = fine_t.cpu().numpy()
    cellid = co[:, 0].astype(np.int64) * model.V + co[:, 1]
    order = np.argsort(cellid, kind="stable")
    sizes = np.bincount(cellid, minlength=model.V ** 2)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return co, fi, order, sizes, starts


def _oracle_model(model):
    from oracle import lopq_oracle as orc
    return orc, orc.OracleModel(model.Cs, model.Rs, model.mus, model.subquantizers)


def oracle_spot_check(a, cfg, model, coarse_t, fine_t, Qdev, out, quota, k, nqs=3):
    """N > 1: the first queries of a timed batch against the oracle (ids, visited, distances)"""
    orc, omodel = _oracle_model(model)
    co, fi, order, sizes, starts = _host_index(model, coarse_t, fine_t)
    qs = Qdev[:nqs].cpu().numpy()
    s = oracle_index_for(omodel, orc, co, fi, sizes, starts, order, qs, quota)
    _, res = time_oracle_queries(s, qs, quota, k)
    same = all([r.id for r in res[i][0]] == out["ids"][i][:len(res[i][0])].tolist() and res[i][1] == int(out["visited"][i]) for i in range(nqs))
    maxd = max(float(np.max(np.abs(np.array([r.dist for r in res[i][0]]) - out["dist"][i][:len(res[i][0])]))) for i in range(nqs))
    return {"queries": nqs, "ids_match_oracle": bool(same), "max_abs_dist_diff": maxd}


def cpu_baseline_search(a, cfg, model, coarse_t, fine_t, Qall, searcher, quota, k):
    orc, omodel = _oracle_model(model)
    co, fi, order, sizes, starts = _host_index(model, coarse_t, fine_t)
    qs = Qall[:a.cpu_queries].cpu().numpy()
    s = oracle_index_for(omodel, orc, co, fi, sizes, starts, order, qs, quota)
    dt, res = time_oracle_queries(s, qs, quota, k)
    g = searcher.search_batch(qs, quota=quota, limit=k)
    same = all([r.id for r in res[i][0]] == g["ids"][i][:len(res[i][0])].tolist() and res[i][1] == int(g["visited"][i])
               for i in range(len(qs)))
    maxd = max(float(np.max(np.abs(np.array([r.dist for r in res[i][0]]) - g["dist"][i][:len(res[i][0])]))) for i in range(len(qs)))
    return {"value": len(qs) / dt, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": "%d queries of the same workload, reference algorithm (per-item Python ADC loop, oracle port), %0.1f s" % (len(qs), dt),
            "ids_match_gpu": bool(same), "max_abs_dist_diff": maxd, "host_cores": os.cpu_count()}


# ------------------------------------------------------------------------------------------------
def run_product(a):
    """--config pv: the product-shaped configuration (V = 2048, M = 8, 128-d, quota 10000, top-100) on one GPU through the
    large-V path (csrc/largev.cuh).  One step = one batch of `--batch` queries through b2l_search."""
    env = Env()
    torch = env.torch
    from columbiaimagesearch_b200 import synth
    import columbiaimagesearch_b200.lopq as lopq
    cfg = a.cfg
    assert env.world == 1, "the large-V path is single-GPU"
    D, V, M, n, quota, k, nq = cfg["D"], cfg["V"], cfg["M"], cfg["n_db"], cfg["quota"], max(a.k, 100), a.batch
    # model: trained here with the package's trainer on seeded vectors (data preparation; models are inputs)
    Xt = synth.dlib_style_torch(a.ntrain * 5, D, seed=a.seed + 5, device=env.dev).cpu().numpy().astype(np.float64)
    model = lopq.LOPQModel(V=V, M=M, subquantizer_clusters=256)
    t0 = time.perf_counter()
    model.fit(Xt, n_init=1, kmeans_coarse_iters=8, kmeans_local_iters=8, random_state=0)
    train_s = time.perf_counter() - t0
    coarse_t, fine_t, Qall, gt, enc_stats = build_database(env, a, cfg, model, synth)
    s = lopq.LOPQSearcher(model, device=env.local, keep_host_copy=False)
    h = s._handle
    h.index_add_device(coarse_t.data_ptr(), fine_t.data_ptr(), n)
    s.nb_indexed = n
    s._row_ids = [np.arange(n, dtype=np.int64)]
    nb = a.warmup + a.steps
    outs = dict(rowid=torch.empty((nq, k), dtype=torch.int64, device=env.dev), dist=torch.empty((nq, k), dtype=torch.float64, device=env.dev),
                coarse=torch.empty((nq, k, 2), dtype=torch.int32, device=env.dev), fine=torch.empty((nq, k, M), dtype=torch.uint8, device=env.dev),
                count=torch.empty(nq, dtype=torch.int32, device=env.dev), visited=torch.empty(nq, dtype=torch.int32, device=env.dev))

    def dev_step(b):
        q = Qall[b * nq:(b + 1) * nq]
        h.search_device(q.data_ptr(), nq, quota, k, outs["rowid"].data_ptr(), outs["dist"].data_ptr(), outs["coarse"].data_ptr(),
                        outs["fine"].data_ptr(), outs["count"].data_ptr(), outs["visited"].data_ptr())

    # recall (eval.get_recall definition) on the first batches, through the host API
    nrec = min(nb, 4)
    rec = np.zeros(2)
    vis = 0
    for b in range(nrec):
        q = Qall[b * nq:(b + 1) * nq].cpu().numpy()
        o = s.search_batch(q, quota=quota, limit=k)
        hit = (o["ids"] == gt[b * nq:(b + 1) * nq, None]) & (np.arange(k)[None, :] < o["count"][:, None])
        rank = np.where(hit.any(1), hit.argmax(1), k)
        rec += [np.count_nonzero(rank < 1), np.count_nonzero(rank < 10)]
        vis += int(o["visited"].sum())
    st0 = h.stats()
    for b in range(a.warmup):
        dev_step(b)
    sampler = ClockSampler(env.local)
    sampler.start()
    stream = torch.cuda.ExternalStream(h.stream(), device=env.dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.reset_stats()
    env.barrier()
    ev0.record(stream)
    for b in range(a.warmup, nb):
        dev_step(b)
    ev1.record(stream)
    env.barrier()
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    st = h.stats()
    Qh = torch.empty((a.steps * nq, D), dtype=torch.float32).pin_memory()
    Qh.copy_(Qall[a.warmup * nq:nb * nq])
    qh = Qh.numpy()
    s.search_batch(qh[:nq], quota=quota, limit=k)
    t0 = time.perf_counter()
    for b in range(a.steps):
        s.search_batch(qh[b * nq:(b + 1) * nq], quota=quota, limit=k)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    cpu = None
    if not a.no_cpu_baseline:
        orc, omodel = _oracle_model(model)
        co, fi = coarse_t.cpu().numpy(), fine_t.cpu().numpy()
        index = orc.ArrayIndex(V, co, fi, np.arange(n, dtype=np.int64))
        qs = Qall[:a.cpu_queries].cpu().numpy()
        t0 = time.perf_counter()
        res = [orc.search_arrays(omodel, index, q, quota, k) for q in qs]
        dt = time.perf_counter() - t0
        g = s.search_batch(qs, quota=quota, limit=k)
        same = all(np.array_equal(g["ids"][i][:len(r[0])], r[0]) and int(g["visited"][i]) == r[4] for i, r in enumerate(res))
        maxd = max(float(np.max(np.abs(r[1] - g["dist"][i][:len(r[1])]))) for i, r in enumerate(res))
        cpu = {"value": len(qs) / dt, "unit": "queries/s", "cores": 1, "kind": "port",
               "sample": "%d queries, reference algorithm (Python heap multi-sequence walk, memoised LUTs; ADC sums vectorised over a cell), %.1f s" % (len(qs), dt),
               "ids_match_gpu": bool(same), "max_abs_dist_diff": maxd, "host_cores": os.cpu_count()}
    tot = nrec * nq
    line = {"metric": cfg["metric"], "value": a.steps * nq / dev_s, "unit": "queries/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dev_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "pv: %dM x %d-d dlib-style synthetic, V=%d M=%d K=256 (model trained at start-up, %d vectors), batch=%d near-duplicate "
                                   "queries (rho=%.2f), quota=%d, top-%d" % (n // 1_000_000, D, V, M, Xt.shape[0], nq, a.rho, quota, k),
                       "recall@10": rec[1] / tot, "recall@1": rec[0] / tot, "cells_visited_per_query": vis / tot, "model_train_s": train_s,
                       "arithmetic": "float64 throughout: device multi-sequence traversal, grouped-GEMM projections, exact ADC of every retrieved code"},
            "recall@10": rec[1] / tot, "codes_ranked_per_query": st["acc_codes_scanned"] / max(1, st["acc_calls"]) / nq,
            "projection_slots_per_query": st0["lut_slots"] / nq,
            "e2e": {"value": a.steps * nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * D * 4, "d2h_bytes_per_step": nq * k * (24 + M) + nq * 9,
                    "api": "LOPQSearcher.search_batch (b2l_search, host queries in / host results out)"},
            "gpu_launches": int(st["acc_kernel_launches"]), "encode": enc_stats, "cpu_baseline": cpu, "clocks": clocks,
            "roofline": None}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_encode(a):
    """--config c5: compute_codes over 50M x 128-d rows, row-sharded over the ranks (utils.py:178-200's decomposition, no
    communication).  One step = one encode pass over this rank's resident block of rows; the 50M rows are visited block
    by block (device-resident figure = `value`); `e2e` times the same call with PINNED HOST rows in and host codes out."""
    env = Env()
    torch = env.torch
    from columbiaimagesearch_b200 import synth
    import columbiaimagesearch_b200.lopq as lopq
    cfg = a.cfg
    z = np.load(os.path.join(ROOT, "bench_models", cfg["model"]))
    model = lopq.LOPQModel.from_npz(z)
    M, D, n = model.M, cfg["D"], cfg["n_db"]
    per_rank = n // env.world
    block = min(per_rank, 1 << 22)                              # rows resident per step (2 GB of float32 at 128-d)
    nblocks = (per_rank + block - 1) // block
    h = model._new_handle(env.local)
    stream = torch.cuda.ExternalStream(h.stream(), device=env.dev)
    X = synth.dlib_style_torch(block, D, seed=a.seed + 31 * env.rank, device=env.dev)
    coarse_t = torch.empty((block, 2), dtype=torch.int32, device=env.dev)
    fine_t = torch.empty((block, M), dtype=torch.uint8, device=env.dev)
    for _ in range(max(1, a.warmup)):
        h.encode_device(X.data_ptr(), block, coarse_t.data_ptr(), fine_t.data_ptr())
    sampler = ClockSampler(env.local)
    sampler.start()
    env.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        h.encode_device(X.data_ptr(), block, coarse_t.data_ptr(), fine_t.data_ptr())
    ev1.record(stream)
    env.barrier()
    dev_s = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)[0]
    value = a.steps * block * env.world / dev_s
    # end to end: pinned host rows in, host codes out (H2D-bound: 512 B in, 24 B out per row)
    eb = min(block, 1 << 20)
    Xh = torch.empty((eb, D), dtype=torch.float32).pin_memory()
    Xh.copy_(X[:eb])
    xh = Xh.numpy()
    h.encode(xh)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        co, fi = h.encode(xh)
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)[0]
    clocks = sampler.stop()
    guards = h.encode_guard_count()
    # parity + CPU baseline on rank 0: the reference's compute_codes_notparallel (one model.predict per row) on a bounded sample
    cpu = None
    if env.rank == 0 and not a.no_cpu_baseline:
        orc, omodel = _oracle_model(model)
        ns = 3000
        t0 = time.perf_counter()
        oc = orc.compute_codes(omodel, xh[:ns])
        dt = time.perf_counter() - t0
        same = all(tuple(int(v) for v in c.coarse) == tuple(co[i]) and tuple(int(v) for v in c.fine) == tuple(fi[i]) for i, c in enumerate(oc))
        cpu = {"value": ns / dt, "unit": "codes/s", "cores": 1, "kind": "port",
               "sample": "%d rows, compute_codes_notparallel (model.predict per row, utils.py:203-218), %.1f s" % (ns, dt),
               "codes_match_gpu": bool(same), "host_cores": os.cpu_count()}
    flops_row = 3 * model.V * D + D * D + 3 * 256 * D           # coarse + rotation + fine
    if env.rank == 0:
        line = {"metric": cfg["metric"], "value": value, "unit": "codes/s", "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * dev_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(a, cfg), "n_db": n,
                           "decomposition": "rows row-sharded over %d rank(s): %d rows per rank in %d resident block(s) of %d rows; one step = one "
                                            "pass over a resident block" % (env.world, per_rank, nblocks, block),
                           "rows_per_step": block * env.world, "seconds_for_all_rows_at_this_rate": n / value,
                           "l2": "a resident block (%d MB) exceeds the 126 MB L2" % (block * D * 4 // 1_000_000)},
                "e2e": {"value": a.steps * eb * env.world / e2e_s, "unit": "codes/s", "h2d_bytes_per_step": eb * D * 4 * env.world,
                        "d2h_bytes_per_step": eb * (8 + M) * env.world, "api": "Handle.encode (b2l_encode, pinned host rows in, host codes out)"},
                "gpu_launches": int(h.stats()["kernel_launches"]) * a.steps,
                "roofline": {"bound": "tensor", "kernel": "b2l_encode pipeline (k_coarse_assign, k_rotate_dmma, k_fine_tc + k_fine_redo)",
                             "achieved": value / env.world * flops_row / 1e12, "peak": 75.0, "unit": "TFLOP/s",
                             "frac": value / env.world * flops_row / 1e12 / 75.0, "traffic": None,
                             "note": "SURVEY 8d flops per row (3VD + D^2 + 3KD = %d) over the fp32 FFMA peak of the part (~75 TFLOP/s), kept as "
                                     "the yardstick of earlier rounds: the fine argmin, 83%% of these flops, now runs on the tcgen05 tensor cores "
                                     "(kind::tf32, three TF32 pieces per float32 product, accumulators in tensor memory) with a float64 list pass "
                                     "for what its guard cannot decide; the rotation on the float64 tensor cores (DMMA)" % flops_row},
                "guard_subvectors_total": guards, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(line))
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c5":
        run_encode(args)
    elif args.config == "pv":
        run_product(args)
    else:
        run_search(args)
