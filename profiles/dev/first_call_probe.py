"""Developer probe: time of the first large encode call after a small warm-up call (workspace growth), then steady state."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200.lopq as lopq
from columbiaimagesearch_b200 import synth
z = np.load(os.path.join(ROOT, "bench_models", "dlib128_V8_M16.npz"))
model = lopq.LOPQModel.from_npz(z)
n = 1_000_000
X = synth.dlib_style_torch(n, 128, seed=4321, device="cuda:0")
co = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
fi = torch.empty((n, 16), dtype=torch.uint8, device="cuda:0")
for mode in (0, 2):
    h = model._new_handle(0)
    h.set_fine_mode(mode)
    t = []
    for rows in (1 << 16, n, n, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.encode_device(X.data_ptr(), rows, co.data_ptr(), fi.data_ptr())
        torch.cuda.synchronize()
        t.append((rows, round((time.perf_counter() - t0) * 1e3, 2)))
    print("mode", mode, t, flush=True)
    h.close()
