"""Developer probe: encode time of a product-shaped model (V = 2048, M = 8, 128-d) with the streamed float32 coarse
assignment (default) and with the exact kernels only (fine mode 1)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import columbiaimagesearch_b200.lopq as lopq
from columbiaimagesearch_b200 import synth
from tests.util import random_model_params
V = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
params = random_model_params(128, V, 8, 256, seed=5)
model = lopq.LOPQModel(parameters=params)
h = model._native()
n = 1 << 20
X = synth.dlib_style_torch(n, 128, seed=3, device="cuda:0") * 0.2
co = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
fi = torch.empty((n, 8), dtype=torch.uint8, device="cuda:0")
res = {}
for mode in (0, 1):
    h.set_fine_mode(mode)
    nn = n if mode == 0 else n // 8
    h.encode_guard_count(reset=True)
    h.encode_device(X.data_ptr(), nn, co.data_ptr(), fi.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3 if mode == 0 else 1
    for _ in range(reps):
        h.encode_device(X.data_ptr(), nn, co.data_ptr(), fi.data_ptr())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[mode] = (nn / dt / 1e6, int(co[:nn].sum().item()), int(fi[:nn // 8].sum().item()), h.encode_guard_count(reset=True))
    print("mode %d: %.2f M codes/s (rows %d), coarse checksum %d, fine checksum(first n/64) %d, guards %d" % ((mode, res[mode][0], nn) + res[mode][1:]), flush=True)
h.set_fine_mode(0)
h.encode_device(X.data_ptr(), n // 8, co.data_ptr(), fi.data_ptr())
torch.cuda.synchronize()
print("coarse checksum of the first n/8 rows in mode 0: %d (mode 1: %d)" % (int(co[:n // 8].sum().item()), res[1][1]))
