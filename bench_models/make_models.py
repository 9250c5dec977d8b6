"""Train the benchmark models ONCE in the build container with the REFERENCE's own LOPQModel.fit
(oracle/ref_loader.py) on the seeded synthetic distribution of columbiaimagesearch_b200/synth.py, and
save the parameters as .npz fixtures (training is not part of the parity contract: models are inputs).

    python bench_models/make_models.py            # -> bench_models/dlib128_V8_M16.npz, dlib128_V4_M8.npz
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                        # noqa: E402
from oracle.lopq_oracle import model_to_npz_dict     # noqa: E402
from columbiaimagesearch_b200 import synth           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = {"dlib128_V8_M16": dict(D=128, V=8, M=16, ntrain=50000), "dlib128_V4_M8": dict(D=128, V=4, M=8, ntrain=50000)}

if __name__ == "__main__":
    ref = ref_loader.load()
    for name, c in MODELS.items():
        train = synth.dlib_style(c["ntrain"], c["D"], seed=1234).astype(np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref.LOPQModel(V=c["V"], M=c["M"], subquantizer_clusters=256)
            m.fit(train, n_init=1, random_state=0)
        d = model_to_npz_dict(m)
        d["Rs"] = d["Rs"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "->", path, "%.0f KB" % (os.path.getsize(path) / 1024.0))
