"""Generate pickles of the REFERENCE's own model classes (build container only; needs /root/reference):

    python tests/golden/make_pickles.py

The product stores its models with pickle (cufacesearch storer/local.py:58,75; searcher_lopqhbase.py:113,142,192), so the
class path inside those files is ``lopq.model.LOPQModel`` / ``lopq.model.LOPQModelPCA``.  Here the reference package is
loaded under the name ``lopq`` (oracle/ref_loader.py, arithmetic untouched), instances are rebuilt from the parameters of
golden cases B and C and dumped with protocol 2 (what Python 2's cPickle.HIGHEST_PROTOCOL writes).  The tests load them
through ``install_as_lopq()``: the pickles must resolve to this repo's classes and encode / search as the reference does.
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = ref_loader.load(name="lopq")           # classes then pickle as lopq.model.*
    for case, fname in (("B", "ref_model_B.pkl"), ("C", "ref_model_C_pca.pkl")):
        z = np.load(os.path.join(HERE, "case_%s.npz" % case))
        m = int(z["M"])
        subs = z["subs"]
        params = ((z["C0"], z["C1"]), (z["Rs"][0], z["Rs"][1]), (z["mus"][0], z["mus"][1]),
                  ([subs[j] for j in range(m // 2)], [subs[j] for j in range(m // 2, m)]))
        if "pca_P" in z:
            model = ref.LOPQModelPCA(renorm=bool(z["renorm"]), parameters=params + (z["pca_P"], z["pca_mu"]))
        else:
            model = ref.LOPQModel(parameters=params)
        path = os.path.join(HERE, fname)
        with open(path, "wb") as f:
            pickle.dump(model, f, protocol=2)
        print(path, os.path.getsize(path), type(model).__module__, type(model).__name__)


if __name__ == "__main__":
    main()
