"""Drop-in mirror of the reference ``lopq`` package (lopq/lopq/__init__.py:3-9) on top of
libb200lopq: same module and class names, same method signatures and return shapes; every
arithmetic step runs in the CUDA library (no CPU fallback).

``install_as_lopq()`` registers this package under the name ``lopq`` so that existing callers
(``from lopq.search import LOPQSearcher``) and pickles of ``lopq.model.LOPQModel[PCA]`` resolve
to it unchanged.
"""
import sys

from . import model, search, utils, eval  # noqa: F401
from .model import LOPQModel, LOPQModelPCA, LOPQCode
from .search import LOPQSearcher, LOPQSearcherGPU, LOPQSearcherLMDB, multisequence

__all__ = ["LOPQModel", "LOPQModelPCA", "LOPQCode", "LOPQSearcher", "LOPQSearcherGPU", "LOPQSearcherLMDB", "multisequence",
           "model", "search", "utils", "eval", "install_as_lopq"]


def install_as_lopq(force=False):
    """Alias this package as top-level ``lopq`` (and its sub-modules) in ``sys.modules``."""
    me = sys.modules[__name__]
    if "lopq" in sys.modules and sys.modules["lopq"] is not me and not force:
        raise RuntimeError("another `lopq` package is already imported; pass force=True to shadow it")
    sys.modules["lopq"] = me
    for sub in ("model", "search", "utils", "eval"):
        sys.modules["lopq." + sub] = sys.modules[__name__ + "." + sub]
    return me
