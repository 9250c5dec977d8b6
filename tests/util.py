"""Shared helpers for the parity tests (golden fixtures + oracle models)."""
import os

import numpy as np

from oracle.lopq_oracle import OracleModel
from tests.golden.make_golden import CASES, case_inputs  # noqa: F401  (re-exported)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    return z, OracleModel.from_npz(z)


def searches(z):
    for si in range(int(z["n_searches"])):
        limit = int(z["s%d_limit" % si])
        yield si, int(z["s%d_quota" % si]), (None if limit < 0 else limit)
