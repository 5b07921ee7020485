import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz")))


@pytest.fixture(scope="session")
def dims():
    from texocr_b200 import spec
    return spec.dims_from_config(spec.default_config(max_length=256, vocab_size=1000))


@pytest.fixture(scope="session")
def sd(dims):
    from texocr_b200 import synth
    return synth.seeded_state_dict(dims, seed=0)


def rel_max(a, b):
    """The parity metric of SURVEY.md section 8d: max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def tie_aware_rows(tokens, ref_tokens, ref_gaps, tau):
    """Rows must match exactly or first diverge where the reference's own top-2 gap < tau.
    Returns (n_exact_rows, list of (row, step, gap) divergences, ok)."""
    tokens, ref_tokens = np.asarray(tokens), np.asarray(ref_tokens)
    n = min(tokens.shape[1], ref_tokens.shape[1])
    exact, div, ok = 0, [], True
    for r in range(ref_tokens.shape[0]):
        neq = np.nonzero(tokens[r, :n] != ref_tokens[r, :n])[0]
        if len(neq) == 0:
            exact += 1
            continue
        s = int(neq[0])
        gap = float(ref_gaps[r, s])
        div.append((r, s, gap))
        if gap >= tau:
            ok = False
    return exact, div, ok
