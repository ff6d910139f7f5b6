import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"

torch.set_grad_enabled(False)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(GOLDEN / f"{name}.npz"))
        return cache[name]

    return load


def lg_pair_from_source(golden, src):
    """'sp_real:api:0:1' -> (kpts0 [N,2] f32, desc0 [256,N], kpts1, desc1) numpy."""
    f, conf, i, j = str(src).split(":")
    b = golden(f)
    return (b[f"{conf}/{i}/keypoints"].astype(np.float32), b[f"{conf}/{i}/descriptors"],
            b[f"{conf}/{j}/keypoints"].astype(np.float32), b[f"{conf}/{j}/descriptors"])


def match_f1(m_test, m_ref):
    """F1 of the (i, matches0[i]) sets."""
    a = {(i, int(j)) for i, j in enumerate(np.asarray(m_test)) if j > -1}
    b = {(i, int(j)) for i, j in enumerate(np.asarray(m_ref)) if j > -1}
    if not a and not b:
        return 1.0
    tp = len(a & b)
    return 2 * tp / max(len(a) + len(b), 1)
