"""CPU oracle: a restatement of the reference's hot-path algorithms.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product path (image-matching-webui_b200/)
never does and fails loudly when its CUDA library is missing.

Every function cites the reference file:line it follows.  Arithmetic is fp32 on the CPU through
torch's functional ops (the reference itself is PyTorch eager; this keeps the restatement's
floating-point graph the same as the reference's: same op order, same dtypes).

Pinning: the reference's own tests hold no golden vectors, so the oracle is
pinned against outputs of the reference modules themselves, generated in the build container by
tools/make_golden.py (which imports /root/reference unmodified) and committed under
tests/golden/.  tests/test_oracle_golden.py checks every oracle function against those fixtures.
"""
from pathlib import Path

import torch

WEIGHTS_DIR = Path(__file__).resolve().parent.parent / "weights"


def load_weights(name: str) -> dict:
    """Flat {param name: fp32 tensor} dict written by tools/fetch_weights.py."""
    path = WEIGHTS_DIR / name
    if not path.exists():
        raise FileNotFoundError(
            f"{path} missing: run `python tools/fetch_weights.py` in the build container "
            "(needs /root/reference) before shipping the repo to the GPU box")
    return torch.load(str(path), map_location="cpu")
