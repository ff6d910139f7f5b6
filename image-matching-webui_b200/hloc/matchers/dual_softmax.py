"""Dual-softmax matcher -- drop-in for imcui/hloc/matchers/dual_softmax.py:39-71.  The N x M
probability matrix is never materialised (streamed in shared memory) and nothing but the [N]
outputs leaves the device (the reference copies all of P to the host, dual_softmax.py:28,32)."""
import torch

from ..utils.base_model import BaseModel
from ... import ops
from .nearest_neighbor import _pack_pair


class DualSoftMax(BaseModel):
    default_conf = {
        "match_threshold": 0.2,
        "inv_temperature": 20,
        "tensor_cores": True,  # B200 engine switch (see nearest_neighbor.py)
    }
    required_inputs = ["descriptors0", "descriptors1"]

    def _init(self, conf):
        pass

    def _forward(self, data):
        d0, d1 = data["descriptors0"], data["descriptors1"]
        if d0.size(-1) == 0 or d1.size(-1) == 0:  # dual_softmax.py:51-60
            matches0 = torch.full(d0.shape[:2], -1, device=d0.device)
            return {"matches0": matches0, "matching_scores0": torch.zeros_like(matches0)}
        ds, counts, n, m = _pack_pair(d0, d1)
        m0, s0 = ops.dual_softmax(ds, counts, self.conf["match_threshold"], self.conf["inv_temperature"], self.conf["tensor_cores"])
        # reference dtypes: int64 matches, float64 scores (NumPy round trip, dual_softmax.py:29-35)
        return {"matches0": m0[:, :n].long(), "matching_scores0": s0[:, :n].double()}

    def match_batch(self, batch):
        """see LightGlue.match_batch"""
        return ops.dual_softmax(batch["descriptors"], batch["counts"], self.conf["match_threshold"], self.conf["inv_temperature"],
                                self.conf["tensor_cores"])
