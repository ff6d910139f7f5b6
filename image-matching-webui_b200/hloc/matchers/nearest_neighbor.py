"""Mutual nearest-neighbour matcher -- drop-in for imcui/hloc/matchers/nearest_neighbor.py:27-66."""
import torch

from ..utils.base_model import BaseModel
from ... import ops


def _pack_pair(d0, d1):
    """[1,D,N],[1,D,M] channel-first (reference layout) -> token-major slots [2,cap,D] + counts
    (cap a multiple of 128: the row tile of the tcgen05 similarity kernel)."""
    n, m, dim = d0.shape[-1], d1.shape[-1], d0.shape[1]
    cap = (max(n, m) + 127) // 128 * 128
    ds = torch.zeros(2, cap, dim, device=d0.device)
    ds[0, :n], ds[1, :m] = d0[0].t().float(), d1[0].t().float()
    return ds, torch.tensor([n, m], dtype=torch.int32, device=d0.device), n, m


class NearestNeighbor(BaseModel):
    default_conf = {
        "ratio_threshold": None,
        "distance_threshold": None,
        "do_mutual_check": True,
        # B200 engine switch (not in the reference): similarity tiles on tcgen05 with 3xTF32 split operands
        # (fp32-equivalent; needs dim % 32 == 0, else the fp32 CUDA-core kernel runs)
        "tensor_cores": True,
    }
    required_inputs = ["descriptors0", "descriptors1"]

    def _init(self, conf):
        pass

    def _forward(self, data):
        d0, d1 = data["descriptors0"], data["descriptors1"]
        if d0.size(-1) == 0 or d1.size(-1) == 0:  # nearest_neighbor.py:39-48
            matches0 = torch.full(d0.shape[:2], -1, device=d0.device)
            return {"matches0": matches0, "matching_scores0": torch.zeros_like(matches0)}
        ds, counts, n, m = _pack_pair(d0, d1)
        m0, s0 = ops.nearest_neighbor(ds, counts, self.conf["ratio_threshold"], self.conf["distance_threshold"],
                                      self.conf["do_mutual_check"], self.conf["tensor_cores"])
        return {"matches0": m0[:, :n].long(), "matching_scores0": s0[:, :n]}

    def match_batch(self, batch):
        """see LightGlue.match_batch: descriptors [2P,cap,D] token-major, counts [2P] -> (matches0 [P,cap], scores0 [P,cap])"""
        return ops.nearest_neighbor(batch["descriptors"], batch["counts"], self.conf["ratio_threshold"], self.conf["distance_threshold"],
                                    self.conf["do_mutual_check"], self.conf["tensor_cores"])
