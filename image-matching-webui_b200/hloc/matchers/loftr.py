"""LoFTR dense matcher plugin -- drop-in for imcui/hloc/matchers/loftr.py:12-71 (which wraps kornia.feature.LoFTR).
Same default_conf / required_inputs / output dict (keypoints0, keypoints1, scores); the forward pass runs in
libimw_b200.so (imw_loftr_forward).

Weights: the reference takes kornia's `outdoor` checkpoint (download).  No LoFTR checkpoint exists offline, so the
plugin accepts `conf["state_dict"]` (reference key names: backbone.*, loftr_coarse.*, fine_preprocess.*, loftr_fine.*)
or `weights/loftr_<weights>.pt`."""
import torch

from .. import WEIGHTS_DIR, logger
from ..utils.base_model import BaseModel
from ... import ops


class LoFTR(BaseModel):
    default_conf = {
        "weights": "outdoor",
        "match_threshold": 0.2,
        "sinkhorn_iterations": 20,  # unused by the dual-softmax default (kept for conf compatibility)
        "max_keypoints": -1,
        "state_dict": None,
        "tensor_cores": True,
        "max_matches": None,        # capacity of the match buffers (default: one per coarse cell)
    }
    required_inputs = ["image0", "image1"]

    def _init(self, conf):
        sd = conf.get("state_dict")
        if sd is None:
            path = WEIGHTS_DIR / f"loftr_{conf['weights']}.pt"
            if not path.exists():
                raise FileNotFoundError(f"{path} not found: no LoFTR checkpoint is available offline; pass conf['state_dict']")
            sd = torch.load(str(path), map_location="cpu")
        self.temp_bug_fix = "minima" in str(conf.get("model_name", ""))  # hloc/matchers/loftr.py:27-28
        self._packed = ops.loftr_pack_weights(sd)
        self._dev = None
        self.conf["state_dict"] = None  # do not keep a second copy alive
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)
        logger.info(f"Loaded LoFTR with weights {conf['weights']}")

    def _weights(self, device):
        if self._dev is None or self._dev[0] != device:
            self._dev = (device, ops.loftr_to_device(self._packed, device))
        return self._dev[1]

    def _forward(self, data):
        # hloc refines the keypoints of ITS image0: the LoFTR module sees (image1, image0) (hloc/matchers/loftr.py:43-51)
        im0, im1 = data["image1"], data["image0"]
        assert im0.shape[0] == 1 and im0.shape[1] == 1 and im1.shape[:2] == (1, 1), "grayscale pair expected"
        tc = {False: 0, True: 1, "3xtf32": 1, "tf32": 2}[self.conf["tensor_cores"]]
        kconf = {"match_threshold": self.conf["match_threshold"], "use_tensor_cores": tc}
        if im0.shape == im1.shape:
            out = ops.loftr_forward(self._weights(im0.device), torch.stack([im0[0, 0], im1[0, 0]]).float(), kconf,
                                    self.conf["max_matches"], temp_bug_fix=self.temp_bug_fix)
        else:   # the two images keep their own sizes (kornia LoFTR == SE2LoFTR loftr.py:48-56 runs the backbone per image then)
            out = ops.loftr_forward(self._weights(im0.device), im0[0].float(), kconf, self.conf["max_matches"],
                                    temp_bug_fix=self.temp_bug_fix, images1=im1[0].float())
        n = int(out["counts"][0])
        k0, k1, scores = out["keypoints0"][0, :n], out["keypoints1"][0, :n], out["confidence"][0, :n]
        top_k = self.conf["max_keypoints"]
        if top_k is not None and len(scores) > top_k:  # hloc/matchers/loftr.py:58-65 (sic: with -1 the slice [:-1] drops the lowest-confidence match, as in the reference)
            keep = torch.argsort(scores, descending=True)[:top_k]
            k0, k1, scores = k0[keep], k1[keep], scores[keep]
        # switch back: module keypoints0 belong to hloc's image1
        return {"keypoints1": k0, "keypoints0": k1, "scores": scores,
                "batch_indexes": torch.zeros(len(scores), dtype=torch.long, device=scores.device)}
