"""ALIKED extractor plugin -- drop-in for imcui/hloc/extractors/aliked.py:12-32 (which wraps lightglue.ALIKED).
Same default_conf / required_inputs / output dict (lists of keypoints [N,2], scores [N], descriptors [128,N]); the
forward pass runs in libimw_b200.so (imw_aliked_forward).

Weights: the reference downloads aliked-n16.pth from GitHub (aliked.py:692-695); offline the plugin takes
`conf["state_dict"]` (reference key names) or `weights/<model_name>.pth`."""
import torch

from .. import WEIGHTS_DIR, logger
from ..utils.base_model import BaseModel
from ... import ops


class ALIKED(BaseModel):
    default_conf = {
        "model_name": "aliked-n16",
        "max_num_keypoints": -1,
        "detection_threshold": 0.2,
        "nms_radius": 2,
        "state_dict": None,
    }
    required_inputs = ["image"]
    n_limit_max = 20000

    def _init(self, conf):
        if conf["model_name"] not in ("aliked-n16", "aliked-n16rot"):  # same architecture (cfgs :629-634), different weights
            raise NotImplementedError(f"{conf['model_name']}: only the n16 architecture (c = 16/32/64/128, K = 3, M = 16) is built")
        sd = conf.get("state_dict")
        if sd is None:
            path = WEIGHTS_DIR / f"{conf['model_name']}.pth"
            if not path.exists():
                raise FileNotFoundError(f"{path} not found: no ALIKED checkpoint is available offline; pass conf['state_dict']")
            sd = torch.load(str(path), map_location="cpu")
        for k, v in ops.aliked_pack_weights(sd).items():
            self.register_buffer(k, v, persistent=False)
        self.conf["state_dict"] = None
        self._uncapped_cap = 4096
        logger.info("Load ALIKED model done.")

    def _forward(self, data):
        conf = self.conf
        image = data["image"].float()
        mk = int(conf["max_num_keypoints"])
        cap = mk if mk > 0 else self._uncapped_cap
        bufs = dict(self.named_buffers())
        out = ops.aliked_forward(bufs, image, conf, cap)
        counts = out["counts"].cpu()  # ragged outputs: one host sync
        if int((counts[1] > counts[0]).any()):
            self._uncapped_cap = cap = min(int(counts[1].max()), self.n_limit_max)
            out = ops.aliked_forward(bufs, image, conf, cap)
            counts = out["counts"].cpu()
        n = [int(c) for c in counts[0]]
        B = image.shape[0]
        return {
            "keypoints": [out["keypoints"][b, : n[b]] for b in range(B)],
            "scores": [out["scores"][b, : n[b]] for b in range(B)],
            "descriptors": [out["descriptors"][b, : n[b]].t() for b in range(B)],
        }
