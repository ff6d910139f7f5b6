"""Programmatic pipeline -- mirror of the compute part of the reference's imcui/api/core.py:19-246 (ImageMatchingAPI:
constructor conf handling, _init_models, _forward, extract, forward, _geometry_check).  The plotting half
(visualize, :247-330, matplotlib) stays with the reference: it consumes the dict returned here unchanged.

The models come from the B200 plugin roots through the same registry calls (get_model / get_feature_model ->
dynamic_load), so a zoo entry such as {"matcher": "superpoint-lightglue", "feature": "superpoint_max"} resolves to the
CUDA plugins; geometric verification defaults to the reference's conf but accepts method "B200_MAGSAC"."""
from typing import Any, Dict

import numpy as np
import torch

from ..hloc import extract_features, logger, match_dense, match_features
from ..ui.utils import filter_matches, get_feature_model, get_model


class ImageMatchingAPI(torch.nn.Module):
    default_conf = {
        "ransac": {
            "enable": True,
            "estimator": "poselib",
            "geometry": "homography",
            "method": "RANSAC",
            "reproj_threshold": 3,
            "confidence": 0.9999,
            "max_iter": 10000,
        },
    }

    def __init__(self, conf: dict = {}, device: str = "cuda", detect_threshold: float = 0.015, max_keypoints: int = 1024,
                 match_threshold: float = 0.2) -> None:
        super().__init__()
        self.device = device
        self.conf = {**self.default_conf, **conf}
        self._updata_config(detect_threshold, max_keypoints, match_threshold)
        self._init_models()
        self.pred = None

    def parse_match_config(self, conf):  # core.py:62-76
        if conf["standalone"]:
            return {**conf, "matcher": match_dense.confs.get(conf["matcher"]["model"]["name"]), "standalone": True}
        return {**conf, "feature": extract_features.confs.get(conf["feature"]["model"]["name"]),
                "matcher": match_features.confs.get(conf["matcher"]["model"]["name"]), "standalone": False}

    def _updata_config(self, detect_threshold=0.015, max_keypoints=1024, match_threshold=0.2):  # core.py:78-96 (sic)
        self.standalone = self.conf["standalone"]
        if self.conf["standalone"]:
            try:
                self.conf["matcher"]["model"]["match_threshold"] = match_threshold
            except TypeError as e:
                logger.error(e)
        else:
            self.conf["feature"]["model"]["max_keypoints"] = max_keypoints
            self.conf["feature"]["model"]["keypoint_threshold"] = detect_threshold
            self.extract_conf = self.conf["feature"]
        self.match_conf = self.conf["matcher"]

    def _init_models(self):  # core.py:98-105
        self.matcher = get_model(self.match_conf, self.device)
        self.extractor = None if self.standalone else get_feature_model(self.conf["feature"], self.device)

    def _forward(self, img0, img1):  # core.py:107-127
        if self.standalone:
            return match_dense.match_images(self.matcher, img0, img1, self.match_conf["preprocessing"], device=self.device)
        pred0 = extract_features.extract(self.extractor, img0, self.extract_conf["preprocessing"])
        pred1 = extract_features.extract(self.extractor, img1, self.extract_conf["preprocessing"])
        return match_features.match_images(self.matcher, pred0, pred1)

    def _convert_pred(self, pred):  # core.py:129-138
        ret = {k: v.cpu().detach()[0].numpy() if isinstance(v, torch.Tensor) else v for k, v in pred.items()}
        return {k: v[0].cpu().detach().numpy() if isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], torch.Tensor) else v
                for k, v in ret.items()}

    @torch.inference_mode()
    def extract(self, img0: np.ndarray, **kwargs) -> Dict[str, np.ndarray]:  # core.py:140-172
        self.extractor.conf["max_keypoints"] = kwargs.get("max_keypoints", 512)
        self.extractor.conf["keypoint_threshold"] = kwargs.get("keypoint_threshold", 0.0)
        pred = self._convert_pred(extract_features.extract(self.extractor, img0, self.extract_conf["preprocessing"]))
        s0 = pred["original_size"] / pred["size"]
        pred["keypoints_orig"] = match_features.scale_keypoints(pred["keypoints"] + 0.5, s0) - 0.5
        if kwargs.get("binarize", False):
            assert "descriptors" in pred
            pred["descriptors"] = (pred["descriptors"] > 0).astype(np.uint8).T
        return pred

    @torch.inference_mode()
    def forward(self, img0: np.ndarray, img1: np.ndarray) -> Dict[str, Any]:  # core.py:174-209
        assert isinstance(img0, np.ndarray)
        assert isinstance(img1, np.ndarray)
        self.pred = self._forward(img0, img1)
        if self.conf["ransac"]["enable"]:
            self.pred = self._geometry_check(self.pred)
        return self.pred

    def _geometry_check(self, pred: Dict[str, Any]) -> Dict[str, Any]:  # core.py:211-235
        return filter_matches(pred, ransac_method=self.conf["ransac"]["method"],
                              ransac_reproj_threshold=self.conf["ransac"]["reproj_threshold"],
                              ransac_confidence=self.conf["ransac"]["confidence"], ransac_max_iter=self.conf["ransac"]["max_iter"])
