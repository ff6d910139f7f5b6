"""Geometric-verification host code -- mirror of the reference's imcui/ui/utils.py:326-610
(_filter_matches_opencv, proc_ransac_matches, filter_matches, compute_geometry, set_null_pred, ransac_zoo).

Adds the method key "B200_MAGSAC" (default here): cv2.findHomography / cv2.findFundamentalMat(USAC_MAGSAC) replaced
by the batched GPU MAGSAC++ of libimw_b200.so (imw_magsac).  The "CV2_*" keys keep calling OpenCV exactly as the
reference does (they are the reference's own CPU path, kept for A/B comparison)."""
from typing import Any, Dict, Optional

import cv2
import numpy as np
import torch

from .. import ops
from ..hloc import extract_features, extractors, match_dense, match_features, matchers
from ..hloc.utils.base_model import dynamic_load

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"   # ui/utils.py:38


def parse_match_config(conf):
    """ui/utils.py:87-109: zoo entry {matcher, feature, dense/standalone} -> conf dicts of the registry."""
    if conf["standalone"]:
        return {"matcher": match_dense.confs.get(conf["matcher"]), "standalone": True, "info": conf.get("info", {})}
    return {"feature": extract_features.confs.get(conf["feature"]), "matcher": match_features.confs.get(conf["matcher"]),
            "standalone": False, "info": conf.get("info", {})}


def get_matcher_zoo(matcher_zoo):
    """ui/utils.py:66-84."""
    return {k: parse_match_config(v) for k, v in matcher_zoo.items()}


def get_model(match_conf: Dict[str, Any], device=None):
    """ui/utils.py:112-124 (the registry call site: dynamic_load over the plugin root)."""
    Model = dynamic_load(matchers, match_conf["model"]["name"])
    return Model(match_conf["model"]).eval().to(device or DEVICE)


def get_feature_model(conf: Dict[str, Dict[str, Any]], device=None):
    """ui/utils.py:127-139."""
    Model = dynamic_load(extractors, conf["model"]["name"])
    return Model(conf["model"]).eval().to(device or DEVICE)


DEFAULT_RANSAC_METHOD = "B200_MAGSAC"
DEFAULT_RANSAC_REPROJ_THRESHOLD = 8
DEFAULT_RANSAC_CONFIDENCE = 0.9999
DEFAULT_RANSAC_MAX_ITER = 10000
DEFAULT_MIN_NUM_MATCHES = 4

ransac_zoo = {  # ui/utils.py:1100-1110 plus the GPU entry
    "B200_MAGSAC": "b200",
    "POSELIB": "LO-RANSAC",
    "CV2_RANSAC": cv2.RANSAC,
    "CV2_USAC_MAGSAC": cv2.USAC_MAGSAC,
    "CV2_USAC_DEFAULT": cv2.USAC_DEFAULT,
    "CV2_USAC_FM_8PTS": cv2.USAC_FM_8PTS,
    "CV2_USAC_PROSAC": cv2.USAC_PROSAC,
    "CV2_USAC_FAST": cv2.USAC_FAST,
    "CV2_USAC_ACCURATE": cv2.USAC_ACCURATE,
    "CV2_USAC_PARALLEL": cv2.USAC_PARALLEL,
}


def _filter_matches_b200(kp0, kp1, method=None, reproj_threshold=3.0, confidence=0.99, max_iter=2000,
                         geometry_type="Homography", device="cuda"):
    """Same signature / return convention as _filter_matches_opencv (ui/utils.py:326-379):
    (M [3,3] float64 or None, mask bool [K] or None)."""
    if geometry_type not in ("Homography", "Fundamental"):
        raise NotImplementedError
    k = len(kp0)
    if k < (4 if geometry_type == "Homography" else 8):
        return None, None
    cap = max(128, (k + 127) // 128 * 128)
    p0 = torch.zeros(1, cap, 2, device=device)
    p1 = torch.zeros(1, cap, 2, device=device)
    p0[0, :k] = torch.as_tensor(np.asarray(kp0, dtype=np.float32).reshape(-1, 2), device=device)
    p1[0, :k] = torch.as_tensor(np.asarray(kp1, dtype=np.float32).reshape(-1, 2), device=device)
    counts = torch.tensor([k], dtype=torch.int32, device=device)
    M, mask, n_inl, _ = ops.magsac(p0, p1, counts, geometry_type, reproj_threshold, confidence, max_iter)
    if int(n_inl[0]) == 0:
        return None, None
    return M[0].cpu().numpy(), mask[0, :k].cpu().numpy().astype(bool)


def _filter_matches_opencv(kp0, kp1, method=cv2.RANSAC, reproj_threshold=3.0, confidence=0.99, max_iter=2000,
                           geometry_type="Homography"):
    """ui/utils.py:326-379 (unchanged behaviour)."""
    try:
        if geometry_type == "Homography":
            M, mask = cv2.findHomography(kp0, kp1, method=method, ransacReprojThreshold=reproj_threshold,
                                         confidence=confidence, maxIters=max_iter)
        elif geometry_type == "Fundamental":
            M, mask = cv2.findFundamentalMat(kp0, kp1, method=method, ransacReprojThreshold=reproj_threshold,
                                             confidence=confidence, maxIters=max_iter)
    except cv2.error:
        return None, None
    if mask is None:
        return None, None
    return M, np.array(mask.ravel().astype("bool"), dtype="bool")


def proc_ransac_matches(mkpts0, mkpts1, ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=3.0,
                        ransac_confidence=0.99, ransac_max_iter=2000, geometry_type="Homography"):
    """ui/utils.py:424-456."""
    if ransac_method.startswith("B200"):
        return _filter_matches_b200(mkpts0, mkpts1, None, ransac_reproj_threshold, ransac_confidence, ransac_max_iter, geometry_type)
    if ransac_method.startswith("CV2"):
        return _filter_matches_opencv(mkpts0, mkpts1, ransac_zoo[ransac_method], ransac_reproj_threshold, ransac_confidence,
                                      ransac_max_iter, geometry_type)
    raise NotImplementedError  # POSELIB: not installed in this image


def set_null_pred(feature_type: str, pred: dict):
    """ui/utils.py:382-398."""
    if feature_type == "KEYPOINT":
        pred["mmkeypoints0_orig"] = np.array([])
        pred["mmkeypoints1_orig"] = np.array([])
        pred["mmconf"] = np.array([])
    elif feature_type == "LINE":
        pred["mline_keypoints0_orig"] = np.array([])
        pred["mline_keypoints1_orig"] = np.array([])
    pred["H"] = np.eye(3)
    pred["geom_info"] = {}
    return pred


def compute_geometry(pred: Dict[str, Any], ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=DEFAULT_RANSAC_REPROJ_THRESHOLD,
                     ransac_confidence=DEFAULT_RANSAC_CONFIDENCE, ransac_max_iter=DEFAULT_RANSAC_MAX_ITER):
    """ui/utils.py:532-610: F first, then H, then the uncalibrated rectification from ALL matches."""
    mkpts0 = mkpts1 = None
    if "mkeypoints0_orig" in pred and "mkeypoints1_orig" in pred:
        mkpts0, mkpts1 = pred["mkeypoints0_orig"], pred["mkeypoints1_orig"]
    elif "line_keypoints0_orig" in pred and "line_keypoints1_orig" in pred:
        mkpts0, mkpts1 = pred["line_keypoints0_orig"], pred["line_keypoints1_orig"]
    if mkpts0 is None or mkpts1 is None:
        return {}
    if len(mkpts0) < 2 * DEFAULT_MIN_NUM_MATCHES:
        return {}
    geo_info = {}
    F, mask_f = proc_ransac_matches(mkpts0, mkpts1, ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter,
                                    geometry_type="Fundamental")
    if F is not None:
        geo_info["Fundamental"] = F.tolist()
        geo_info["mask_f"] = mask_f
    H, mask_h = proc_ransac_matches(mkpts0, mkpts1, ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter,
                                    geometry_type="Homography")
    h0, w0 = pred["image0_orig"].shape[:2]
    if H is not None:
        geo_info["Homography"] = H.tolist()
        geo_info["mask_h"] = mask_h
        try:
            _, H1, H2 = cv2.stereoRectifyUncalibrated(np.asarray(mkpts0, np.float64).reshape(-1, 2), np.asarray(mkpts1, np.float64).reshape(-1, 2),
                                                      F, imgSize=(w0, h0))
            geo_info["H1"] = H1.tolist()
            geo_info["H2"] = H2.tolist()
        except cv2.error:
            pass
    return geo_info


def filter_matches(pred: Dict[str, Any], ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=DEFAULT_RANSAC_REPROJ_THRESHOLD,
                   ransac_confidence=DEFAULT_RANSAC_CONFIDENCE, ransac_max_iter=DEFAULT_RANSAC_MAX_ITER, ransac_estimator=None):
    """ui/utils.py:459-529."""
    mkpts0: Optional[np.ndarray] = None
    feature_type: Optional[str] = None
    if "mkeypoints0_orig" in pred and "mkeypoints1_orig" in pred:
        mkpts0, mkpts1, feature_type = pred["mkeypoints0_orig"], pred["mkeypoints1_orig"], "KEYPOINT"
    elif "line_keypoints0_orig" in pred and "line_keypoints1_orig" in pred:
        mkpts0, mkpts1, feature_type = pred["line_keypoints0_orig"], pred["line_keypoints1_orig"], "LINE"
    else:
        return set_null_pred(feature_type, pred)
    if mkpts0 is None:
        return set_null_pred(feature_type, pred)
    if ransac_method not in ransac_zoo:
        ransac_method = DEFAULT_RANSAC_METHOD
    if len(mkpts0) < DEFAULT_MIN_NUM_MATCHES:
        return set_null_pred(feature_type, pred)
    geom_info = compute_geometry(pred, ransac_method=ransac_method, ransac_reproj_threshold=ransac_reproj_threshold,
                                 ransac_confidence=ransac_confidence, ransac_max_iter=ransac_max_iter)
    if "Homography" in geom_info:
        mask = geom_info["mask_h"]
        if feature_type == "KEYPOINT":
            pred["mmkeypoints0_orig"] = mkpts0[mask]
            pred["mmkeypoints1_orig"] = mkpts1[mask]
            pred["mmconf"] = pred["mconf"][mask]
        elif feature_type == "LINE":
            pred["mline_keypoints0_orig"] = mkpts0[mask]
            pred["mline_keypoints1_orig"] = mkpts1[mask]
        pred["H"] = np.array(geom_info["Homography"])
    else:
        set_null_pred(feature_type, pred)
    geom_info.pop("mask_h", None)
    geom_info.pop("mask_f", None)
    pred["geom_info"] = geom_info
    return pred
