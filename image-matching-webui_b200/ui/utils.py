"""Registry call sites and geometric verification -- drop-in for the hot-path part of imcui/ui/utils.py:
`get_matcher_zoo` / `parse_match_config` / `get_model` / `get_feature_model` (:66-139) and `filter_matches` /
`compute_geometry` / `proc_ransac_matches` / `set_null_pred` / `ransac_zoo` (:326-610, :1100-1110).

Verification is organised around BATCHES: `verify_pairs` takes the match dicts of many pairs, runs ONE MAGSAC++ launch for
all fundamental matrices and ONE for all homographies (imw_magsac: one CTA per pair) and fills every dict with the
reference's keys.  `filter_matches(pred, ...)` -- the reference's per-pair signature -- is the batch of one.  The method key
"B200_MAGSAC" (default here) is the GPU estimator; the reference's "CV2_*" keys still call OpenCV exactly as the reference
does (its own CPU path, kept for A/B comparison)."""
from typing import Any, Dict, List, Optional, Sequence

import cv2
import numpy as np
import torch

from .. import ops
from .._lib import ImwError
from ..hloc import extract_features, extractors, match_dense, match_features, matchers
from ..hloc.utils.base_model import dynamic_load

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"   # ui/utils.py:38

DEFAULT_RANSAC_METHOD = "B200_MAGSAC"
DEFAULT_RANSAC_REPROJ_THRESHOLD = 8
DEFAULT_RANSAC_CONFIDENCE = 0.9999
DEFAULT_RANSAC_MAX_ITER = 10000
DEFAULT_MIN_NUM_MATCHES = 4
MAGSAC_MAX_POINTS = 12288          # correspondences one CTA of imw_magsac holds in shared memory

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
MIN_SAMPLE = {"Homography": 4, "Fundamental": 7}


# ---- registry ---------------------------------------------------------------------------------------------------------
def parse_match_config(conf):
    """ui/utils.py:87-109: zoo entry {matcher, feature, dense/standalone} -> conf dicts of the registry."""
    out = {"matcher": (match_dense if conf["standalone"] else match_features).confs.get(conf["matcher"]),
           "standalone": bool(conf["standalone"]), "info": conf.get("info", {})}
    if not conf["standalone"]:
        out["feature"] = extract_features.confs.get(conf["feature"])
    return out


def get_matcher_zoo(matcher_zoo):
    """ui/utils.py:66-84."""
    return {name: parse_match_config(entry) for name, entry in matcher_zoo.items()}


def _load_plugin(root, conf, device):
    return dynamic_load(root, conf["model"]["name"])(conf["model"]).eval().to(device or DEVICE)


def get_model(match_conf: Dict[str, Any], device=None):
    """ui/utils.py:112-124 (the registry call site: dynamic_load over the matcher plugins)."""
    return _load_plugin(matchers, match_conf, device)


def get_feature_model(conf: Dict[str, Dict[str, Any]], device=None):
    """ui/utils.py:127-139."""
    return _load_plugin(extractors, conf, device)


# ---- estimators -------------------------------------------------------------------------------------------------------
def _residual_mask(M, p0, p1, geometry_type, thr):
    """inlier mask of a model over ALL correspondences with the estimator's residuals (re-projection error for H,
    Sampson distance for F; csrc/magsac.cu resid_h / resid_f), used when the set exceeds one CTA's capacity"""
    x0, y0, x1, y1 = (a.astype(np.float64) for a in (p0[:, 0], p0[:, 1], p1[:, 0], p1[:, 1]))
    if geometry_type == "Homography":
        w = M[2, 0] * x0 + M[2, 1] * y0 + M[2, 2]
        r2 = ((M[0, 0] * x0 + M[0, 1] * y0 + M[0, 2]) / w - x1) ** 2 + ((M[1, 0] * x0 + M[1, 1] * y0 + M[1, 2]) / w - y1) ** 2
    else:
        a = M[0, 0] * x0 + M[0, 1] * y0 + M[0, 2]; b = M[1, 0] * x0 + M[1, 1] * y0 + M[1, 2]; c = M[2, 0] * x0 + M[2, 1] * y0 + M[2, 2]
        at = M[0, 0] * x1 + M[1, 0] * y1 + M[2, 0]; bt = M[0, 1] * x1 + M[1, 1] * y1 + M[2, 1]
        r2 = (x1 * a + y1 * b + c) ** 2 / (a * a + b * b + at * at + bt * bt)
    return r2 <= thr * thr


def magsac_batch(point_sets: Sequence, geometry_type, reproj_threshold, confidence, max_iter, device="cuda", weights=None):
    """[(kp0 [K,2], kp1 [K,2]), ...] -> [(M [3,3] float64 | None, mask bool [K] | None), ...] with ONE imw_magsac launch.
    Sets larger than one CTA's shared memory are estimated on their `MAGSAC_MAX_POINTS` most confident correspondences
    (`weights`, else an even stride) and the mask is then evaluated over all of them."""
    if geometry_type not in MIN_SAMPLE:
        raise NotImplementedError(geometry_type)
    n = len(point_sets)
    res: List = [(None, None)] * n
    live, subsets = [], {}
    for i, (a, b) in enumerate(point_sets):
        k = len(a)
        if k < MIN_SAMPLE[geometry_type]:
            continue
        if k > MAGSAC_MAX_POINTS:
            w = None if weights is None else weights[i]
            subsets[i] = (np.argsort(-np.asarray(w), kind="stable")[:MAGSAC_MAX_POINTS] if w is not None and len(w) == k
                          else np.linspace(0, k - 1, MAGSAC_MAX_POINTS).astype(np.int64))
        live.append(i)
    if not live:
        return res
    sizes = [len(subsets[i]) if i in subsets else len(point_sets[i][0]) for i in live]
    cap = max(128, (max(sizes) + 127) // 128 * 128)
    h0 = np.zeros((len(live), cap, 2), np.float32); h1 = np.zeros_like(h0)
    for r, i in enumerate(live):
        a, b = (np.asarray(x, dtype=np.float32).reshape(-1, 2) for x in point_sets[i])
        if i in subsets:
            a, b = a[subsets[i]], b[subsets[i]]
        h0[r, :len(a)], h1[r, :len(b)] = a, b
    counts = torch.tensor(sizes, dtype=torch.int32, device=device)
    try:
        M, masks, n_inl, _ = ops.magsac(torch.from_numpy(h0).to(device), torch.from_numpy(h1).to(device), counts, geometry_type,
                                        reproj_threshold, confidence, max_iter)
    except (ImwError, ValueError):
        return res                                            # like cv2.error in the reference: no model (:360-362)
    M, masks, n_inl = M.cpu().numpy(), masks.cpu().numpy(), n_inl.cpu().numpy()      # one D2H per output for the whole batch
    for r, i in enumerate(live):
        if n_inl[r] == 0:
            continue
        if i in subsets:
            a, b = (np.asarray(x, dtype=np.float32).reshape(-1, 2) for x in point_sets[i])
            res[i] = (M[r], _residual_mask(M[r], a, b, geometry_type, reproj_threshold))
        else:
            res[i] = (M[r], masks[r, :sizes[r]].astype(bool))
    return res


def opencv_estimate(kp0, kp1, method, reproj_threshold, confidence, max_iter, geometry_type):
    """the reference's CPU estimator call (ui/utils.py:326-379): (M, mask bool) or (None, None)"""
    fn = {"Homography": cv2.findHomography, "Fundamental": cv2.findFundamentalMat}.get(geometry_type)
    if fn is None:
        raise NotImplementedError(geometry_type)
    try:
        M, mask = fn(kp0, kp1, method=method, ransacReprojThreshold=reproj_threshold, confidence=confidence, maxIters=max_iter)
    except cv2.error:
        return None, None
    return (None, None) if mask is None else (M, mask.ravel().astype(bool))


def estimate_batch(point_sets, ransac_method, reproj_threshold, confidence, max_iter, geometry_type, weights=None):
    if ransac_method.startswith("B200"):
        return magsac_batch(point_sets, geometry_type, reproj_threshold, confidence, max_iter, weights=weights)
    if ransac_method.startswith("CV2"):
        return [opencv_estimate(a, b, ransac_zoo[ransac_method], reproj_threshold, confidence, max_iter, geometry_type) for a, b in point_sets]
    raise NotImplementedError(f"{ransac_method}: not available in this build (poselib is not installed)")   # ui/utils.py:111


def proc_ransac_matches(mkpts0, mkpts1, ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=3.0,
                        ransac_confidence=0.99, ransac_max_iter=2000, geometry_type="Homography"):
    """ui/utils.py:424-456 for one correspondence set."""
    return estimate_batch([(mkpts0, mkpts1)], ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter, geometry_type)[0]


# ---- the pred-dict contract -------------------------------------------------------------------------------------------------
def set_null_pred(feature_type: Optional[str], pred: dict):
    """ui/utils.py:382-398: the 'no geometry' result (H is None, as in the reference :321)."""
    empty = {"KEYPOINT": ("mmkeypoints0_orig", "mmkeypoints1_orig", "mmconf"), "LINE": ("mline_keypoints0_orig", "mline_keypoints1_orig")}
    for key in empty.get(feature_type, ()):
        pred[key] = np.array([])
    pred["H"] = None
    pred["geom_info"] = {}
    return pred


def _correspondences(pred):
    for kind, k0, k1 in (("KEYPOINT", "mkeypoints0_orig", "mkeypoints1_orig"), ("LINE", "line_keypoints0_orig", "line_keypoints1_orig")):
        if k0 in pred and k1 in pred:
            return kind, pred[k0], pred[k1]
    return None, None, None


def verify_pairs(preds: List[Dict[str, Any]], ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=DEFAULT_RANSAC_REPROJ_THRESHOLD,
                 ransac_confidence=DEFAULT_RANSAC_CONFIDENCE, ransac_max_iter=DEFAULT_RANSAC_MAX_ITER):
    """filter_matches + compute_geometry (ui/utils.py:459-610) for a list of match dicts: pairs with >= 8 correspondences
    (:566) get F, then H, the uncalibrated rectification from ALL matches (:598-603) and the homography-inlier subset
    (:513-522); the others the null result.  Two estimator launches for the whole list."""
    if ransac_method not in ransac_zoo:
        ransac_method = DEFAULT_RANSAC_METHOD                                   # :499-500
    kinds, todo = [], []
    for i, pred in enumerate(preds):
        kind, m0, m1 = _correspondences(pred)
        kinds.append(kind)
        if m0 is None or len(m0) < DEFAULT_MIN_NUM_MATCHES:
            set_null_pred(kind, pred)
        elif len(m0) < 2 * DEFAULT_MIN_NUM_MATCHES:                             # compute_geometry returns {} (:566-567)
            set_null_pred(kind, pred)
        else:
            todo.append(i)
    if not todo:
        return preds
    sets = [(_correspondences(preds[i])[1], _correspondences(preds[i])[2]) for i in todo]
    wts = [preds[i].get("mconf") for i in todo]
    args = (ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter)
    Fs = estimate_batch(sets, *args, geometry_type="Fundamental", weights=wts)
    Hs = estimate_batch(sets, *args, geometry_type="Homography", weights=wts)
    for i, (m0, m1), (F, _), (Hm, mask_h) in zip(todo, sets, Fs, Hs):
        pred, info = preds[i], {}
        if F is not None:
            info["Fundamental"] = F.tolist()
        if Hm is None:
            set_null_pred(kinds[i], pred)
            pred["geom_info"] = info
            continue
        info["Homography"] = Hm.tolist()
        h0, w0 = pred["image0_orig"].shape[:2]
        try:
            if F is None:
                raise cv2.error("no fundamental matrix")
            _, H1, H2 = cv2.stereoRectifyUncalibrated(np.asarray(m0, np.float64).reshape(-1, 2), np.asarray(m1, np.float64).reshape(-1, 2),
                                                      F, imgSize=(w0, h0))
            info["H1"], info["H2"] = H1.tolist(), H2.tolist()
        except cv2.error:
            pass
        if kinds[i] == "KEYPOINT":
            pred["mmkeypoints0_orig"], pred["mmkeypoints1_orig"], pred["mmconf"] = m0[mask_h], m1[mask_h], pred["mconf"][mask_h]
        else:
            pred["mline_keypoints0_orig"], pred["mline_keypoints1_orig"] = m0[mask_h], m1[mask_h]
        pred["H"] = np.array(info["Homography"])
        pred["geom_info"] = info
    return preds


def compute_geometry(pred: Dict[str, Any], ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=DEFAULT_RANSAC_REPROJ_THRESHOLD,
                     ransac_confidence=DEFAULT_RANSAC_CONFIDENCE, ransac_max_iter=DEFAULT_RANSAC_MAX_ITER):
    """ui/utils.py:532-610: the geometry dict of one pair (incl. the masks the reference keeps until filter_matches pops them)."""
    _, m0, m1 = _correspondences(pred)
    if m0 is None or m1 is None or len(m0) < 2 * DEFAULT_MIN_NUM_MATCHES:
        return {}
    work = dict(pred)
    work.setdefault("mconf", np.ones(len(m0), np.float32))
    verify_pairs([work], ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter)
    return work.get("geom_info", {})


def filter_matches(pred: Dict[str, Any], ransac_method=DEFAULT_RANSAC_METHOD, ransac_reproj_threshold=DEFAULT_RANSAC_REPROJ_THRESHOLD,
                   ransac_confidence=DEFAULT_RANSAC_CONFIDENCE, ransac_max_iter=DEFAULT_RANSAC_MAX_ITER, ransac_estimator=None):
    """ui/utils.py:459-529: the reference's per-pair entry point = verify_pairs on a batch of one."""
    return verify_pairs([pred], ransac_method, ransac_reproj_threshold, ransac_confidence, ransac_max_iter)[0]
