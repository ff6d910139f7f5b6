"""Dense matching driver -- drop-in for `match_images` of imcui/hloc/match_dense.py:577-724 for the detector-free matchers
of the hot path (LoFTR), with image preparation and result preparation on the GPU (see extract_features.py /
match_features.py).  The line-matching branch (:687-722) belongs to matchers outside the path."""
import numpy as np
import torch

from .configs import confs_dict
from .pipeline import FramePrep, PairResult

confs = confs_dict["matchers"]


@torch.no_grad()
def match_images(model, image_0, image_1, conf, device="cuda"):
    """match_dense.py:577-724: two uint8 RGB [H,W,3] (or gray [H,W]) frames -> the reference's result dict (NumPy)."""
    if image_0.ndim == 2 and image_1.ndim == 3 and dict(conf).get("grayscale", True):
        # match_dense.py:626-633 tests image_0's rank for BOTH images: an RGB image_1 next to a gray image_0 stays RGB there
        # and trips the reference's own `assert image.ndim == 2` (:604)
        raise AssertionError(image_1.shape)
    (img0, orig0, size0), (img1, orig1, size1) = FramePrep(conf, device)([image_0, image_1])
    pred = model({"image0": img0, "image1": img1})
    if "keypoints0" not in pred or "keypoints1" not in pred:
        return {}
    s0, s1 = orig0 / size0, orig1 / size1
    conf_t = pred["mconf"] if "mconf" in pred else pred.get("scores")      # "adapting loftr" (:683)
    res = PairResult.dense(pred["keypoints0"], pred["keypoints1"], conf_t, s0, s1)
    mk0, mk1 = pred.get("mkeypoints0"), pred.get("mkeypoints1")
    if mk0 is None or mk1 is None:
        m = {"mkeypoints0": res["keypoints0"], "mkeypoints1": res["keypoints1"],
             "mkeypoints0_orig": res["keypoints0_orig"], "mkeypoints1_orig": res["keypoints1_orig"]}
    else:
        r2 = PairResult.dense(mk0, mk1, None, s0, s1)
        m = {"mkeypoints0": r2["keypoints0"], "mkeypoints1": r2["keypoints1"],
             "mkeypoints0_orig": r2["keypoints0_orig"], "mkeypoints1_orig": r2["keypoints1_orig"]}
    return {"image0": img0.squeeze().cpu().numpy(), "image1": img1.squeeze().cpu().numpy(),
            "image0_orig": image_0, "image1_orig": image_1, **res, **m,
            "original_size0": orig0, "original_size1": orig1, "new_size0": size0, "new_size1": size1,
            "scale0": s0, "scale1": s1}
