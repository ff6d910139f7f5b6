"""In-memory dense matching driver -- mirror of the reference's imcui/hloc/match_dense.py:577-724 (match_images) for
the detector-free matchers of the hot path (LoFTR).  Pre-processing is the reference's: cv2 RGB->gray on uint8,
INTER_AREA resizes, /255, torchvision resize to a multiple of dfactor; then the matcher, then keypoints rescaled to the
original frame with (k + 0.5) * s - 0.5.  The line-matching branch (:687-722) belongs to matchers outside the path."""
from types import SimpleNamespace

import cv2
import numpy as np
import torch
import torchvision.transforms.functional as F

from .configs import confs_dict
from .extract_features import resize_image
from .match_features import scale_keypoints

confs = confs_dict["matchers"]


def preprocess(image: np.ndarray, conf: SimpleNamespace):
    """match_dense.py:588-620 (the returned scale is recomputed by the caller exactly as the reference does)."""
    image = image.astype(np.float32, copy=False)
    size = image.shape[:2][::-1]
    scale = np.array([1.0, 1.0])
    if conf.resize_max:
        scale = conf.resize_max / max(size)
        if scale < 1.0:
            size_new = tuple(int(round(x * scale)) for x in size)
            image = resize_image(image, size_new, "cv2_area")
            scale = np.array(size) / np.array(size_new)
    if conf.force_resize:
        size = image.shape[:2][::-1]
        image = resize_image(image, (conf.width, conf.height), "cv2_area")
        size_new = (conf.width, conf.height)
        scale = np.array(size) / np.array(size_new)
    if conf.grayscale:
        assert image.ndim == 2, image.shape
        image = image[None]
    else:
        image = image.transpose((2, 0, 1))
    image = torch.from_numpy(image / 255.0).float()
    size_new = tuple(map(lambda x: int(x // conf.dfactor * conf.dfactor), image.shape[-2:]))
    image = F.resize(image, size=size_new)
    scale = np.array(size) / np.array(size_new)[::-1]
    return image, scale


@torch.no_grad()
def match_images(model, image_0, image_1, conf, device="cuda"):
    """match_dense.py:577-724.  image_0/1: uint8 RGB [H,W,3] (or gray [H,W]); returns the reference's dict."""
    default_conf = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "cache_images": False, "force_resize": False,
                    "width": 320, "height": 240}
    conf = SimpleNamespace(**{**default_conf, **conf})
    image0 = cv2.cvtColor(image_0, cv2.COLOR_RGB2GRAY) if len(image_0.shape) == 3 and conf.grayscale else image_0
    image1 = cv2.cvtColor(image_1, cv2.COLOR_RGB2GRAY) if len(image_0.shape) == 3 and conf.grayscale else image_1  # sic: :628
    image0, _ = preprocess(image0, conf)
    image1, _ = preprocess(image1, conf)
    image0 = image0.to(device)[None]
    image1 = image1.to(device)[None]
    pred = model({"image0": image0, "image1": image1})
    s0 = np.array(image_0.shape[:2][::-1]) / np.array(image0.shape[-2:][::-1])
    s1 = np.array(image_1.shape[:2][::-1]) / np.array(image1.shape[-2:][::-1])
    ret = {}
    if "keypoints0" in pred and "keypoints1" in pred:
        # one D2H for everything the caller reads
        kpts0, kpts1 = pred["keypoints0"].cpu(), pred["keypoints1"].cpu()
        mkpts0, mkpts1 = pred.get("mkeypoints0"), pred.get("mkeypoints1")
        if mkpts0 is None or mkpts1 is None:
            mkpts0, mkpts1 = kpts0, kpts1
        else:
            mkpts0, mkpts1 = mkpts0.cpu(), mkpts1.cpu()
        # scale_keypoints works in place on "kpts + 0.5" temporaries, like the reference
        ret = {
            "image0": image0.squeeze().cpu().numpy(), "image1": image1.squeeze().cpu().numpy(),
            "image0_orig": image_0, "image1_orig": image_1,
            "keypoints0": kpts0.numpy(), "keypoints1": kpts1.numpy(),
            "keypoints0_orig": (scale_keypoints(kpts0 + 0.5, s0) - 0.5).numpy(),
            "keypoints1_orig": (scale_keypoints(kpts1 + 0.5, s1) - 0.5).numpy(),
            "mkeypoints0": mkpts0.numpy(), "mkeypoints1": mkpts1.numpy(),
            "mkeypoints0_orig": (scale_keypoints(mkpts0 + 0.5, s0) - 0.5).numpy(),
            "mkeypoints1_orig": (scale_keypoints(mkpts1 + 0.5, s1) - 0.5).numpy(),
            "original_size0": np.array(image_0.shape[:2][::-1]), "original_size1": np.array(image_1.shape[:2][::-1]),
            "new_size0": np.array(image0.shape[-2:][::-1]), "new_size1": np.array(image1.shape[-2:][::-1]),
            "scale0": s0, "scale1": s1,
        }
        if "mconf" in pred:
            ret["mconf"] = pred["mconf"].cpu().numpy()
        elif "scores" in pred:  # adapting loftr (:683)
            ret["mconf"] = pred["scores"].cpu().numpy()
        else:
            ret["mconf"] = np.ones_like(kpts0.numpy()[:, 0])
    # (the reference calls torch.cuda.empty_cache() per pair, :723; workspaces are cached on purpose here)
    return ret
