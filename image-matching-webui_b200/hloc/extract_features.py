"""In-memory extraction driver -- mirror of the reference's imcui/hloc/extract_features.py:27-41 (resize_image)
and :106-170 (extract): RGB->gray with cv2 on uint8, INTER_AREA resizes, /255, torchvision antialias resize to a
multiple of dfactor, H2D, call the extractor, merge dicts.  Pre-processing stays on the host with the very same
cv2 / torchvision calls so that the model input is bit-identical to the reference's (SURVEY.md 8(f) rank 1 lists
a GPU version as the next row)."""
from types import SimpleNamespace

import cv2
import numpy as np
import torch
import torchvision.transforms.functional as F

from .configs import confs_dict

confs = confs_dict["extractors"]


def resize_image(image, size, interp):
    """extract_features.py:27-41 (cv2 branch; the PIL branch is not used by extract())."""
    if not interp.startswith("cv2_"):
        raise ValueError(f"Unknown interpolation {interp}.")
    interp = getattr(cv2, "INTER_" + interp[len("cv2_"):].upper())
    h, w = image.shape[:2]
    if interp == cv2.INTER_AREA and (w < size[0] or h < size[1]):
        interp = cv2.INTER_LINEAR
    return cv2.resize(image, size, interpolation=interp)


def preprocess(image: np.ndarray, conf: SimpleNamespace, device):
    """extract_features.py:120-156."""
    image = image.astype(np.float32, copy=False)
    size = image.shape[:2][::-1]
    if conf.resize_max:
        scale = conf.resize_max / max(size)
        if scale < 1.0:
            size_new = tuple(int(round(x * scale)) for x in size)
            image = resize_image(image, size_new, "cv2_area")
    if conf.force_resize:
        image = resize_image(image, (conf.width, conf.height), "cv2_area")
    if conf.grayscale:
        assert image.ndim == 2, image.shape
        image = image[None]
    else:
        image = image.transpose((2, 0, 1))
    image = torch.from_numpy(image / 255.0).float()
    size_new = tuple(map(lambda x: int(x // conf.dfactor * conf.dfactor), image.shape[-2:]))
    image = F.resize(image, size=size_new, antialias=True)
    input_ = image.to(device, non_blocking=True)[None]
    return {"image": input_, "original_size": np.array(size), "size": np.array(image.shape[1:][::-1])}


def extract(model, image_0, conf):
    """extract_features.py:106-170.  image_0: uint8 RGB [H,W,3] (or gray [H,W])."""
    default_conf = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "cache_images": False, "force_resize": False,
                    "width": 320, "height": 240, "interpolation": "cv2_area"}
    conf = SimpleNamespace(**{**default_conf, **conf})
    device = next(model.buffers()).device if any(True for _ in model.buffers()) else ("cuda" if torch.cuda.is_available() else "cpu")
    if len(image_0.shape) == 3 and conf.grayscale:
        image0 = cv2.cvtColor(image_0, cv2.COLOR_RGB2GRAY)
    else:
        image0 = image_0
    data = preprocess(image0, conf, device)
    data["image_orig"] = image_0
    pred = model({"image": data["image"]})
    pred["image_size"] = data["original_size"]
    return {**pred, **data}
