"""Extraction driver -- drop-in for `extract` of imcui/hloc/extract_features.py:106-170, with the image preparation on
the GPU.

The reference converts, resizes and normalises every image with cv2 / torchvision on the host and then uploads a float
tensor; here the decoded uint8 frame is uploaded as it is and `imw_preprocess` (csrc/prepost.cu) produces the model input
on the device -- bit-identical to the reference's tensor for gray conversion, INTER_AREA resizes, /255 and the dfactor
alignment (tests/test_gpu_prepost.py).  `extract_batch` is the same call for many frames at once: frames of equal size
share one upload, one pre-processing pass and one extractor forward."""
import torch

from .configs import confs_dict
from .pipeline import FramePrep

confs = confs_dict["extractors"]


def _model_device(model):
    for t in model.buffers():
        return t.device
    for t in model.parameters():
        return t.device
    return torch.device("cuda")


def _feature_dict(pred, b, image, frame, original_size, size):
    """One image's entry of a (possibly batched) extractor output, in the reference's container types (:164-170)."""
    one = {}
    for k, v in pred.items():
        if isinstance(v, (list, tuple)):
            one[k] = type(v)([v[b]])
        elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] > b:
            one[k] = v[b:b + 1]
        else:
            one[k] = v
    one["image_size"] = original_size
    return {**one, "image": image, "original_size": original_size, "size": size, "image_orig": frame}


@torch.no_grad()
def extract_batch(model, frames, conf):
    """frames: list of uint8 RGB [H,W,3] / gray [H,W] arrays (any sizes).  Returns one feature dict per frame with the keys
    `extract` returns.  Same-size frames are pre-processed and extracted together."""
    prepped = FramePrep(conf, _model_device(model))(frames)
    groups = {}
    for i, (img, _, _) in enumerate(prepped):
        groups.setdefault(tuple(img.shape), []).append(i)
    out = [None] * len(frames)
    for _, idx in groups.items():
        batch = torch.cat([prepped[i][0] for i in idx]) if len(idx) > 1 else prepped[idx[0]][0]
        pred = model({"image": batch})
        for b, i in enumerate(idx):
            out[i] = _feature_dict(pred, b, prepped[i][0], frames[i], prepped[i][1], prepped[i][2])
    return out


def extract(model, image_0, conf):
    """extract_features.py:106-170: image_0 uint8 RGB [H,W,3] (or gray [H,W]) -> {keypoints, scores, descriptors, image,
    original_size, size, image_orig, image_size}."""
    return extract_batch(model, [image_0], conf)[0]
