"""In-memory matching driver -- mirror of the reference's imcui/hloc/match_features.py:189-275
(scale_keypoints, match_images): pack the pair dict, call the matcher, one D2H of the outputs, gather the
matched keypoints and rescale them to the original image with (k + 0.5) * s - 0.5."""
import numpy as np
import torch

from .configs import confs_dict

confs = confs_dict["matchers"]


def scale_keypoints(kpts, scale):
    """match_features.py:189-201."""
    if isinstance(scale, (list, tuple, np.ndarray)) and len(scale) == 2 and np.any(scale != np.array([1.0, 1.0])):
        kpts[:, 0] *= scale[0]
        kpts[:, 1] *= scale[1]
    return kpts


@torch.no_grad()
def match_images(model, feat0, feat1):
    """match_features.py:204-275."""
    desc0, desc1 = feat0["descriptors"][0], feat1["descriptors"][0]
    if len(desc0.shape) == 2:
        desc0 = desc0.unsqueeze(0)
    if len(desc1.shape) == 2:
        desc1 = desc1.unsqueeze(0)
    if isinstance(feat0["keypoints"], list):
        feat0["keypoints"] = feat0["keypoints"][0][None]
    if isinstance(feat1["keypoints"], list):
        feat1["keypoints"] = feat1["keypoints"][0][None]
    input_dict = {
        "image0": feat0["image"], "keypoints0": feat0["keypoints"], "scores0": feat0["scores"][0].unsqueeze(0),
        "descriptors0": desc0,
        "image1": feat1["image"], "keypoints1": feat1["keypoints"], "scores1": feat1["scores"][0].unsqueeze(0),
        "descriptors1": desc1,
    }
    for k in ("scales", "oris"):
        if k in feat0:
            input_dict[k + "0"] = feat0[k]
        if k in feat1:
            input_dict[k + "1"] = feat1[k]
    pred = model(input_dict)
    pred = {k: v.cpu().detach()[0] if isinstance(v, torch.Tensor) else v for k, v in pred.items()}
    kpts0, kpts1 = feat0["keypoints"][0].cpu().numpy(), feat1["keypoints"][0].cpu().numpy()
    matches, confid = pred["matches0"], pred["matching_scores0"]
    valid = matches > -1
    mkpts0 = kpts0[valid]
    mkpts1 = kpts1[matches[valid]]
    mconfid = confid[valid]
    s0 = feat0["original_size"] / feat0["size"]
    s1 = feat1["original_size"] / feat1["size"]
    kpts0_origin = scale_keypoints(torch.from_numpy(kpts0 + 0.5), s0) - 0.5
    kpts1_origin = scale_keypoints(torch.from_numpy(kpts1 + 0.5), s1) - 0.5
    mkpts0_origin = scale_keypoints(torch.from_numpy(mkpts0 + 0.5), s0) - 0.5
    mkpts1_origin = scale_keypoints(torch.from_numpy(mkpts1 + 0.5), s1) - 0.5
    # (the reference calls torch.cuda.empty_cache() here for every pair, match_features.py:273; the engine's
    # workspaces are cached on purpose, so this is dropped)
    return {
        "image0_orig": feat0["image_orig"], "image1_orig": feat1["image_orig"],
        "keypoints0": kpts0, "keypoints1": kpts1,
        "keypoints0_orig": kpts0_origin.numpy(), "keypoints1_orig": kpts1_origin.numpy(),
        "mkeypoints0": mkpts0, "mkeypoints1": mkpts1,
        "mkeypoints0_orig": mkpts0_origin.numpy(), "mkeypoints1_orig": mkpts1_origin.numpy(),
        "mconf": mconfid.numpy(),
    }
