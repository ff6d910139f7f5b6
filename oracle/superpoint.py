"""Oracle for SuperPoint.

Restates third_party/SuperGluePretrainedNetwork/models/superpoint.py:47-206 as plain functions over
a flat weight dict.  Test infrastructure only (see oracle/__init__.py).
"""
import torch
import torch.nn.functional as F

DEFAULT_CONF = {  # hloc/extractors/superpoint.py:34-41
    "nms_radius": 4,
    "keypoint_threshold": 0.005,
    "max_keypoints": -1,
    "remove_borders": 4,
}


def _conv(x, w, name, relu=True, pad=1):
    y = F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=1, padding=pad)
    return F.relu(y) if relu else y


def encoder(w, image):
    """Shared VGG encoder, superpoint.py:152-162.  image [B,1,H,W] -> [B,128,H/8,W/8]."""
    x = _conv(image, w, "conv1a")
    x = _conv(x, w, "conv1b")
    x = F.max_pool2d(x, 2, 2)
    x = _conv(x, w, "conv2a")
    x = _conv(x, w, "conv2b")
    x = F.max_pool2d(x, 2, 2)
    x = _conv(x, w, "conv3a")
    x = _conv(x, w, "conv3b")
    x = F.max_pool2d(x, 2, 2)
    x = _conv(x, w, "conv4a")
    x = _conv(x, w, "conv4b")
    return x


def detector_logits(w, feat):
    """convPa+ReLU, convPb (1x1, 65 ch), superpoint.py:165-166."""
    cPa = _conv(feat, w, "convPa")
    return _conv(cPa, w, "convPb", relu=False, pad=0)


def dense_scores(logits):
    """softmax-65, drop dustbin, 8x8 depth-to-space, superpoint.py:167-170. -> [B,8h,8w]."""
    s = F.softmax(logits, 1)[:, :-1]
    b, _, h, w = s.shape
    s = s.permute(0, 2, 3, 1).reshape(b, h, w, 8, 8)
    return s.permute(0, 1, 3, 2, 4).reshape(b, h * 8, w * 8)


def simple_nms(scores, r):
    """superpoint.py:47-62: three rounds of max-pool NMS with exact float equality."""
    assert r >= 0

    def mp(x):
        return F.max_pool2d(x, kernel_size=2 * r + 1, stride=1, padding=r)

    zeros = torch.zeros_like(scores)
    max_mask = scores == mp(scores)
    for _ in range(2):
        supp = mp(max_mask.float()) > 0
        supp_scores = torch.where(supp, zeros, scores)
        new_max = supp_scores == mp(supp_scores)
        max_mask = max_mask | (new_max & (~supp))
    return torch.where(max_mask, scores, zeros)


def select_keypoints(nms_scores, conf):
    """Threshold -> row-major nonzero -> border filter -> top-k, superpoint.py:174-191.
    nms_scores [H,W] of one image.  Returns kpts [N,2] float (x,y), scores [N]."""
    H, W = nms_scores.shape
    k = torch.nonzero(nms_scores > conf["keypoint_threshold"])  # (y,x) row-major
    s = nms_scores[tuple(k.t())]
    b = conf["remove_borders"]
    m = (k[:, 0] >= b) & (k[:, 0] < H - b) & (k[:, 1] >= b) & (k[:, 1] < W - b)
    k, s = k[m], s[m]
    mk = conf["max_keypoints"]
    if mk >= 0 and mk < len(k):
        s, idx = torch.topk(s, mk, dim=0)
        k = k[idx]
    return torch.flip(k, [1]).float(), s


def dense_descriptors(w, feat):
    """convDa+ReLU, convDb (1x1), channel L2 norm, superpoint.py:194-196."""
    cDa = _conv(feat, w, "convDa")
    d = _conv(cDa, w, "convDb", relu=False, pad=0)
    return F.normalize(d, p=2, dim=1)


def sample_descriptors(kpts, desc, s=8):
    """superpoint.py:80-92 (fix_sampling=False): bilinear, align_corners=True, then L2 norm.
    kpts [1,N,2] (x,y) pixels, desc [1,C,h,w] -> [1,C,N]."""
    b, c, h, w = desc.shape
    k = kpts - s / 2 + 0.5
    k = k / torch.tensor([(w * s - s / 2 - 0.5), (h * s - s / 2 - 0.5)]).to(k)[None]
    k = k * 2 - 1
    d = F.grid_sample(desc, k.view(b, 1, -1, 2), mode="bilinear", align_corners=True)
    return F.normalize(d.reshape(b, c, -1), p=2, dim=1)


def sample_descriptors_fix_sampling(kpts, desc, s=8):
    """hloc/extractors/superpoint.py:16-30 (fix_sampling=True): (k + 0.5) / ([w, h] * s), align_corners=False."""
    b, c, h, w = desc.shape
    k = (kpts + 0.5) / (torch.tensor([w, h]).to(kpts) * s)
    k = k * 2 - 1
    d = F.grid_sample(desc, k.view(b, 1, -1, 2), mode="bilinear", align_corners=False)
    return F.normalize(d.reshape(b, c, -1), p=2, dim=1)


def forward(w, image, conf=None, return_dense=False):
    """superpoint.py:145-206.  image [B,1,H,W] fp32 in [0,1].
    Returns {"keypoints": [ [N,2] ], "scores": ( [N] ), "descriptors": [ [256,N] ]}."""
    conf = {**DEFAULT_CONF, **(conf or {})}
    mk = conf["max_keypoints"]
    if mk == 0 or mk < -1:
        raise ValueError('"max_keypoints" must be positive or "-1"')  # superpoint.py:139-141
    feat = encoder(w, image)
    dense = dense_scores(detector_logits(w, feat))
    nms = simple_nms(dense, conf["nms_radius"])
    kpts, scores = [], []
    for b in range(image.shape[0]):
        k, s = select_keypoints(nms[b], conf)
        kpts.append(k)
        scores.append(s)
    dd = dense_descriptors(w, feat)
    sample = sample_descriptors_fix_sampling if conf.get("fix_sampling") else sample_descriptors
    descs = [sample(k[None], d[None], 8)[0] for k, d in zip(kpts, dd)]
    out = {"keypoints": kpts, "scores": tuple(scores), "descriptors": descs}
    if return_dense:
        out["dense_scores"] = dense
        out["nms_scores"] = nms
        out["dense_descriptors"] = dd
    return out
