"""CPU restatement of ALIKED (aliked-n16) -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Follows the reference's third_party/LightGlue/lightglue/aliked.py:
  extract_dense_map :709-740, ConvBlock :386-419, ResBlock :422-476, DeformableConv2d :291-349 (torchvision
  deform_conv2d restated as a bilinear gather + contraction), InputPadder :264-288, simple_nms :68-91, DKD :94-261,
  SDDH :479-609, get_patches :49-65, ALIKED.forward :757-775.
Pinned by tests/golden/aliked.npz, minted from the UNMODIFIED reference module (tools/make_golden.py) loaded with the
deterministic random weights of `random_weights` (no aliked-n16.pth exists offline: parity is random-weight parity)."""
import math

import torch
import torch.nn.functional as F

CFG = {"c": (16, 32, 64, 128), "dim": 128, "K": 3, "M": 16}
SELU = F.selu


def _bn(w, p, x):
    return F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"], False, 0.0, 1e-5)


def deform_conv3x3(x, offset, weight):
    """torchvision.ops.deform_conv2d (3x3, stride 1, pad 1, dilation 1, one offset group, no mask): the sample of tap
    k = (ky, kx) at output (y, x) is x[:, y - 1 + ky + offset[2k], x - 1 + kx + offset[2k+1]], bilinear with zeros
    outside the map; out = sum_k W[:, :, ky, kx] @ sample_k."""
    B, C, H, W = x.shape
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    cols = []
    for k in range(9):
        ky, kx = divmod(k, 3)
        py = ys - 1 + ky + offset[:, 2 * k]
        px = xs - 1 + kx + offset[:, 2 * k + 1]
        y0, x0 = torch.floor(py), torch.floor(px)
        ly, lx = py - y0, px - x0
        val = torch.zeros(B, C, H, W, dtype=x.dtype)
        for dy, wy in ((0, 1 - ly), (1, ly)):
            for dx, wx in ((0, 1 - lx), (1, lx)):
                yy, xx = (y0 + dy).long(), (x0 + dx).long()
                ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).view(B, 1, -1).expand(B, C, -1)
                g = torch.gather(x.reshape(B, C, -1), 2, idx).view(B, C, H, W)
                val = val + g * (wy * wx * ok).unsqueeze(1)
        cols.append(val)
    cols = torch.stack(cols, 2)  # B, C, 9, H, W
    return torch.einsum("ock,bckhw->bohw", weight.reshape(weight.shape[0], C, 9), cols)


def _conv(w, p, x, dcn):
    if not dcn:
        return F.conv2d(x, w[p + "weight"], None, padding=1)
    h, wd = x.shape[2:]
    max_offset = max(h, wd) / 4.0
    off = F.conv2d(x, w[p + "offset_conv.weight"], w[p + "offset_conv.bias"], padding=1).clamp(-max_offset, max_offset)
    return deform_conv3x3(x, off, w[p + "regular_conv.weight"])


def _resblock(w, p, x, dcn):
    out = SELU(_bn(w, p + "bn1.", _conv(w, p + "conv1.", x, dcn)))
    out = _bn(w, p + "bn2.", _conv(w, p + "conv2.", out, dcn))
    idt = F.conv2d(x, w[p + "downsample.weight"], w[p + "downsample.bias"])
    return SELU(out + idt)


def _up(x, s):
    return F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True)


def extract_dense_map(w, image):
    """aliked.py:709-740 -> (feature_map [B,128,H,W] L2-normalised, score_map [B,1,H,W])."""
    h, wd = image.shape[-2:]
    ph, pw = ((h // 32 + 1) * 32 - h) % 32, ((wd // 32 + 1) * 32 - wd) % 32
    pad = [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]
    x = F.pad(image, pad, mode="replicate")
    x1 = SELU(_bn(w, "block1.bn1.", F.conv2d(x, w["block1.conv1.weight"], None, padding=1)))
    x1 = SELU(_bn(w, "block1.bn2.", F.conv2d(x1, w["block1.conv2.weight"], None, padding=1)))
    x2 = _resblock(w, "block2.", F.avg_pool2d(x1, 2, 2), False)
    x3 = _resblock(w, "block3.", F.avg_pool2d(x2, 4, 4), True)
    x4 = _resblock(w, "block4.", F.avg_pool2d(x3, 4, 4), True)
    a1 = SELU(F.conv2d(x1, w["conv1.weight"]))
    a2 = SELU(F.conv2d(x2, w["conv2.weight"]))
    a3 = SELU(F.conv2d(x3, w["conv3.weight"]))
    a4 = SELU(F.conv2d(x4, w["conv4.weight"]))
    x1234 = torch.cat([a1, _up(a2, 2), _up(a3, 8), _up(a4, 32)], 1)
    s = SELU(F.conv2d(x1234, w["score_head.0.weight"]))
    s = SELU(F.conv2d(s, w["score_head.2.weight"], padding=1))
    s = SELU(F.conv2d(s, w["score_head.4.weight"], padding=1))
    score = torch.sigmoid(F.conv2d(s, w["score_head.6.weight"], padding=1))
    feat = F.normalize(x1234, p=2, dim=1)
    H2, W2 = feat.shape[-2:]
    sl = (slice(pad[2], H2 - pad[3]), slice(pad[0], W2 - pad[1]))
    return feat[..., sl[0], sl[1]], score[..., sl[0], sl[1]]


def simple_nms(scores, r):
    """aliked.py:68-91 (same scheme as SuperPoint's)."""
    mp = lambda t: F.max_pool2d(t, 2 * r + 1, 1, r)
    zeros = torch.zeros_like(scores)
    max_mask = scores == mp(scores)
    for _ in range(2):
        supp = mp(max_mask.float()) > 0
        ss = torch.where(supp, zeros, scores)
        max_mask = max_mask | ((ss == mp(ss)) & ~supp)
    return torch.where(max_mask, scores, zeros)


def dkd(score_map, radius=2, scores_th=0.2, n_limit=20000, top_k=-1):
    """DKD.forward :127-222 for B = 1 (threshold / top-k selection, 5x5 soft-argmax at T = 0.1, bilinear score).
    Returns keypoints in [-1, 1] [N,2], scores [N], integer cell indices [N]."""
    _, _, h, w = score_map.shape
    nms = simple_nms(score_map, radius)
    nms[:, :, :radius] = 0; nms[:, :, :, :radius] = 0; nms[:, :, -radius:] = 0; nms[:, :, :, -radius:] = 0
    flat = score_map.reshape(-1)
    if top_k > 0:
        idx = torch.topk(nms.view(-1), top_k).indices
    else:
        if scores_th > 0:
            mask = nms > scores_th
            if mask.sum() == 0:
                mask = nms > flat.mean()
        else:
            mask = nms > flat.mean()
        idx = mask.reshape(-1).nonzero()[:, 0]
        if len(idx) > n_limit:
            idx = idx[flat[idx].sort(descending=True)[1][:n_limit]]
    k = 2 * radius + 1
    lin = torch.linspace(-radius, radius, k)
    grid = torch.stack(torch.meshgrid(lin, lin, indexing="ij")).view(2, -1).t()[:, [1, 0]]  # (x, y) offsets, row-major window
    patches = F.unfold(score_map, k, padding=radius)[0].t()[idx]                             # N x k*k (zero padded)
    xy_nms = torch.stack([idx % w, idx // w], 1)
    x_exp = ((patches - patches.max(1, keepdim=True).values) / 0.1).exp()
    resid = x_exp @ grid / x_exp.sum(1, keepdim=True)
    wh = torch.tensor([w - 1, h - 1], dtype=score_map.dtype)
    kp = (xy_nms + resid) / wh * 2 - 1
    sc = F.grid_sample(score_map, kp.view(1, 1, -1, 2), mode="bilinear", align_corners=True)[0, 0, 0]
    return kp, sc, idx


def sddh(w, feat, kpts, K=3, M=16):
    """SDDH.forward :536-609 for B = 1: KxK patch at the truncated keypoint -> offsets -> M bilinear samples ->
    1x1 conv + SELU -> per-position aggregation -> L2 norm."""
    _, c, h, wd = feat.shape
    wh = torch.tensor([[wd - 1, h - 1]], dtype=feat.dtype)
    max_offset = max(h, wd) / 4.0
    kwh = (kpts / 2 + 0.5) * wh
    corner = (kwh.long() - K / 2 + 1).long()   # get_patches :52-55
    corner[:, 0] = corner[:, 0].clamp(0, wd - 1 - K)
    corner[:, 1] = corner[:, 1].clamp(0, h - 1 - K)
    o = torch.arange(K)
    # get_patches :57-65 yields patch[n, c, i, j] = feat[c, cy + i, cx + j] (checked against the reference function)
    yy = corner[:, 1, None, None] + o[None, :, None]
    xx = corner[:, 0, None, None] + o[None, None, :]
    patch = feat[0][:, yy, xx].permute(1, 0, 2, 3)        # N, C, K, K
    off = F.conv2d(patch, w["desc_head.offset_conv.0.weight"], w["desc_head.offset_conv.0.bias"])
    off = F.conv2d(SELU(off), w["desc_head.offset_conv.2.weight"], w["desc_head.offset_conv.2.bias"]).clamp(-max_offset, max_offset)
    off = off[:, :, 0, 0].view(-1, 2, M).permute(0, 2, 1)  # N, M, 2
    pos = 2.0 * (kwh.unsqueeze(1) + off) / wh[None] - 1
    f = F.grid_sample(feat, pos.reshape(1, -1, 1, 2), mode="bilinear", align_corners=True)   # 1, C, N*M, 1
    f = f.reshape(c, -1, M, 1).permute(1, 0, 2, 3)                                              # N, C, M, 1
    f = SELU(F.conv2d(f, w["desc_head.sf_conv.weight"])).squeeze(-1)                            # N, C, M
    d = torch.einsum("ncp,pcd->nd", f, w["desc_head.agg_weights"])
    return F.normalize(d, p=2.0, dim=1), off


def forward(w, image, detection_threshold=0.2, max_num_keypoints=-1, nms_radius=2):
    """ALIKED.forward :757-775 (B = 1).  image [1,3,H,W] or [1,1,H,W] in [0,1]."""
    if image.shape[1] == 1:
        image = image.expand(-1, 3, -1, -1)
    feat, score = extract_dense_map(w, image)
    kp, sc, idx = dkd(score, nms_radius, detection_threshold, max_num_keypoints if max_num_keypoints > 0 else 20000,
                      -1 if detection_threshold > 0 else max_num_keypoints)
    desc, off = sddh(w, feat, kp)
    _, _, h, wd = image.shape
    wh = torch.tensor([wd - 1, h - 1], dtype=image.dtype)
    return {"keypoints": wh * (kp + 1) / 2.0, "descriptors": desc, "keypoint_scores": sc, "indices": idx,
            "score_map": score, "feature_map": feat, "offsets": off}


def _synth_weights():
    """image-matching-webui_b200/utils/synth_weights.py loaded by path (no import of the product package)."""
    import importlib.util
    from pathlib import Path
    p = Path(__file__).resolve().parent.parent / "image-matching-webui_b200" / "utils" / "synth_weights.py"
    spec = importlib.util.spec_from_file_location("_imw_synth_weights", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def random_weights(seed=0):
    """Deterministic random aliked-n16 parameters (reference key names): see utils/synth_weights.py."""
    return _synth_weights().aliked_random_weights(seed)
