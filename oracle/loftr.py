"""Oracle for LoFTR.

hloc's `loftr` matcher calls kornia.feature.LoFTR (un-vendored, not installed); the in-tree source it ports is
third_party/SE2LoFTR/src/loftr/*, which this file restates as plain functions over a flat weight dict (BatchNorm in
eval mode, default cvpr_ds_config).  No standard LoFTR checkpoint exists in the tree, so parity is pinned with SEEDED
RANDOM weights of that copy (real-weight parity: unpinned).  Test infrastructure only."""
import math

import torch
import torch.nn.functional as F

CFG = {  # utils/cvpr_ds_config.py:10-50
    "d_model": 256, "nhead": 8, "coarse_layers": ["self", "cross"] * 4, "fine_layers": ["self", "cross"],
    "d_fine": 128, "thr": 0.2, "border_rm": 2, "temperature": 0.1, "fine_window": 5, "temp_bug_fix": False,
}


def _bn(w, p, x):
    return F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"], False, 0.0, 1e-5)


def _block(w, p, x, stride):
    """BasicBlock, backbone/resnet_fpn.py:15-41."""
    y = F.relu(_bn(w, p + "bn1.", F.conv2d(x, w[p + "conv1.weight"], None, stride, 1)))
    y = _bn(w, p + "bn2.", F.conv2d(y, w[p + "conv2.weight"], None, 1, 1))
    if stride != 1:
        x = _bn(w, p + "downsample.1.", F.conv2d(x, w[p + "downsample.0.weight"], None, stride, 0))
    return F.relu(x + y)


def backbone(w, x):
    """ResNetFPN_8_2.forward, backbone/resnet_fpn.py:100-118.  x [N,1,H,W] -> coarse [N,256,H/8,W/8], fine [N,128,H/2,W/2]."""
    p = "backbone."
    x0 = F.relu(_bn(w, p + "bn1.", F.conv2d(x, w[p + "conv1.weight"], None, 2, 3)))
    x1 = _block(w, p + "layer1.1.", _block(w, p + "layer1.0.", x0, 1), 1)
    x2 = _block(w, p + "layer2.1.", _block(w, p + "layer2.0.", x1, 2), 1)
    x3 = _block(w, p + "layer3.1.", _block(w, p + "layer3.0.", x2, 2), 1)
    x3_out = F.conv2d(x3, w[p + "layer3_outconv.weight"])
    x3_2x = F.interpolate(x3_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x2_out = F.conv2d(x2, w[p + "layer2_outconv.weight"])

    def outconv2(pp, t):
        t = F.conv2d(t, w[pp + "0.weight"], None, 1, 1)
        t = F.leaky_relu(_bn(w, pp + "1.", t), 0.01)
        return F.conv2d(t, w[pp + "3.weight"], None, 1, 1)

    x2_out = outconv2(p + "layer2_outconv2.", x2_out + x3_2x)
    x2_2x = F.interpolate(x2_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x1_out = F.conv2d(x1, w[p + "layer1_outconv.weight"])
    x1_out = outconv2(p + "layer1_outconv2.", x1_out + x2_2x)
    return x3_out, x1_out


def position_encoding(d_model, h, w_, temp_bug_fix=False):
    """utils/position_encoding.py:6-42 (the default is the 'buggy' div_term)."""
    pe = torch.zeros((d_model, h, w_))
    y_pos = torch.ones((h, w_)).cumsum(0).float().unsqueeze(0)
    x_pos = torch.ones((h, w_)).cumsum(1).float().unsqueeze(0)
    if temp_bug_fix:
        div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))
    else:
        div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))
    div = div[:, None, None]
    pe[0::4] = torch.sin(x_pos * div); pe[1::4] = torch.cos(x_pos * div)
    pe[2::4] = torch.sin(y_pos * div); pe[3::4] = torch.cos(y_pos * div)
    return pe


def linear_attention(q, k, v, eps=1e-6):
    """loftr_module/linear_attention.py:20-47.  q [N,L,H,D], k,v [N,S,H,D]."""
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    S = v.size(1)
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S).contiguous()


def encoder_layer(w, p, x, source, nhead):
    """LoFTREncoderLayer.forward, loftr_module/transformer.py:35-58 (all Linear layers bias-free)."""
    bs, d = x.size(0), x.size(2)
    q = F.linear(x, w[p + "q_proj.weight"]).view(bs, -1, nhead, d // nhead)
    k = F.linear(source, w[p + "k_proj.weight"]).view(bs, -1, nhead, d // nhead)
    v = F.linear(source, w[p + "v_proj.weight"]).view(bs, -1, nhead, d // nhead)
    msg = linear_attention(q, k, v)
    msg = F.linear(msg.view(bs, -1, d), w[p + "merge.weight"])
    msg = F.layer_norm(msg, (d,), w[p + "norm1.weight"], w[p + "norm1.bias"])
    msg = F.linear(F.relu(F.linear(torch.cat([x, msg], 2), w[p + "mlp.0.weight"])), w[p + "mlp.2.weight"])
    msg = F.layer_norm(msg, (d,), w[p + "norm2.weight"], w[p + "norm2.bias"])
    return x + msg


def transformer(w, prefix, f0, f1, names, nhead):
    """LocalFeatureTransformer.forward, transformer.py:83-101: cross layers update feat0 FIRST, feat1 sees the new feat0."""
    for i, name in enumerate(names):
        p = f"{prefix}layers.{i}."
        if name == "self":
            f0, f1 = encoder_layer(w, p, f0, f0, nhead), encoder_layer(w, p, f1, f1, nhead)
        else:
            f0 = encoder_layer(w, p, f0, f1, nhead)
            f1 = encoder_layer(w, p, f1, f0, nhead)
    return f0, f1


def coarse_matching(f0, f1, hw0, hw1, thr, border, temperature=0.1):
    """CoarseMatching.forward + get_coarse_match (eval), utils/coarse_matching.py:108-119,151-250.  B = 1."""
    f0, f1 = f0 / f0.shape[-1] ** 0.5, f1 / f1.shape[-1] ** 0.5
    sim = torch.einsum("nlc,nsc->nls", f0, f1) / temperature
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    mask = (conf > thr).view(1, hw0[0], hw0[1], hw1[0], hw1[1]).clone()
    b = border
    if b > 0:
        mask[:, :b] = 0; mask[:, :, :b] = 0; mask[:, :, :, :b] = 0; mask[:, :, :, :, :b] = 0
        mask[:, -b:] = 0; mask[:, :, -b:] = 0; mask[:, :, :, -b:] = 0; mask[:, :, :, :, -b:] = 0
    mask = mask.view(1, hw0[0] * hw0[1], hw1[0] * hw1[1])
    mask = mask * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j = mask.max(dim=2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    return conf, i_ids, j_ids, conf[b_ids, i_ids, j_ids]


def fine_preprocess(w, ff0, ff1, fc0, fc1, i_ids, j_ids, stride, W=5):
    """FinePreprocess.forward, loftr_module/fine_preprocess.py:29-59.  B = 1."""
    if len(i_ids) == 0:
        return torch.empty(0, W * W, ff0.shape[1]), torch.empty(0, W * W, ff0.shape[1])

    def unfold(f, ids):
        u = F.unfold(f, kernel_size=(W, W), stride=stride, padding=W // 2)        # [1, C*WW, L]
        u = u.view(1, f.shape[1], W * W, -1).permute(0, 3, 2, 1)                  # n l ww c
        return u[0, ids]

    u0, u1 = unfold(ff0, i_ids), unfold(ff1, j_ids)
    cwin = F.linear(torch.cat([fc0[0, i_ids], fc1[0, j_ids]], 0), w["fine_preprocess.down_proj.weight"], w["fine_preprocess.down_proj.bias"])
    cat = torch.cat([torch.cat([u0, u1], 0), cwin[:, None, :].expand(-1, W * W, -1)], -1)
    out = F.linear(cat, w["fine_preprocess.merge_feat.weight"], w["fine_preprocess.merge_feat.bias"])
    return torch.chunk(out, 2, dim=0)


def fine_matching(u0, u1, mk0_c, mk1_c, scale, W=5):
    """FineMatching.forward + get_fine_match, utils/fine_matching.py:18-77."""
    M, WW, C = u0.shape
    if M == 0:
        return mk0_c, mk1_c
    sim = torch.einsum("mc,mrc->mr", u0[:, WW // 2, :], u1)
    heat = torch.softmax(sim / C ** 0.5, dim=1).view(-1, W, W)
    lin = torch.linspace(-1, 1, W)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    coords = torch.stack([(heat * gx).sum((-1, -2)), (heat * gy).sum((-1, -2))], -1)  # dsnt.spatial_expectation2d
    return mk0_c, mk1_c + coords * (W // 2) * scale


def forward(w, image0, image1, thr=0.2, cfg=None):
    """LoFTR.forward, loftr.py:29-75 (B = 1; the two images may differ in size: the backbone then runs per image, :48-56).
    Returns dict with keypoints0/1 [M,2], confidence [M], plus intermediates for staged parity checks."""
    c = {**CFG, **(cfg or {})}
    if image0.shape == image1.shape:
        fc, ff = backbone(w, torch.cat([image0, image1], 0))
        fc0, fc1, ff0, ff1 = fc[:1], fc[1:], ff[:1], ff[1:]
    else:
        (fc0, ff0), (fc1, ff1) = backbone(w, image0), backbone(w, image1)
        fc = None
    hc, wc = fc0.shape[2:]
    hc1, wc1 = fc1.shape[2:]
    t0 = (fc0 + position_encoding(c["d_model"], hc, wc, c["temp_bug_fix"])[None]).flatten(2).transpose(1, 2)
    t1 = (fc1 + position_encoding(c["d_model"], hc1, wc1, c["temp_bug_fix"])[None]).flatten(2).transpose(1, 2)
    t0, t1 = transformer(w, "loftr_coarse.", t0, t1, c["coarse_layers"], c["nhead"])
    conf, i_ids, j_ids, mconf = coarse_matching(t0, t1, (hc, wc), (hc1, wc1), thr, c["border_rm"], c["temperature"])
    scale_c = image0.shape[2] / hc
    mk0_c = torch.stack([i_ids % wc, i_ids // wc], 1) * scale_c
    mk1_c = torch.stack([j_ids % wc1, j_ids // wc1], 1) * scale_c
    stride = ff0.shape[2] // hc
    u0, u1 = fine_preprocess(w, ff0, ff1, t0, t1, i_ids, j_ids, stride, c["fine_window"])
    if u0.size(0) != 0:
        u0, u1 = transformer(w, "loftr_fine.", u0, u1, c["fine_layers"], c["nhead"])
    mk0, mk1 = fine_matching(u0, u1, mk0_c, mk1_c, image0.shape[2] / ff0.shape[2], c["fine_window"])
    return {"keypoints0": mk0, "keypoints1": mk1, "confidence": mconf, "i_ids": i_ids, "j_ids": j_ids,
            "feat_c0": t0, "feat_c1": t1, "feat_f0": ff0, "feat_f1": ff1, "backbone_c": fc, "conf_matrix": conf}


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
    """Deterministic random LoFTR parameters (no standard checkpoint exists offline): see utils/synth_weights.py."""
    return _synth_weights().loftr_random_weights(seed)
