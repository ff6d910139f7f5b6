"""Oracle for SuperGlue.  Restates third_party/SuperGluePretrainedNetwork/models/
superglue.py:50-283 as plain functions over a flat weight dict (BatchNorm in eval mode).  Test infrastructure only."""
import torch
import torch.nn.functional as F


def _mlp(w, prefix, x, n_layers, bn_last=False):
    """MLP(): Conv1d(k=1) [+ BatchNorm1d + ReLU] ... last layer plain (superglue.py:50-62).  x [1,C,N].
    nn.Sequential indices: conv at 3*l, bn at 3*l+1."""
    for l in range(n_layers):
        x = F.conv1d(x, w[f"{prefix}.{3 * l}.weight"], w[f"{prefix}.{3 * l}.bias"])
        if l < n_layers - 1:
            p = f"{prefix}.{3 * l + 1}."
            x = F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"], False, 0.0, 1e-5)
            x = F.relu(x)
    return x


def normalize_keypoints(kpts, image_shape):
    """superglue.py:65-72."""
    _, _, height, width = image_shape
    one = kpts.new_tensor(1)
    size = torch.stack([one * width, one * height])[None]
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center[:, None, :]) / scaling[:, None, :]


def attentional_propagation(w, i, x, source, heads=4):
    """AttentionalPropagation + MultiHeadedAttention (superglue.py:87-121)."""
    p = f"gnn.layers.{i}."
    b, d_model, _ = x.shape
    dim = d_model // heads
    q, k, v = [F.conv1d(t, w[p + f"attn.proj.{j}.weight"], w[p + f"attn.proj.{j}.bias"]).view(b, dim, heads, -1)
               for j, t in enumerate((x, source, source))]
    scores = torch.einsum("bdhn,bdhm->bhnm", q, k) / dim ** 0.5
    prob = F.softmax(scores, dim=-1)
    msg = torch.einsum("bhnm,bdhm->bdhn", prob, v)
    msg = F.conv1d(msg.contiguous().view(b, dim * heads, -1), w[p + "attn.merge.weight"], w[p + "attn.merge.bias"])
    return _mlp(w, p + "mlp", torch.cat([x, msg], dim=1), 2)


def log_sinkhorn_iterations(Z, log_mu, log_nu, iters):
    """superglue.py:143-149."""
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1)


def log_optimal_transport(scores, alpha, iters):
    """superglue.py:152-172."""
    b, m, n = scores.shape
    one = scores.new_tensor(1)
    ms, ns = (m * one).to(scores), (n * one).to(scores)
    bins0, bins1, alpha = alpha.expand(b, m, 1), alpha.expand(b, 1, n), alpha.expand(b, 1, 1)
    couplings = torch.cat([torch.cat([scores, bins0], -1), torch.cat([bins1, alpha], -1)], 1)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])
    log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])
    Z = log_sinkhorn_iterations(couplings, log_mu[None].expand(b, -1), log_nu[None].expand(b, -1), iters)
    return Z - norm


def forward(w, data, sinkhorn_iterations=100, match_threshold=0.2, layer_names=("self", "cross") * 9):
    """SuperGlue.forward (superglue.py:228-283).  data: keypoints0/1 [1,N,2], scores0/1 [1,N], descriptors0/1
    [1,256,N], image0/1 (shape only)."""
    desc0, desc1 = data["descriptors0"], data["descriptors1"]
    kpts0, kpts1 = data["keypoints0"], data["keypoints1"]
    if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:
        s0, s1 = kpts0.shape[:-1], kpts1.shape[:-1]
        return {"matches0": kpts0.new_full(s0, -1, dtype=torch.int), "matches1": kpts1.new_full(s1, -1, dtype=torch.int),
                "matching_scores0": kpts0.new_zeros(s0), "matching_scores1": kpts1.new_zeros(s1)}
    kpts0 = normalize_keypoints(kpts0, data["image0"].shape)
    kpts1 = normalize_keypoints(kpts1, data["image1"].shape)

    def kenc(k, s):
        return _mlp(w, "kenc.encoder", torch.cat([k.transpose(1, 2), s.unsqueeze(1)], dim=1), 5)

    desc0 = desc0 + kenc(kpts0, data["scores0"])
    desc1 = desc1 + kenc(kpts1, data["scores1"])
    for i, name in enumerate(layer_names):
        src0, src1 = (desc1, desc0) if name == "cross" else (desc0, desc1)
        d0, d1 = attentional_propagation(w, i, desc0, src0), attentional_propagation(w, i, desc1, src1)
        desc0, desc1 = desc0 + d0, desc1 + d1
    m0 = F.conv1d(desc0, w["final_proj.weight"], w["final_proj.bias"])
    m1 = F.conv1d(desc1, w["final_proj.weight"], w["final_proj.bias"])
    scores = torch.einsum("bdn,bdm->bnm", m0, m1) / 256 ** 0.5
    scores = log_optimal_transport(scores, w["bin_score"], sinkhorn_iterations)
    max0, max1 = scores[:, :-1, :-1].max(2), scores[:, :-1, :-1].max(1)
    i0, i1 = max0.indices, max1.indices
    ar0, ar1 = torch.arange(i0.shape[1])[None], torch.arange(i1.shape[1])[None]
    mutual0, mutual1 = ar0 == i1.gather(1, i0), ar1 == i0.gather(1, i1)
    zero = scores.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values.exp(), zero)
    ms1 = torch.where(mutual1, ms0.gather(1, i1), zero)
    valid0 = mutual0 & (ms0 > match_threshold)
    valid1 = mutual1 & valid0.gather(1, i1)
    return {"matches0": torch.where(valid0, i0, i0.new_tensor(-1)), "matches1": torch.where(valid1, i1, i1.new_tensor(-1)),
            "matching_scores0": ms0, "matching_scores1": ms1}
