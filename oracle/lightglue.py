"""Oracle for LightGlue.

Restates third_party/LightGlue/lightglue/lightglue.py:31-661 (the fp32 einsum/SDPA CPU path, B=1)
as plain functions over a flat weight dict.  Test infrastructure only (see oracle/__init__.py).

`pruning_min_kpts` reproduces the reference's device-dependent switch (lightglue.py:339-344,
663-667): -1 is what the reference does on CPU (prune at every layer), 1536 what it does on
CUDA with flash SDPA (the deployment mode), 1024 CUDA without flash.
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CONF = {  # lightglue.py:322-335 merged with hloc/matchers/lightglue.py:15-25,50
    "n_layers": 9,
    "num_heads": 4,
    "descriptor_dim": 256,
    "depth_confidence": 0.95,
    "width_confidence": 0.99,
    "filter_threshold": 0.2,
    "pruning_min_kpts": -1,
}


def normalize_keypoints(kpts, size=None):
    """lightglue.py:31-43.  kpts [1,N,2]."""
    if size is None:
        size = 1 + kpts.max(-2).values - kpts.min(-2).values
    size = size.to(kpts)
    shift = size / 2
    scale = size.max(-1).values / 2
    return (kpts - shift[..., None, :]) / scale[..., None, None]


def posenc(w, kpts):
    """LearnableFourierPositionalEncoding, lightglue.py:68-81 -> [2,1,1,N,64]."""
    proj = F.linear(kpts, w["posenc.Wr.weight"])
    emb = torch.stack([torch.cos(proj), torch.sin(proj)], 0).unsqueeze(-3)
    return emb.repeat_interleave(2, dim=-1)


def rotate_half(x):
    """lightglue.py:58-61."""
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(start_dim=-2)


def rotary(freqs, t):
    """lightglue.py:64-65."""
    return (t * freqs[0]) + (rotate_half(t) * freqs[1])


def _ffn(w, p, x, msg):
    """nn.Sequential(Linear(512,512), LayerNorm(512), GELU, Linear(512,256)), lightglue.py:152-157."""
    h = F.linear(torch.cat([x, msg], -1), w[p + "ffn.0.weight"], w[p + "ffn.0.bias"])
    h = F.layer_norm(h, (h.shape[-1],), w[p + "ffn.1.weight"], w[p + "ffn.1.bias"], 1e-5)
    h = F.gelu(h)
    return F.linear(h, w[p + "ffn.3.weight"], w[p + "ffn.3.bias"])


def self_block(w, i, x, enc, heads=4):
    """SelfBlock.forward, lightglue.py:159-172 (SDPA default scale 1/sqrt(64))."""
    p = f"transformers.{i}.self_attn."
    qkv = F.linear(x, w[p + "Wqkv.weight"], w[p + "Wqkv.bias"])
    qkv = qkv.unflatten(-1, (heads, -1, 3)).transpose(1, 2)
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
    q, k = rotary(enc, q), rotary(enc, k)
    ctx = F.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous())
    msg = F.linear(ctx.transpose(1, 2).flatten(start_dim=-2), w[p + "out_proj.weight"], w[p + "out_proj.bias"])
    return x + _ffn(w, p, x, msg)


def cross_block(w, i, x0, x1, heads=4):
    """CrossBlock.forward einsum path, lightglue.py:199-230."""
    p = f"transformers.{i}.cross_attn."

    def proj(name, x):
        y = F.linear(x, w[p + name + ".weight"], w[p + name + ".bias"])
        return y.unflatten(-1, (heads, -1)).transpose(1, 2)

    qk0, qk1 = proj("to_qk", x0), proj("to_qk", x1)
    v0, v1 = proj("to_v", x0), proj("to_v", x1)
    scale = (x0.shape[-1] // heads) ** -0.5
    qk0, qk1 = qk0 * scale ** 0.5, qk1 * scale ** 0.5
    sim = torch.einsum("bhid, bhjd -> bhij", qk0, qk1)
    attn01 = F.softmax(sim, dim=-1)
    attn10 = F.softmax(sim.transpose(-2, -1).contiguous(), dim=-1)
    m0 = torch.einsum("bhij, bhjd -> bhid", attn01, v1)
    m1 = torch.einsum("bhji, bhjd -> bhid", attn10.transpose(-2, -1), v0)

    def out(m):
        return F.linear(m.transpose(1, 2).flatten(start_dim=-2), w[p + "to_out.weight"], w[p + "to_out.bias"])

    m0, m1 = out(m0), out(m1)
    return x0 + _ffn(w, p, x0, m0), x1 + _ffn(w, p, x1, m1)


def token_confidence(w, i, x):
    """TokenConfidence, lightglue.py:84-94."""
    return torch.sigmoid(F.linear(x, w[f"token_confidence.{i}.token.0.weight"], w[f"token_confidence.{i}.token.0.bias"])).squeeze(-1)


def matchability(w, i, x):
    """MatchAssignment.get_matchability, lightglue.py:298-299."""
    return torch.sigmoid(F.linear(x, w[f"log_assignment.{i}.matchability.weight"], w[f"log_assignment.{i}.matchability.bias"])).squeeze(-1)


def confidence_threshold(i, n_layers=9):
    """lightglue.py:636-639."""
    return float(min(max(0.8 + 0.1 * math.exp(-4.0 * i / n_layers), 0.0), 1.0))


def log_assignment(w, i, d0, d1):
    """MatchAssignment.forward + sigmoid_log_double_softmax, lightglue.py:265-296."""
    p = f"log_assignment.{i}."
    md0 = F.linear(d0, w[p + "final_proj.weight"], w[p + "final_proj.bias"])
    md1 = F.linear(d1, w[p + "final_proj.weight"], w[p + "final_proj.bias"])
    d = md0.shape[-1]
    md0, md1 = md0 / d ** 0.25, md1 / d ** 0.25
    sim = torch.einsum("bmd,bnd->bmn", md0, md1)
    z0 = F.linear(d0, w[p + "matchability.weight"], w[p + "matchability.bias"])
    z1 = F.linear(d1, w[p + "matchability.weight"], w[p + "matchability.bias"])
    b, m, n = sim.shape
    cert = F.logsigmoid(z0) + F.logsigmoid(z1).transpose(1, 2)
    s0 = F.log_softmax(sim, 2)
    s1 = F.log_softmax(sim.transpose(-1, -2).contiguous(), 2).transpose(-1, -2)
    scores = sim.new_full((b, m + 1, n + 1), 0)
    scores[:, :m, :n] = s0 + s1 + cert
    scores[:, :-1, -1] = F.logsigmoid(-z0.squeeze(-1))
    scores[:, -1, :-1] = F.logsigmoid(-z1.squeeze(-1))
    return scores


def filter_matches(scores, th):
    """lightglue.py:302-318."""
    max0, max1 = scores[:, :-1, :-1].max(2), scores[:, :-1, :-1].max(1)
    m0, m1 = max0.indices, max1.indices
    i0 = torch.arange(m0.shape[1])[None]
    i1 = torch.arange(m1.shape[1])[None]
    mutual0 = i0 == m1.gather(1, m0)
    mutual1 = i1 == m0.gather(1, m1)
    max0_exp = max0.values.exp()
    zero = max0_exp.new_tensor(0)
    ms0 = torch.where(mutual0, max0_exp, zero)
    ms1 = torch.where(mutual1, ms0.gather(1, m1), zero)
    valid0 = mutual0 & (ms0 > th)
    valid1 = mutual1 & valid0.gather(1, m1)
    m0 = torch.where(valid0, m0, -1)
    m1 = torch.where(valid1, m1, -1)
    return m0, m1, ms0, ms1


def forward(w, kpts0, desc0, kpts1, desc1, conf=None, trace=None, scale_ori=None):
    """LightGlue._forward for one pair, lightglue.py:488-634.
    kpts* [1,N,2] pixels, desc* [1,N,256].  Returns the reference's dict (tensors with batch dim 1).
    `trace`, if a list, receives per-layer dicts (descriptor states, token confidences)."""
    c = {**DEFAULT_CONF, **(conf or {})}
    L = c["n_layers"]
    b, m, _ = kpts0.shape
    _, n, _ = kpts1.shape
    k0 = normalize_keypoints(kpts0).clone()
    k1 = normalize_keypoints(kpts1).clone() if n > 0 else kpts1
    if m == 0:
        k0 = kpts0
    d0, d1 = desc0.contiguous(), desc1.contiguous()
    in_dim = w["input_proj.weight"].shape[1] if "input_proj.weight" in w else 256   # lightglue.py:392-395,510-511
    assert d0.shape[-1] == in_dim and d1.shape[-1] == in_dim
    if "input_proj.weight" in w:
        d0 = F.linear(d0, w["input_proj.weight"], w["input_proj.bias"])
        d1 = F.linear(d1, w["input_proj.weight"], w["input_proj.bias"])
    if scale_ori is not None:   # add_scale_ori (lightglue.py:500-506): (scales0, oris0, scales1, oris1), each [1,N]
        s0, o0, s1, o1 = scale_ori
        k0, k1 = torch.cat([k0, s0[..., None], o0[..., None]], -1), torch.cat([k1, s1[..., None], o1[..., None]], -1)
    e0, e1 = posenc(w, k0), posenc(w, k1)
    do_stop = c["depth_confidence"] > 0
    do_prune = c["width_confidence"] > 0
    pth = c["pruning_min_kpts"]
    if do_prune:
        ind0, ind1 = torch.arange(m)[None], torch.arange(n)[None]
        prune0, prune1 = torch.ones_like(ind0), torch.ones_like(ind1)
    t0 = t1 = None
    i = 0
    for i in range(L):
        if d0.shape[1] == 0 or d1.shape[1] == 0:
            break
        d0, d1 = self_block(w, i, d0, e0), self_block(w, i, d1, e1)
        d0, d1 = cross_block(w, i, d0, d1)
        if trace is not None:
            trace.append({"layer": i, "desc0": d0.clone(), "desc1": d1.clone()})
        if i == L - 1:
            continue
        if do_stop:
            t0, t1 = token_confidence(w, i, d0), token_confidence(w, i, d1)
            conf_all = torch.cat([t0[..., :m], t1[..., :n]], -1)
            ratio = 1.0 - (conf_all < confidence_threshold(i, L)).float().sum() / (m + n)
            if trace is not None:
                trace[-1].update(token0=t0.clone(), token1=t1.clone(), ratio=float(ratio))
            if ratio > c["depth_confidence"]:
                break
        if do_prune and d0.shape[-2] > pth:
            keep = matchability(w, i, d0) > (1 - c["width_confidence"])
            if t0 is not None:
                keep |= t0 <= confidence_threshold(i, L)
            kp = torch.where(keep)[1]
            ind0, d0, e0 = ind0.index_select(1, kp), d0.index_select(1, kp), e0.index_select(-2, kp)
            prune0[:, ind0] += 1
        if do_prune and d1.shape[-2] > pth:
            keep = matchability(w, i, d1) > (1 - c["width_confidence"])
            if t1 is not None:
                keep |= t1 <= confidence_threshold(i, L)
            kp = torch.where(keep)[1]
            ind1, d1, e1 = ind1.index_select(1, kp), d1.index_select(1, kp), e1.index_select(-2, kp)
            prune1[:, ind1] += 1

    if d0.shape[1] == 0 or d1.shape[1] == 0:  # lightglue.py:573-593
        out = {
            "matches0": torch.full((b, m), -1, dtype=torch.long),
            "matches1": torch.full((b, n), -1, dtype=torch.long),
            "matching_scores0": torch.zeros((b, m)),
            "matching_scores1": torch.zeros((b, n)),
            "stop": i + 1,
            "matches": torch.empty((b, 0, 2), dtype=torch.long),
            "scores": torch.empty((b, 0)),
        }
        out["prune0"] = prune0 if do_prune else torch.ones((b, m)) * L
        out["prune1"] = prune1 if do_prune else torch.ones((b, n)) * L
        return out

    scores = log_assignment(w, i, d0, d1)
    m0, m1, ms0, ms1 = filter_matches(scores, c["filter_threshold"])
    valid = m0[0] > -1
    mi0 = torch.where(valid)[0]
    mi1 = m0[0][valid]
    if do_prune:
        mi0, mi1 = ind0[0, mi0], ind1[0, mi1]
    matches = [torch.stack([mi0, mi1], -1)]
    mscores = [ms0[0][valid]]
    if do_prune:  # lightglue.py:610-619
        m0_ = torch.full((b, m), -1, dtype=m0.dtype)
        m1_ = torch.full((b, n), -1, dtype=m1.dtype)
        m0_[:, ind0] = torch.where(m0 == -1, -1, ind1.gather(1, m0.clamp(min=0)))
        m1_[:, ind1] = torch.where(m1 == -1, -1, ind0.gather(1, m1.clamp(min=0)))
        ms0_, ms1_ = torch.zeros((b, m)), torch.zeros((b, n))
        ms0_[:, ind0] = ms0
        ms1_[:, ind1] = ms1
        m0, m1, ms0, ms1 = m0_, m1_, ms0_, ms1_
    else:
        prune0 = torch.ones_like(ms0) * L
        prune1 = torch.ones_like(ms1) * L
    return {
        "matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1,
        "stop": i + 1, "matches": matches, "scores": mscores, "prune0": prune0, "prune1": prune1,
    }
