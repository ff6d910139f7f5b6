"""Deterministic synthetic model parameters for the two networks whose checkpoints are not available offline (LoFTR,
ALIKED) -- synthetic-data utilities like utils/synth.py: used by bench tooling, by tools/make_golden.py (which loads
the very same tensors into the unmodified reference modules) and, through oracle/, by the tests.  Pure torch, generated
key by key from a CPU torch.Generator so that every machine reproduces them bit for bit."""
import math

import torch


def loftr_random_weights(seed=0):
    """Deterministic random LoFTR parameters (no standard checkpoint exists offline).  Generated key by key from a CPU
    torch.Generator so that the GPU box reproduces them without shipping a 46 MB file; tools/make_golden.py loads the
    very same dict into the reference module.  BatchNorm statistics / affine terms are randomised (not the identity
    defaults) so that BN folding is exercised."""
    g = torch.Generator().manual_seed(1234 + seed)
    w = {}

    def conv(name, co, ci, k):
        w[name] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / (k * k * co))

    def bn(p, c):
        w[p + "weight"] = 0.9 + 0.2 * torch.rand(c, generator=g)
        w[p + "bias"] = 0.02 * torch.randn(c, generator=g)
        w[p + "running_mean"] = 0.02 * torch.randn(c, generator=g)
        w[p + "running_var"] = 0.9 + 0.2 * torch.rand(c, generator=g)

    def lin(name, o, i, bias=False):
        w[name + ".weight"] = torch.randn(o, i, generator=g) * math.sqrt(2.0 / (o + i))
        if bias:
            w[name + ".bias"] = 0.05 * torch.randn(o, generator=g)

    conv("backbone.conv1.weight", 128, 1, 7); bn("backbone.bn1.", 128)
    dims = [(128, 128), (128, 196), (196, 256)]
    for li, (ci, co) in enumerate(dims, 1):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}."
            cin = ci if bi == 0 else co
            conv(p + "conv1.weight", co, cin, 3); conv(p + "conv2.weight", co, co, 3)
            bn(p + "bn1.", co); bn(p + "bn2.", co)
            if bi == 0 and li > 1:
                conv(p + "downsample.0.weight", co, cin, 1); bn(p + "downsample.1.", co)
    conv("backbone.layer3_outconv.weight", 256, 256, 1)
    conv("backbone.layer2_outconv.weight", 256, 196, 1)
    conv("backbone.layer2_outconv2.0.weight", 256, 256, 3); bn("backbone.layer2_outconv2.1.", 256)
    conv("backbone.layer2_outconv2.3.weight", 196, 256, 3)
    conv("backbone.layer1_outconv.weight", 196, 128, 1)
    conv("backbone.layer1_outconv2.0.weight", 196, 196, 3); bn("backbone.layer1_outconv2.1.", 196)
    conv("backbone.layer1_outconv2.3.weight", 128, 196, 3)
    for prefix, n, d in (("loftr_coarse.", 8, 256), ("loftr_fine.", 2, 128)):
        for i in range(n):
            p = f"{prefix}layers.{i}."
            for nm in ("q_proj", "k_proj", "v_proj", "merge"):
                lin(p + nm, d, d)
            lin(p + "mlp.0", 2 * d, 2 * d); lin(p + "mlp.2", d, 2 * d)
            for nm in ("norm1", "norm2"):
                w[p + nm + ".weight"] = 1.0 + 0.05 * torch.randn(d, generator=g)
                w[p + nm + ".bias"] = 0.02 * torch.randn(d, generator=g)
    lin("fine_preprocess.down_proj", 128, 256, bias=True)
    lin("fine_preprocess.merge_feat", 128, 256, bias=True)
    return w


def aliked_random_weights(seed=0):
    """Deterministic random aliked-n16 parameters (reference key names), generated key by key from a CPU Generator."""
    g = torch.Generator().manual_seed(4321 + seed)
    w = {}
    c1, c2, c3, c4 = 16, 32, 64, 128   # aliked-n16 (aliked.py:627)
    dim, K, M = 128, 3, 16

    def conv(name, co, ci, k, bias=False, gain=1.0):
        w[name + ".weight"] = torch.randn(co, ci, k, k, generator=g) * (gain * math.sqrt(1.0 / (k * k * ci)))
        if bias:
            w[name + ".bias"] = 0.1 * torch.randn(co, generator=g)

    def bn(p, c):
        w[p + "weight"] = 0.9 + 0.2 * torch.rand(c, generator=g)
        w[p + "bias"] = 0.05 * torch.randn(c, generator=g)
        w[p + "running_mean"] = 0.05 * torch.randn(c, generator=g)
        w[p + "running_var"] = 0.9 + 0.2 * torch.rand(c, generator=g)
        w[p + "num_batches_tracked"] = torch.tensor(0)

    conv("block1.conv1", c1, 3, 3, gain=3.0); bn("block1.bn1.", c1); conv("block1.conv2", c1, c1, 3); bn("block1.bn2.", c1)
    for name, ci, co, dcn in (("block2", c1, c2, False), ("block3", c2, c3, True), ("block4", c3, c4, True)):
        for j, cin in ((1, ci), (2, co)):
            if dcn:
                conv(f"{name}.conv{j}.offset_conv", 18, cin, 3, bias=True, gain=2.0)
                conv(f"{name}.conv{j}.regular_conv", co, cin, 3)
            else:
                conv(f"{name}.conv{j}", co, cin, 3)
            bn(f"{name}.bn{j}.", co)
        conv(f"{name}.downsample", co, ci, 1, bias=True)
    conv("conv1", dim // 4, c1, 1); conv("conv2", dim // 4, c2, 1); conv("conv3", dim // 4, c3, 1); conv("conv4", dim // 4, dim, 1)
    conv("score_head.0", 8, dim, 1); conv("score_head.2", 4, 8, 3); conv("score_head.4", 4, 4, 3)
    conv("score_head.6", 1, 4, 3, gain=-1.0)   # sign chosen so that the score map is sparse (mean ~0.09, a few hundred maxima > 0.2)
    conv("desc_head.offset_conv.0", 2 * M, dim, K, bias=True, gain=6.0); conv("desc_head.offset_conv.2", 2 * M, 2 * M, 1, bias=True, gain=4.0)
    conv("desc_head.sf_conv", dim, dim, 1, gain=3.0)
    w["desc_head.agg_weights"] = torch.rand(M, dim, dim, generator=g)
    return w
