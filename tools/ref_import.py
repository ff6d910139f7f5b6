"""Import the UNMODIFIED reference modules from /root/reference (this container only).

Test/fixture infrastructure: used by tools/make_golden.py and tools/fetch_weights.py to
mint golden vectors and convert weights.  Nothing here travels to the GPU box and
nothing in the product path imports it.  Recipe follows SURVEY.md section 8(c).
"""
import importlib.util
import os
import sys
import types
from pathlib import Path

REF = Path(os.environ.get("IMW_REFERENCE", "/root/reference"))
TP = REF / "imcui" / "third_party"
SP_WEIGHTS = TP / "SE2LoFTR/third_party/SuperGluePretrainedNetwork/models/weights/superpoint_v1.pth"
SG_WEIGHTS = TP / "SE2LoFTR/third_party/SuperGluePretrainedNetwork/models/weights"
LG_GIM_CKPT = TP / "gim/weights/gim_lightglue_100h.ckpt"


def available() -> bool:
    return (REF / "imcui").is_dir()


def _load_file(name, path):
    if name in sys.modules:  # one module object per process (class attributes are patched by tools)
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def superpoint_module():
    """third_party/SuperGluePretrainedNetwork/models/superpoint.py (the one hloc uses)."""
    return _load_file("_ref_superpoint", TP / "SuperGluePretrainedNetwork/models/superpoint.py")


def superglue_module():
    return _load_file("_ref_superglue", TP / "SuperGluePretrainedNetwork/models/superglue.py")


def lightglue_module():
    """third_party/LightGlue/lightglue/lightglue.py loaded by file (package __init__ needs kornia)."""
    return _load_file("_ref_lightglue", TP / "LightGlue/lightglue/lightglue.py")


def make_superpoint(conf):
    import contextlib, io
    sp = superpoint_module()
    with contextlib.redirect_stdout(io.StringIO()):
        net = sp.SuperPoint({**conf, "weights_path": str(SP_WEIGHTS)})
    return net.eval()


def lightglue_state_dict():
    import torch
    ck = torch.load(str(LG_GIM_CKPT), map_location="cpu", weights_only=False)
    sd = ck["state_dict"]
    return {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}


def make_lightglue(**conf):
    lg = lightglue_module()
    net = lg.LightGlue(features=None, input_dim=256, weights=None, **conf)
    missing, unexpected = net.load_state_dict(lightglue_state_dict(), strict=False)
    assert not unexpected, unexpected
    assert all("confidence_thresholds" in m for m in missing), missing
    return net.eval()


def lightglue_proj_state(dim=128, seed=7):
    """gim LightGlue weights + a deterministic Linear(dim -> 256) input_proj: an orthonormal lift Q^T (so that
    256-d SuperPoint descriptors compressed with Q keep matching) and a small random bias.  Returns (state, Q [dim,256])."""
    import torch
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(256, dim, generator=g))   # [256, dim], orthonormal columns
    sd = dict(lightglue_state_dict())
    sd["input_proj.weight"] = q.contiguous()
    sd["input_proj.bias"] = 0.01 * torch.randn(256, generator=g)
    return sd, q.t().contiguous()


def make_lightglue_proj(dim=128, **conf):
    lg = lightglue_module()
    net = lg.LightGlue(features=None, input_dim=dim, weights=None, **conf)
    sd, q = lightglue_proj_state(dim)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("confidence_thresholds" in m for m in missing), (missing, unexpected)
    return net.eval(), sd, q


def hloc_matchers():
    """imcui.hloc.matchers.{nearest_neighbor,dual_softmax} import as-is (run from a scratch cwd:
    importing imcui.hloc writes log.txt)."""
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from imcui.hloc.matchers import nearest_neighbor, dual_softmax
    return nearest_neighbor, dual_softmax


def make_superglue(conf):
    import contextlib, io
    sg = superglue_module()
    w = conf.get("weights", "outdoor")
    with contextlib.redirect_stdout(io.StringIO()):
        net = sg.SuperGlue({**conf, "weights_path": str(SG_WEIGHTS / f"superglue_{w}.pth")})
    return net.eval()


def loftr_module():
    """third_party/SE2LoFTR/src/loftr (the in-tree copy of zju3dv LoFTR that kornia.feature.LoFTR ports) with the
    import stubs of SURVEY.md 8(c): e2cnn, yacs CfgNode, kornia dsnt / create_meshgrid."""
    import torch
    if "_ref_loftr_pkg" in sys.modules:
        return sys.modules["_ref_loftr_pkg"]

    class CfgNode(dict):
        def __getattr__(self, k):
            return self[k]

        def __setattr__(self, k, v):
            self[k] = v

    def spatial_expectation2d(heat, normalized_coordinates=True):
        b, c, h, w = heat.shape
        ys, xs = torch.linspace(-1, 1, h), torch.linspace(-1, 1, w)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        ex = (heat * gx).sum((-1, -2))
        ey = (heat * gy).sum((-1, -2))
        return torch.stack([ex, ey], -1)

    def create_meshgrid(h, w, normalized_coordinates=True, device=None):
        ys, xs = torch.linspace(-1, 1, h), torch.linspace(-1, 1, w)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1)[None]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mod("yacs"); mod("yacs.config", CfgNode=CfgNode)
    for n in ("e2cnn", "e2cnn.nn", "e2cnn.gspaces"):
        mod(n)
    dsnt = mod("kornia.geometry.subpix.dsnt", spatial_expectation2d=spatial_expectation2d)
    mod("kornia"); mod("kornia.geometry", create_meshgrid=create_meshgrid); mod("kornia.geometry.subpix", dsnt=dsnt)
    mod("kornia.utils", create_meshgrid=create_meshgrid)
    src = TP / "SE2LoFTR" / "src"
    pkg = types.ModuleType("_ref_loftr_src"); pkg.__path__ = [str(src)]
    sys.modules["_ref_loftr_src"] = pkg
    import importlib
    # the package __init__ of src/loftr imports e2cnn variants; load the plain LoFTR modules directly
    lp = types.ModuleType("_ref_loftr_src.loftr"); lp.__path__ = [str(src / "loftr")]
    sys.modules["_ref_loftr_src.loftr"] = lp
    bb = types.ModuleType("_ref_loftr_src.loftr.backbone"); bb.__path__ = [str(src / "loftr" / "backbone")]
    sys.modules["_ref_loftr_src.loftr.backbone"] = bb
    resnet = importlib.import_module("_ref_loftr_src.loftr.backbone.resnet_fpn")
    bb.build_backbone = lambda config: resnet.ResNetFPN_8_2(config["resnetfpn"])
    loftr = importlib.import_module("_ref_loftr_src.loftr.loftr")
    cfg = importlib.import_module("_ref_loftr_src.loftr.utils.cvpr_ds_config")
    out = types.SimpleNamespace(LoFTR=loftr.LoFTR, default_cfg=cfg.default_cfg)
    sys.modules["_ref_loftr_pkg"] = out
    return out


def make_loftr(seed=0, thr=0.2):
    import copy, torch
    m = loftr_module()
    cfg = copy.deepcopy(dict(m.default_cfg))
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    cfg["match_coarse"]["thr"] = thr
    torch.manual_seed(seed)
    net = m.LoFTR(cfg).eval()
    return net


def aliked_module():
    """third_party/LightGlue/lightglue/aliked.py under a stand-in `lightglue` package (its __init__ pulls every model)
    with kornia.color.grayscale_to_rgb stubbed (kornia is not installed; the function is a channel repeat)."""
    import torch
    if "_ref_lg_pkg.aliked" in sys.modules:
        return sys.modules["_ref_lg_pkg.aliked"]
    if "kornia" not in sys.modules or not hasattr(sys.modules["kornia"], "color"):
        k = sys.modules.get("kornia") or types.ModuleType("kornia")
        k.__path__ = []
        col = types.ModuleType("kornia.color")
        col.grayscale_to_rgb = lambda x: x.repeat(1, 3, 1, 1) if x.dim() == 4 else x.repeat(3, 1, 1)
        k.color = col
        sys.modules["kornia"] = k
        sys.modules["kornia.color"] = col
    pkg = types.ModuleType("_ref_lg_pkg"); pkg.__path__ = [str(TP / "LightGlue/lightglue")]
    sys.modules["_ref_lg_pkg"] = pkg
    import importlib
    return importlib.import_module("_ref_lg_pkg.aliked")


def make_aliked(state_dict, **conf):
    """The unmodified ALIKED module with `state_dict` loaded in place of the GitHub download (aliked.py:692-695)."""
    import torch
    m = aliked_module()
    orig = torch.hub.load_state_dict_from_url
    torch.hub.load_state_dict_from_url = lambda *a, **k: state_dict
    try:
        net = m.ALIKED(**conf)
    finally:
        torch.hub.load_state_dict_from_url = orig
    return net.eval()
