"""LoFTR on the GPU vs the goldens minted from the in-tree LoFTR module (deterministic random weights) and, stage by
stage, vs the CPU oracle."""
import importlib.util

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _pair(H, W):
    spec = importlib.util.spec_from_file_location("synth", ROOT / "image-matching-webui_b200/utils/synth.py")
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    a, b, _ = synth.make_pair(0, H, W)
    return torch.from_numpy(a.astype(np.float32) / 255.0), torch.from_numpy(b.astype(np.float32) / 255.0)


@pytest.mark.parametrize("tag", ["s", "m"])
@pytest.mark.parametrize("tc", [0, 1], ids=["linears-fp32", "linears-3xtf32"])
def test_loftr_matches_reference(golden, tag, tc):
    from imcui_b200 import ops
    from oracle import loftr as ol
    dev = torch.device("cuda:0")
    g = golden("loftr")
    H, W = (int(v) for v in g[tag + "/hw"])
    x0, x1 = _pair(H, W)
    sd = ol.random_weights(0)
    wd = ops.loftr_to_device(ops.loftr_pack_weights(sd), dev)
    out = ops.loftr_forward(wd, torch.stack([x0, x1]).to(dev), {"match_threshold": float(g[tag + "/thr"]), "use_tensor_cores": tc}, debug=True)
    torch.cuda.synchronize()
    if tag == "s":  # staged check against the oracle (cheap at this size)
        o = ol.forward(sd, x0[None, None], x1[None, None], thr=float(g[tag + "/thr"]))
        eb = float((out["backbone_c"].cpu() - o["backbone_c"].permute(0, 2, 3, 1).reshape(2, -1, 256)).abs().max())
        ec = float((out["feat_c"].cpu() - torch.cat([o["feat_c0"], o["feat_c1"]])).abs().max())
        print(f"[loftr] {tag} tc={tc}: backbone max err {eb:.2e} (|x| max {float(o['backbone_c'].abs().max()):.2f}), coarse feature max err {ec:.2e}")
        assert eb < 2e-4 and ec < 2e-3
    n = int(out["counts"][0])
    k0, k1, cf = out["keypoints0"][0, :n].cpu().numpy(), out["keypoints1"][0, :n].cpu().numpy(), out["confidence"][0, :n].cpu().numpy()
    ref_set = {tuple(r) for r in g[tag + "/keypoints0"].tolist()}
    got_set = {tuple(r) for r in k0.tolist()}
    print(f"[loftr] {tag} tc={tc}: matches {n} (reference {len(g[tag + '/confidence'])}), common {len(ref_set & got_set)}")
    assert n == len(g[tag + "/confidence"]) and np.array_equal(k0, g[tag + "/keypoints0"])       # same cells, same order
    np.testing.assert_allclose(k1, g[tag + "/keypoints1"], atol=2e-3)                               # sub-pixel refinement
    np.testing.assert_allclose(cf, g[tag + "/confidence"], rtol=2e-3)


def test_loftr_plugin_contract(golden):
    from imcui_b200.hloc import matchers
    from imcui_b200.hloc.utils.base_model import dynamic_load
    from oracle import loftr as ol
    dev = torch.device("cuda:0")
    g = golden("loftr")
    x0, x1 = _pair(240, 320)
    model = dynamic_load(matchers, "loftr")({"state_dict": ol.random_weights(0), "match_threshold": float(g["s/thr"]), "max_keypoints": 2000}).eval().to(dev)
    out = model({"image0": x1[None, None].to(dev), "image1": x0[None, None].to(dev)})   # hloc order: the module sees (image1, image0)
    assert set(out) >= {"keypoints0", "keypoints1", "scores"}
    assert np.array_equal(out["keypoints1"].cpu().numpy(), g["s/keypoints0"]) and out["keypoints0"].shape == out["keypoints1"].shape


def test_match_dense_with_loftr_end_to_end(golden):
    """hloc.match_dense.match_images (row a10) driving the LoFTR plugin (row a11) from uint8 host images."""
    import importlib.util
    from imcui_b200.hloc import match_dense, matchers
    from imcui_b200.hloc.configs import confs_dict
    from imcui_b200.hloc.utils.base_model import dynamic_load
    from oracle import loftr as ol
    spec = importlib.util.spec_from_file_location("synth", ROOT / "image-matching-webui_b200/utils/synth.py")
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    a, b, _ = synth.make_pair(0, 480, 640)
    g = golden("loftr")
    conf = confs_dict["matchers"]["loftr"]
    model = dynamic_load(matchers, "loftr")({**conf["model"], "state_dict": ol.random_weights(0), "match_threshold": float(g["m/thr"]), "max_keypoints": 30}).eval().to("cuda:0")
    out = match_dense.match_images(model, b, a, conf["preprocessing"], device="cuda:0")   # the module sees (image_1, image_0)
    order = np.argsort(-g["m/confidence"], kind="stable")[:30]
    assert np.array_equal(out["mkeypoints1"], g["m/keypoints0"][order])
    np.testing.assert_allclose(out["mkeypoints0"], g["m/keypoints1"][order], atol=2e-3)
    np.testing.assert_allclose(out["mconf"], g["m/confidence"][order], rtol=2e-3)
    assert np.all(np.diff(out["mconf"]) <= 1e-7)          # top-k keeps them sorted by confidence (loftr.py:58-65)
    assert out["scale0"].tolist() == [1.0, 1.0] and np.array_equal(out["mkeypoints0_orig"], out["mkeypoints0"])


@pytest.mark.parametrize("hw", [(200, 264), (136, 328)], ids=lambda v: f"{v[0]}x{v[1]}")
def test_loftr_ragged_sizes_vs_oracle(hw):
    """Sizes whose coarse grid is not a multiple of the 128-token tile (partial GEMM / conv / similarity tiles)."""
    from imcui_b200 import ops
    from oracle import loftr as ol
    dev = torch.device("cuda:0")
    H, W = hw
    x0, x1 = _pair(H, W)
    sd = ol.random_weights(0)
    thr = 1e-5
    o = ol.forward(sd, x0[None, None], x1[None, None], thr=thr)
    out = ops.loftr_forward(ops.loftr_to_device(ops.loftr_pack_weights(sd), dev), torch.stack([x0, x1]).to(dev), {"match_threshold": thr})
    n = int(out["counts"][0])
    print(f"[loftr ragged] {hw}: matches {n} (oracle {len(o['confidence'])})")
    assert n == len(o["confidence"]) and n > 0
    assert np.array_equal(out["keypoints0"][0, :n].cpu().numpy(), o["keypoints0"].numpy())
    np.testing.assert_allclose(out["keypoints1"][0, :n].cpu().numpy(), o["keypoints1"].numpy(), atol=2e-3)
    np.testing.assert_allclose(out["confidence"][0, :n].cpu().numpy(), o["confidence"].numpy(), rtol=2e-3)


@pytest.mark.parametrize("tag", ["d", "e"])
def test_loftr_different_sizes_vs_reference(golden, tag):
    """The two images of a pair keep their own sizes (no force_resize in the `minima_loftr` / `loftr_aachen`-style confs): the
    module runs its backbone per image (loftr.py:48-56).  Library entry (imw_loftr_forward_hw) and plugin vs the goldens of the
    unmodified module; a batch of two such pairs == the pairs one by one."""
    from imcui_b200 import ops
    from imcui_b200.hloc import matchers
    from imcui_b200.hloc.utils.base_model import dynamic_load
    from oracle import loftr as ol
    dev = torch.device("cuda:0")
    g = golden("loftr_hw")
    thr = float(g[tag + "/thr"])
    x0 = torch.from_numpy(g[tag + "/image0"].astype(np.float32) / 255.0).to(dev)
    x1 = torch.from_numpy(g[tag + "/image1"].astype(np.float32) / 255.0).to(dev)
    sd = ol.random_weights(0)
    wd = ops.loftr_to_device(ops.loftr_pack_weights(sd), dev)
    out = ops.loftr_forward(wd, torch.stack([x0, x0]), {"match_threshold": thr}, images1=torch.stack([x1, x1]))
    for p in range(2):
        n = int(out["counts"][p])
        print(f"[loftr hw] {tag} pair {p}: {tuple(x0.shape)} x {tuple(x1.shape)} -> {n} matches (reference {len(g[tag + '/confidence'])})")
        assert n == len(g[tag + "/confidence"]) and n > 0
        assert np.array_equal(out["keypoints0"][p, :n].cpu().numpy(), g[tag + "/keypoints0"])
        np.testing.assert_allclose(out["keypoints1"][p, :n].cpu().numpy(), g[tag + "/keypoints1"], atol=2e-3)
        np.testing.assert_allclose(out["confidence"][p, :n].cpu().numpy(), g[tag + "/confidence"], rtol=2e-3)
    # plugin: hloc hands the module (image1, image0) (hloc/matchers/loftr.py:43-51)
    model = dynamic_load(matchers, "loftr")({"state_dict": sd, "match_threshold": thr, "max_keypoints": None}).eval().to(dev)
    pred = model({"image0": x1[None, None], "image1": x0[None, None]})
    assert np.array_equal(pred["keypoints1"].cpu().numpy(), g[tag + "/keypoints0"])
    np.testing.assert_allclose(pred["keypoints0"].cpu().numpy(), g[tag + "/keypoints1"], atol=2e-3)
