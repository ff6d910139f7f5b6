"""Host-side mirrors of the hloc drivers -- CPU tests."""
from types import SimpleNamespace

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_preprocess_is_bit_identical_to_reference(golden):
    """extract() pre-processing (gray, INTER_AREA force-resize, /255, dfactor alignment) == the tensors the
    reference's own calls produced (stored as the golden SuperPoint inputs)."""
    from imcui_b200.hloc import extract_features as ef
    from imcui_b200.hloc.configs import confs_dict
    g = golden("sp_real")
    pre = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "force_resize": False, "width": 320, "height": 240,
           **confs_dict["extractors"]["superpoint_max"]["preprocessing"]}
    for i, name in enumerate(["02928139_3448003521", "17295357_9106075285"]):
        rgb = np.load(GOLDEN / "data" / f"{name}.npz")["rgb"]
        d = ef.preprocess(cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), SimpleNamespace(**pre), "cpu")
        assert torch.equal(d["image"][0], torch.from_numpy(g["images"][i]))
        assert tuple(d["size"]) == (640, 480) and tuple(d["original_size"]) == rgb.shape[:2][::-1]


def test_resize_image_switches_to_linear_when_upsampling():
    from imcui_b200.hloc.extract_features import resize_image
    img = np.random.RandomState(0).rand(10, 12).astype(np.float32)
    up = resize_image(img, (24, 20), "cv2_area")
    assert np.array_equal(up, cv2.resize(img, (24, 20), interpolation=cv2.INTER_LINEAR))


def test_match_images_postprocessing_with_a_fake_matcher():
    """match_images: valid mask, gather and the (k + 0.5) * s - 0.5 rescale (match_features.py:244-257)."""
    from imcui_b200.hloc import match_features as mf

    class Fake(torch.nn.Module):
        def forward(self, data):
            n = data["keypoints0"].shape[1]
            m = torch.full((1, n), -1, dtype=torch.long)
            m[0, 0], m[0, 2] = 1, 0
            return {"matches0": m, "matching_scores0": torch.linspace(0, 1, n)[None], "stop": 3}

    def feat(n, orig, size):
        return {"keypoints": [torch.arange(2 * n, dtype=torch.float32).view(n, 2)], "scores": (torch.ones(n),),
                "descriptors": [torch.zeros(4, n)], "image": torch.zeros(1, 1, 8, 8), "image_orig": np.zeros((2, 2, 3)),
                "original_size": np.array(orig), "size": np.array(size)}
    out = mf.match_images(Fake(), feat(3, (1280, 960), (640, 480)), feat(2, (640, 480), (640, 480)))
    assert out["mkeypoints0"].tolist() == [[0, 1], [4, 5]] and out["mkeypoints1"].tolist() == [[2, 3], [0, 1]]
    np.testing.assert_allclose(out["mkeypoints0_orig"], (np.array([[0, 1], [4, 5]]) + 0.5) * 2 - 0.5)
    np.testing.assert_allclose(out["mkeypoints1_orig"], [[2, 3], [0, 1]])
    assert out["mconf"].shape == (2,)


def test_match_dense_postprocessing_with_a_fake_matcher():
    """match_dense.match_images (reference :577-686): force-resize to 640x480, dfactor alignment, the swap-free output
    keys, 'scores' adopted as mconf and the (k + 0.5) * s - 0.5 rescale with s = original / network size."""
    from imcui_b200.hloc import match_dense as md
    from imcui_b200.hloc.configs import confs_dict
    seen = {}

    class Fake(torch.nn.Module):
        def forward(self, data):
            seen["shape"] = (tuple(data["image0"].shape), tuple(data["image1"].shape))
            return {"keypoints0": torch.tensor([[0.0, 0.0], [8.0, 16.0]]), "keypoints1": torch.tensor([[1.0, 2.0], [9.5, 18.25]]),
                    "scores": torch.tensor([0.9, 0.4])}
    rgb0 = np.random.RandomState(0).randint(0, 255, (960, 1280, 3), dtype=np.uint8)
    rgb1 = np.random.RandomState(1).randint(0, 255, (480, 640, 3), dtype=np.uint8)
    out = md.match_images(Fake(), rgb0, rgb1, confs_dict["matchers"]["loftr"]["preprocessing"], device="cpu")
    assert seen["shape"] == ((1, 1, 480, 640), (1, 1, 480, 640))
    assert out["scale0"].tolist() == [2.0, 2.0] and out["scale1"].tolist() == [1.0, 1.0]
    np.testing.assert_allclose(out["mkeypoints0_orig"], (np.array([[0, 0], [8, 16]]) + 0.5) * 2 - 0.5)
    np.testing.assert_allclose(out["mkeypoints1_orig"], [[1, 2], [9.5, 18.25]])
    assert np.array_equal(out["keypoints0"], out["mkeypoints0"]) and out["mconf"].tolist() == pytest.approx([0.9, 0.4])
    assert out["new_size0"].tolist() == [640, 480] and out["original_size0"].tolist() == [1280, 960]
    # gray conversion + area resize + /255 of image 1 (no resize needed) is the plain cv2 result
    assert np.array_equal(out["image1"], (cv2.cvtColor(rgb1, cv2.COLOR_RGB2GRAY).astype(np.float32) / 255.0))


def test_loftr_weight_packing_layout():
    """BN folding, 196 -> 256 zero padding and the three-plane bf16 split of the LoFTR backbone convolutions."""
    from imcui_b200 import ops
    from oracle import loftr as ol
    sd = ol.random_weights(0)
    pk = ops.loftr_pack_weights(sd)
    c = pk["convs"]["l2.0"]
    assert (c["cin"], c["cout"], c["ksize"], c["stride"]) == (128, 256, 3, 2) and tuple(c["w"].shape) == (3, 9, 256, 128)
    g = sd["backbone.layer2.0.bn1.weight"] / torch.sqrt(sd["backbone.layer2.0.bn1.running_var"] + 1e-5)
    w = (sd["backbone.layer2.0.conv1.weight"] * g[:, None, None, None]).permute(2, 3, 0, 1).reshape(9, 196, 128)
    rec = c["w"].float().sum(0)
    assert float((rec[:, :196] - w).abs().max()) < 1e-6 * float(w.abs().max()) + 1e-7 and float(rec[:, 196:].abs().max()) == 0.0
    assert float(c["b"][196:].abs().max()) == 0.0 and pk["convs"]["l3_out"]["b"] is None
    assert tuple(pk["conv1_w"].shape) == (49, 128) and tuple(pk["c0.qkv_w"].shape) == (768, 256) and tuple(pk["f1.mlp0_w"].shape) == (256, 256)
    pe = ops.loftr_position_encoding(256, 30, 40)
    assert torch.allclose(pe, ol.position_encoding(256, 30, 40).permute(1, 2, 0).reshape(1200, 256))


def test_conf_registry_equals_reference():
    """Every conf name of the package has the reference's output / model / preprocessing values (golden: the
    reference's imcui/hloc/configs entries dumped by tools/make_golden.py confs)."""
    import json
    from imcui_b200.hloc.configs import confs_dict
    ref = json.loads((GOLDEN / "confs.json").read_text())
    for kind in ("extractors", "matchers"):
        assert set(confs_dict[kind]) == set(ref[kind]), kind   # no invented names, none missing from the dump
        for name, c in confs_dict[kind].items():
            for k in ("output", "model", "preprocessing", "max_error", "cell_size"):
                assert c.get(k) == ref[kind][name].get(k), (kind, name, k)


def test_plugin_classes_keep_the_reference_contract():
    """Class names, default_conf values and required_inputs of the plugins == the reference's (golden: read from the
    reference sources with `ast` by tools/make_golden.py plugins).  Extra keys are allowed only for engine switches."""
    import json
    from imcui_b200.hloc import extractors, matchers
    from imcui_b200.hloc.utils.base_model import dynamic_load
    ref = json.loads((GOLDEN / "plugins.json").read_text())
    engine_keys = {"tensor_cores", "state_dict", "max_matches", "pruning_min_kpts", "n_layers"}  # n_layers: lightglue.py default
    for key, r in ref.items():
        kind, mod = key.split("/")
        cls = dynamic_load({"extractors": extractors, "matchers": matchers}[kind], mod)
        assert cls.__name__ == r["class"]
        assert list(cls.required_inputs) == r["required_inputs"], key
        for k, v in r["default_conf"].items():
            assert cls.default_conf[k] == v, (key, k, cls.default_conf.get(k), v)
        assert set(cls.default_conf) - set(r["default_conf"]) <= engine_keys, (key, set(cls.default_conf) - set(r["default_conf"]))


def test_api_pipeline_builds_from_the_registry_and_refuses_cpu():
    """ImageMatchingAPI mirror: zoo entry -> conf dicts -> plugins through dynamic_load; the product path has no CPU
    fallback, so a forward on a machine without CUDA raises instead of silently computing elsewhere."""
    from imcui_b200.api import ImageMatchingAPI
    from imcui_b200.ui.utils import get_matcher_zoo
    zoo = get_matcher_zoo({"superpoint+mnn": {"matcher": "NN-mutual", "feature": "superpoint_max", "dense": False, "standalone": False}})
    conf = zoo["superpoint+mnn"]
    assert conf["feature"]["model"]["name"] == "superpoint" and conf["matcher"]["model"]["name"] == "nearest_neighbor"
    api = ImageMatchingAPI(conf={**conf, "ransac": {"enable": False}}, device="cpu", max_keypoints=512)
    assert type(api.extractor).__name__ == "SuperPoint" and type(api.matcher).__name__ == "NearestNeighbor"
    assert api.extractor.conf["max_keypoints"] == 512 and api.extractor.conf["keypoint_threshold"] == 0.015   # core.py:78-96
    rgb = np.load(GOLDEN / "data" / "02928139_3448003521.npz")["rgb"]
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="CUDA|cuda"):
            api(rgb, rgb)
