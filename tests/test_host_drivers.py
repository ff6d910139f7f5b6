"""Host-side mirrors of the hloc drivers -- CPU tests."""
from types import SimpleNamespace

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_preprocess_plan_equals_the_reference_size_logic():
    """Output geometry of imw_preprocess (host arithmetic of the C ABI, no GPU needed) == the size logic of
    extract_features.py:120-156 as restated by the oracle, incl. Python's round-half-even and the dfactor floor."""
    from oracle import preprocess as op
    from imcui_b200 import ops
    for (h, w, c) in ((1063, 780, 3), (673, 1013, 3), (480, 640, 1), (1001, 1500, 3), (250, 1000, 1)):
        for conf in ({"grayscale": True, "resize_max": 1024, "dfactor": 8}, {"grayscale": False, "resize_max": 500, "dfactor": 16},
                     {"grayscale": True, "resize_max": 1600, "force_resize": True, "width": 640, "height": 480, "dfactor": 8},
                     {"grayscale": True, "resize_max": 375, "dfactor": 1}):
            frame = np.zeros((h, w, 3) if c == 3 else (h, w), np.uint8)
            if c == 1 and not conf["grayscale"]:
                continue
            x, orig, size = op.preprocess(frame[:8, :8] if False else frame, conf) if h * w < 400000 else (None, None, None)
            oc, oh, ow = ops.preprocess_plan(conf, h, w, c)
            assert oc == (1 if (conf["grayscale"] or c == 1) else 3)
            if x is not None:
                assert (oc, oh, ow) == x.shape, (h, w, conf)
            sc = conf["resize_max"] / max(h, w)
            hh, ww = (int(round(h * sc)), int(round(w * sc))) if sc < 1.0 else (h, w)
            if conf.get("force_resize"):
                hh, ww = conf["height"], conf["width"]
            assert (oh, ow) == (hh // conf["dfactor"] * conf["dfactor"], ww // conf["dfactor"] * conf["dfactor"])


def test_drivers_refuse_cpu():
    """extract / match_dense.match_images prepare images on the GPU: a CPU device raises instead of computing elsewhere."""
    from imcui_b200.hloc import extract_features as ef, match_dense as md

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.register_buffer("anchor", torch.zeros(1))
    rgb = np.zeros((48, 64, 3), np.uint8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ef.extract(Fake(), rgb, {"grayscale": True})
    with pytest.raises(RuntimeError, match="CUDA"):
        md.match_images(Fake(), rgb, rgb, {"grayscale": True}, device="cpu")


def test_pair_inputs_containers():
    """match_features.pair_inputs: list / tuple / batched-tensor feature containers -> [1,N,2], [1,N], [1,D,N] (:207-234)."""
    from imcui_b200.hloc import match_features as mf

    def feat(n, as_list):
        k = torch.arange(2 * n, dtype=torch.float32).view(n, 2)
        return {"keypoints": [k] if as_list else k[None], "scores": (torch.ones(n),), "descriptors": [torch.zeros(4, n)],
                "image": torch.zeros(1, 1, 8, 8), "scales": torch.ones(1, n)}
    d = mf.pair_inputs(feat(3, True), feat(2, False))
    assert d["keypoints0"].shape == (1, 3, 2) and d["keypoints1"].shape == (1, 2, 2)
    assert d["scores0"].shape == (1, 3) and d["descriptors1"].shape == (1, 4, 2) and "scales0" in d and "oris0" not in d


def test_set_null_pred_and_registry_parsing():
    from imcui_b200.ui import utils as U
    p = U.set_null_pred("KEYPOINT", {})
    assert p["H"] is None and p["geom_info"] == {} and p["mmkeypoints0_orig"].size == 0      # reference ui/utils.py:382-398
    out = U.filter_matches({"mkeypoints0_orig": np.zeros((3, 2)), "mkeypoints1_orig": np.zeros((3, 2)), "mconf": np.ones(3)})
    assert out["H"] is None and out["geom_info"] == {}                                       # < 4 matches (:502-503)
    zoo = U.get_matcher_zoo({"superpoint+mnn": {"matcher": "NN-mutual", "feature": "superpoint_max", "dense": False, "standalone": False},
                             "loftr": {"matcher": "loftr", "dense": True, "standalone": True}})
    assert zoo["superpoint+mnn"]["feature"]["model"]["name"] == "superpoint" and zoo["superpoint+mnn"]["matcher"]["model"]["name"] == "nearest_neighbor"
    assert zoo["loftr"]["standalone"] and "feature" not in zoo["loftr"] and zoo["loftr"]["matcher"]["model"]["name"] == "loftr"
    with pytest.raises(NotImplementedError):
        U.proc_ransac_matches(np.zeros((9, 2)), np.zeros((9, 2)), "POSELIB")


def test_loftr_weight_packing_layout():
    """BN folding, 196 -> 256 zero padding and the two-plane fp16 split of the LoFTR backbone convolutions."""
    from imcui_b200 import ops
    from oracle import loftr as ol
    sd = ol.random_weights(0)
    pk = ops.loftr_pack_weights(sd)
    c = pk["convs"]["l2.0"]
    assert (c["cin"], c["cout"], c["ksize"], c["stride"]) == (128, 256, 3, 2) and tuple(c["w"].shape) == (2, 9, 256, 128)
    g = sd["backbone.layer2.0.bn1.weight"] / torch.sqrt(sd["backbone.layer2.0.bn1.running_var"] + 1e-5)
    w = (sd["backbone.layer2.0.conv1.weight"] * g[:, None, None, None]).permute(2, 3, 0, 1).reshape(9, 196, 128)
    rec = c["w"][0].float() + c["w"][1].float() / 2048.0
    assert float((rec[:, :196] - w).abs().max()) < 1e-6 * float(w.abs().max()) + 1e-7 and float(rec[:, 196:].abs().max()) == 0.0
    assert float(c["b"][196:].abs().max()) == 0.0 and pk["convs"]["l3_out"]["b"] is None
    assert tuple(pk["conv1_w"].shape) == (49, 128) and tuple(pk["c0.qkv_w"].shape) == (2 * 768, 256) and tuple(pk["f1.mlp0_w"].shape) == (2 * 256, 256)   # [W ; split-fp16 planes]
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
