"""Host-side mirrors of the hloc drivers -- CPU tests."""
from types import SimpleNamespace

import cv2
import numpy as np
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
