"""GPU pre-/post-processing through the C ABI: imw_preprocess against the
libraries the reference calls (cv2 / torchvision, bit-exact) and the NumPy oracle; imw_gather_matches against the
reference's NumPy statements (match_features.py:244-257)."""
import cv2
import numpy as np
import pytest
import torch
import torchvision.transforms.functional as TF

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
NAMES = ["02928139_3448003521", "17295357_9106075285"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rgb(i):
    return np.load(GOLDEN / "data" / f"{NAMES[i]}.npz")["rgb"]


def _reference_preprocess(rgb, conf):
    """The reference's own statements (extract_features.py:120-162) through cv2 / torchvision on the host."""
    c = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "force_resize": False, "width": 320, "height": 240, **conf}
    img = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY) if (c["grayscale"] and rgb.ndim == 3) else rgb
    img = img.astype(np.float32)
    size = img.shape[:2][::-1]

    def rs(im, sz):
        h, w = im.shape[:2]
        interp = cv2.INTER_LINEAR if (w < sz[0] or h < sz[1]) else cv2.INTER_AREA
        return cv2.resize(im, sz, interpolation=interp)
    if c["resize_max"]:
        scale = c["resize_max"] / max(size)
        if scale < 1.0:
            img = rs(img, tuple(int(round(x * scale)) for x in size))
    if c["force_resize"]:
        img = rs(img, (c["width"], c["height"]))
    img = img[None] if img.ndim == 2 else img.transpose(2, 0, 1)
    t = torch.from_numpy(img / 255.0).float()
    new = tuple(int(x // c["dfactor"] * c["dfactor"]) for x in t.shape[-2:])
    return TF.resize(t, size=list(new), antialias=True).numpy()


CONFS = [
    {"grayscale": True, "force_resize": True, "resize_max": 1024, "width": 640, "height": 480, "dfactor": 8},   # test_one / API conf
    {"grayscale": True, "resize_max": 1600, "dfactor": 8},                                                     # superpoint_max: size kept, antialias to x8
    {"grayscale": True, "resize_max": 1024, "dfactor": 8},                                                     # area resize_max + antialias
    {"grayscale": False, "resize_max": 512, "dfactor": 16},                                                    # RGB branch (ALIKED-style)
    {"grayscale": True, "resize_max": 500, "force_resize": True, "width": 250, "height": 187, "dfactor": 1},   # two area stages, odd sizes
]


@pytest.mark.parametrize("conf", CONFS, ids=lambda c: "-".join(f"{k[:2]}{v}" for k, v in c.items()))
@pytest.mark.parametrize("i", [0, 1])
def test_preprocess_bit_identical_to_cv2_and_torchvision(dev, conf, i):
    from imcui_b200 import ops
    rgb = _rgb(i)
    ref = _reference_preprocess(rgb, conf)
    out = ops.preprocess(torch.from_numpy(rgb).to(dev)[None], conf)[0].cpu().numpy()
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert np.array_equal(out, ref), f"max |diff| {np.abs(out - ref).max():.3e} on {(out != ref).sum()} px"


@pytest.mark.parametrize("hw,dst", [((480, 640), (320, 240)), ((480, 640), (160, 120)), ((90, 122), (61, 45)), ((34, 46), (23, 17)),
                                    ((301, 403), (133, 77)), ((480, 640), (640, 480))])
def test_area_resize_odd_sizes_and_integer_scales(dev, hw, dst):
    """integer scales (scalar unrolled order), the 2x2 SIMD order incl. its row tail, general tables, identity -- gray uint8 in."""
    from imcui_b200 import ops
    g = np.random.default_rng(5).integers(0, 256, hw, dtype=np.uint8)
    conf = {"grayscale": True, "resize_max": 0, "force_resize": True, "width": dst[0], "height": dst[1], "dfactor": 1}
    ref = (cv2.resize(g.astype(np.float32), dst, interpolation=cv2.INTER_AREA) / 255.0).astype(np.float32)
    out = ops.preprocess(torch.from_numpy(g).to(dev)[None], conf)[0, 0].cpu().numpy()
    assert np.array_equal(out, ref), f"max |diff| {np.abs(out - ref).max():.3e}"


def test_batch_equals_single_and_oracle(dev):
    from oracle import preprocess as op
    from imcui_b200 import ops
    rng = np.random.default_rng(6)
    frames = rng.integers(0, 256, (5, 211, 317, 3), dtype=np.uint8)
    conf = {"grayscale": True, "resize_max": 200, "dfactor": 8}
    out = ops.preprocess(torch.from_numpy(frames).to(dev), conf).cpu().numpy()
    for b in range(5):
        x, _, _ = op.preprocess(frames[b], conf)
        assert np.array_equal(out[b], x), b


def test_upsampling_branch_follows_opencv_linear(dev):
    """extract_features.py:30-31: INTER_AREA silently becomes INTER_LINEAR when up-sampling.  OpenCV's own arithmetic is
    matched exactly (IPP off); with IPP (pip default) the library result itself moves by ~4e-6 relative."""
    from imcui_b200 import ops
    g = np.random.default_rng(7).integers(0, 256, (100, 120), dtype=np.uint8)
    conf = {"grayscale": True, "resize_max": 0, "force_resize": True, "width": 240, "height": 200, "dfactor": 1}
    out = ops.preprocess(torch.from_numpy(g).to(dev)[None], conf)[0, 0].cpu().numpy()
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        ref = (cv2.resize(g.astype(np.float32), (240, 200), interpolation=cv2.INTER_LINEAR) / 255.0).astype(np.float32)
    finally:
        cv2.ipp.setUseIPP(was)
    assert np.abs(out - ref).max() <= 6e-8, np.abs(out - ref).max()
    ref_ipp = (cv2.resize(g.astype(np.float32), (240, 200), interpolation=cv2.INTER_LINEAR) / 255.0).astype(np.float32)
    assert np.abs(out - ref_ipp).max() < 4e-6


def test_gather_matches_equals_numpy(dev):
    """match_features.py:244-257 restated in NumPy on random matches: order, gather, (k + 0.5) * s - 0.5 in fp32."""
    from imcui_b200 import ops
    rng = np.random.default_rng(8)
    P, cap = 5, 256
    kp = (rng.random((2 * P, cap, 2)) * 600).astype(np.float32)
    counts = rng.integers(0, cap + 1, 2 * P).astype(np.int32)
    counts[2], counts[3] = 0, 17
    m = np.full((2 * P, cap), -1, np.int32)
    sc = rng.random((2 * P, cap)).astype(np.float32)
    for p in range(P):
        n0, n1 = counts[2 * p], counts[2 * p + 1]
        if n0 and n1:
            sel = rng.random(n0) < 0.6
            m[2 * p, :n0][sel] = rng.integers(0, n1, int(sel.sum()))
    orig = rng.integers(300, 2000, (2 * P, 2)); size = rng.integers(200, 800, (2 * P, 2))
    scales = (orig / size)                       # float64, as match_features.py:249-250
    t = lambda a: torch.from_numpy(a).to(dev)
    out = {k: torch.zeros(P, cap, 2, device=dev) for k in ("mkpts0", "mkpts1", "mkpts0_orig", "mkpts1_orig")}
    out["mconf"] = torch.zeros(P, cap, device=dev); out["mcount"] = torch.zeros(P, dtype=torch.int32, device=dev)
    ops.gather_matches(t(kp), t(m), t(counts), scores=t(sc), scales=t(scales.astype(np.float32)), out=out)
    for p in range(P):
        n0 = counts[2 * p]
        k0, k1, m0 = kp[2 * p, :n0], kp[2 * p + 1], m[2 * p, :n0]
        valid = m0 > -1
        mk0, mk1, mconf = k0[valid], k1[m0[valid]], sc[2 * p, :n0][valid]
        n = int(valid.sum())
        assert int(out["mcount"][p]) == n
        assert np.array_equal(out["mkpts0"][p, :n].cpu().numpy(), mk0) and np.array_equal(out["mkpts1"][p, :n].cpu().numpy(), mk1)
        assert np.array_equal(out["mconf"][p, :n].cpu().numpy(), mconf)
        for key, mk, s in (("mkpts0_orig", mk0, scales[2 * p]), ("mkpts1_orig", mk1, scales[2 * p + 1])):
            ref = torch.from_numpy(mk + 0.5)
            ref[:, 0] *= s[0]; ref[:, 1] *= s[1]          # scale_keypoints: fp32 tensor times a scalar
            assert np.array_equal(out[key][p, :n].cpu().numpy(), (ref - 0.5).numpy()), (p, key)


def test_match_images_postprocessing_with_a_fake_matcher(dev):
    """match_images: valid mask, gather and the (k + 0.5) * s - 0.5 rescale (match_features.py:244-257), all container types."""
    from imcui_b200.hloc import match_features as mf

    class Fake(torch.nn.Module):
        def forward(self, data):
            n = data["keypoints0"].shape[1]
            m = torch.full((1, n), -1, dtype=torch.long, device=dev)
            m[0, 0], m[0, 2] = 1, 0
            return {"matches0": m, "matching_scores0": torch.linspace(0, 1, n, device=dev)[None], "stop": 3}

    def feat(n, orig, size):
        return {"keypoints": [torch.arange(2 * n, dtype=torch.float32, device=dev).view(n, 2)], "scores": (torch.ones(n, device=dev),),
                "descriptors": [torch.zeros(4, n, device=dev)], "image": torch.zeros(1, 1, 8, 8, device=dev), "image_orig": np.zeros((2, 2, 3)),
                "original_size": np.array(orig), "size": np.array(size)}
    out = mf.match_images(Fake(), feat(3, (1280, 960), (640, 480)), feat(2, (640, 480), (640, 480)))
    assert out["mkeypoints0"].tolist() == [[0, 1], [4, 5]] and out["mkeypoints1"].tolist() == [[2, 3], [0, 1]]
    assert np.array_equal(out["mkeypoints0_orig"], ((np.array([[0, 1], [4, 5]], np.float32) + 0.5) * 2 - 0.5).astype(np.float32))
    assert out["mkeypoints1_orig"].tolist() == [[2, 3], [0, 1]]
    assert out["keypoints0"].shape == (3, 2) and out["keypoints1_orig"].tolist() == [[0, 1], [2, 3]]
    assert out["mconf"].tolist() == [0.0, 1.0] and out["image0_orig"].shape == (2, 2, 3)
    # no matches at all: empty arrays, not an error
    class Nope(torch.nn.Module):
        def forward(self, data):
            n = data["keypoints0"].shape[1]
            return {"matches0": torch.full((1, n), -1, dtype=torch.long, device=dev), "matching_scores0": torch.zeros(1, n, device=dev)}
    out = mf.match_images(Nope(), feat(3, (640, 480), (640, 480)), feat(2, (640, 480), (640, 480)))
    assert out["mkeypoints0"].shape == (0, 2) and out["mconf"].shape == (0,)


def test_match_dense_postprocessing_with_a_fake_matcher(dev):
    """match_dense.match_images (reference :577-686): force-resize to 640x480 on the GPU, 'scores' adopted as mconf and the
    (k + 0.5) * s - 0.5 rescale with s = original / network size."""
    from imcui_b200.hloc import match_dense as md
    from imcui_b200.hloc.configs import confs_dict
    seen = {}

    class Fake(torch.nn.Module):
        def forward(self, data):
            seen["shape"] = (tuple(data["image0"].shape), tuple(data["image1"].shape))
            return {"keypoints0": torch.tensor([[0.0, 0.0], [8.0, 16.0]], device=dev), "keypoints1": torch.tensor([[1.0, 2.0], [9.5, 18.25]], device=dev),
                    "scores": torch.tensor([0.9, 0.4], device=dev)}
    rgb0 = np.random.RandomState(0).randint(0, 255, (960, 1280, 3), dtype=np.uint8)
    rgb1 = np.random.RandomState(1).randint(0, 255, (480, 640, 3), dtype=np.uint8)
    out = md.match_images(Fake(), rgb0, rgb1, confs_dict["matchers"]["loftr"]["preprocessing"], device=dev)
    assert seen["shape"] == ((1, 1, 480, 640), (1, 1, 480, 640))
    assert out["scale0"].tolist() == [2.0, 2.0] and out["scale1"].tolist() == [1.0, 1.0]
    np.testing.assert_allclose(out["mkeypoints0_orig"], (np.array([[0, 0], [8, 16]]) + 0.5) * 2 - 0.5)
    np.testing.assert_allclose(out["mkeypoints1_orig"], [[1, 2], [9.5, 18.25]])
    assert np.array_equal(out["keypoints0"], out["mkeypoints0"]) and out["mconf"].tolist() == pytest.approx([0.9, 0.4])
    assert out["new_size0"].tolist() == [640, 480] and out["original_size0"].tolist() == [1280, 960]
    # gray conversion + /255 of image 1 (no resize needed) and the two area resizes of image 0 (resize_max 1024: 1280x960 ->
    # 1024x768, then force_resize -> 640x480) are the plain cv2 results
    assert np.array_equal(out["image1"], (cv2.cvtColor(rgb1, cv2.COLOR_RGB2GRAY).astype(np.float32) / 255.0))
    g0 = cv2.resize(cv2.cvtColor(rgb0, cv2.COLOR_RGB2GRAY).astype(np.float32), (1024, 768), interpolation=cv2.INTER_AREA)
    g0 = cv2.resize(g0, (640, 480), interpolation=cv2.INTER_AREA)
    assert np.array_equal(out["image0"], (g0 / 255.0).astype(np.float32))


def test_verify_pairs_batched_and_oversized(dev):
    """verify_pairs: many pairs through two estimator launches == one pair at a time; a set beyond one CTA's capacity
    (LoFTR without a cap can return 16 k matches) still gets a model and a full-length mask."""
    from imcui_b200.ui import utils as U
    rng = np.random.default_rng(9)
    Hm = np.array([[1.02, 0.03, 5.0], [-0.02, 0.98, -3.0], [1e-5, 2e-5, 1.0]])

    def pair(n, outl=0.3):
        x = rng.uniform([0, 0], [640, 480], (n, 2)).astype(np.float32)
        q = np.c_[x, np.ones(n)] @ Hm.T
        y = (q[:, :2] / q[:, 2:]).astype(np.float32) + rng.normal(0, 0.5, (n, 2)).astype(np.float32)
        bad = rng.uniform(size=n) < outl
        y[bad] = rng.uniform([0, 0], [640, 480], (int(bad.sum()), 2)).astype(np.float32)
        return {"mkeypoints0_orig": x, "mkeypoints1_orig": y, "mconf": rng.uniform(size=n).astype(np.float32),
                "image0_orig": np.zeros((480, 640, 3), np.uint8), "image1_orig": np.zeros((480, 640, 3), np.uint8)}, ~bad
    preds, gts = zip(*[pair(n) for n in (300, 900, 5, 2, 14000)])
    out = U.verify_pairs([dict(p) for p in preds], "B200_MAGSAC", 3.0, 0.9999, 10000)
    for p, gt, n in zip(out, gts, (300, 900, 5, 2, 14000)):
        if n < 8:
            assert p["H"] is None and p["geom_info"] == {}
            continue
        assert p["H"].shape == (3, 3) and "Fundamental" in p["geom_info"] and "Homography" in p["geom_info"]
        k = len(p["mmkeypoints0_orig"])
        assert abs(k - int(gt.sum())) <= 0.06 * gt.sum(), (n, k, int(gt.sum()))
        assert p["mmconf"].shape == (k,)
    single = U.filter_matches(dict(preds[1]), "B200_MAGSAC", 3.0, 0.9999, 10000)
    assert abs(len(single["mmkeypoints0_orig"]) - len(out[1]["mmkeypoints0_orig"])) <= 0.03 * 900
