"""MAGSAC++ on the GPU vs OpenCV's USAC_MAGSAC (the arithmetic the reference delegates to, imcui/ui/utils.py:352-372).
Parity is statistical: inlier-mask F1, inlier counts, model error on the ground-truth inliers."""
import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _f1(a, b):
    tp = float(np.sum(a & b))
    return 2 * tp / max(float(a.sum() + b.sum()), 1.0)


def _homography_set(seed, n=1000, outliers=0.4, noise=0.5):
    rng = np.random.default_rng(seed)
    src = np.array([[0, 0], [639, 0], [639, 479], [0, 479]], np.float32)
    H = cv2.getPerspectiveTransform(src, src + rng.uniform(-60, 60, (4, 2)).astype(np.float32))
    p0 = rng.uniform([0, 0], [640, 480], (n, 2))
    q = np.c_[p0, np.ones(n)] @ H.T
    p1 = q[:, :2] / q[:, 2:] + rng.normal(0, noise, (n, 2))
    out = rng.uniform(size=n) < outliers
    p1[out] = rng.uniform([0, 0], [640, 480], (int(out.sum()), 2))
    return p0.astype(np.float32), p1.astype(np.float32), ~out, H


def _fundamental_set(seed, n=1000, outliers=0.4, noise=0.5):
    rng = np.random.default_rng(seed)
    K = np.array([[500.0, 0, 320], [0, 500, 240], [0, 0, 1]])
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(4, 10, n)]
    ang = rng.uniform(-0.2, 0.2, 3)
    R, _ = cv2.Rodrigues(ang)
    t = rng.uniform(-1, 1, 3)
    x0 = X @ K.T
    x1 = (X @ R.T + t) @ K.T
    p0 = x0[:, :2] / x0[:, 2:] + rng.normal(0, noise, (n, 2))
    p1 = x1[:, :2] / x1[:, 2:] + rng.normal(0, noise, (n, 2))
    out = rng.uniform(size=n) < outliers
    p1[out] = rng.uniform([0, 0], [640, 480], (int(out.sum()), 2))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F = np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)
    return p0.astype(np.float32), p1.astype(np.float32), ~out, F


def _sampson(F, p0, p1):
    x0, x1 = np.c_[p0, np.ones(len(p0))], np.c_[p1, np.ones(len(p1))]
    Fx0, Ftx1 = x0 @ F.T, x1 @ F
    e = np.sum(x1 * Fx0, 1)
    return np.abs(e) / np.sqrt(Fx0[:, 0] ** 2 + Fx0[:, 1] ** 2 + Ftx1[:, 0] ** 2 + Ftx1[:, 1] ** 2)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("kind", ["Homography", "Fundamental"])
def test_magsac_matches_opencv(kind, seed):
    from imcui_b200.ui import utils as U
    p0, p1, gt, M_true = (_homography_set if kind == "Homography" else _fundamental_set)(seed)
    M, mask = U.proc_ransac_matches(p0, p1, "B200_MAGSAC", 3.0, 0.9999, 10000, kind)
    Mc, maskc = U.proc_ransac_matches(p0, p1, "CV2_USAC_MAGSAC", 3.0, 0.9999, 10000, kind)
    assert M is not None and M.shape == (3, 3) and M.dtype == np.float64 and mask.dtype == bool and mask.shape == (len(p0),)
    f1_cv, f1_gt, f1_cv_gt = _f1(mask, maskc), _f1(mask, gt), _f1(maskc, gt)
    print(f"[magsac] {kind} seed {seed}: inliers ours {int(mask.sum())} cv2 {int(maskc.sum())} gt {int(gt.sum())}; "
          f"F1 vs cv2 {f1_cv:.3f}, vs gt {f1_gt:.3f} (cv2 vs gt {f1_cv_gt:.3f})")
    assert f1_cv > 0.95 and f1_gt >= f1_cv_gt - 0.02
    assert abs(int(mask.sum()) - int(maskc.sum())) <= 0.05 * maskc.sum()
    if kind == "Homography":
        q = np.c_[p0[gt], np.ones(int(gt.sum()))] @ M.T
        err = np.linalg.norm(q[:, :2] / q[:, 2:] - p1[gt], axis=1)
        assert np.median(err) < 1.0
    else:
        assert np.median(_sampson(M, p0[gt], p1[gt])) < 1.0
        assert abs(np.linalg.det(M / np.linalg.norm(M))) < 1e-6  # rank 2


def test_magsac_batched_and_degenerate(dev=None):
    from imcui_b200 import ops
    dev = torch.device("cuda:0")
    sets = [_homography_set(s, n=300 + 100 * s) for s in range(3)]
    cap = 512
    p0 = torch.zeros(4, cap, 2, device=dev); p1 = torch.zeros(4, cap, 2, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    for i, (a, b, _, _) in enumerate(sets):
        p0[i, :len(a)] = torch.from_numpy(a).to(dev); p1[i, :len(a)] = torch.from_numpy(b).to(dev); counts[i] = len(a)
    counts[3] = 3  # too few points: no model, not an error
    M, mask, n_inl, n_it = ops.magsac(p0, p1, counts, "Homography", 3.0, 0.9999, 10000)
    for i, (_, _, gt, _) in enumerate(sets):
        assert _f1(mask[i, :len(gt)].cpu().numpy(), gt) > 0.95
        assert not mask[i, len(gt):].any()
    assert int(n_inl[3]) == 0 and not mask[3].any() and float(M[3].abs().sum()) == 0.0
    assert (n_it[:3] > 0).all() and (n_it[:3] <= 10000 + 8 * 256).all()   # up to 8 CTAs x 256 hypotheses per round


def test_filter_matches_contract_on_real_matches(golden):
    """filter_matches / compute_geometry dict contract (ui/utils.py:459-610) on the tests/data pair (config 1)."""
    from imcui_b200.ui import utils as U
    g, m = golden("sp_real"), golden("matchers")
    k0, k1 = g["api/0/keypoints"].astype(np.float32), g["api/1/keypoints"].astype(np.float32)
    m0 = m["nn/0/matches0"]
    v = m0 > -1
    pred = {"mkeypoints0_orig": k0[v], "mkeypoints1_orig": k1[m0[v]], "mconf": np.ones(int(v.sum()), np.float32),
            "image0_orig": np.zeros((480, 640, 3), np.uint8), "image1_orig": np.zeros((480, 640, 3), np.uint8)}
    ref = U.filter_matches(dict(pred), "CV2_USAC_MAGSAC", 3.0, 0.9999, 10000)
    out = U.filter_matches(dict(pred), "B200_MAGSAC", 3.0, 0.9999, 10000)
    assert set(out) == set(ref) and out["H"].shape == (3, 3)
    assert set(out["geom_info"]) == set(ref["geom_info"])
    n, nr = len(out["mmkeypoints0_orig"]), len(ref["mmkeypoints0_orig"])
    print(f"[magsac] real pair ({int(v.sum())} NN matches): homography inliers ours {n}, cv2 {nr}")
    assert n >= 4 and abs(n - nr) <= max(6, 0.5 * nr)  # 21 +- a few inliers out of 260: low-inlier regime, sampling noise
