"""Seeded synthetic image-pair generator for the BASELINE configs.

Uniform noise is useless for a corner detector, so each pair is a canvas of random filled convex
polygons / ellipses plus band-limited noise, and a homography-warped, photometrically perturbed
copy.  Pure function of (seed, height, width): streams are reproducible on any box.
"""
import cv2
import numpy as np


def make_image(seed: int, height: int = 480, width: int = 640, n_shapes: int = 400) -> np.ndarray:
    """Grayscale uint8 [H,W] canvas for image0 of pair `seed`."""
    rng = np.random.default_rng(1000 + seed)
    img = np.full((height, width), float(rng.uniform(0.2, 0.8)), np.float32)
    for _ in range(n_shapes):
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        r = rng.uniform(4, 60)
        g = float(rng.uniform(0, 1))
        if rng.uniform() < 0.5:
            axes = (int(r), int(max(2, r * rng.uniform(0.3, 1.0))))
            cv2.ellipse(img, (int(cx), int(cy)), axes, float(rng.uniform(0, 180)), 0, 360, g, -1, cv2.LINE_AA)
        else:
            k = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            rad = r * rng.uniform(0.5, 1.0, k)
            pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)
            cv2.fillConvexPoly(img, pts.astype(np.int32), g, cv2.LINE_AA)
    noise = rng.normal(0, 1, (height, width)).astype(np.float32)
    noise = cv2.GaussianBlur(noise, (0, 0), 2.0) * 0.08 * 8.0  # blur shrinks sigma; rescale
    img = np.clip(img + noise, 0, 1)
    return np.round(img * 255).astype(np.uint8)


def make_pair(seed: int, height: int = 480, width: int = 640):
    """(image0, image1, H) with image1 = photometric(warp_H(image0)); uint8 [H,W] each."""
    img0 = make_image(seed, height, width)
    rng = np.random.default_rng(2000 + seed)
    src = np.array([[0, 0], [width - 1, 0], [width - 1, height - 1], [0, height - 1]], np.float32)
    dst = src + rng.uniform(-48, 48, (4, 2)).astype(np.float32) * (min(height, width) / 480.0)
    Hm = cv2.getPerspectiveTransform(src, dst)
    w1 = cv2.warpPerspective(img0, Hm, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    f = w1.astype(np.float32) / 255.0
    f = f * float(rng.uniform(0.8, 1.2)) + float(rng.uniform(-0.1, 0.1))
    f = f + rng.normal(0, 2.0 / 255.0, f.shape).astype(np.float32)
    img1 = np.round(np.clip(f, 0, 1) * 255).astype(np.uint8)
    return img0, img1, Hm


def to_rgb(gray: np.ndarray) -> np.ndarray:
    """Three correlated channels from one gray image (BASELINE config 4 uses RGB input): uint8 [...,H,W] -> [...,H,W,3]."""
    g = gray.astype(np.float32)
    r = np.clip(g * 1.05 + 4.0, 0, 255)
    b = np.clip(g * 0.90 + 12.0, 0, 255)
    return np.round(np.stack([r, g, b], -1)).astype(np.uint8)


def make_pair_batch(seeds, height: int = 480, width: int = 640):
    """uint8 arrays [P,H,W], [P,H,W] for a list of pair seeds."""
    a, b = zip(*[make_pair(s, height, width)[:2] for s in seeds])
    return np.stack(a), np.stack(b)


def make_descriptor_pair(seed: int, n: int = 4096, dim: int = 128, frac: float = 0.5, noise: float = 0.3):
    """Config 5: unit-norm descriptor sets with a planted partial permutation.  -> d0,d1 [dim,n] fp32."""
    rng = np.random.default_rng(3000 + seed)
    d0 = rng.normal(0, 1, (dim, n)).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=0, keepdims=True)
    d1 = rng.normal(0, 1, (dim, n)).astype(np.float32)
    perm = rng.permutation(n)
    sel = rng.uniform(size=n) < frac
    d1[:, perm[sel]] = d0[:, sel] + noise * rng.normal(0, 1, (dim, int(sel.sum()))).astype(np.float32) / np.sqrt(dim)
    d1 /= np.linalg.norm(d1, axis=0, keepdims=True)
    return d0, d1
