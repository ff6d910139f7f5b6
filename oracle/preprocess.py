"""CPU restatement of the image pre-processing of imcui/hloc/extract_features.py:120-162 (and match_dense.py:588-640,
same calls) in plain NumPy -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic lives in third-party libraries that are not in /root/reference: OpenCV (`opencv-python`, unpinned in
requirements.txt; 4.13.0 in this image) for cvtColor / resize and torchvision/ATen for the antialiased dfactor resize.
Each function below restates the published algorithm of the library routine the reference calls and is pinned
bit-for-bit against the installed library by tests/test_preprocess_oracle.py (cv2 / torch are importable on the GPU box
too, so the GPU parity tests check the kernels against the libraries directly as well).

  rgb2gray_u8      cv2.cvtColor(RGB2GRAY) on uint8: 15-bit fixed point (R 9798, G 19235, B 3735, round half up)
  resize_area_f32  cv2.resize(float32, INTER_AREA), both scales >= 1: integer scales -> ResizeAreaFast (sum * 1/area,
                   2x2 through the SIMD path), otherwise the DecimateAlpha tables of computeResizeAreaTab
  resize_linear_f32  cv2.resize(float32, INTER_LINEAR) (the reference falls back to it when up-sampling, :30-31)
  resize_aa_f32    torchvision F.resize(antialias=True) on a CPU float tensor = ATen _upsample_bilinear2d_aa
"""
import math

import numpy as np

F32 = np.float32


def rgb2gray_u8(rgb):
    """cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), uint8 [H,W,3] -> [H,W] (extract_features.py:159-162)."""
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def area_tab(ssize, dsize):
    """computeResizeAreaTab: list of (di, si, alpha float32) for one axis, scale = ssize / dsize as a double."""
    scale = 1.0 / (float(dsize) / float(ssize))
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, F32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, F32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, F32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _is_int_scale(ssize, dsize):
    scale = 1.0 / (float(dsize) / float(ssize))
    i = int(math.floor(scale + 0.5))     # saturate_cast<int>(double) = round
    return abs(scale - i) < np.finfo(np.float64).eps, i


def resize_area_f32(img, dsize, cn=1):
    """cv2.resize(img float32 [H,W], (w, h), interpolation=cv2.INTER_AREA) for w <= W and h <= H.  `cn`: channel count of
    the cv2 call this plane belongs to (the 2x2 SIMD path exists for 1 and 4 channels only, and only for whole
    4-pixel vectors; everything else runs the scalar loop, whose summation order differs)."""
    img = np.ascontiguousarray(img, dtype=F32)
    Hs, Ws = img.shape
    Wd, Hd = dsize
    assert Wd <= Ws and Hd <= Hs, "INTER_AREA restatement covers down-scaling only (the reference switches to LINEAR otherwise)"
    if (Wd, Hd) == (Ws, Hs):
        return img.copy()
    ix, sx = _is_int_scale(Ws, Wd)
    iy, sy = _is_int_scale(Hs, Hd)
    if ix and iy:   # ResizeAreaFast: sum of the sy x sx block (row-major order) times float(1/area); 2x2 goes through the SIMD path
        v = img[: Hd * sy, : Wd * sx].reshape(Hd, sy, Wd, sx)
        simd = None
        if sx == 2 and sy == 2 and cn == 1:
            simd = ((v[:, 0, :, 0] + v[:, 0, :, 1]) + (v[:, 1, :, 0] + v[:, 1, :, 1])) * F32(0.25)
        # generic ResizeAreaFast loop, unrolled by four (CV_ENABLE_UNROLLED): sum += ((S0 + S1) + S2) + S3, then the tail
        flat = v.transpose(0, 2, 1, 3).reshape(Hd, Wd, sy * sx)
        s = np.zeros((Hd, Wd), F32)
        k, area = 0, sx * sy
        while k <= area - 4:
            s = s + (((flat[..., k] + flat[..., k + 1]) + flat[..., k + 2]) + flat[..., k + 3])
            k += 4
        while k < area:
            s = s + flat[..., k]
            k += 1
        res = (s * F32(1.0 / area)).astype(F32)
        if simd is not None:
            nv = Wd // 4 * 4            # whole 128-bit vectors (universal intrinsics at the SSE baseline); the tail is scalar
            res[:, :nv] = simd[:, :nv]
        return res
    xt, yt = area_tab(Ws, Wd), area_tab(Hs, Hd)
    xdi = np.array([t[0] for t in xt]); xsi = np.array([t[1] for t in xt]); xa = np.array([t[2] for t in xt], F32)
    out = np.zeros((Hd, Wd), F32)
    first = np.ones(Hd, bool)
    # horizontal: buf[dx] += S[si] * alpha in table order (separate multiply and add, fp32)
    order = {}
    for k, d in enumerate(xdi):
        order.setdefault(int(d), []).append(k)
    maxlen = max(len(v) for v in order.values())
    for (dy, sy_, beta) in yt:
        S = img[sy_]
        buf = np.zeros(Wd, F32)
        for j in range(maxlen):
            ks = np.array([order[d][j] if j < len(order[d]) else -1 for d in range(Wd)])
            m = ks >= 0
            buf[m] = buf[m] + S[xsi[ks[m]]] * xa[ks[m]]
        if first[dy]:
            out[dy] = beta * buf
            first[dy] = False
        else:
            out[dy] = out[dy] + beta * buf
    return out


def resize_linear_f32(img, dsize):
    """cv2.resize(img float32 [H,W], (w, h), interpolation=cv2.INTER_LINEAR): horizontal pass then vertical pass,
    coefficients (1 - f, f) in fp32 from fx = (dx + 0.5) * scale - 0.5 computed in float."""
    img = np.ascontiguousarray(img, dtype=F32)
    Hs, Ws = img.shape
    Wd, Hd = dsize

    def tab(ssize, dsize_):
        scale = 1.0 / (float(dsize_) / float(ssize))
        idx, a0, a1 = [], [], []
        for d in range(dsize_):
            f = F32((d + 0.5) * scale - 0.5)
            s = int(math.floor(f))
            f = F32(f - s)
            if s < 0:
                s, f = 0, F32(0)
            if s >= ssize - 1:
                s, f = ssize - 1, F32(0)
            idx.append(s); a0.append(F32(1.0) - f); a1.append(f)
        return np.array(idx), np.array(a0, F32), np.array(a1, F32)
    xi, xa0, xa1 = tab(Ws, Wd)
    yi, ya0, ya1 = tab(Hs, Hd)
    xi1 = np.minimum(xi + 1, Ws - 1)
    rows = img[:, xi] * xa0 + img[:, xi1] * xa1               # [Hs, Wd]
    yi1 = np.minimum(yi + 1, Hs - 1)
    return (rows[yi] * ya0[:, None] + rows[yi1] * ya1[:, None]).astype(F32)


def aa_weights(in_size, out_size):
    """ATen _compute_indices_weights_aa for the bilinear (triangle) filter, align_corners=False: per output index
    (xmin, [weights float32])."""
    # every quantity is scalar_t = float in the CPU kernel (area_pixel_compute_scale<float>, aa_filter<float>)
    scale = F32(in_size) / F32(out_size)
    support = F32(1.0) * scale if scale >= 1.0 else F32(1.0)
    invscale = F32(1.0) / scale if scale >= 1.0 else F32(1.0)
    out = []
    for i in range(out_size):
        center = scale * F32(i + 0.5)
        xmin = max(int(center - support + F32(0.5)), 0)
        xsize = min(int(center + support + F32(0.5)), in_size) - xmin
        ws = []
        total = F32(0)
        for j in range(xsize):
            x = abs(F32(F32(j + xmin) - center + F32(0.5)) * invscale)
            w = F32(1.0) - x if x < 1.0 else F32(0)
            ws.append(F32(w))
            total = F32(total + w)
        ws = [F32(w / total) if total != 0.0 else F32(0) for w in ws]
        out.append((xmin, ws))
    return out


def resize_aa_f32(img, size_hw, order="hv"):
    """torchvision.transforms.functional.resize(tensor [C,H,W] float32, size, antialias=True) on the CPU
    (ATen separable_upsample_generic_Nd_kernel_impl: one 1-D pass per axis, fp32 accumulation in tap order)."""
    img = np.ascontiguousarray(img, dtype=F32)
    Hd, Wd = size_hw

    def pass_w(x, Wd_):
        # t = src[0] * w[0]; t = fma(src[j], w[j], t) for j >= 1: the ATen loop `t += src * w` is contracted to FMA in the
        # AVX2 / AVX512 dispatch builds (emulated in float64: the product is exact there, one rounding to fp32 follows)
        tab = aa_weights(x.shape[-1], Wd_)
        out = np.zeros(x.shape[:-1] + (Wd_,), F32)
        for i, (xmin, ws) in enumerate(tab):
            acc = x[..., xmin] * ws[0]
            for j in range(1, len(ws)):
                acc = (x[..., xmin + j].astype(np.float64) * np.float64(ws[j]) + acc.astype(np.float64)).astype(F32)
            out[..., i] = acc
        return out
    for ax in order:
        if ax == "h" and img.shape[-1] != Wd:
            img = pass_w(img, Wd)
        if ax == "v" and img.shape[-2] != Hd:
            img = np.swapaxes(pass_w(np.swapaxes(img, -1, -2), Hd), -1, -2)
    return np.ascontiguousarray(img)


def preprocess(image_u8, conf):
    """extract_features.py:120-162 for one image: uint8 RGB [H,W,3] or gray [H,W] -> (float32 [C,H,W] in [0,1],
    original_size (w,h), size (w,h)).  conf: grayscale, resize_max, force_resize, width, height, dfactor."""
    img = image_u8
    if img.ndim == 3 and conf.get("grayscale", True):
        img = rgb2gray_u8(img)
    chans = [img.astype(F32)] if img.ndim == 2 else [img[..., c].astype(F32) for c in range(img.shape[2])]
    h, w = chans[0].shape
    orig = (w, h)

    def resize(ch, size):
        hh, ww = ch[0].shape
        if ww < size[0] or hh < size[1]:
            return [resize_linear_f32(c, size) for c in ch]
        return [resize_area_f32(c, size, cn=len(ch)) for c in ch]
    if conf.get("resize_max"):
        scale = conf["resize_max"] / max(w, h)
        if scale < 1.0:
            chans = resize(chans, tuple(int(round(x * scale)) for x in (w, h)))
    if conf.get("force_resize"):
        chans = resize(chans, (conf["width"], conf["height"]))
    x = np.stack(chans).astype(F32) / F32(255.0)          # numpy float32 / python float stays float32 (NEP 50)
    df = conf.get("dfactor", 8)
    hh, ww = x.shape[1:]
    new = (int(hh // df * df), int(ww // df * df))
    if new != (hh, ww):
        x = resize_aa_f32(x, new)
    return x, np.array(orig), np.array(x.shape[1:][::-1])
