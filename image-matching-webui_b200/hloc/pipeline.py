"""Device-side pre-/post-processing for the hloc drivers.

The reference prepares every image with cv2 / torchvision on the host and post-processes every pair with `.cpu()` +
NumPy (extract_features.py:120-162, match_features.py:236-257, match_dense.py:588-686).  Here the decoded uint8 frame
is the only thing that crosses PCIe on the way in (imw_preprocess does gray / INTER_AREA / /255 / dfactor alignment on the
device, bit-identical to those libraries) and ONE packed buffer crosses it on the way out (imw_gather_matches /
imw_rescale_keypoints write matched keypoints, rescaled coordinates and confidences next to each other).

  FramePrep      frames (any sizes) -> model inputs; same-size frames share one H2D and one kernel pass
  PairResult     matcher outputs of one pair -> the reference's `match_images` dict through one D2H
"""
from collections import defaultdict

import numpy as np
import torch

from .. import ops

PRE_KEYS = ("grayscale", "resize_max", "dfactor", "force_resize", "width", "height")
PRE_DEFAULT = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "force_resize": False, "width": 320, "height": 240}


def pre_conf(conf):
    """extract_features.py:107-116 / match_dense.py:578-586 defaults merged with the caller's preprocessing conf."""
    return {**PRE_DEFAULT, **{k: v for k, v in dict(conf or {}).items() if k in PRE_KEYS}}


class PinnedPool:
    """Pinned staging buffers keyed by (dtype, size bucket): allocating pinned memory costs milliseconds, reusing it nothing."""

    def __init__(self):
        self.bufs, self.busy = {}, {}

    def get(self, n, dtype):
        key = (dtype, 1 << max(10, int(n - 1).bit_length()))
        b = self.bufs.get(key)
        if b is None:
            b = self.bufs[key] = torch.empty(key[1], dtype=dtype).pin_memory()
        ev = self.busy.pop(key, None)
        if ev is not None:
            ev.synchronize()          # an asynchronous H2D of the previous user may still be reading the buffer
        return b[:n]

    def mark_in_flight(self, n, dtype, device):
        """call after enqueueing an async copy OUT of the buffer returned by get(n, dtype)"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.busy[(dtype, 1 << max(10, int(n - 1).bit_length()))] = ev


_pool = PinnedPool()


class FramePrep:
    """Decoded frames -> the tensors the extractor / dense matcher takes.

    frames: list of uint8 arrays [H,W,3] (RGB) or [H,W] (gray).  Frames of the same shape travel together: one pinned
    staging copy, one H2D, one imw_preprocess call.  Returns, per frame and in input order,
    (image [1,C,H',W'] fp32 on `device`, original_size (w,h), size (w,h))."""

    def __init__(self, conf, device):
        self.conf, self.device = pre_conf(conf), torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"FramePrep: device {self.device}; the B200 engine only runs on CUDA devices (no CPU fallback)")

    def __call__(self, frames):
        groups = defaultdict(list)
        for i, f in enumerate(frames):
            if f.dtype != np.uint8 or f.ndim not in (2, 3):
                raise ValueError(f"frame {i}: expected a decoded uint8 image [H,W] or [H,W,3], got {f.dtype} {f.shape}")
            groups[f.shape].append(i)
        out = [None] * len(frames)
        for shape, idx in groups.items():
            n = int(np.prod(shape))
            stage = _pool.get(n * len(idx), torch.uint8).view(len(idx), *shape)
            for j, i in enumerate(idx):
                stage[j].copy_(torch.from_numpy(np.ascontiguousarray(frames[i])))
            dev_u8 = stage.to(self.device, non_blocking=True)
            _pool.mark_in_flight(n * len(idx), torch.uint8, self.device)
            x = ops.preprocess(dev_u8, self.conf)
            orig = np.array(shape[:2][::-1])
            size = np.array([x.shape[3], x.shape[2]])
            for j, i in enumerate(idx):
                out[i] = (x[j:j + 1], orig, size)
        return out


class PairResult:
    """Post-processing of one matched pair on the device (match_features.py:236-257 / match_dense.py:642-686).

    All keypoints of both images, their coordinates in the original frames, the matched subsets (ascending keypoint
    order, as boolean-mask indexing gives), their original-frame coordinates and the confidences are produced by two
    kernels into ONE packed buffer and fetched with one D2H; the only host synchronisation of the pair."""

    @staticmethod
    def sparse(kpts0, kpts1, matches0, scores0, scale0, scale1):
        """kpts* [N*,2] fp32, matches0 [N0] int (−1 = none), scores0 [N0]; scale* = original_size / size (float64 pair)."""
        dev = kpts0.device
        n0, n1 = int(kpts0.shape[0]), int(kpts1.shape[0])
        cap = max(n0, n1, 1)
        kp = torch.zeros(2, cap, 2, device=dev)
        kp[0, :n0], kp[1, :n1] = kpts0.float(), kpts1.float()
        m = torch.full((2, cap), -1, dtype=torch.int32, device=dev)
        m[0, :n0] = matches0.to(torch.int32)
        sc = torch.zeros(2, cap, device=dev)
        sc[0, :n0] = scores0.float()
        counts = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
        scales = torch.tensor(np.stack([scale0, scale1]).astype(np.float32), device=dev)   # torch multiplies fp32 tensors by the scalar in fp32
        # packed layout (floats): kp [2,cap,2] | kp_orig [2,cap,2] | mk0 | mk1 | mk0_orig | mk1_orig [cap,2] each | mconf [cap] | count (int32 bits)
        packed = torch.empty(17 * cap + 1, device=dev)
        packed[:4 * cap] = kp.view(-1)
        kp_orig = packed[4 * cap:8 * cap].view(2, cap, 2)
        o, views = 8 * cap, {}
        for k in ("mkpts0", "mkpts1", "mkpts0_orig", "mkpts1_orig"):
            views[k] = packed[o:o + 2 * cap].view(1, cap, 2); o += 2 * cap
        views["mconf"] = packed[o:o + cap].view(1, cap); o += cap
        views["mcount"] = packed[o:o + 1].view(torch.int32)
        ops.rescale_keypoints(kp, scales, counts, out=kp_orig)
        ops.gather_matches(kp, m, counts, scores=sc, scales=scales, out=views)
        host = _pool.get(packed.numel(), torch.float32)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h = host.numpy()
        k = int(host[-1:].view(torch.int32)[0])
        g = lambda a, n: h[a:a + 2 * cap].reshape(cap, 2)[:n].copy()
        return {"keypoints0": g(0, n0), "keypoints1": g(2 * cap, n1), "keypoints0_orig": g(4 * cap, n0), "keypoints1_orig": g(6 * cap, n1),
                "mkeypoints0": g(8 * cap, k), "mkeypoints1": g(10 * cap, k),
                "mkeypoints0_orig": g(12 * cap, k), "mkeypoints1_orig": g(14 * cap, k),
                "mconf": h[16 * cap:16 * cap + k].copy()}

    @staticmethod
    def dense(kpts0, kpts1, conf, scale0, scale1):
        """Detector-free matcher output (keypoints0/1 [K,2] already paired): original-frame coordinates through one D2H."""
        dev = kpts0.device
        k = int(kpts0.shape[0])
        cap = max(k, 1)
        kp = torch.zeros(2, cap, 2, device=dev)
        kp[0, :k], kp[1, :k] = kpts0.float(), kpts1.float()
        scales = torch.tensor(np.stack([scale0, scale1]).astype(np.float32), device=dev)
        packed = torch.empty(8 * cap + cap, device=dev)
        packed[:4 * cap] = kp.view(-1)
        ops.rescale_keypoints(kp, scales, None, out=packed[4 * cap:8 * cap].view(2, cap, 2))
        packed[8 * cap:8 * cap + k] = conf.float() if conf is not None else 1.0
        host = _pool.get(packed.numel(), torch.float32)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h = host.numpy()
        g = lambda a: h[a:a + 2 * cap].reshape(cap, 2)[:k].copy()
        return {"keypoints0": g(0), "keypoints1": g(2 * cap), "keypoints0_orig": g(4 * cap), "keypoints1_orig": g(6 * cap),
                "mconf": h[8 * cap:8 * cap + k].copy()}
