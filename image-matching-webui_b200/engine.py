"""Batched SuperPoint -> LightGlue pair pipeline (the BASELINE config-2 hot path as one device-resident
stream): images of P pairs in, matches out, no host synchronisation between the stages.

This is what bench.py times and what a stream driver (hloc match_from_paths-style) would call; the
single-pair plugin classes under hloc/ go through the same C-ABI entry points."""
import torch

from . import _lib as L
from . import ops
from .hloc import WEIGHTS_DIR

SP_CONF_DEFAULT = {"nms_radius": 3, "keypoint_threshold": 0.005, "max_keypoints": 1024, "remove_borders": 4}
LG_CONF_DEFAULT = {"depth_confidence": 0.95, "width_confidence": 0.99, "filter_threshold": 0.2, "pruning_min_kpts": 1536,
                   "use_tensor_cores": 1}


class PairEngine:
    def __init__(self, device, n_pairs, height=480, width=640, sp_conf=None, lg_conf=None, n_layers=9):
        self.device = torch.device(device)
        self.P, self.H, self.W = n_pairs, height, width
        self.sp_conf = {**SP_CONF_DEFAULT, **(sp_conf or {})}
        self.lg_conf = {**LG_CONF_DEFAULT, **(lg_conf or {})}
        self.n_layers = n_layers
        L.lib()  # fail loudly now if the CUDA library is missing
        sp_sd = torch.load(str(WEIGHTS_DIR / "superpoint_v1.pt"), map_location="cpu")
        lg_sd = torch.load(str(WEIGHTS_DIR / "superpoint_lightglue.pt"), map_location="cpu")
        self.sp_w = {k: v.to(self.device) for k, v in ops.sp_pack_weights(sp_sd).items()}
        self.lg_w = {k: v.to(self.device) for k, v in ops.lg_pack_weights(lg_sd, n_layers).items()}
        mk = self.sp_conf["max_keypoints"]
        assert mk > 0, "the batched engine needs a keypoint cap"
        self.cap = (mk + 127) // 128 * 128
        S, dev = 2 * n_pairs, self.device
        self.sp_out = {
            "keypoints": torch.zeros(S, self.cap, 2, device=dev),
            "scores": torch.zeros(S, self.cap, device=dev),
            "descriptors": torch.zeros(S, self.cap, 256, device=dev),
            "counts": torch.zeros(2, S, dtype=torch.int32, device=dev),
        }
        self.lg_out = {
            "matches": torch.empty(S, self.cap, dtype=torch.int32, device=dev),
            "scores": torch.empty(S, self.cap, device=dev),
            "stop": torch.empty(n_pairs, dtype=torch.int32, device=dev),
            "prune": torch.empty(S, self.cap, dtype=torch.int32, device=dev),
        }
        # pinned host staging for the end-to-end path
        self.h_images = torch.empty(S, height, width, dtype=torch.uint8).pin_memory()
        self.d_images_u8 = torch.empty(S, height, width, dtype=torch.uint8, device=dev)
        self.h_matches = torch.empty(S, self.cap, dtype=torch.int32).pin_memory()
        self.h_mscores = torch.empty(S, self.cap, dtype=torch.float32).pin_memory()
        self.h_kpts = torch.empty(S, self.cap, 2, dtype=torch.float32).pin_memory()
        self.h_counts = torch.empty(2, S, dtype=torch.int32).pin_memory()
        self.h_stop = torch.empty(n_pairs, dtype=torch.int32).pin_memory()

    def match_device(self, images):
        """images [2P,1,H,W] fp32 on the device, slot 2p+side.  Everything stays on the device."""
        sp = ops.superpoint_forward(self.sp_w, images, self.sp_conf, self.cap, out=self.sp_out)
        lg = ops.lightglue_forward(self.lg_w, self.n_layers, sp["keypoints"], sp["descriptors"], sp["counts"][0],
                                   self.lg_conf, out=self.lg_out)
        return sp, lg

    def to_float(self, images_u8):
        """uint8 -> fp32 in [0,1] exactly as the reference does (extract_features.py:139:
        float64 division, then .float())."""
        return (images_u8.double() / 255.0).float()[:, None]

    def match_host(self, images_u8_pinned=None):
        """End-to-end: pinned uint8 host images -> H2D -> SuperPoint -> LightGlue -> D2H of matches,
        scores, keypoints, counts, stop into pinned host buffers.  One stream synchronisation at the end."""
        src = self.h_images if images_u8_pinned is None else images_u8_pinned
        self.d_images_u8.copy_(src, non_blocking=True)
        sp, lg = self.match_device(self.to_float(self.d_images_u8))
        self.h_matches.copy_(lg["matches"], non_blocking=True)
        self.h_mscores.copy_(lg["scores"], non_blocking=True)
        self.h_kpts.copy_(sp["keypoints"], non_blocking=True)
        self.h_counts.copy_(sp["counts"], non_blocking=True)
        self.h_stop.copy_(lg["stop"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_matches, self.h_mscores, self.h_kpts, self.h_counts, self.h_stop

    @property
    def h2d_bytes(self):
        return self.h_images.numel()

    @property
    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.h_matches, self.h_mscores, self.h_kpts, self.h_counts, self.h_stop))
