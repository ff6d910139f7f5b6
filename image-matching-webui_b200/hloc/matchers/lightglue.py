"""LightGlue matcher plugin -- drop-in for imcui/hloc/matchers/lightglue.py:14-75.
Same default_conf / required_inputs / output dict; the forward pass runs in libimw_b200.so
(imw_lightglue_forward) instead of third_party/LightGlue/lightglue/lightglue.py."""
from pathlib import Path

import torch

from .. import MODEL_REPO_ID, logger
from ..utils.base_model import BaseModel
from ... import ops


class LightGlue(BaseModel):
    default_conf = {
        "match_threshold": 0.2,
        "filter_threshold": 0.2,
        "width_confidence": 0.99,  # for point pruning
        "depth_confidence": 0.95,  # for early stopping,
        "features": "superpoint",
        "model_name": "superpoint_lightglue.pth",
        "flash": True,  # reference: selects the CUDA pruning threshold 1536 (flash) vs 1024
        "mp": False,
        "add_scale_ori": False,
        "n_layers": 9,
        "state_dict": None,  # B200 engine: in-memory checkpoint (reference key names) instead of model_name
        # B200 engine switch (not in the reference): linear layers on tcgen05.
        # True / "3xtf32": split-precision TF32 (fp32-equivalent, parity-grade); "tf32": single TF32 (fast mode,
        # match-F1 >= 0.99 but scores only within ~1.5e-2); False: fp32 CUDA cores
        "tensor_cores": True,
    }
    required_inputs = ["image0", "keypoints0", "scores0", "descriptors0", "image1", "keypoints1", "scores1",
                       "descriptors1"]
    input_dims = {"superpoint": 256, "aliked": 128, "disk": 128, "raco-aliked": 128, "sift": 128, "doghardnet": 128}   # lightglue.py:350-377
    scale_ori = {"sift", "doghardnet"}                                                                          # add_scale_ori features

    def _init(self, conf):
        logger.info("Loading lightglue model, {}".format(conf["model_name"]))
        if conf["features"] not in self.input_dims:
            raise ValueError(f"Unsupported features: {conf['features']} (B200 engine build: {sorted(self.input_dims)})")
        sd = conf.get("state_dict")   # offline: no aliked_lightglue.pth / disk_lightglue.pth in the tree
        if sd is None:
            model_path = self._download_model(
                repo_id=MODEL_REPO_ID, filename="{}/{}".format(Path(__file__).stem, self.conf["model_name"]))
            sd = torch.load(str(model_path), map_location="cpu")
        self.conf["state_dict"] = None
        self.input_dim = self.input_dims[conf["features"]]
        if (self.input_dim != 256) != ("input_proj.weight" in sd):
            raise ValueError(f"features={conf['features']} needs {'an' if self.input_dim != 256 else 'no'} input_proj in the checkpoint")
        want_so = bool(conf["add_scale_ori"]) or conf["features"] in self.scale_ori
        if (sd["posenc.Wr.weight"].shape[1] == 4) != want_so:
            raise ValueError(f"features={conf['features']}: add_scale_ori={want_so} but posenc.Wr is {tuple(sd['posenc.Wr.weight'].shape)}")
        self.add_scale_ori = want_so
        conf["filter_threshold"] = conf["match_threshold"]  # hloc/matchers/lightglue.py:50
        self.conf["filter_threshold"] = conf["match_threshold"]
        for k, v in ops.lg_pack_weights(sd, conf["n_layers"]).items():
            self.register_buffer(k.replace(".", "__"), v, persistent=False)
        logger.info("Load lightglue model done.")

    def _bufs(self):
        return {k.replace("__", "."): v for k, v in self.named_buffers()}

    def _kernel_conf(self):
        c = self.conf
        pth = c.get("pruning_min_kpts")
        if pth is None:  # lightglue.py:339-344,663-667 on a CUDA device
            pth = 1536 if c["flash"] else 1024
        return {"depth_confidence": c["depth_confidence"], "width_confidence": c["width_confidence"],
                "filter_threshold": c["filter_threshold"], "pruning_min_kpts": pth,
                "use_tensor_cores": {False: 0, True: 1, "3xtf32": 1, "tf32": 2}[c["tensor_cores"]]}

    def _forward(self, data):
        k0, k1 = data["keypoints0"], data["keypoints1"]
        d0, d1 = data["descriptors0"].permute(0, 2, 1), data["descriptors1"].permute(0, 2, 1)  # [1,N,D]
        assert k0.shape[0] == 1 and k1.shape[0] == 1, "one pair per call (reference semantics are B=1)"
        D = self.input_dim
        assert d0.shape[-1] == D and d1.shape[-1] == D  # lightglue.py:510-511
        m, n = k0.shape[1], k1.shape[1]
        dev = k0.device
        cap = max(128, (max(m, n) + 127) // 128 * 128)  # 128-row tiles of the tcgen05 path
        kp = torch.zeros(2, cap, 2, device=dev)
        ds = torch.zeros(2, cap, D, device=dev)
        kp[0, :m], kp[1, :n] = k0[0].float(), k1[0].float()
        ds[0, :m], ds[1, :n] = d0[0].float(), d1[0].float()
        counts = torch.tensor([m, n], dtype=torch.int32, device=dev)
        so = {}
        if self.add_scale_ori:   # hloc/matchers/lightglue.py:61-73 forwards scales0/1, oris0/1 when the extractor provides them
            for name in ("scales", "oris"):
                assert name + "0" in data and name + "1" in data, f"features with add_scale_ori need {name}0 / {name}1"
                t = torch.zeros(2, cap, device=dev)
                t[0, :m], t[1, :n] = data[name + "0"].reshape(-1).float(), data[name + "1"].reshape(-1).float()
                so[name] = t
        out = ops.lightglue_forward(self._bufs(), self.conf["n_layers"], kp, ds, counts, self._kernel_conf(), **so)
        m0, m1 = out["matches"][0, :m].long()[None], out["matches"][1, :n].long()[None]
        ms0, ms1 = out["scores"][0, :m][None], out["scores"][1, :n][None]
        valid = m0[0] > -1
        mi0 = torch.where(valid)[0]
        do_prune = self.conf["width_confidence"] > 0
        prune0, prune1 = out["prune"][0, :m][None], out["prune"][1, :n][None]
        if not do_prune:
            prune0, prune1 = prune0.float(), prune1.float()  # reference: ones_like(mscores) * n_layers
        else:
            prune0, prune1 = prune0.long(), prune1.long()
        return {
            "matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1,
            "stop": int(out["stop"][0]),
            "matches": [torch.stack([mi0, m0[0][valid]], -1)],
            "scores": [ms0[0][valid]],
            "prune0": prune0, "prune1": prune1,
        }

    def match_batch(self, batch):
        """Many pairs in one library call (hloc/pairs_stream.py): batch = {keypoints [2P,cap,2], descriptors [2P,cap,D] token-major,
        scores [2P,cap], counts [2P] int32, image_wh [2P,2]} (slot 2p+side, cap % 128 == 0) -> (matches0 [P,cap] int32,
        matching_scores0 [P,cap]).  Every pair keeps the reference's B = 1 semantics (own early exit, own pruning)."""
        so = {k: batch[k] for k in ("scales", "oris") if self.add_scale_ori}
        out = ops.lightglue_forward(self._bufs(), self.conf["n_layers"], batch["keypoints"], batch["descriptors"], batch["counts"],
                                    self._kernel_conf(), **so)
        return out["matches"][0::2], out["scores"][0::2]
