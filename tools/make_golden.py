"""Mint golden vectors by running the UNMODIFIED reference modules (build container only).

    cd /tmp && python /root/repo/tools/make_golden.py

Writes tests/golden/*.npz.  Imports /root/reference through tools/ref_import.py (SURVEY.md 8(c)
recipe); runs single-threaded-deterministic CPU fp32.  The reference's own tests hold no golden
vectors, so these fixtures are what pins the oracle and the CUDA path.
"""
import importlib.util
import sys
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import ref_import as R  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", ROOT / "image-matching-webui_b200/utils/synth.py")
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)
spec = importlib.util.spec_from_file_location("synth_weights", ROOT / "image-matching-webui_b200/utils/synth_weights.py")
synth_weights = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth_weights)

OUT = ROOT / "tests" / "golden"
torch.set_grad_enabled(False)

SP_CONFS = {
    # ImageMatchingAPI defaults written over superpoint_max (api/core.py:36-38,80-97): test_one path
    "api": {"nms_radius": 3, "max_keypoints": 1024, "keypoint_threshold": 0.015, "remove_borders": 4},
    # hloc/configs/extractors.py:31-37 with the API's 1024 cap (BASELINE config 2)
    "max1024": {"nms_radius": 3, "max_keypoints": 1024, "keypoint_threshold": 0.005, "remove_borders": 4},
    # no cap: row-major keypoint order
    "nocap": {"nms_radius": 4, "max_keypoints": -1, "keypoint_threshold": 0.005, "remove_borders": 4},
    "max2048": {"nms_radius": 3, "max_keypoints": 2048, "keypoint_threshold": 0.005, "remove_borders": 4},
}


def preprocess_like_extract(rgb):
    """hloc/extract_features.py:106-170 with the superpoint_max preprocessing conf
    (grayscale, resize_max 1600, force_resize 640x480, dfactor 8), executed with the same cv2 /
    torchvision calls the reference makes."""
    import torchvision.transforms.functional as TF
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    image = gray.astype(np.float32, copy=False)
    size = image.shape[:2][::-1]
    scale = 1600 / max(size)
    if scale < 1.0:
        size_new = tuple(int(round(x * scale)) for x in size)
        image = cv2.resize(image, size_new, interpolation=cv2.INTER_AREA)
    h, w = image.shape[:2]
    interp = cv2.INTER_AREA if not (w < 640 or h < 480) else cv2.INTER_LINEAR
    image = cv2.resize(image, (640, 480), interpolation=interp)
    image = torch.from_numpy(image[None] / 255.0).float()
    size_new = tuple(int(x // 8 * 8) for x in image.shape[-2:])
    image = TF.resize(image, size=size_new, antialias=True)
    return image[None], np.array(size)


def run_sp(net, image, conf):
    out = net({"image": image}, conf)
    return out


def sp_case(name, images, confs, dense_for=()):
    """images: fp32 [B,1,H,W]"""
    sp_mod = R.superpoint_module()
    net = R.make_superpoint({"nms_radius": 4, "max_keypoints": -1, "keypoint_threshold": 0.005})
    blob = {"images_u8": None, "images": images.numpy()}
    for cname, conf in confs.items():
        for b in range(images.shape[0]):
            out = net({"image": images[b:b + 1]}, conf)
            blob[f"{cname}/{b}/keypoints"] = out["keypoints"][0].numpy().astype(np.int16)
            blob[f"{cname}/{b}/scores"] = out["scores"][0].numpy()
            blob[f"{cname}/{b}/descriptors"] = out["descriptors"][0].numpy()
    # dense intermediates (conf-independent up to NMS radius): score map before NMS
    for b in dense_for:
        x = images[b:b + 1]
        r = net.relu
        h = r(net.conv1a(x)); h = r(net.conv1b(h)); h = net.pool(h)
        h = r(net.conv2a(h)); h = r(net.conv2b(h)); h = net.pool(h)
        h = r(net.conv3a(h)); h = r(net.conv3b(h)); h = net.pool(h)
        h = r(net.conv4a(h)); h = r(net.conv4b(h))
        logits = net.convPb(r(net.convPa(h)))
        s = torch.nn.functional.softmax(logits, 1)[:, :-1]
        bb, _, hh, ww = s.shape
        s = s.permute(0, 2, 3, 1).reshape(bb, hh, ww, 8, 8).permute(0, 1, 3, 2, 4).reshape(bb, hh * 8, ww * 8)
        blob[f"dense/{b}/scores"] = s[0].numpy()
        blob[f"dense/{b}/nms3"] = sp_mod.simple_nms(s, 3)[0].numpy()
        d = torch.nn.functional.normalize(net.convDb(r(net.convDa(h))), p=2, dim=1)
        blob[f"dense/{b}/desc_sub"] = d[0, :, ::4, ::4].numpy()
        blob[f"dense/{b}/feat_sub"] = h[0, :, ::4, ::4].numpy()
    del blob["images_u8"]
    np.savez_compressed(OUT / f"{name}.npz", **blob)
    print("wrote", name, {k: v.shape for k, v in list(blob.items())[:6]})
    return blob


def sp_fix_case(name, images):
    """The reference PLUGIN (imcui/hloc/extractors/superpoint.py, unmodified) with fix_sampling=True: _init swaps the module's
    sample_descriptors for sample_descriptors_fix_sampling (:16-30,46-47)."""
    if str(R.REF) not in sys.path:
        sys.path.insert(0, str(R.REF))
    from imcui.hloc.extractors import superpoint as plug
    plug.SuperPoint._download_model = lambda self, repo_id=None, filename=None: str(R.SP_WEIGHTS)
    import contextlib, io
    conf = {"nms_radius": 3, "max_keypoints": 256, "keypoint_threshold": 0.005, "remove_borders": 4, "fix_sampling": True}
    with contextlib.redirect_stdout(io.StringIO()):
        net = plug.SuperPoint(dict(conf)).eval()
    blob = {"images": images.numpy()}
    for b in range(images.shape[0]):
        out = net({"image": images[b:b + 1]})
        blob[f"{b}/keypoints"] = out["keypoints"][0].numpy().astype(np.int16)
        blob[f"{b}/scores"] = out["scores"][0].numpy()
        blob[f"{b}/descriptors"] = out["descriptors"][0].numpy()
    np.savez_compressed(OUT / f"{name}.npz", **blob)
    print("wrote", name, {k: v.shape for k, v in list(blob.items())[:4]})
    plug.superpoint.sample_descriptors = None  # never reuse the patched module in this process


LG_MODES = {
    # oracle A: full depth, no pruning (deterministic compute graph)
    "full": dict(depth_confidence=-1, width_confidence=-1, prune_th=-1),
    # oracle B: what the reference does on CUDA+flash for <=1536 kpts: early stop only
    "cuda": dict(depth_confidence=0.95, width_confidence=0.99, prune_th=1536),
    # oracle C: reference CPU defaults: early stop + prune at every layer
    "cpu": dict(depth_confidence=0.95, width_confidence=0.99, prune_th=-1),
}


def lg_case(name, pairs, sources):
    """pairs: list of (kpts0 [N,2], desc0 [256,N], kpts1, desc1) numpy; sources: "file:conf:i:j" each."""
    lgm = R.lightglue_module()
    blob = {}
    for mname, mode in LG_MODES.items():
        lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = mode["prune_th"]
        net = R.make_lightglue(filter_threshold=0.2, depth_confidence=mode["depth_confidence"],
                               width_confidence=mode["width_confidence"])
        for p, (k0, d0, k1, d1) in enumerate(pairs):
            data = {
                "image0": {"keypoints": torch.from_numpy(k0).float()[None],
                           "descriptors": torch.from_numpy(d0).t().contiguous()[None]},
                "image1": {"keypoints": torch.from_numpy(k1).float()[None],
                           "descriptors": torch.from_numpy(d1).t().contiguous()[None]},
            }
            out = net(data)
            pre = f"{mname}/{p}/"
            blob[pre + "matches0"] = out["matches0"][0].numpy().astype(np.int32)
            blob[pre + "matches1"] = out["matches1"][0].numpy().astype(np.int32)
            blob[pre + "matching_scores0"] = out["matching_scores0"][0].numpy()
            blob[pre + "matching_scores1"] = out["matching_scores1"][0].numpy()
            blob[pre + "stop"] = np.int32(out["stop"])
            blob[pre + "prune0"] = out["prune0"][0].numpy().astype(np.int32)
            blob[pre + "prune1"] = out["prune1"][0].numpy().astype(np.int32)
            print(name, mname, p, "stop", out["stop"], "matches", int((out["matches0"] > -1).sum()))
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = -1
    # inputs are the reference SuperPoint outputs already stored in sp_*.npz: record where
    blob["sources"] = np.array(sources)
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def lg_proj_case(name, pairs, sources):
    """LightGlue with a Linear input_proj (the aliked / disk architecture, lightglue.py:392-395): 128-d inputs made by
    compressing the stored SuperPoint descriptors with the orthonormal Q of ref_import.lightglue_proj_state."""
    lgm = R.lightglue_module()
    blob = {"sources": np.array(sources)}
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = LG_MODES["cuda"]["prune_th"]
    net, sd, q = R.make_lightglue_proj(128, filter_threshold=0.2, depth_confidence=0.95, width_confidence=0.99)
    blob["input_proj_w"], blob["input_proj_b"], blob["q"] = sd["input_proj.weight"].numpy(), sd["input_proj.bias"].numpy(), q.numpy()
    for p, (k0, d0, k1, d1) in enumerate(pairs):
        c0 = torch.nn.functional.normalize(torch.from_numpy(d0).t() @ q.t(), dim=-1)   # [N,128]
        c1 = torch.nn.functional.normalize(torch.from_numpy(d1).t() @ q.t(), dim=-1)
        out = net({"image0": {"keypoints": torch.from_numpy(k0).float()[None], "descriptors": c0[None]},
                   "image1": {"keypoints": torch.from_numpy(k1).float()[None], "descriptors": c1[None]}})
        pre = f"{p}/"
        blob[pre + "descriptors0"], blob[pre + "descriptors1"] = c0.numpy(), c1.numpy()
        for k in ("matches0", "matches1", "prune0", "prune1"):
            blob[pre + k] = out[k][0].numpy().astype(np.int32)
        blob[pre + "matching_scores0"] = out["matching_scores0"][0].numpy()
        blob[pre + "matching_scores1"] = out["matching_scores1"][0].numpy()
        blob[pre + "stop"] = np.int32(out["stop"])
        print(name, p, "stop", out["stop"], "matches", int((out["matches0"] > -1).sum()))
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = -1
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def lg_so_case(name, pairs, sources):
    """LightGlue with add_scale_ori (the sift / doghardnet architecture, lightglue.py:366-377,500-506): 128-d inputs through
    input_proj as in lg_proj_case, posenc over (x, y, scale, orientation) with Wr = [GIM Wr | seeded extra columns]."""
    lgm = R.lightglue_module()
    blob = {"sources": np.array(sources)}
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = LG_MODES["cuda"]["prune_th"]
    net = lgm.LightGlue(features=None, input_dim=128, add_scale_ori=True, weights=None, filter_threshold=0.2, depth_confidence=0.95, width_confidence=0.99)
    sd, q = R.lightglue_proj_state(128)
    g = torch.Generator().manual_seed(11)
    sd["posenc.Wr.weight"] = torch.cat([sd["posenc.Wr.weight"], 0.3 * torch.randn(32, 2, generator=g)], 1).contiguous()
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("confidence_thresholds" in m for m in missing), (missing, unexpected)
    net.eval()
    blob["input_proj_w"], blob["input_proj_b"], blob["posenc_wr"] = sd["input_proj.weight"].numpy(), sd["input_proj.bias"].numpy(), sd["posenc.Wr.weight"].numpy()
    rng = np.random.default_rng(5)
    for p, (k0, d0, k1, d1) in enumerate(pairs):
        c0 = torch.nn.functional.normalize(torch.from_numpy(d0).t() @ q.t(), dim=-1)
        c1 = torch.nn.functional.normalize(torch.from_numpy(d1).t() @ q.t(), dim=-1)
        so = [torch.from_numpy(rng.uniform(lo, hi, n).astype(np.float32))[None] for n in (len(k0), len(k1)) for lo, hi in ((1.0, 4.0), (-3.14, 3.14))]
        out = net({"image0": {"keypoints": torch.from_numpy(k0).float()[None], "descriptors": c0[None], "scales": so[0], "oris": so[1]},
                   "image1": {"keypoints": torch.from_numpy(k1).float()[None], "descriptors": c1[None], "scales": so[2], "oris": so[3]}})
        pre = f"{p}/"
        blob[pre + "descriptors0"], blob[pre + "descriptors1"] = c0.numpy(), c1.numpy()
        for nm, t in zip(("scales0", "oris0", "scales1", "oris1"), so):
            blob[pre + nm] = t[0].numpy()
        for k in ("matches0", "matches1"):
            blob[pre + k] = out[k][0].numpy().astype(np.int32)
        blob[pre + "matching_scores0"] = out["matching_scores0"][0].numpy()
        blob[pre + "stop"] = np.int32(out["stop"])
        print(name, p, "stop", out["stop"], "matches", int((out["matches0"] > -1).sum()))
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = -1
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def sg_case(name, pairs):
    """pairs: list of (kpts0 [N,2], scores0 [N], desc0 [256,N], kpts1, scores1, desc1, source)."""
    blob = {"sources": np.array([p[-1] for p in pairs])}
    for it in (50, 20):  # hloc conf "superglue": 50 (configs/matchers.py:15); a shorter run pins the iteration count
        net = R.make_superglue({"weights": "outdoor", "sinkhorn_iterations": it, "match_threshold": 0.2})
        for p, (k0, s0, d0, k1, s1, d1, _) in enumerate(pairs):
            data = {"image0": torch.empty(1, 1, 480, 640), "image1": torch.empty(1, 1, 480, 640),
                    "keypoints0": torch.from_numpy(k0)[None], "keypoints1": torch.from_numpy(k1)[None],
                    "scores0": torch.from_numpy(s0)[None], "scores1": torch.from_numpy(s1)[None],
                    "descriptors0": torch.from_numpy(d0)[None], "descriptors1": torch.from_numpy(d1)[None]}
            out = net(data)
            pre = f"it{it}/{p}/"
            blob[pre + "matches0"] = out["matches0"][0].numpy().astype(np.int32)
            blob[pre + "matches1"] = out["matches1"][0].numpy().astype(np.int32)
            blob[pre + "matching_scores0"] = out["matching_scores0"][0].numpy()
            blob[pre + "matching_scores1"] = out["matching_scores1"][0].numpy()
            print(name, it, p, "matches", int((out["matches0"] > -1).sum()))
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def loftr_case(name):
    """LoFTR with the deterministic random weights of utils/synth_weights.py loaded into the UNMODIFIED in-tree
    LoFTR module (third_party/SE2LoFTR/src/loftr); thr lowered so that random weights still yield coarse matches."""
    w = synth_weights.loftr_random_weights(0)
    blob = {}
    for tag, (H, W), thr in (("s", (240, 320), 1e-5), ("m", (480, 640), 1e-6)):
        net = R.make_loftr(0, thr=thr)
        missing, unexpected = net.load_state_dict(w, strict=False)
        assert not unexpected and all("num_batches_tracked" in m for m in missing), (missing, unexpected)
        a, b, _ = synth.make_pair(0, H, W)
        x0 = torch.from_numpy(a.astype(np.float32) / 255.0)[None, None]
        x1 = torch.from_numpy(b.astype(np.float32) / 255.0)[None, None]
        d = {"image0": x0, "image1": x1}
        net(d)
        blob[tag + "/thr"] = np.float32(thr)
        blob[tag + "/hw"] = np.array([H, W])
        blob[tag + "/keypoints0"] = d["mkpts0_f"].numpy(); blob[tag + "/keypoints1"] = d["mkpts1_f"].numpy()
        blob[tag + "/confidence"] = d["mconf"].numpy()
        blob[tag + "/i_ids"] = d["i_ids"].numpy().astype(np.int32); blob[tag + "/j_ids"] = d["j_ids"].numpy().astype(np.int32)
        cm = d["conf_matrix"][0]
        blob[tag + "/conf_rowmax"] = cm.max(1)[0].numpy(); blob[tag + "/conf_colmax"] = cm.max(0)[0].numpy()
        print(name, tag, "matches", len(d["mconf"]), "conf max", float(cm.max()))
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def loftr_hw_case(name):
    """LoFTR on a pair whose two images have DIFFERENT sizes (the module then runs the backbone per image, loftr.py:48-56)."""
    w = synth_weights.loftr_random_weights(0)
    blob = {}
    for tag, hw0, hw1, thr in (("d", (240, 320), (256, 288), 1e-5), ("e", (192, 256), (320, 240), 1e-5)):
        net = R.make_loftr(0, thr=thr)
        net.load_state_dict(w, strict=False)
        a = synth.make_pair(0, *hw0)[0]
        b = synth.make_pair(1, *hw1)[1]
        d = {"image0": torch.from_numpy(a.astype(np.float32) / 255.0)[None, None], "image1": torch.from_numpy(b.astype(np.float32) / 255.0)[None, None]}
        net(d)
        blob[tag + "/thr"] = np.float32(thr); blob[tag + "/hw0"] = np.array(hw0); blob[tag + "/hw1"] = np.array(hw1)
        blob[tag + "/image0"] = a; blob[tag + "/image1"] = b
        blob[tag + "/keypoints0"] = d["mkpts0_f"].numpy(); blob[tag + "/keypoints1"] = d["mkpts1_f"].numpy()
        blob[tag + "/confidence"] = d["mconf"].numpy()
        blob[tag + "/i_ids"] = d["i_ids"].numpy().astype(np.int32); blob[tag + "/j_ids"] = d["j_ids"].numpy().astype(np.int32)
        print(name, tag, "matches", len(d["mconf"]))
    np.savez_compressed(OUT / f"{name}.npz", **blob)


sys.path.insert(0, str(ROOT / "tests"))
from aliked_cases import ALIKED_CASES  # noqa: E402


def aliked_case(name):
    """ALIKED (aliked-n16) with utils/synth_weights.aliked_random_weights loaded into the UNMODIFIED reference module."""
    w = synth_weights.aliked_random_weights(0)
    blob = {}
    for tag, (seed, H, W, rgb, conf) in ALIKED_CASES.items():
        net = R.make_aliked(w, **conf)
        a, _, _ = synth.make_pair(seed, H, W)
        if rgb:
            img = torch.from_numpy(synth.to_rgb(a).astype(np.float32) / 255.0).permute(2, 0, 1)[None]
        else:
            img = torch.from_numpy(a.astype(np.float32) / 255.0)[None, None]
        out = net({"image": img})
        blob[tag + "/keypoints"] = out["keypoints"][0].numpy()
        blob[tag + "/descriptors"] = out["descriptors"][0].numpy()[:: (8 if tag == "dense" else 1)]   # dense: every 8th row
        blob[tag + "/scores"] = out["keypoint_scores"][0].numpy()
        if tag in ("s", "pad"):
            fm, sm = net.extract_dense_map(img if rgb else img.repeat(1, 3, 1, 1))
            blob[tag + "/score_map"] = sm[0, 0].numpy()
            blob[tag + "/feature_map_rows"] = fm[0, :, ::40].numpy()   # every 40th row of the 128-channel map
        print(name, tag, "keypoints", out["keypoints"].shape[1], "score mean", float(out["keypoint_scores"].mean()))
    np.savez_compressed(OUT / f"{name}.npz", **blob)


ALIKED_LG_CONF = {"detection_threshold": 0.1, "max_num_keypoints": 1024, "nms_radius": 2}   # BASELINE configs[3] (bench_configs.Config4)


def aliked_lg_case(name, train_seeds=range(5000, 5016), eval_seeds=(2, 65)):
    """BASELINE configs[3] without its checkpoints (aliked-n16.pth / aliked_lightglue.pth are not in the tree): the
    LightGlue `input_proj` (Linear 128 -> 256, lightglue.py:392-395) is FITTED so that the GIM SuperPoint-LightGlue weights
    see SuperPoint-like descriptors -- ridge regression from the descriptors of the reference ALIKED module (seeded random
    weights, utils/synth_weights.py) to the reference SuperPoint descriptors sampled at the same keypoints, on synthetic
    images outside the bench stream.  With it the composition ALIKED -> LightGlue -> MAGSAC produces hundreds of (mostly
    correct) matches per pair instead of none.  Also stores two evaluation pairs run through the UNMODIFIED reference
    modules end to end (ALIKED features -> LightGlue(input_dim=128) with this input_proj): parity goldens."""
    w = synth_weights.aliked_random_weights(0)
    anet = R.make_aliked(w, **ALIKED_LG_CONF)
    sp_mod = R.superpoint_module()
    snet = R.make_superpoint({"nms_radius": 4, "max_keypoints": -1, "keypoint_threshold": 0.005})

    def feats(gray_u8):
        rgb = torch.from_numpy(synth.to_rgb(gray_u8).astype(np.float32) / 255.0).permute(2, 0, 1)[None]
        out = anet({"image": rgb})
        k, d = out["keypoints"][0], out["descriptors"][0]                       # [N,2] pixels, [N,128]
        x = torch.from_numpy(gray_u8.astype(np.float32) / 255.0)[None, None]
        r = snet.relu
        h = r(snet.conv1a(x)); h = r(snet.conv1b(h)); h = snet.pool(h)
        h = r(snet.conv2a(h)); h = r(snet.conv2b(h)); h = snet.pool(h)
        h = r(snet.conv3a(h)); h = r(snet.conv3b(h)); h = snet.pool(h)
        h = r(snet.conv4a(h)); h = r(snet.conv4b(h))
        dd = torch.nn.functional.normalize(snet.convDb(r(snet.convDa(h))), p=2, dim=1)
        return k, d, sp_mod.sample_descriptors(k[None], dd, 8)[0].t()            # SuperPoint descriptors at the ALIKED keypoints

    X, Y = [], []
    for s_ in train_seeds:
        a, c, _ = synth.make_pair(s_, 480, 640)
        for g in (a, c):
            _, d, y = feats(g)
            X.append(d); Y.append(y)
    X = torch.cat(X).double(); Y = torch.cat(Y).double()
    Xb = torch.cat([X, torch.ones(len(X), 1, dtype=torch.float64)], 1)
    Wb = torch.linalg.solve(Xb.t() @ Xb + 1e-3 * torch.eye(129, dtype=torch.float64), Xb.t() @ Y)
    W, b = Wb[:128].t().float().contiguous(), Wb[128].float().contiguous()
    blob = {"input_proj_w": W.numpy(), "input_proj_b": b.numpy(), "train_seeds": np.array(list(train_seeds)), "eval_seeds": np.array(eval_seeds)}
    lgm = R.lightglue_module()
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = LG_MODES["cuda"]["prune_th"]
    net = lgm.LightGlue(features=None, input_dim=128, weights=None, filter_threshold=0.2, depth_confidence=0.95, width_confidence=0.99)
    sd = dict(R.lightglue_state_dict()); sd["input_proj.weight"], sd["input_proj.bias"] = W, b
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("confidence_thresholds" in m for m in missing)
    net = net.eval()
    for p_, s_ in enumerate(eval_seeds):
        a, c, Hm = synth.make_pair(s_, 480, 640)
        (k0, d0, _), (k1, d1, _) = feats(a), feats(c)
        out = net({"image0": {"keypoints": k0[None], "descriptors": d0[None]}, "image1": {"keypoints": k1[None], "descriptors": d1[None]}})
        pre = f"{p_}/"
        blob[pre + "keypoints0"], blob[pre + "keypoints1"] = k0.numpy(), k1.numpy()
        blob[pre + "descriptors0"], blob[pre + "descriptors1"] = d0.numpy(), d1.numpy()
        blob[pre + "matches0"] = out["matches0"][0].numpy().astype(np.int32)
        blob[pre + "matching_scores0"] = out["matching_scores0"][0].numpy()
        blob[pre + "stop"] = np.int32(out["stop"])
        blob[pre + "H"] = np.asarray(Hm, dtype=np.float64)
        m = blob[pre + "matches0"]; v = m > -1
        q = np.concatenate([k0.numpy()[v], np.ones((v.sum(), 1))], 1) @ blob[pre + "H"].T
        err = np.linalg.norm(q[:, :2] / q[:, 2:] - k1.numpy()[m[v]], axis=1)
        print(name, "seed", s_, "stop", out["stop"], "matches", int(v.sum()), "within 3 px of the ground-truth warp", int((err < 3).sum()))
    lgm.LightGlue.pruning_keypoint_thresholds["cpu"] = -1
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def confs_case(name):
    """The reference's registry entries (imcui/hloc/configs) for every conf name the B200 package provides."""
    import json
    if str(R.REF) not in sys.path:
        sys.path.insert(0, str(R.REF))
    from imcui.hloc.configs import confs_dict as ref
    sys.path.insert(0, str(ROOT))
    from imcui_b200.hloc.configs import confs_dict as mine
    out = {kind: {n: ref[kind][n] for n in mine[kind] if n in ref[kind]} for kind in ("extractors", "matchers")}
    (OUT / f"{name}.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(name, {k: len(v) for k, v in out.items()})


PLUGIN_FILES = {"extractors": ["superpoint", "aliked"], "matchers": ["lightglue", "superglue", "loftr", "nearest_neighbor", "dual_softmax"]}


def plugin_contract_case(name):
    """default_conf / required_inputs of the reference plugin classes, read from the source with `ast` (several of the
    modules cannot be imported here: kornia)."""
    import ast, json
    out = {}
    for kind, mods in PLUGIN_FILES.items():
        for m in mods:
            tree = ast.parse((R.REF / "imcui/hloc" / kind / f"{m}.py").read_text())
            for node in ast.walk(tree):
                if isinstance(node, ast.ClassDef) and any(getattr(b, "id", "") == "BaseModel" for b in node.bases):
                    d = {"class": node.name}
                    for st in node.body:
                        if isinstance(st, ast.Assign) and st.targets[0].id in ("default_conf", "required_inputs"):
                            d[st.targets[0].id] = ast.literal_eval(st.value)
                    out[f"{kind}/{m}"] = d
    (OUT / f"{name}.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(name, {k: v["class"] for k, v in out.items()})


def matcher_case(name, pairs):
    nn_mod, ds_mod = R.hloc_matchers()
    blob = {}
    nn = nn_mod.NearestNeighbor({"do_mutual_check": True})
    nn_ratio = nn_mod.NearestNeighbor({"do_mutual_check": True, "ratio_threshold": 0.9, "distance_threshold": 0.9})
    nn_nomut = nn_mod.NearestNeighbor({"do_mutual_check": False})
    ds = ds_mod.DualSoftMax({"match_threshold": 0.01, "inv_temperature": 20})
    for p, (d0, d1) in enumerate(pairs):
        data = {"descriptors0": torch.from_numpy(d0)[None], "descriptors1": torch.from_numpy(d1)[None]}
        for tag, model in (("nn", nn), ("nn_ratio", nn_ratio), ("nn_nomutual", nn_nomut), ("dsm", ds)):
            out = model(data)
            blob[f"{tag}/{p}/matches0"] = out["matches0"][0].numpy().astype(np.int32)
            blob[f"{tag}/{p}/matching_scores0"] = out["matching_scores0"][0].numpy()
            print(name, tag, p, int((out["matches0"] > -1).sum()), out["matching_scores0"].dtype)
        if p > 0:  # pair 0 = sp_real api descriptors, already stored there
            blob[f"in/{p}/descriptors0"] = d0
            blob[f"in/{p}/descriptors1"] = d1
    np.savez_compressed(OUT / f"{name}.npz", **blob)


def dense_agg_case(name):
    """hloc/match_dense.py aggregate_matches / assign_keypoints / kpids_to_matches0 of the UNMODIFIED reference on seeded
    semi-dense matches of four images (six pairs), run against an in-memory stand-in for h5py.File (h5py is not installed)."""
    import importlib
    import types
    store = {}

    class DS:
        def __init__(self, a): self.a = np.asarray(a)
        def __array__(self, *a, **k): return self.a

    class Grp(dict):
        def create_dataset(self, k, data=None): self[k] = DS(data)

    class File:
        def __init__(self, path, *a, **k): self.d = store.setdefault(str(path), {})
        def __enter__(self): return self
        def __exit__(self, *e): return False
        def __getitem__(self, k): return self.d[k]
        def __contains__(self, k): return k in self.d
        def __delitem__(self, k): del self.d[k]
        def create_group(self, k): self.d[k] = Grp(); return self.d[k]
        def visititems(self, fn): [fn(g, v) for g, grp in self.d.items() for v in grp.values()]
    h5 = types.ModuleType("h5py"); h5.File = File; h5.Dataset = DS
    sys.modules["h5py"] = h5
    sys.path.insert(0, str(R.REF))
    importlib.import_module("imcui.hloc")
    pc = types.ModuleType("pycolmap"); pc.__version__ = "0.0"; sys.modules.setdefault("pycolmap", pc)
    md = importlib.import_module("imcui.hloc.match_dense")
    md.list_h5_names = lambda path: list(store.get(str(path), {}).keys())
    rng = np.random.default_rng(11)
    names = ["q/a.jpg", "db/b.jpg", "db/c.jpg", "db/d.jpg"]
    pairs = [(names[i], names[j]) for i in range(4) for j in range(i + 1, 4)]
    out = {}
    for case, conf, max_kps, fixed in (("sfm", {"max_error": 1, "cell_size": 1}, None, False), ("coarse", {"max_error": 2, "cell_size": 8}, 300, False),
                                       ("loc", {"max_error": 4, "cell_size": 4}, None, True)):
        store.clear()
        mfile = File("m.h5")
        base = {n: rng.uniform([0, 0], [320, 240], (500, 2)).astype(np.float32) for n in names}
        for a, b in pairs:
            k = int(rng.integers(200, 420))
            ia, ib = rng.integers(0, 500, k), rng.integers(0, 500, k)
            g = mfile.create_group(md.names_to_pair(a, b))
            k0 = (base[a][ia] + rng.normal(0, 0.4, (k, 2))).astype(np.float32); k1 = (base[b][ib] + rng.normal(0, 0.4, (k, 2))).astype(np.float32)
            sc = rng.uniform(0.2, 1.0, k).astype(np.float32)
            sc[::7] = sc[1::7][: len(sc[::7])] if len(sc[1::7]) >= len(sc[::7]) else sc[::7]     # some exactly equal scores
            g.create_dataset("keypoints0", data=k0); g.create_dataset("keypoints1", data=k1); g.create_dataset("scores", data=sc)
            for key, v in (("keypoints0", k0), ("keypoints1", k1), ("scores", sc)):
                out[f"{case}/in/{md.names_to_pair(a, b)}/{key}"] = v
        kw = {}
        if fixed:   # localisation: db keypoints are fixed arrays, only the query (name0 of its pairs) is aggregated, unbinned
            cp = {n: rng.uniform([0, 0], [320, 240], (400, 2)).astype(np.float32) for n in names[1:]}
            for n, v in cp.items():
                out[f"{case}/fixed/{n}"] = v
            cpd = md.defaultdict(list); cpd.update(cp)
            kw = dict(required_queries={names[0]}, cpdict=cpd, bindict=md.defaultdict(list))
            use_pairs = [p for p in pairs if p[0] == names[0]]
        else:
            kw = dict(cpdict=md.defaultdict(list), bindict=md.defaultdict(list))
            use_pairs = pairs
        md.aggregate_matches(conf, list(use_pairs), "m.h5", "f.h5", max_kps=max_kps, **kw)
        for a, b in use_pairs:
            g = store["m.h5"][md.names_to_pair(a, b)]
            out[f"{case}/out/{md.names_to_pair(a, b)}/matches0"] = np.asarray(g["matches0"]); out[f"{case}/out/{md.names_to_pair(a, b)}/matching_scores0"] = np.asarray(g["matching_scores0"])
        for n, g in store.get("f.h5", {}).items():
            out[f"{case}/feat/{n}/keypoints"] = np.asarray(g["keypoints"]); out[f"{case}/feat/{n}/score"] = np.asarray(g["score"])
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(f"[golden] {name}: {len(out)} arrays")


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "data").mkdir(exist_ok=True)
    # ---- real pair (tests/data of the reference; BASELINE config 1) -------------------------------
    names = ["02928139_3448003521.jpg", "17295357_9106075285.jpg"]
    imgs = []
    for n in names:
        bgr = cv2.imread(str(R.REF / "tests/data" / n))
        rgb = np.ascontiguousarray(bgr[:, :, ::-1])
        x, orig = preprocess_like_extract(rgb)
        imgs.append(x)
        # keep a small decoded copy of the fixture image for the a1 (pre-processing) parity test
        np.savez_compressed(OUT / "data" / (Path(n).stem + ".npz"), rgb=rgb)
    real = torch.cat(imgs, 0)
    rb = sp_case("sp_real", real, {k: SP_CONFS[k] for k in ("api", "max1024", "nocap")}, dense_for=(0, 1))
    # ---- synthetic pairs (BASELINE config 2 stream, seeds 0..1) -----------------------------------
    a, b = synth.make_pair_batch([0, 1])
    syn_u8 = np.stack([a[0], b[0], a[1], b[1]])
    syn = torch.from_numpy(syn_u8.astype(np.float32) / 255.0)[:, None]
    sb = sp_case("sp_synth", syn, {k: SP_CONFS[k] for k in ("max1024", "max2048")})
    sp_fix_case("sp_fix", torch.cat([real[:1, :, :240, :320], syn[1:2, :, 100:340, 64:384]], 0))

    def pair(blob, conf, i, j):
        return (blob[f"{conf}/{i}/keypoints"].astype(np.float32), blob[f"{conf}/{i}/descriptors"],
                blob[f"{conf}/{j}/keypoints"].astype(np.float32), blob[f"{conf}/{j}/descriptors"])

    lg_case("lg_real", [pair(rb, "api", 0, 1), pair(rb, "nocap", 0, 1)], ["sp_real:api:0:1", "sp_real:nocap:0:1"])
    lg_case("lg_synth", [pair(sb, "max1024", 0, 1), pair(sb, "max1024", 2, 3), pair(sb, "max2048", 0, 1)],
            ["sp_synth:max1024:0:1", "sp_synth:max1024:2:3", "sp_synth:max2048:0:1"])
    lg_proj_case("lg_proj", [pair(sb, "max1024", 0, 1), pair(rb, "api", 0, 1)], ["sp_synth:max1024:0:1", "sp_real:api:0:1"])
    lg_so_case("lg_so", [pair(sb, "max1024", 0, 1), pair(rb, "api", 0, 1)], ["sp_synth:max1024:0:1", "sp_real:api:0:1"])

    def sg_pair(blob, src):
        f, conf, i, j = src.split(":")
        return (blob[f"{conf}/{i}/keypoints"].astype(np.float32), blob[f"{conf}/{i}/scores"], blob[f"{conf}/{i}/descriptors"],
                blob[f"{conf}/{j}/keypoints"].astype(np.float32), blob[f"{conf}/{j}/scores"], blob[f"{conf}/{j}/descriptors"], src)

    sg_case("sg", [sg_pair(rb, "sp_real:api:0:1"), sg_pair(sb, "sp_synth:max1024:0:1")])
    loftr_case("loftr")
    aliked_case("aliked")
    confs_case("confs")
    plugin_contract_case("plugins")
    d0, d1 = synth.make_descriptor_pair(0, n=768, dim=128)
    matcher_case("matchers", [(rb["api/0/descriptors"], rb["api/1/descriptors"]), (d0, d1[:, :700].copy())])
    dense_agg_case("dense_agg")
    loftr_hw_case("loftr_hw")
    aliked_lg_case("aliked_lg")


if __name__ == "__main__":
    if len(sys.argv) > 1:   # python make_golden.py aliked -> only that family
        with torch.no_grad():
            if sys.argv[1] == "sp_fix":
                gr, gs = np.load(OUT / "sp_real.npz")["images"], np.load(OUT / "sp_synth.npz")["images"]
                sp_fix_case("sp_fix", torch.from_numpy(np.concatenate([gr[:1, :, :240, :320], gs[1:2, :, 100:340, 64:384]], 0)))
            elif sys.argv[1] in ("lg_proj", "lg_so"):   # inputs come from the stored SuperPoint goldens
                gb = {n: np.load(OUT / f"{n}.npz") for n in ("sp_synth", "sp_real")}
                pr = lambda f, c, i, j: (gb[f][f"{c}/{i}/keypoints"].astype(np.float32), gb[f][f"{c}/{i}/descriptors"],
                                         gb[f][f"{c}/{j}/keypoints"].astype(np.float32), gb[f][f"{c}/{j}/descriptors"])
                {"lg_proj": lg_proj_case, "lg_so": lg_so_case}[sys.argv[1]](sys.argv[1], [pr("sp_synth", "max1024", 0, 1), pr("sp_real", "api", 0, 1)], ["sp_synth:max1024:0:1", "sp_real:api:0:1"])
            else:
                {"aliked": aliked_case, "loftr": loftr_case, "confs": confs_case, "plugins": plugin_contract_case, "dense_agg": dense_agg_case, "loftr_hw": loftr_hw_case, "aliked_lg": aliked_lg_case}[sys.argv[1]](sys.argv[1])
    else:
        main()
